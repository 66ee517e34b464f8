"""The conv kernel ships as CTA pairs (tcgen05 cta_group::2, IST_B200_CONV=pair, default); the single-CTA variant of the same
source (IST_B200_CONV=halo) and the forward chain lengths (IST_B200_PROMOTE_FWD) stay selectable. The switches are read once
per process, so each configuration runs the closure parity tests in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"IST_B200_CONV": "halo"}, {"IST_B200_CONV": "pair", "IST_B200_PROMOTE_FWD": "1"},
                                 {"IST_B200_NO_STREAMK": "1"}, {"IST_B200_NO_PDL": "1"}, {"IST_B200_CFD": "cuda", "IST_B200_CFF": "cuda"},
                                 # round 2 switches: weight tiles reloaded per tile, one Gram launch per layer, no side stream
                                 {"IST_B200_B_RESIDENT": "0", "IST_B200_GRAM_MULTI": "0", "IST_B200_NO_OVERLAP": "1"}])
def test_closure_parity_in_other_configurations(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_closure_gpu.py"), "-x", "-q",
                        "-k", "golden or live_oracle or bitwise"], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
