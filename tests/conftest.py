import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; the parity tests proper, run through the C ABI")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected with -m gpu on a box without a GPU; in a plain run on a CPU box they
    # are skipped so that `pytest tests/` stays usable.
    import torch
    if torch.cuda.is_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if selected_gpu:
        return
    skip = pytest.mark.skip(reason="no GPU in this container (run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    import ist_b200
    ist_b200.build()
    return ist_b200.load()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {
        "64": np.load(os.path.join(GOLDEN, "closure_64.npz")),
        "48x80": np.load(os.path.join(GOLDEN, "closure_48x80.npz")),
        "modules": np.load(os.path.join(GOLDEN, "modules.npz")),
    }
