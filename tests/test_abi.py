"""The C-ABI library builds, loads and exports every symbol include/ist_b200.h declares; without a GPU every compute
entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import ist_b200
from ist_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ist_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ist_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in ist_b200.h but not exported by libist_b200.so"


def test_python_binding_covers_header(lib):
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()


def test_version_and_error_string(lib):
    assert lib.ist_version() >= 1
    assert isinstance(lib.ist_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    assert lib.ist_device_check() == 3          # IST_ERR_DEVICE
    h = ctypes.c_void_p()
    layers = (_lib.LayerDesc * 1)()
    layers[0].kind, layers[0].cin, layers[0].cout = 0, 3, 64
    rc = lib.ist_plan_create(ctypes.byref(h), 1, layers, 1, 16, 16)
    assert rc == 3 and not h.value
    with pytest.raises(_lib.IstError):
        _lib.check(rc)
    # the Python mirror refuses CPU tensors instead of computing on the host
    from ist_b200.config import get_cfg_defaults
    from ist_b200.model import build_model
    vgg = build_model(get_cfg_defaults())
    with pytest.raises(_lib.IstError):
        vgg(torch.zeros(1, 3, 16, 16), ["relu1_1"])


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "can-image-style-transfer-save-automotive-radar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
