"""ctypes binding of libist_b200.so (C ABI declared in include/ist_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``python -m <package>.build``). There is no CPU or
PyTorch fallback on this path: if the shared library is missing, or the device is not an sm_100 GPU, every compute
call raises.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libist_b200.so")
_LIB_OVERRIDE = os.environ.get("IST_B200_LIB")      # A/B runs of two builds on one GPU box (tools only)
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ist_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [HEADER]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/ist_b200.cu for sm_100a into libist_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "ist_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


class LayerDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("cin", ctypes.c_int), ("cout", ctypes.c_int)]


LAYER_CONV3X3_RELU = 0
LAYER_MAXPOOL2X2 = 1

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol of include/ist_b200.h (tests/test_abi.py checks that)
SIGNATURES = {
    "ist_last_error": (ctypes.c_char_p, []),
    "ist_version": (ctypes.c_int, []),
    "ist_device_check": (ctypes.c_int, []),
    "ist_launch_count": (ctypes.c_ulonglong, []),
    "ist_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "ist_profile_begin": (ctypes.c_int, []),
    "ist_profile_end": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), _c_float_p, _c_int_p]),
    "ist_plan_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.POINTER(LayerDesc), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "ist_plan_destroy": (ctypes.c_int, [_vp]),
    "ist_plan_set_weights": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp]),
    "ist_plan_bytes": (ctypes.c_size_t, [_vp]),
    "ist_plan_forward": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "ist_plan_get_feature": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "ist_plan_feature_shape": (ctypes.c_int, [_vp, ctypes.c_int, _c_int_p, _c_int_p, _c_int_p]),
    "ist_plan_get_pool_index": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "ist_plan_gram": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "ist_plan_set_loss": (ctypes.c_int, [_vp, ctypes.c_int, _c_int_p, _c_float_p, ctypes.c_int, _c_int_p, _c_float_p]),
    "ist_plan_set_style_target": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "ist_plan_capture_content_target": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "ist_plan_loss_and_grad": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "ist_plan_backward": (ctypes.c_int, [_vp, ctypes.c_int, _c_int_p, ctypes.POINTER(_vp), _vp, _vp]),
    "ist_lbfgs_create": (ctypes.c_int, [ctypes.POINTER(_vp), _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_double]),
    "ist_lbfgs_destroy": (ctypes.c_int, [_vp]),
    "ist_lbfgs_reset": (ctypes.c_int, [_vp, _vp]),
    "ist_lbfgs_step": (ctypes.c_int, [_vp, _vp, _c_int_p, _c_float_p, _vp]),
    "ist_lbfgs_last_losses": (ctypes.c_int, [_vp, _c_float_p]),
    "ist_lbfgs_frame_state": (ctypes.c_int, [_vp, ctypes.c_int, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p]),
    "ist_lbfgs_set_trace": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int]),
    "ist_lbfgs_trace_count": (ctypes.c_int, [_vp, _c_int_p]),
    "ist_lbfgs_create_test": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_double]),
    "ist_image_resize_target": (ctypes.c_int, [ctypes.c_int] * 3 + [_c_int_p, _c_int_p]),
    "ist_image_post_u8": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 3 + [_c_double_p, _vp]),
    "ist_image_prep_u8": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 3 + [_c_double_p, _vp]),
    "ist_image_resize_u8": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 5 + [_vp]),
    "ist_image_handoff_workspace": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "ist_image_handoff": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_size_t] + [ctypes.c_int] * 5 + [_c_double_p, _vp]),
    "ist_op_conv3x3_relu_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "ist_op_conv3x3_dgrad": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "ist_op_maxpool2x2_fwd": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_maxpool2x2_bwd": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_relu_bwd": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_gram": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_gram_bwd": (ctypes.c_int, [_vp, _vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_gram_mse": (ctypes.c_int, [_vp, _vp, ctypes.c_float, _vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
    "ist_op_mse": (ctypes.c_int, [_vp, _vp, ctypes.c_float, _vp, _vp] + [ctypes.c_int] * 4 + [_vp]),
}

_lib = None


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "This package has no CPU / PyTorch fallback for the style-transfer path.")
    lib = ctypes.CDLL(_LIB_OVERRIDE if _LIB_OVERRIDE else LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if _LIB_OVERRIDE and not hasattr(lib, name):
            continue                      # an older build compared A/B may lack the newest entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class IstError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().ist_last_error()
        raise IstError(f"ist_b200 error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device). Objects bound to a device (plans,
    optimisers) pass their own device, so a plan on cuda:1 works while cuda:0 is current (the C side switches devices itself)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor of `dtype` (float32 unless stated)."""
    import torch
    dtype = torch.float32 if dtype is None else dtype
    if not t.is_cuda:
        raise IstError("expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise IstError(f"expected a contiguous {dtype} tensor")
    return ctypes.c_void_p(t.data_ptr())
