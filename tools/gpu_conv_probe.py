"""Phase-stamp probe of single conv launches (run under gpurun with IST_B200_DBG_TIMES=1):
python tools/gpu_conv_probe.py  -> the library prints per-launch average clock stamps to stderr."""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ist_lib", os.path.join(ROOT, "can-image-style-transfer-save-automotive-radar_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
lib = L.load()
dev = torch.device("cuda:0")
cases = [(64, 64, 512), (128, 128, 256), (256, 256, 128), (512, 512, 64), (512, 512, 32)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (cin, cout, S) in cases:
    x = torch.relu(torch.randn(1, cin, S, S, device=dev) * 40)
    w = torch.randn(cout, cin, 3, 3, device=dev) * (2.0 / (9 * cin)) ** 0.5
    b = torch.zeros(cout, device=dev)
    y = torch.empty(1, cout, S, S, device=dev)
    for _ in range(3):
        L.check(lib.ist_op_conv3x3_relu_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), 1, cin, cout, S, S, 1, L.stream_ptr()))
    torch.cuda.synchronize()
