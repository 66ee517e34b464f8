"""Synthetic job directory for the batch-of-frames workload (BASELINE.json configs[3]): N radar-like 512x512 PNG frames
(seeds 1000 .. 1000+N-1), one lidar-like style PNG (seed 2) and a synthetic vgg_conv.pth (Kaiming-normal, seed 0) with the
reference's state-dict keys. (test / bench infrastructure: uses oracle.synth generators only)

    python tools/make_frames.py DIR [N=256] [SIZE=512]
"""
import os
import sys

import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
os.makedirs(os.path.join(out, "radar"), exist_ok=True)
for i in range(n):
    Image.fromarray(synth.radar_frame(size, 1000 + i), "RGB").save(os.path.join(out, "radar", "%05d.png" % i))
Image.fromarray(synth.lidar_frame(size, 2), "RGB").save(os.path.join(out, "style.png"))
torch.save({k: torch.from_numpy(v) for k, v in synth.vgg_state_dict(0).items()}, os.path.join(out, "vgg_conv.pth"))
print("wrote", n, "frames,", "style.png, vgg_conv.pth to", out)
