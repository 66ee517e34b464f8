"""Frame pipeline for the batch-of-frames loop of IST/main.py:184-238.

The reference handles one frame at a time, serially: PNG decode -> preprocess on the host -> upload -> 300 L-BFGS
evaluations -> download -> PNG encode -> next frame; the GPU idles during every decode / encode and the host idles during
every optimisation. Here the three stages overlap:

  loader thread   PNG decode (PIL), copy into pinned memory, H2D of the 8-bit image on a side stream   frame i + 1 .. i + depth
  main thread     resize / preprocess on the device, targets, the device L-BFGS (one host sync per       frame i
                  optimizer.step), 8-bit post-processing on the device, D2H into a pinned buffer
  writer threads  wait for the frame's D2H event, PNG encode, write                                        frame i - 1 ..

Frames are independent optimisation problems (every frame its own L-BFGS state): `frames_per_batch` > 1 runs
that many equally sized frames side by side in one plan, bit-identical to running them one by one. Results are written as
`<output_dir>/<frame name>.png` like the reference (main.py:230); with `skip_existing` frames whose output already exists
are not recomputed (a restarted job continues where it stopped).
"""
import collections
import os
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image
from torch.autograd import Variable

from .data import DeviceImageTransform
from .model.engine.utils import optimize


class FramePipeline:
    def __init__(self, cfg, model, device, style_image, output_dir, frames_per_batch=1, prefetch=2, writers=2,
                 keep_results=False, max_iterations=None):
        self.cfg, self.model = cfg, model
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.output_dir = output_dir
        self.fpb = max(1, int(frames_per_batch))
        self.depth = max(1, int(prefetch)) * self.fpb
        self.keep_results = keep_results
        self.max_iterations = int(max_iterations if max_iterations is not None else cfg.LOSS.MAX_ITER)
        self.tf = DeviceImageTransform(cfg.DATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN, self.device)
        self.copy_stream = torch.cuda.Stream(self.device)
        # worker threads start with device 0 current: bind them to this pipeline's device (pinned allocations and event waits
        # would otherwise create a context on GPU 0 from every rank)
        bind = lambda: torch.cuda.set_device(self.device)
        self.loader = ThreadPoolExecutor(max_workers=1, thread_name_prefix="ist-load", initializer=bind)
        self.writer = ThreadPoolExecutor(max_workers=max(1, int(writers)), thread_name_prefix="ist-write", initializer=bind)
        # one shared style image (main.py:184-185): transformed once; its Gram targets are cached by optimize()
        self.style = self.tf.preparation(style_image).unsqueeze(0)
        self.results = {}            # frame index -> uint8 [H,W,3] device tensor (keep_results)
        self.stats = collections.Counter()
        self._lock = threading.Lock()
        os.makedirs(output_dir, exist_ok=True)

    # ---- stage 1: decode + upload (loader thread) ---------------------------------------------------------------------------
    def _load(self, path):
        t0 = time.perf_counter()
        arr = np.asarray(Image.open(path).convert('RGB'))                 # main.py:205-206
        host = torch.from_numpy(np.array(arr, copy=True, order="C")).pin_memory()
        with torch.cuda.stream(self.copy_stream):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        with self._lock:
            self.stats["decode_s"] += time.perf_counter() - t0
            self.stats["h2d_bytes"] += host.numel()
        return dev, ev, host                                               # `host` stays referenced until the copy is done

    # ---- stage 3: download + encode (writer threads) ------------------------------------------------------------------------
    def _write(self, host_u8, ev, paths):
        ev.synchronize()
        t0 = time.perf_counter()
        arr = host_u8.numpy()
        for k, p in enumerate(paths):
            Image.fromarray(arr[k], "RGB").save(p)                          # main.py:230
        with self._lock:
            self.stats["encode_s"] += time.perf_counter() - t0
            self.stats["written"] += len(paths)

    def out_path(self, frame_path):
        return os.path.join(self.output_dir, os.path.splitext(os.path.basename(frame_path))[0] + ".png")

    # ---- stage 2 + orchestration -------------------------------------------------------------------------------------------------
    def run(self, frame_paths, indices=None, skip_existing=False):
        """Processes `frame_paths[i]` for i in `indices` (default: all), in that order. Returns the list of indices done."""
        indices = list(range(len(frame_paths))) if indices is None else list(indices)
        todo = [i for i in indices if not (skip_existing and os.path.exists(self.out_path(frame_paths[i])))]
        self.stats["skipped"] += len(indices) - len(todo)
        pending = collections.deque()              # (index, future) in submission order
        nxt = 0
        main_stream = torch.cuda.current_stream(self.device)
        writes = []
        t_gpu = 0.0
        while nxt < len(todo) or pending:
            while nxt < len(todo) and len(pending) < self.depth:
                pending.append((todo[nxt], self.loader.submit(self._load, frame_paths[todo[nxt]])))
                nxt += 1
            # a group = up to frames_per_batch consecutive frames of equal (transformed) size
            group, tensors = [], []
            while pending and len(group) < self.fpb:
                i, fut = pending[0]
                dev, ev, host = fut.result()
                main_stream.wait_event(ev)
                dev.record_stream(main_stream)                             # allocated on the copy stream, consumed here
                x = self.tf.preparation(dev)                               # resize + BGR/mean/x255 on the device
                if tensors and tuple(x.shape) != tuple(tensors[0].shape):
                    break                                                  # different size: starts the next group
                pending.popleft()
                group.append(i)
                tensors.append(x)
                del host
            t0 = time.perf_counter()
            content = torch.stack(tensors).contiguous()
            optimized = Variable(content.clone(), requires_grad=True)
            optimize(self.model, content, self.style, optimized, self.cfg, self.max_iterations)
            rgb = self.tf.post_u8(optimized.data)                          # uint8 [B,H,W,3] on the device
            host = torch.empty(rgb.shape, dtype=torch.uint8).pin_memory()
            host.copy_(rgb, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main_stream)
            t_gpu += time.perf_counter() - t0
            with self._lock:
                self.stats["d2h_bytes"] += host.numel()
                self.stats["evals"] += int(self.model.last_evals) * len(group)
            if self.keep_results:
                for k, i in enumerate(group):
                    self.results[i] = rgb[k]
            writes.append(self.writer.submit(self._write, host, ev, [self.out_path(frame_paths[i]) for i in group]))
        for w in writes:
            w.result()
        with self._lock:
            self.stats["optimize_s"] += t_gpu
            self.stats["frames"] += len(todo)
        return todo

    def close(self):
        self.loader.shutdown(wait=True)
        self.writer.shutdown(wait=True)
