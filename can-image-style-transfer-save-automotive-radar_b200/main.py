"""Entry point with the reference's functions (IST/main.py:23-248): get_model, transfer_style and the batch-of-frames
driver. Differences, all on the host side: `--config-file` is honoured (the reference parses and ignores it,
main.py:103-116), data locations are arguments instead of hard-coded /home/dj paths (main.py:119,142-143), and when
launched under torchrun the sorted frame list is sharded round-robin over the GPUs with one final gather.

    python -m ist_b200.main --content-dir radar/ --style-img lidar/09033.png --output-dir out/
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m ist_b200.main --content-dir ... --style-img ...
"""
import argparse
import glob
import os
import time

import numpy as np
import torch
import torch.nn as nn
from PIL import Image

from .config import get_cfg_defaults
from .model import build_model
from .model.engine import do_transfer_style
from .model.engine.hr_transfer_style import do_hr_transfer_style
from .model.meta_arch import GramMSELoss, StyleTransfer
from .parallel import gather_results, init_distributed, shard_indices
from .pipeline import FramePipeline
from .util.logger import setup_logger


def get_model(cfg, state_dict=None):
    """main.py:23-44. `state_dict` overrides cfg.MODEL.WEIGHTS (there is no network here to fetch vgg_conv.pth)."""
    vgg_model = build_model(cfg)
    device = torch.device(cfg.MODEL.DEVICE)
    vgg_model.to(device)
    if state_dict is None:
        if not os.path.exists(cfg.MODEL.WEIGHTS):
            raise FileNotFoundError(f"{cfg.MODEL.WEIGHTS} not found: place vgg_conv.pth there (IST/util/download_models.sh) "
                                    "or pass a state_dict")
        state_dict = torch.load(cfg.MODEL.WEIGHTS, map_location="cpu")
    vgg_model.load_state_dict(state_dict)
    for param in vgg_model.parameters():
        param.requires_grad = False

    loss_layers = cfg.LOSS.STYLE_LAYERS + cfg.LOSS.CONTENT_LAYERS
    loss_functions = [GramMSELoss()] * len(cfg.LOSS.STYLE_LAYERS) + [nn.MSELoss()] * len(cfg.LOSS.CONTENT_LAYERS)
    loss_functions = [loss_function.to(device) for loss_function in loss_functions]
    loss_weights = cfg.LOSS.STYLE_WEIGHTS + cfg.LOSS.CONTENT_WEIGHTS

    model = StyleTransfer(vgg_model, loss_layers, loss_functions, loss_weights)
    return model, device


def transfer_style(cfg, high_resolution=False, state_dict=None):
    """main.py:47-75: one content / style pair from the config; optional coarse-to-fine second stage."""
    model, device = get_model(cfg, state_dict)
    content_image = Image.open(cfg.DATA.CONTENT_IMG_PATH)
    style_image = Image.open(cfg.DATA.STYLE_IMG_PATH)
    out_image = do_transfer_style(cfg, model, content_image, style_image, device)
    if high_resolution:
        out_image = do_hr_transfer_style(cfg, model, content_image, style_image, out_image, device)
    return out_image


def transfer_style_schedule(cfg, sizes, state_dict=None):
    """Coarse-to-fine schedule over several sizes, e.g. [512, 1024, 2048]: the reference's single hr stage
    (hr_transfer_style.py:11-33) applied once per extra size, each hand-off through the 8-bit clamped image."""
    model, device = get_model(cfg, state_dict)
    content_image = Image.open(cfg.DATA.CONTENT_IMG_PATH)
    style_image = Image.open(cfg.DATA.STYLE_IMG_PATH)
    cfg = cfg.clone()
    cfg.DATA.IMG_SIZE = sizes[0]
    out_image, x = do_transfer_style(cfg, model, content_image, style_image, device, return_tensor=True)
    for s in sizes[1:]:
        cfg.HRDATA.IMG_SIZE = s
        # the previous stage's result stays on the device: 8-bit clamp + bilinear up-scaling + re-preprocessing as kernels
        out_image, x = do_hr_transfer_style(cfg, model, content_image, style_image, x, device, return_tensor=True)
    return out_image


def main(argv=None):
    parser = argparse.ArgumentParser(description="PyTorch Style Transfer -- Content and Style Reconstruction (B200 path)")
    parser.add_argument("--config-file", default="", metavar="FILE", help="path to config file", type=str)
    parser.add_argument("--content-dir", default="", help="directory of content (radar) frames, *.png")
    parser.add_argument("--style-img", default="", help="the shared style (lidar) image")
    parser.add_argument("--output-dir", default="./output/full_transfer/", help="where the stylised frames go")
    parser.add_argument("--max-frames", type=int, default=0, help="process at most this many frames (0 = all)")
    parser.add_argument("--high-resolution", action="store_true", help="run the coarse-to-fine second stage (serial per frame)")
    parser.add_argument("--frames-per-batch", type=int, default=4,
                        help="optimise this many (independent, equally sized) frames side by side on each GPU")
    parser.add_argument("--prefetch", type=int, default=2, help="batches decoded and uploaded ahead of the GPU")
    parser.add_argument("--skip-existing", action="store_true", help="do not recompute frames whose output file exists")
    parser.add_argument("--no-gather", action="store_true", help="skip the end-of-run gather of the finished frames (files only)")
    parser.add_argument("--gather-limit-mb", type=int, default=4096,
                        help="skip the gather (files only) when the finished 8-bit frames of the whole job exceed this many MiB")
    parser.add_argument("--summary-json", default="", help="rank 0 writes a one-line JSON summary (frames/s incl. I/O) here")
    parser.add_argument("opts", help="Modify config options using the command-line", default=None, nargs=argparse.REMAINDER)
    args = parser.parse_args(argv)

    t_all = time.time()
    cfg = get_cfg_defaults()
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    if args.opts:
        cfg.merge_from_list(args.opts)
    rank, world, local_rank = init_distributed()
    if cfg.MODEL.DEVICE == "cuda" and world > 1:
        cfg.MODEL.DEVICE = "cuda:%d" % local_rank
    cfg.OUTPUT.DIR = args.output_dir if args.output_dir.endswith("/") else args.output_dir + "/"
    cfg.freeze()
    os.makedirs(cfg.OUTPUT.DIR, exist_ok=True)
    logger = setup_logger("style-transfer", cfg.OUTPUT.DIR if rank == 0 else False, rank)
    logger.info(args)

    model, device = get_model(cfg)
    if device.type == 'cuda' and device.index is None:      # MODEL.DEVICE = 'cuda' (defaults.py:14): the current device
        device = torch.device('cuda', torch.cuda.current_device())
    torch.cuda.set_device(device)
    style_image = Image.open(args.style_img or cfg.DATA.STYLE_IMG_PATH).convert('RGB')      # one shared style, main.py:184-185
    frames = sorted(glob.glob(os.path.join(args.content_dir, "*.png"))) if args.content_dir else [cfg.DATA.CONTENT_IMG_PATH]
    if args.max_frames > 0:
        frames = frames[: args.max_frames]
    mine = shard_indices(len(frames), rank, world)
    # the gathered batch lives on every GPU: the same decision on every rank, from the job size alone
    job_mib = len(frames) * 3 * int(cfg.DATA.IMG_SIZE) ** 2 * 2 / 2 ** 20          # x2: Scale keeps the aspect ratio (non-square frames)
    if job_mib > args.gather_limit_mb:
        args.no_gather = True

    if args.high_resolution:
        # the coarse-to-fine second stage re-reads the content image at HRDATA.IMG_SIZE: plain per-frame loop (main.py:59-75)
        done = []
        for i in mine:
            out_path = os.path.join(cfg.OUTPUT.DIR, os.path.splitext(os.path.basename(frames[i]))[0] + ".png")
            if args.skip_existing and os.path.exists(out_path):
                continue
            content_image = Image.open(frames[i]).convert('RGB')
            out_image, x = do_transfer_style(cfg, model, content_image, style_image, device, return_tensor=True, save=False)
            out_image = do_hr_transfer_style(cfg, model, content_image, style_image, x, device, save=False)
            out_image.save(out_path)
            done.append(i)
        results, stats = {}, {"frames": len(done), "skipped": len(mine) - len(done)}
    else:
        pipe = FramePipeline(cfg, model, device, style_image, cfg.OUTPUT.DIR, frames_per_batch=args.frames_per_batch,
                             prefetch=args.prefetch, keep_results=(not args.no_gather) and (world > 1 or bool(args.summary_json)))
        done = pipe.run(frames, mine, skip_existing=args.skip_existing)
        pipe.close()
        results, stats = pipe.results, dict(pipe.stats)
    logger.info("rank %d: %d frames done, %d skipped (%s)" % (rank, len(done), stats.get("skipped", 0),
                                                             ", ".join("%s=%.3g" % kv for kv in sorted(stats.items()))))

    # every rank enters the collectives, also with an empty shard (n_frames < world) or mixed frame sizes
    gathered = None
    if not args.no_gather and not args.high_resolution:
        gathered = gather_results(results, len(frames), rank, world, device)
        if rank == 0:
            logger.info("gathered %s" % ("%d frames of shape %s on rank 0" % (gathered.shape[0], tuple(gathered.shape[1:]))
                                          if gathered is not None else "nothing (frames differ in size or were skipped: the files are the result)"))
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    wall = time.time() - t_all
    if rank == 0:
        n = sum(1 for _ in frames)
        logger.info("avg time per frame: %f (%d frames on %d GPU(s), %.2f frames/s incl. model load, style target, I/O and gather)"
                    % (wall / max(1, n), n, world, n / wall))
        if args.summary_json:
            import hashlib
            import json
            digest = hashlib.sha256(gathered.cpu().numpy().tobytes()).hexdigest() if gathered is not None else None
            with open(args.summary_json, "w") as f:
                f.write(json.dumps({"frames": n, "n_gpus": world, "wall_s": wall, "frames_per_s": n / wall,
                                    "frames_per_batch": args.frames_per_batch, "gathered_sha256": digest,
                                    "rank0_stats": {k: float(v) for k, v in stats.items()}}) + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
