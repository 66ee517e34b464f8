"""Frame sharding across the GPUs of one box: frames are independent optimisation problems, so the only
collective is one end-of-batch gather of the finished images (NCCL over NVLink on GPUs, gloo on CPU for the tests).
The reference processes frames in a serial loop on one GPU (IST/main.py:186-238)."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world_size, local_rank).
    Without RANK/WORLD_SIZE in the environment this is a single-process run and nothing is initialised."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_indices(n_frames, rank, world):
    """Static round-robin: frame i -> rank i mod world (every frame costs the same number of closure evaluations)."""
    return list(range(rank, n_frames, world))


def frames_per_rank(n_frames, world):
    return (n_frames + world - 1) // world


def gather_frames(local, n_frames, rank, world):
    """local: [n_local, ...] results of shard_indices(n_frames, rank, world) in that order.
    Returns [n_frames, ...] in the original frame order on every rank (one all_gather_into_tensor)."""
    if world == 1:
        return local
    per = frames_per_rank(n_frames, world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    out = out.view((world, per) + tuple(local.shape[1:]))
    # frame i lives at [i % world, i // world]
    idx = torch.arange(n_frames, device=local.device)
    return out[idx % world, idx // world]


def gather_results(results, n_frames, rank, world, device):
    """End-of-run gather of finished frames for the batch driver. `results`: {frame index: tensor [H,W,3]} of this rank
    (possibly empty). EVERY rank must call this (ranks with an empty shard included): the ranks first agree, through one
    all_reduce, on whether all frames have one common shape; only then the fixed-shape all_gather runs. Returns the
    [n_frames,H,W,3] batch in frame order, or None when the frames differ in shape (Scale keeps the aspect ratio, so a
    directory may mix sizes) or some frame is missing (skip-existing) — the per-frame files are the result in that case."""
    if world == 1:
        if len(results) != n_frames or len({tuple(t.shape) for t in results.values()}) != 1:
            return None
        return torch.stack([results[i] for i in range(n_frames)])
    big = 1 << 30
    hs = [int(t.shape[0]) for t in results.values()]
    ws = [int(t.shape[1]) for t in results.values()]
    lo = torch.tensor([min(hs) if hs else big, min(ws) if ws else big], dtype=torch.int64, device=device)
    hi = torch.tensor([max(hs) if hs else 0, max(ws) if ws else 0], dtype=torch.int64, device=device)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([len(results)], dtype=torch.int64, device=device)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if int(cnt) != n_frames or int(lo[0]) != int(hi[0]) or int(lo[1]) != int(hi[1]) or int(hi[0]) == 0:
        return None
    mine = shard_indices(n_frames, rank, world)
    if sorted(results.keys()) != mine:
        raise ValueError("gather_results expects the round-robin shard of this rank")
    h, w = int(hi[0]), int(hi[1])
    dtype = next(iter(results.values())).dtype if results else torch.uint8
    local = torch.stack([results[i] for i in mine]) if mine else torch.zeros((0, h, w, 3), dtype=dtype, device=device)
    return gather_frames(local.to(device), n_frames, rank, world)


def max_over_ranks(value, device):
    """Max of a python float over ranks (for device-timed numbers, which are reported as the slowest rank)."""
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
