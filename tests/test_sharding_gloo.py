"""N > 1 path on CPU: world_size-2 gloo processes shard a frame list round-robin and gather the results in frame order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ist_b200.parallel import gather_frames, shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_frames, rank, world)
    # "result" of frame i is a tensor filled with i (stands for the optimised image)
    local = torch.stack([torch.full((3, 4, 5), float(i)) for i in mine]) if mine else torch.zeros(0, 3, 4, 5)
    full = gather_frames(local, n_frames, rank, world)
    ok = full.shape[0] == n_frames and all(bool((full[i] == i).all()) for i in range(n_frames))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather(tmp_path):
    for n_frames in (7, 4):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
        assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


def _worker_results(rank, world, port, n_frames, mixed, out_dir):
    """gather_results (the batch driver's end-of-run gather, IST/main.py:184-238 has none): every rank enters the collectives,
    also with an empty shard (n_frames < world) or when the frames differ in size."""
    from ist_b200.parallel import gather_results
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_frames, rank, world)
    results = {i: torch.full((4 + (i % 2 if mixed else 0), 5, 3), i, dtype=torch.uint8) for i in mine}
    full = gather_results(results, n_frames, rank, world, torch.device("cpu"))
    if mixed:
        ok = full is None
    else:
        ok = full is not None and full.shape == (n_frames, 4, 5, 3) and all(bool((full[i] == i).all()) for i in range(n_frames))
    open(os.path.join(out_dir, f"r{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_results_empty_shard_and_mixed_sizes(tmp_path):
    for n_frames, mixed in ((1, False), (5, False), (4, True)):        # 1 frame on 2 ranks: rank 1's shard is empty
        port = _free_port()
        mp.spawn(_worker_results, args=(2, port, n_frames, mixed, str(tmp_path)), nprocs=2, join=True)
        assert open(tmp_path / "r0").read() == "1" and open(tmp_path / "r1").read() == "1", (n_frames, mixed)
