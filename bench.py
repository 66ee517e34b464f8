"""bench.py — L-BFGS iterations/s of the Gatys style-transfer loop at 512x512 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One *step* = one frame of the reference workload: targets for a fresh synthetic 512x512 radar frame + 300 closure
evaluations (15 optimizer.step() of torch-L-BFGS semantics, IST/config/defaults.py:71, IST/model/engine/utils.py:17-45)
through the public `optimize()` of this package. `value` = closure evaluations ("L-BFGS iters") per second over all GPUs
with the frames resident in HBM; `e2e` = the same through host buffers (pinned H2D of the frame, D2H of the result, final
NCCL gather when N > 1). Frames are independent problems: N GPUs process N x K frames (weak scaling), no per-step collective.

`--impl reference` times the reference's own CPU path (the oracle restatement, oracle/ist_oracle.py — the reference is pure
Python/PyTorch and its tree does not exist on the GPU box) on the host cores: each step there is a bounded sample of the
same workload (one optimizer.step = 20 closure evaluations of a 512x512 frame).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 512
EVALS_PER_FRAME = 300
GF_PER_EVAL = 396.95          # algorithmic GFLOP per closure evaluation at 512^2 (SURVEY 8d / BASELINE.md 4)
METRIC = "L-BFGS iters/sec at 512^2 per B200"
UNIT = "closure-evals/s"


def config_dict(n_gpus):
    return {
        "workload": "IST Gatys 512x512 single frame, 300 L-BFGS iters on 1xB200 (BASELINE configs[1]); one frame per step per GPU",
        "size": SIZE, "evals_per_step": EVALS_PER_FRAME, "frames_per_step": n_gpus,
        "weights": "synthetic Kaiming-normal VGG19 (vgg_conv.pth unavailable offline), seed 0",
        "style_layers": "relu1_1..relu5_1", "content_layers": "relu4_2", "lbfgs": "torch defaults (lr 1, max_iter 20, history 100, no line search)",
        "precision": "fp16 hi/lo split operands (3 tcgen05 MMAs per product) forward, bf16 hi/lo split data-gradient, fp32 accumulation promoted from tensor memory to registers every 3 taps forward (CTA pairs, tcgen05 cta_group::2), every 64-channel chunk backward",
        "l2": "working set per step (~0.35 GB activations + 0.63 GB L-BFGS history) exceeds the 126 MB L2; no explicit flush",
        "parallelism": "dp%d (independent frames, one end-of-run gather)" % n_gpus,
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), float(d.get("hbm_gbs", 6650.0)), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle edges of the region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """Reference arm: the oracle's optimize() on the host CPU (all threads), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import ist_oracle as O
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    state = O.state_to_torch(synth.vgg_state_dict(0), torch.float32)
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(SIZE, 2)))
    evals_per_step = 20
    times = []
    for i in range(args.warmup + args.steps):
        content = torch.from_numpy(synth.preprocess(synth.radar_frame(SIZE, 1000 + i)))
        x = content.clone().requires_grad_(True)
        # utils.py:19-20 once per frame, then utils.py:28-43; the sample runs ONE optimizer.step (20 of the frame's 300
        # evaluations) and is scaled to the frame: t_frame = t_targets + 15 * t_step
        t0 = time.perf_counter()
        targets = O.compute_targets(state, content, style, full=True)
        t_targets = time.perf_counter() - t0
        opt = torch.optim.LBFGS([x])
        n = [0]

        def closure():
            opt.zero_grad()
            loss = sum(O.layer_losses(state, x, targets, full=True))
            loss.backward()
            n[0] += 1
            return loss
        t0 = time.perf_counter()
        opt.step(closure)
        t_step = time.perf_counter() - t0
        assert n[0] == evals_per_step
        if i >= args.warmup:
            times.append((t_targets, t_step))
    steps_per_frame = EVALS_PER_FRAME // evals_per_step
    t_frame = sum(a + steps_per_frame * b for a, b in times) / len(times)
    total = sum(a + b for a, b in times)
    value = EVALS_PER_FRAME / t_frame
    sample = ("per step: the two target passes of a 512x512 frame + ONE optimizer.step (20 closure evaluations) of its 15, reference CPU path "
              "(oracle port), fp32; value = 300 / (t_targets + 15 * t_step), i.e. the target passes are amortised over the frame's 300 evaluations")
    cfgd = config_dict(args.gpus)
    cfgd["evals_timed_per_step"] = evals_per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def gpu_reference(dev, evals=EVALS_PER_FRAME):
    """The reference's own PyTorch GPU path (MODEL.DEVICE='cuda': cuDNN convolutions, cuBLAS bmm, autograd, torch.optim.LBFGS on
    the host — IST/model/engine/utils.py:17-45 through the oracle restatement) on the same B200 and the same frame, with the
    default flags (cudnn.allow_tf32=True: TF32 convolutions, gradient ~5e-2 from fp32 truth) and with TF32 off (the accuracy
    class of this implementation). north_star's ">= 10x the reference's own PyTorch GPU path" is judged against these."""
    import torch
    from oracle import ist_oracle as O
    from oracle import synth
    state = O.state_to_torch(synth.vgg_state_dict(0), torch.float32, dev)
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(SIZE, 2))).to(dev)
    old_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {"kind": "port: oracle restatement of the reference's closure + torch.optim.LBFGS on cuda (cuDNN/cuBLAS), full 21-layer forward as the reference runs it",
           "evals_per_frame": evals, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    try:
        for name, tf32 in (("tf32_default", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False          # torch's default
            dt = None
            for rep in range(2):                                   # first frame warms cuDNN's algorithm selection
                content = torch.from_numpy(synth.preprocess(synth.radar_frame(SIZE, 1000 + rep))).to(dev)
                x = content.clone().requires_grad_(True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _, n = O.optimize(state, content, style, x, evals if rep else 40, full=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            targets = O.compute_targets(state, content, style, full=True)
            xx = content + 20 * torch.randn_like(content)
            for _ in range(3):
                O.loss_and_grad(state, xx, targets, full=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                O.loss_and_grad(state, xx, targets, full=True)
            torch.cuda.synchronize()
            closure_ms = (time.perf_counter() - t0) / 20 * 1e3
            out[name] = {"value": n / dt, "unit": UNIT, "ms_per_eval": 1e3 * dt / n, "closure_ms_per_eval": closure_ms,
                         "optimizer_and_host_ms_per_eval": 1e3 * dt / n - closure_ms, "evals": n}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_flags
    return out


def profile_closure(ist_b200, plan, x, reps=3):
    """Per-launch CUDA-event timing of eager closures (same kernels as the graph replays); returns {name: [flops, bytes, ms, n]}."""
    import ctypes
    lib = ist_b200.load()
    maxr = 4096
    agg = {}
    # one stream while profiling: with the side stream a kernel's event time includes waiting for SMs another stream holds
    ist_b200._lib.check(lib.ist_set_option(b"overlap", 0))
    plan.loss_and_grad(x)
    for _ in range(reps):
        lib.ist_profile_begin()
        plan.loss_and_grad(x)
        names = ctypes.create_string_buffer(maxr * 40)
        flops = (ctypes.c_double * maxr)()
        nbytes = (ctypes.c_double * maxr)()
        ms = (ctypes.c_float * maxr)()
        n = ctypes.c_int(0)
        ist_b200._lib.check(lib.ist_profile_end(maxr, names, flops, nbytes, ms, ctypes.byref(n)))
        for i in range(n.value):
            nm = names.raw[i * 40:(i + 1) * 40].split(b"\0")[0].decode()
            a = agg.setdefault(nm, [0.0, 0.0, 0.0, 0])
            a[0] += flops[i]; a[1] += nbytes[i]; a[2] += ms[i]; a[3] += 1
    ist_b200._lib.check(lib.ist_set_option(b"overlap", 1))
    return agg


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ist_b200
    from ist_b200.config import get_cfg_defaults
    from ist_b200.main import get_model
    from ist_b200.model.engine.utils import optimize
    from ist_b200.parallel import gather_frames, init_distributed
    from oracle import synth          # synthetic frames / weights only (no oracle compute on this arm)

    rank, world, local_rank = init_distributed("nccl")
    if world != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = ist_b200.load()
    ist_b200._lib.check(lib.ist_device_check())

    cfg = get_cfg_defaults()
    cfg.MODEL.DEVICE = str(dev)
    model, _ = get_model(cfg, {k: torch.from_numpy(v) for k, v in synth.vgg_state_dict(0).items()})
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(SIZE, 2))).to(dev)
    n_steps = args.warmup + args.steps
    frames_host = [torch.from_numpy(synth.preprocess(synth.radar_frame(SIZE, 1000 + rank * 1000 + i))).pin_memory() for i in range(n_steps)]
    frames_dev = [f.to(dev) for f in frames_host]
    result_host = torch.empty(1, 3, SIZE, SIZE).pin_memory()
    frame_bytes = frames_host[0].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        x = frames_dev[i].clone().requires_grad_(True)
        optimize(model, frames_dev[i], style, x, cfg, EVALS_PER_FRAME)
        return x

    def step_e2e(i):
        c = frames_host[i].to(dev, non_blocking=True)
        x = c.clone().requires_grad_(True)
        optimize(model, c, style, x, cfg, EVALS_PER_FRAME)
        result_host.copy_(x.data, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return x

    # ---- device-resident arm -------------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.ist_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evals = 0
    for i in range(args.warmup, n_steps):
        step_resident(i)
        evals += model.last_evals
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.ist_launch_count() - launches0
    clocks = sampler.stop() if sampler is not None else None
    t = torch.tensor([ms, float(evals), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, evals, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
    value = evals / (ms * 1e-3)

    # ---- end-to-end arm (host buffers) ------------------------------------------------------------------------------------
    if world > 1:        # communicator set-up (lazy in NCCL) is not part of a frame's cost: run the gather once untimed
        gather_frames(torch.zeros(args.steps, 3, SIZE, SIZE, device=dev), world * args.steps, rank, world)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    evals_e = 0
    outs = []
    for i in range(args.warmup, n_steps):
        outs.append(step_e2e(i).data)
        evals_e += model.last_evals
    if world > 1:
        gather_frames(torch.cat(outs), world * len(outs), rank, world)
    e3.record()
    barrier()
    ms_e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3 if world == 1 else 0.0)
    te = torch.tensor([ms_e, float(evals_e)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = te.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms_e, evals_e = float(tm[0]), float(ts[1])
    e2e_value = evals_e / (ms_e * 1e-3)

    # ---- batched arm (informational): BATCH independent frames per optimize() call on every GPU (the 256-frame workload of
    # BASELINE configs[3] runs like this); resident inputs, one warm-up batch, one timed batch
    BATCH = max(1, args.batch)
    batched = None
    if not args.no_batched:
        cb = torch.cat([frames_dev[i % n_steps] for i in range(BATCH)]).contiguous()

        def step_batched():
            xb = cb.clone().requires_grad_(True)
            optimize(model, cb, style, xb, cfg, EVALS_PER_FRAME)
            return model.last_evals

        step_batched()
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        ev_b = step_batched() * BATCH
        e5.record()
        barrier()
        tb = torch.tensor([e4.elapsed_time(e5), float(ev_b)], dtype=torch.float64, device=dev)
        if world > 1:
            tbm = tb.clone()
            dist.all_reduce(tbm, op=dist.ReduceOp.MAX)
            tbs = tb.clone()
            dist.all_reduce(tbs, op=dist.ReduceOp.SUM)
            tb = torch.stack([tbm[0], tbs[1]])
        batched = {"frames_per_batch_per_gpu": BATCH, "value": float(tb[1]) / (float(tb[0]) * 1e-3), "unit": UNIT,
                   "frames_per_s": float(tb[1]) / (float(tb[0]) * 1e-3) / EVALS_PER_FRAME}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0, eager per-launch events) ------------------------------------------------
    plan = model.vgg_model.plan(1, SIZE, SIZE, "relu5_1")
    xprof = frames_dev[0] + 20.0 * torch.randn_like(frames_dev[0])
    agg = profile_closure(ist_b200, plan, xprof)
    groups = {}
    for nm, (fl, by, tms, n) in agg.items():
        key = "conv_halo_kernel" if nm.startswith(("conv_halo", "conv_igemm")) else nm
        g = groups.setdefault(key, [0.0, 0.0, 0.0, 0])
        g[0] += fl; g[1] += by; g[2] += tms; g[3] += n
    total_ms = sum(g[2] for g in groups.values())
    dom = max(groups.items(), key=lambda kv: kv[1][2])
    peak_tf, peak_hbm, peak_src = peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom[0])
        except Exception:
            traffic = None
    fl, by, tms, n = dom[1]
    if fl > 0:
        achieved = fl / (tms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dom[0], "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src + " bf16 sustained",
                "launches_per_eval": n // 3, "share_of_closure": tms / total_ms,
                "traffic_scope": "DRAM bytes (read + write) summed over the same launches of one closure, from the ncu capture under profiles/",
                "note": "algorithmic FLOPs (2*M*N*K, single pass) of all conv_halo launches (13 forward + 13 data-gradient convs incl. the fused Gram backward) of one closure / their summed CUDA-event time; "
                        "the kernel issues 3 MMAs per product (hi/lo split), so tensor-pipe activity is ~3x this fraction"}
    else:
        achieved = by / (tms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak_hbm, "unit": "GB/s", "frac": achieved / peak_hbm,
                "traffic": traffic, "peak_source": peak_src, "share_of_closure": tms / total_ms}
    kernel_table = {k: {"ms_per_eval": v[2] / 3, "launches_per_eval": v[3] // 3} for k, v in sorted(groups.items(), key=lambda kv: -kv[1][2])}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ist_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        state = O.state_to_torch(synth.vgg_state_dict(0), torch.float32)
        c = frames_host[0].clone()
        s = style.cpu()
        x = c.clone().requires_grad_(True)
        t0 = time.perf_counter()
        _, nev = O.optimize(state, c, s, x, 20, full=True)
        dt = time.perf_counter() - t0
        cpu = {"value": nev / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "one optimizer.step (20 closure evaluations + 2 target passes) of one 512x512 frame, oracle port of the reference, fp32"}

    # ---- closure / optimiser split of one evaluation (rank 0): eager closures back to back vs the whole step ----------------
    for _ in range(5):
        plan.loss_and_grad(xprof)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    for _ in range(50):
        plan.loss_and_grad(xprof)
    s1.record()
    torch.cuda.synchronize()
    closure_ms = s0.elapsed_time(s1) / 50
    ms_per_eval = ms / max(1.0, evals / world)
    split = {"ms_per_eval": ms_per_eval, "closure_ms_per_eval": closure_ms, "optimizer_targets_host_ms_per_eval": ms_per_eval - closure_ms,
             "note": "closure = 50 eager ist_plan_loss_and_grad calls back to back on this GPU; the rest = device L-BFGS kernels, "
                     "per-frame target passes, graph launches and the one host sync per optimizer.step"}

    # ---- the reference's PyTorch GPU path on this GPU (rank 0, N = 1 only) -----------------------------------------------------
    gpu_ref, vs_gpu_ref = None, None
    if world == 1 and not args.no_gpu_reference:
        model.vgg_model.release_plans()
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference(dev)
        vs_gpu_ref = {k: e2e_value / gpu_ref[k]["value"] for k in ("tf32_default", "fp32")}
        vs_gpu_ref["numerator"] = "e2e.value"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16x2-split/f32-acc",
        "data": "synthetic", "config": config_dict(world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": frame_bytes * world, "d2h_bytes_per_step": frame_bytes * world},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "frames_per_s": value / EVALS_PER_FRAME, "batched": batched, "algorithmic_tflops": value * GF_PER_EVAL / 1e3, "kernels": kernel_table,
        "split": split, "gpu_reference": gpu_ref, "vs_gpu_reference": vs_gpu_ref,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the informational frames-per-call measurement")
    ap.add_argument("--batch", type=int, default=4, help="frames per optimize() call of the informational batched measurement")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's PyTorch GPU path (N = 1 only, ~25 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
