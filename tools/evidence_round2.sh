#!/bin/bash
# Final evidence of round 2 (one gpurun call, one GPU); every program runs (and exits 0) WITHOUT ncu before it is profiled.
# usage: bash tools/evidence_round2.sh TAG      (outputs under gpurun_out/, copied to profiles/ afterwards)
TAG=${1:-r02}
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || exit 1
tail -1 $O/${TAG}_bench.json | cut -c1-300
python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_reference_arm.json 2>/dev/null
# parity logs of the shipped build
python -m pytest tests/test_parity_fullsize_gpu.py -m gpu -q -s > $O/${TAG}_parity_fullsize.log 2>&1
python -m pytest tests/test_lbfgs_teacher_forced_gpu.py tests/test_optimize_gpu.py -m gpu -q -s > $O/${TAG}_lbfgs_teacher_forced.log 2>&1
tail -1 $O/${TAG}_parity_fullsize.log; tail -1 $O/${TAG}_lbfgs_teacher_forced.log
# timings
python tools/gpu_layer_times.py 512 > $O/${TAG}_layer_times.log 2>&1
for s in 256 512 1024 2048; do python tools/gpu_closure_bench.py $s 50 2>&1 | tail -1 | cut -c1-70; done > $O/${TAG}_closure_sizes.log
IST_B200_NO_GRAPH=1 python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_lbfgs_times.log 2>&1
python tools/gpu_schedule.py > $O/${TAG}_schedule_512_1024_2048.log 2>&1
# ncu: launch list of the bench command, per-launch metrics of the closure's kernels, L-BFGS kernels at full history
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 450 --csv --log-file $O/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batched --no-gpu-reference > $O/${TAG}_ncu_bench.log 2>&1
ncu --metrics $M --clock-control none -k regex:"gram_syrk|content_partial|maxpool|grad_route|gram_reduce|gram_dmat|loss_total|conv_first" \
    -c 160 --csv --log-file $O/${TAG}_ncu_hbm_kernels.csv python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_hbm.log 2>&1
ncu --metrics $M --clock-control none -k regex:conv_halo -s 96 -c 25 --csv --log-file $O/${TAG}_ncu_conv_25launches_metrics.csv \
    python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_conv25.log 2>&1
IST_B200_NO_GRAPH=1 ncu --metrics $M --clock-control none -k regex:lbfgs -s 520 -c 24 --csv --log-file $O/${TAG}_ncu_lbfgs_kernels.csv \
    python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_ncu_lbfgs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o $O/${TAG}_conv_pair_c42 \
    python tools/gpu_conv_probe.py 512,512,64 > $O/${TAG}_ncu_full.log 2>&1
ls $O | grep ${TAG}_ | tr '\n' ' '
