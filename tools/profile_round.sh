#!/bin/bash
# Evidence run for profiles/ (one gpurun call): bench line, ncu launch list of the bench command, per-launch metrics of the
# 25 conv launches of one 512^2 closure, one full capture of a conv4_2-shaped forward launch, per-layer event times.
# usage: bash tools/profile_round.sh TAG     (outputs under gpurun_out/)
TAG=${1:-r01_final}
O=gpurun_out
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || exit 1
tail -1 $O/${TAG}_bench.json | cut -c1-400
python tools/gpu_layer_times.py 512 > $O/${TAG}_layer_times.log 2>&1
# launch list of the bench command (first closures of the first frame after setup)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 450 --csv --log-file $O/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batched > $O/${TAG}_ncu_bench.log 2>&1
# the 25 conv launches of one closure (skip: 12 style + 9 content target launches + 3 warm-up closures)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__cycles_active.avg \
    --clock-control none -k regex:conv_halo -s 96 -c 25 --csv --log-file $O/${TAG}_ncu_conv_25launches_metrics.csv \
    python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_conv25.log 2>&1
# full capture of one conv4_2-shaped forward launch (512 -> 512 channels at 64 x 64)
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o $O/${TAG}_conv_pair_c42 \
    python tools/gpu_conv_probe.py 512,512,64 > $O/${TAG}_ncu_full.log 2>&1
ls -la $O | grep ${TAG}
