#!/bin/bash
# Second evidence run (one gpurun call): closure times per size, 1024^2 kernel microbench vs the cuDNN/cuBLAS PyTorch ops,
# coarse-to-fine schedule 512 -> 1024 -> 2048, A/B of the conv variants. usage: bash tools/evidence_extra.sh TAG
TAG=${1:-r01_final}
O=gpurun_out
for s in 256 512 1024 2048; do python tools/gpu_closure_bench.py $s 50 2>&1 | tail -1 | cut -c1-70; done > $O/${TAG}_closure_sizes.log
for c in pair halo; do echo "IST_B200_CONV=$c"; for s in 512 1024; do IST_B200_CONV=$c python tools/gpu_closure_bench.py $s 50 2>&1 | tail -1 | cut -c1-70; done; done > $O/${TAG}_conv_variants.log
python tools/gpu_microbench.py > $O/${TAG}_microbench_1024.log 2>&1
python tools/gpu_schedule.py > $O/${TAG}_schedule.log 2>&1
tail -3 $O/${TAG}_closure_sizes.log; tail -2 $O/${TAG}_microbench_1024.log; tail -6 $O/${TAG}_schedule.log
