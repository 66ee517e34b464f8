#!/bin/bash
# Round-2 evidence (one gpurun call, one GPU): ncu metrics of every NON-conv kernel of one 512^2 closure (HBM GB/s, DRAM %, tensor
# pipe % for the Gram SYRK) and of the device L-BFGS kernels at full history, a full source-level capture of the solve kernel,
# the per-launch metrics of the 25 conv launches, and the ncu launch list of the bench command. Each program first runs (and
# exits 0) without ncu. usage: bash tools/profile_round2.sh TAG
TAG=${1:-r02}
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_closure_plain.log 2>&1 || exit 1
# 4 closures run after the target passes (3 warm-up + 1 timed): keep the kernels of the last one (post-processed by name order)
ncu --metrics $M --clock-control none -k regex:"gram_syrk|content_partial|maxpool|grad_route|gram_reduce|gram_dmat|loss_total|conv_first" \
    -c 160 --csv --log-file $O/${TAG}_ncu_hbm_kernels_raw.csv python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_hbm.log 2>&1
ncu --metrics $M --clock-control none -k regex:conv_halo -s 96 -c 25 --csv --log-file $O/${TAG}_ncu_conv_25launches_metrics.csv \
    python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_conv25.log 2>&1
# L-BFGS kernels at full history (6 steps = 120 iterations before): metrics of the last step's kernels
IST_B200_NO_GRAPH=1 python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_lbfgs_times.log 2>&1 || exit 1
IST_B200_NO_GRAPH=1 ncu --metrics $M --clock-control none -k regex:lbfgs -s 520 -c 24 --csv --log-file $O/${TAG}_ncu_lbfgs_kernels.csv \
    python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_ncu_lbfgs.log 2>&1
IST_B200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:lbfgs_solve -s 130 -c 1 -o $O/${TAG}_lbfgs_solve \
    python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_ncu_solve_full.log 2>&1
IST_B200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:lbfgs_update -s 130 -c 1 -o $O/${TAG}_lbfgs_update \
    python tools/gpu_lbfgs_times.py 512 6 > $O/${TAG}_ncu_update_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_syrk -s 15 -c 1 -o $O/${TAG}_gram_syrk_relu1_1 \
    python tools/gpu_closure_bench.py 512 1 > $O/${TAG}_ncu_gram_full.log 2>&1
ls -la $O | grep ${TAG}
