timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_closure_gpu.py -x -q 2>&1 | tail -2
IST_B200_DBG_TIMES=1 python tools/gpu_conv_probe.py 64,64,512 128,128,256 256,256,128 512,512,64 2>&1 | grep dbg | awk "NR%3==0" | sed 's/avg clk since entry://; s/first_mma [-0-9]* last_issue [-0-9]*//'
for s in 512 1024; do timeout 120 python tools/gpu_closure_bench.py $s 100 2>&1 | tail -1 | cut -c1-60; done
