# forward chain length A/B in pair mode: accuracy (self-test conv fwd lines, closure parity prints) and closure time
for pr in 1 2 3 5; do
  echo "=== PROMOTE_FWD=$pr"
  IST_B200_PROMOTE_FWD=$pr timeout 200 python tools/gpu_selftest.py 2>&1 | grep -i "conv.*fwd\|conv3x3 fwd\|FAIL" | head -8
  IST_B200_PROMOTE_FWD=$pr timeout 200 python -m pytest tests/test_closure_gpu.py -q -s -k "live_oracle or golden" 2>&1 | grep -E "rel-L2|passed|failed"
  for s in 512 1024; do IST_B200_PROMOTE_FWD=$pr timeout 120 python tools/gpu_closure_bench.py $s 100 2>&1 | tail -1 | cut -c1-60; done
done
