for pr in 1 2 3 5; do
  echo "=== PROMOTE_FWD=$pr"
  for s in 512 1024; do IST_B200_PROMOTE_FWD=$pr timeout 120 python tools/gpu_closure_bench.py $s 100 2>&1 | tail -1 | cut -c1-60; done
done
