# A/B of the conv kernel variants (IST_B200_CONV = pair | halo): closure-only loop and the bench line
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), 'conv ms/eval', round(d['kernels']['conv_halo_kernel']['ms_per_eval'],3), 'frac', round(d['roofline']['frac'],3))"; }
for c in pair halo; do
  echo "CONV=$c"
  for s in 256 512 1024; do IST_B200_CONV=$c timeout 120 python tools/gpu_closure_bench.py $s 100 2>&1 | tail -1 | cut -c1-60; done
  IST_B200_CONV=$c run
done
IST_B200_CONV=pair timeout 200 python tools/gpu_layer_times.py 512 2>&1 | grep "conv_halo" | head -30
