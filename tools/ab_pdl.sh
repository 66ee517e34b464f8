run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))"; }
for m in 0 1 3 7 4; do echo "MASK=$m graph"; IST_B200_PDL_MASK=$m run; done
for m in 0 7; do echo "MASK=$m nograph"; IST_B200_NO_GRAPH=1 IST_B200_PDL_MASK=$m run; done
