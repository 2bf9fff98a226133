for M in 4096 16384 65536 262144; do for rep in 1 2; do for m in 0 1 3; do
PCV_MMAOPT=$m timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --e2e-steps 0 --M $M 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('M', $M, 'mmaopt', $m, round(d['value'],1), round(d['ms_per_step'],4))"
done; done; done
