#!/bin/bash
# GPU box: debug cases (each in a subprocess with a timeout), GPU tests, then bench with the split-tile kernel on/off
mkdir -p gpurun_out
timeout 900 python tools/tc_debug.py tiny1tile tiny2tile kv4 heads ragged d64 d32_96 d24 pad causal peaked ramp ramp_1tile_segments ramp_causal seg_many long wide256 2>&1 | tee gpurun_out/split_debug.log
if grep -q '"hang"\|"error"' gpurun_out/split_debug.log; then echo "debug cases failed; stopping"; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for s in 1 0; do
  echo "== bench PCV_SPLIT=$s"
  PCV_SPLIT=$s timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --e2e-steps 2 2>gpurun_out/bench_split$s.err | tee gpurun_out/bench_split$s.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
done
