#!/bin/bash
# GPU box: validates the opt-in split hand-off kernel (PCV_SPLIT=1) — debug cases (each in a subprocess with a
# timeout), the GPU test-suite — then benches the kernel variants
mkdir -p gpurun_out
export PCV_SPLIT=1
timeout 900 python tools/tc_debug.py tiny1tile tiny2tile kv4 heads ragged d64 d32_96 d24 pad causal peaked ramp ramp_1tile_segments ramp_causal seg_many long wide256 2>&1 | tee gpurun_out/split_debug.log
if grep -q '"hang"\|"error"\|"rc"' gpurun_out/split_debug.log; then echo "debug cases failed; stopping"; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_split.log
unset PCV_SPLIT
for cfg in "PCV_SPLIT=1" "PCV_SPLIT=0" ${EXTRA_CFGS}; do
  echo "== bench $cfg"
  env $cfg timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --e2e-steps 2 2>gpurun_out/bench_$cfg.err | tee gpurun_out/bench_$cfg.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
