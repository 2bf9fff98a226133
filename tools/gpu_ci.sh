#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, short bench. Everything logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py ${BENCH_ARGS:---steps 5 --warmup 3 --e2e-steps 3 --cpu-seconds 5} 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
