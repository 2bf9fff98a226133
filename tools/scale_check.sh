#!/bin/bash
# one 8-GPU box: the 4- and 8-rank points of the scaling run for both rank grids (auto = batch first, m = key axis only)
mkdir -p gpurun_out
for cfg in "4 auto" "8 auto" "8 m" "4 m"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$1 bench.py \
      --gpus $1 --steps 20 --warmup 5 --decomp $2 2>gpurun_out/bench$1_$2.err | tail -1 > gpurun_out/r02_bench_$1gpu_$2.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$1gpu_$2.json"))
print("$1 $2", round(d["value"],1), "TF/s", round(d["ms_per_step"],4), "ms |", d["config"]["parallelism"], "| e2e", round(d["e2e"]["ms_per_step"],3), "ms | kernel", round(d["roofline"]["kernel_ms"],4), "ms | launches", d["gpu_launches"])
PY
done
