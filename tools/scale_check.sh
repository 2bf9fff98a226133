#!/bin/bash
# scaling points for both rank grids (auto = batch first, m = key axis only); usage: scale_check.sh "4 auto" "4 m" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 bench.py \
      --gpus $1 --steps 20 --warmup 5 --decomp $2 2>gpurun_out/bench$1_$2.err | tail -1 > gpurun_out/r02_bench_$1gpu_$2.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$1gpu_$2.json"))
print("$1 $2", round(d["value"],1), "TF/s", round(d["ms_per_step"],4), "ms |", d["config"]["parallelism"], "| e2e", round(d["e2e"]["ms_per_step"],3), "ms | kernel", round(d["roofline"]["kernel_ms"],4), "ms | launches", d["gpu_launches"])
PY
done
