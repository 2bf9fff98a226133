#!/bin/bash
# Synthetic cross-attention sweep of BASELINE.json configs[4]: M in {4096, 16384, 65536, 262144} x GPUS (one box),
# key axis sharded across the ranks.  One JSON line per point into gpurun_out/sweep_${TAG}.jsonl.
# usage: GPUS="1 2 4 8" MS="4096 16384 65536 262144" TAG=r02 bash tools/sweep_mgpu.sh
mkdir -p gpurun_out
TAG=${TAG:-r02}
OUT=gpurun_out/sweep_${TAG}.jsonl
PORT=29600
for n in ${GPUS:-1 2 4 8}; do
  for m in ${MS:-4096 16384 65536 262144}; do
    PORT=$((PORT + 1))
    if [ "$n" = "1" ]; then
      timeout 300 python bench.py --gpus 1 --steps ${STEPS:-30} --warmup 5 --M $m --skip-cpu --skip-module --e2e-steps 3 2>/dev/null >> $OUT
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --steps ${STEPS:-30} --warmup 5 --M $m --skip-cpu --e2e-steps 3 --merge ${MERGE:-auto} 2>/dev/null >> $OUT
    fi
    tail -1 $OUT | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gpus', d['n_gpus'], 'M', d['config']['M'], 'TF/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'merge', d['config']['merge'])"
  done
done
