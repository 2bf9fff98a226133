#!/bin/bash
# ncu passes on the GPU box: (1) launch list of a short bench run, (2) full capture of the attention kernel.
mkdir -p gpurun_out
TAG=${TAG:-prof}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --skip-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG} \
    python tools/profile_target.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/
