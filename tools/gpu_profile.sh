#!/bin/bash
# ncu passes on the GPU box (one GPU): (1) launch list of a short bench run (attention core, e2e and module legs),
# (2) full capture of the attention kernel, (3) full capture of the fused K/V producer GEMM at the module shape.
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --skip-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_attn \
    python tools/profile_target.py > gpurun_out/${TAG}_attn_ncu.log 2>&1
PCV_ROWS=524288 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kvproj_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_kvproj \
    python tools/kvproj_profile_target.py > gpurun_out/${TAG}_kvproj_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_attn_ncu.log gpurun_out/${TAG}_kvproj_ncu.log
ls -la gpurun_out/
