#!/bin/bash
# Profiling pass on the GPU box (run through gpurun).  $1 = tag.  Writes gpurun_out/<tag>_*.
# 1. plain run (must exit 0), 2. per-launch device times of every kernel of one step, 3. `ncu --set full` of selected launches.
TAG=${1:-r}
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# one full step of the steady state: skip the weight-repack / context / warm-up launches (see ${TAG}_launches.csv)
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|groupnorm_apply|layernorm|attn_flash|im2col" -s ${SKIP:-150} -c ${COUNT:-100} -o gpurun_out/${TAG}_step -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/ | tail -12
