#!/bin/bash
# Profiling pass on the GPU box (run through gpurun).  $1 = tag.  Writes gpurun_out/<tag>_*.
#  1. plain run (must exit 0)   2. per-launch device times of every kernel (launch list)
#  3. `ncu --set full` of a window of consecutive GEMM launches of a steady-state step + one launch of each non-GEMM kernel
TAG=${1:-r}
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --train-steps 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# GEMM launches: the weight-repack / context kernels are not GEMMs, so "-s" counts gemm launches only (55 per step)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s ${SKIP:-170} -c ${COUNT:-16} -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"groupnorm_apply_bulk|im2col|upsample" -s 30 -c 3 -o gpurun_out/${TAG}_misc -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
