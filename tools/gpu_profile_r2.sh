#!/bin/bash
# Round-2 profiling pass on the GPU box (run through gpurun).  $1 = tag.  Writes gpurun_out/<tag>_*.
#  1. plain run (must exit 0)   2. launch list   3. ncu --set full of the fused transformer-block kernel, of the large
#  GroupNorm launches (81 MB and 163 MB per launch) and of a window of GEMM launches
TAG=${1:-R2}
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --no-extra-legs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tblock_unet -s 5 -c 2 -o gpurun_out/${TAG}_tblock -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:groupnorm_apply_bulk -s ${GN_SKIP:-77} -c 3 -o gpurun_out/${TAG}_gn -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s ${SKIP:-92} -c ${COUNT:-23} -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
