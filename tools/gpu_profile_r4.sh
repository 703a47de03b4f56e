#!/bin/bash
# Round-2 (session 3) profiling pass on the GPU box (run through gpurun).  $1 = tag.  Writes gpurun_out/<tag>_*.
#  1. plain run (must exit 0)   2. launch list   3. ncu --set full of the fused transformer-block kernel, the large GroupNorm
#  launches, one step's window of GEMM launches (22 per step: gemm_tc_kernel + gemm_pair_kernel) and the output-head kernel
TAG=${1:-R4}
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tblock_unet -s 5 -c 2 -o gpurun_out/${TAG}_tblock -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:groupnorm_apply_bulk -s ${GN_SKIP:-60} -c 5 -o gpurun_out/${TAG}_gn -f $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s ${SKIP:-88} -c ${COUNT:-22} -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:out_head -s 4 -c 1 -o gpurun_out/${TAG}_outhead -f $CMD > gpurun_out/${TAG}_ncu5.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
# summaries on the box (the reports themselves exceed what gpurun_out may carry back); keep only the small reports
for k in tblock gn gemm outhead; do
  python tools/summarize_ncu.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_ncu_$k.txt 2>&1
done
python tools/ncu_stalls.py gpurun_out/${TAG}_tblock.ncu-rep 1 30 >> gpurun_out/${TAG}_ncu_tblock.txt 2>&1
python tools/ncu_stalls.py gpurun_out/${TAG}_gn.ncu-rep 0 20 >> gpurun_out/${TAG}_ncu_gn.txt 2>&1
python tools/ncu_stalls.py gpurun_out/${TAG}_outhead.ncu-rep 0 30 >> gpurun_out/${TAG}_ncu_outhead.txt 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
rm -f gpurun_out/${TAG}_gemm.ncu-rep gpurun_out/${TAG}_tblock.ncu-rep gpurun_out/${TAG}_gn.ncu-rep
du -sh gpurun_out
