run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs --ops-out gpurun_out/R4G_ops_$tag.json > gpurun_out/R4G_$tag.json 2> gpurun_out/R4G_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R4G_$tag.json > gpurun_out/R4G_$tag.txt 2>&1; grep -E "ms/step|gemm_tc|upsample" gpurun_out/R4G_$tag.txt; tail -2 gpurun_out/R4G_$tag.err; }
timeout 400 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -4
run up2 WD_UP_PHASE=2
run up1 WD_UP_PHASE=1
run up2b WD_UP_PHASE=2
