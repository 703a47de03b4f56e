run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs > gpurun_out/R4o_$tag.json 2> gpurun_out/R4o_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R4o_$tag.json > gpurun_out/R4o_$tag.txt 2>&1; grep -E "ms/step|out_head|groupnorm" gpurun_out/R4o_$tag.txt; }
timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
run oh1 WD_OUT_HEAD=1
run oh0 WD_OUT_HEAD=0
