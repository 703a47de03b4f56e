run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs > gpurun_out/R4r_$tag.json 2> gpurun_out/R4r_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R4r_$tag.json > gpurun_out/R4r_$tag.txt 2>&1; head -3 gpurun_out/R4r_$tag.txt; }
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "conv3x3" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -4
run ts1 WD_PAIR_TAIL_SPLIT=1
run ts0 WD_PAIR_TAIL_SPLIT=0
run ts1b WD_PAIR_TAIL_SPLIT=1
run ts0b WD_PAIR_TAIL_SPLIT=0
