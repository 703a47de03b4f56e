run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs --ops-out gpurun_out/R4b_ops_$tag.json > gpurun_out/R4b_$tag.json 2> gpurun_out/R4b_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R4b_$tag.json 2>&1 | head -1; }
run gnp1 WD_GN_PRODUCER=1
run gnp0 WD_GN_PRODUCER=0
run gnp1b WD_GN_PRODUCER=1
run gnp0b WD_GN_PRODUCER=0
