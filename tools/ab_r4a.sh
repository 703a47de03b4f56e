# producer-side GroupNorm A/B (round 2, session 3)
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --vae-batch 0 --no-extra-legs > gpurun_out/R4d_$tag.json 2> gpurun_out/R4d_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R4d_$tag.json 2>&1 | head -4; }
python tools/op_gnfuse.py 20
run gnp2 WD_GN_PRODUCER=2
run gnp1 WD_GN_PRODUCER=1
run gnp0 WD_GN_PRODUCER=0
run gnp2_mink45 WD_GN_PRODUCER=2 WD_GEMM_PAIR_MINK=45
run gnp1_mink45 WD_GN_PRODUCER=1 WD_GEMM_PAIR_MINK=45
run gnp0b WD_GN_PRODUCER=0
run gnp2b WD_GN_PRODUCER=2
run gnp1b WD_GN_PRODUCER=1
