run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0 --train-steps 0 --fp32-steps 0 --eager-steps 0 --no-extra-legs > gpurun_out/R2i_$tag.json 2> gpurun_out/R2i_$tag.err; echo "== $tag rc=$?"; python tools/bench_summary.py gpurun_out/R2i_$tag.json 2>&1 | head -4; }
run base A=1
run gnrev WD_GN_REVERSE=1
run keep WD_GEMM_DBG=64
run keep_rev WD_GEMM_DBG=64 WD_GN_REVERSE=1
run mink45 WD_GEMM_PAIR_MINK=45
