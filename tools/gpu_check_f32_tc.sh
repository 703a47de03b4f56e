#!/bin/bash
# One GPU call that decides whether the split-TF32 tcgen05 GEMM (csrc/f32_gemm_tc.cu) may become the fp32 mode's default:
#   1. its operator test (K = 32 / 320 ran in round 1; K = 2880 exercises the split-K second accumulation level for the first time)
#   2. the whole fp32-mode parity suite (1e-4 vs the reference goldens) with every Linear / 1x1 / 3x3 routed through it
#   3. the fp32-mode leg of bench.py with and without it (ms per batch-256 step)
# Usage (from the repo root): gpurun --timeout 300 -- 'bash tools/gpu_check_f32_tc.sh'
set -u
mkdir -p gpurun_out
echo "== 1. operator tests"; timeout 120 python -m pytest tests/test_gpu_zfp32.py -q -s -k f32_tc 2>&1 | tail -8
echo "== 2. fp32 parity suite through the tensor-core GEMM"; WD_F32_TC=1 timeout 200 python -m pytest tests/test_gpu_zfp32.py -q -s 2>&1 | grep -E "err|passed|failed|Error" | tail -30
for tc in 0 1; do
  echo "== 3. fp32 leg, WD_F32_TC=$tc"
  WD_F32_TC=$tc timeout 200 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --train-steps 0 > gpurun_out/f32tc_bench_$tc.json 2> gpurun_out/f32tc_bench_$tc.err
  python -c "import json; d=json.load(open('gpurun_out/f32tc_bench_$tc.json')); print(json.dumps(d.get('fp32_mode')))"
done
