timeout 600 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train.py -x -q 2>&1 | tail -4
for b in 28 224; do for v in 1 0; do
  echo "batch=$b WD_GN_BWD_ROWS=$v: $(WD_GN_BWD_ROWS=$v python tools/train_bench.py --batch $b --steps 30 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])')"
done; done
