for b in 28 224; do for t in 0 111; do
  echo "batch=$b target_items=$t: $(WD_WGRAD_TARGET_ITEMS=$t python tools/train_bench.py --batch $b --steps 30 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])')"
done; done
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd_ops.py -x -q 2>&1 | tail -3
