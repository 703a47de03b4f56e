#!/bin/bash
# A/B of the trainer's one-launch weight pack (WD_TRAIN_MULTI_PACK) at the single-GPU batch and at one 8-GPU rank's share
timeout 500 python -m pytest tests/test_gpu_train.py tests/test_gpu_bwd_ops.py -x -q 2>&1 | tail -3
for b in 224 28; do for m in 1 0; do
  echo -n "batch $b multi_pack $m: "
  WD_TRAIN_MULTI_PACK=$m timeout 200 python tools/train_bench.py --batch $b --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms/step")'
done; done
