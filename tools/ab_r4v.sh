# gradient all-reduce: one all-reduce after the backward pass (WD_GRAD_BUCKETS=1) vs 4 buckets launched between the backward stages
for nb in 4 1 4 1; do
  WD_GRAD_BUCKETS=$nb python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2953$nb tools/train_bench.py --steps 20 > gpurun_out/R4v_train_${1}gpu_nb$nb.log 2>&1
  echo "buckets=$nb: $(tail -1 gpurun_out/R4v_train_${1}gpu_nb$nb.log | cut -c1-300)"
done
