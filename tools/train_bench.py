"""Training-step benchmark (BASELINE.json configs[3]: train.py noise-prediction step, global batch 224, bf16 tensor-core
compute with fp32 master weights, AdamW + EMA, gradient all-reduce over NCCL when launched with torchrun).

    python tools/train_bench.py [--batch 224] [--steps 10] [--warmup 3]          # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py --batch 224

Prints one JSON line (rank 0).  WD_TRAIN_PROF=1 adds a per-kernel-name device-time table on stderr."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist

import weights as W  # synthetic weights / inputs only (not the oracle's compute)
from worddiffusion_b200.diffusion import Diffusion
from worddiffusion_b200.training import FusedTrainStep, shard_batch
from worddiffusion_b200.unet import UNetModel, default_args

KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)
FWD_GFLOP_PER_LATENT = 9.153 + 0.039  # step-dependent + context work (recomputed every training step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=224, help="GLOBAL batch (BASELINE config 4: 224)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lo, hi = shard_batch(a.batch, rank, world)
    B = hi - lo
    m = UNetModel(args=default_args(dev), **KW)
    m.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
    m = m.to(dev).train()
    step = FusedTrainStep(m, lr=1e-4)
    diff = Diffusion(device=dev)
    inp = W.make_inputs(a.batch, seed=1234)
    g = torch.Generator().manual_seed(99)
    latents = (torch.randn((a.batch, 4, 8, 32), generator=g) * 0.18215)[lo:hi].to(dev)   # synthetic VAE latents (train.py:271-273)
    ctx, y = inp["context"][lo:hi].to(dev), inp["y"][lo:hi].to(dev)

    def one():
        t = diff.sample_timesteps(B).to(dev)                       # train.py:277
        x_t, noise = diff.noise_images(latents, t)                 # train.py:278
        return step.step(x_t, t, ctx, y, noise)

    losses = []
    for _ in range(a.warmup):
        losses.append(float(one()))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(a.steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    wall = time.time() - t0
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    f, b = step.eng.launch_counts
    if rank == 0:
        flops = 3.0 * FWD_GFLOP_PER_LATENT * 1e9 * a.batch   # forward + data gradients + weight gradients
        print(json.dumps({
            "metric": "train_latents_per_sec", "value": a.batch / (ms / 1e3), "unit": "latents/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "train_steps_per_sec": 1e3 / ms, "higher_is_better": True,
            "scaling": "strong", "dtype": "bf16 (fp32 master weights, fp32 accumulation)", "data": "synthetic",
            "config": {"workload": "unet.UNetModel noise-prediction training step (forward + backward + gradient all-reduce + "
                                   "AdamW + EMA), global batch %d, latents 4x8x32" % a.batch, "global_batch": a.batch,
                       "batch_per_gpu": B},
            "model_tflops": flops / (ms / 1e3) / 1e12, "launches_fwd": f, "launches_bwd": b,
            "loss_first": losses[0] if losses else None, "loss_last": float(loss), "wall_ms_per_step": wall / a.steps * 1e3,
            "workspace_gb": step.eng.workspace_bytes / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
