"""BASELINE.json configs[4]: DDIM (eta = 0) 50-step sampling, batch sweep, unetPhosc2.UNetModelPhosc (ctx = 10 chars + 769
PHOSC tokens), 1..N GPUs (torchrun: the global batch is sharded, one all-gather of the latents per trajectory), and
configs[2]: PHOSC-conditioned sampling of a global batch of 1024 sharded over the ranks.

    python tools/ddim_sweep.py --batches 1,4,16,64,256,1024 [--cpu-batch 4 --cpu-steps 2]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ddim_sweep.py --batches 1024,4096

One JSON line per batch (rank 0): trajectory time (device, max over ranks), word-latents/s, UNet latent-steps/s."""
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

import weights as W
from worddiffusion_b200.diffusion import Diffusion
from worddiffusion_b200.unetPhosc2 import UNetModelPhosc
from worddiffusion_b200.unet_base import default_args

KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64,256,1024")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cpu-batch", type=int, default=0, help="also time the CPU oracle port at this batch (rank 0)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    m = UNetModelPhosc(args=default_args(dev), **KW)
    sd = W.make_state_dict(W.load_spec("unetPhosc"), 1234)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    d = Diffusion(device=dev)
    for N in [int(b) for b in a.batches.split(",")]:
        if N < world:
            continue
        inp = W.make_inputs(N, seed=1234)
        ctx, y, ph = inp["context"].to(dev), inp["y"].to(dev), inp["phosc"].to(dev)
        times = []
        for rep in range(2):  # first trajectory = warm-up (plan + arena allocation)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x = d.sample_latents_sharded(m, ctx, y, phosc=ph, seed=7, ddim_steps=a.steps)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t))
        assert x.shape[0] == N and torch.isfinite(x).all()
        if rank == 0:
            ms = times[-1]
            print(json.dumps({"workload": "unetPhosc2 DDIM eta=0", "ddim_steps": a.steps, "global_batch": N, "n_gpus": world,
                              "trajectory_ms": ms, "word_latents_per_sec": N / (ms / 1e3),
                              "latent_steps_per_sec": N * a.steps / (ms / 1e3), "ctx_tokens": 779}), flush=True)
    if rank == 0 and a.cpu_batch > 0:
        import unet_oracle as UO
        from diffusion_oracle import DiffusionOracle
        torch.set_num_threads(os.cpu_count() or 1)
        inp = W.make_inputs(a.cpu_batch, seed=1234)
        do = DiffusionOracle(1000)
        ts = do.ddim_timesteps(a.steps)
        x = inp["x"].clone()
        t0 = None
        with torch.no_grad():
            for k in range(a.cpu_steps + 1):
                t = torch.full((a.cpu_batch,), ts[k], dtype=torch.long)
                eps = UO.unet_forward(sd, x, t, inp["context"], inp["y"], phosc=inp["phosc"], variant="unetPhosc")
                x = do.ddim_step(x, eps, ts[k], ts[k + 1])
                if t0 is None:
                    t0 = time.perf_counter()
        el = (time.perf_counter() - t0) / a.cpu_steps
        print(json.dumps({"workload": "CPU oracle port of unetPhosc (fp32, torch CPU), DDIM step", "batch": a.cpu_batch,
                          "cores": os.cpu_count(), "sec_per_step": el, "latent_steps_per_sec": a.cpu_batch / el,
                          "word_latents_per_sec_50_steps": a.cpu_batch / (el * a.steps)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
