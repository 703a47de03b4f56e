"""Data-parallel training check, launched by torchrun with N >= 2 ranks (one per GPU, NCCL): the gradient of the global
batch computed as all-reduced shards equals the single-GPU gradient of the whole batch, and after one fused AdamW step every
rank holds identical parameters.  Prints `DDP_CHECK_OK` on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist

import weights as W
from worddiffusion_b200.training import FusedTrainStep, allreduce_sum_, shard_batch
from worddiffusion_b200.unet import UNetModel, default_args

KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)


def build(dev):
    m = UNetModel(args=default_args(dev), **KW)
    m.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
    return m.to(dev).train()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    B = 4 * world
    inp = W.make_inputs(B, seed=77)
    noise = torch.randn((B, 4, 8, 32), generator=torch.Generator().manual_seed(78))
    lo, hi = shard_batch(B, rank, world)

    def grads(step, sl):
        eng = step.eng
        x, t, c, y, n = (v[sl].to(dev) for v in (inp["x"], inp["t"], inp["context"], inp["y"], noise))
        eps = eng.forward(x, t, y, c)
        diff = eps - n
        eng.backward(diff * (2.0 / diff.numel()))
        return eng.flat_grad

    step = FusedTrainStep(build(dev), lr=1e-4)
    g = grads(step, slice(lo, hi))
    ws = allreduce_sum_(g)
    assert ws == world
    g_avg = (g / world).clone()
    ok = True
    # the same exchange as buckets launched between the backward stages (what FusedTrainStep.step does at world > 1)
    eng = step.eng
    x, t, c, y, n = (v[lo:hi].to(dev) for v in (inp["x"], inp["t"], inp["context"], inp["y"], noise))
    diff = eng.forward(x, t, y, c) - n
    d_eps = diff * (2.0 / diff.numel())
    works, s0 = [], 0
    assert len(step.buckets) > 1 and step.buckets[-1][0] == eng.n_stages and step.buckets[-1][2] == eng.flat_grad.numel()
    for s1, b_lo, b_hi in step.buckets:
        eng.backward_stages(d_eps, s0, s1)
        works.append(dist.all_reduce(eng.flat_grad[b_lo:b_hi], op=dist.ReduceOp.SUM, async_op=True))
        s0 = s1
    for w in works:
        w.wait()
    rel_b = float((eng.flat_grad / world - g_avg).norm() / g_avg.norm())
    if rank == 0:
        print(f"bucketed ({len(step.buckets)} buckets, overlapped) vs single all-reduce: gradient rel-L2 {rel_b:.3e}")
    ok = ok and rel_b < 1e-5
    if rank == 0:
        ref_step = FusedTrainStep(build(dev), lr=1e-4)
        g_full = grads(ref_step, slice(0, B)).clone()
        rel = float((g_avg - g_full).norm() / g_full.norm())
        print(f"sharded-vs-full gradient rel-L2 {rel:.3e} (world {world}, global batch {B})")
        ok = rel < 5e-3
    # one fused step on every rank: parameters must stay bit-identical across ranks
    x, t, c, y, n = (v[lo:hi].to(dev) for v in (inp["x"], inp["t"], inp["context"], inp["y"], noise))
    step.step(x, t, c, y, n)
    p = step.flat_param
    pmax, pmin = p.clone(), p.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    same = bool((pmax == pmin).all())
    flag = torch.tensor([1 if (ok and same) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("parameters identical across ranks:", same)
        print("DDP_CHECK_OK" if int(flag) == 1 else "DDP_CHECK_FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
