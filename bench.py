#!/usr/bin/env python
"""Benchmark of the WordDiffusion denoising hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model unet|unetPhosc] [--batch B]

A "step" is one DDPM denoising step of `Diffusion.sampling` (train.py:221-236) over one batch of synthetic latents:
one conditional-UNet evaluation (unet.py:1499-1836) plus the sampler update, per GPU.  N = 1 workload: BASELINE.json
configs[1] -- unet.UNetModel, batch 256 latents 4x8x32, bf16 tensor-core compute, fp32 sampler state.  N > 1: every rank
runs its own batch of 256 (weak scaling, no per-step collective; one NCCL all-gather of the final latents closes the
timed trajectory).

`value` = latent-steps/s over all ranks with everything resident in HBM; `e2e` = the same through the drop-in
nn.Module forward with HOST (pinned) inputs and a host read of the predicted noise every step.
`--impl reference`: the reference's CPU path (the oracle port of it -- /root/reference cannot travel to the GPU box) on the
host cores, bounded sample, same metric and unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unet_denoise_latent_steps_per_sec"
UNIT = "latent-steps/s"
# algorithmic FLOPs per latent per UNet evaluation, step-dependent work only
GFLOP_PER_LATENT = {"unet": 9.153, "unetPhosc": 10.559}
MODEL_KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
                attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
                vocab_size=53, max_seq_len=10)


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/traffic.json, from the tools/summarize_ncu.py numbers of profiles/R4q_ncu_gemm.txt)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def _load_oracle():
    """The oracle (test infrastructure) is imported HERE and only here: by the CPU-baseline / `--impl reference` legs and by the
    labelled torch-eager comparator below -- never by the product path being measured."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet_oracle as UO
    import weights as W
    from diffusion_oracle import DiffusionOracle
    return UO, W, DiffusionOracle


def _cpu_sample(UO, W, DiffusionOracle, variant, batch, seconds):
    """latent-steps/s of the oracle port at one batch size over about `seconds` of host time (first evaluation = warm-up)."""
    import torch
    sd = W.make_state_dict(W.load_spec(variant), 1234)
    inp = W.make_inputs(batch, seed=1234)
    d = DiffusionOracle(1000)
    x = inp["x"].clone()
    phosc = inp["phosc"] if variant != "unet" else None
    n, t0, el = 0, None, 0.0
    i = 999
    with torch.no_grad():
        while True:
            t = torch.full((batch,), i, dtype=torch.long)
            # the reference re-encodes the context every step (unet.py:1626-1636): so does its port
            eps = UO.unet_forward(sd, x, t, inp["context"], inp["y"], phosc=phosc, variant=variant)
            x = d.ddpm_step(x, eps, i, torch.randn_like(x))
            i -= 1
            if t0 is None:       # first evaluation = warm-up
                t0 = time.perf_counter()
                continue
            n += 1
            el = time.perf_counter() - t0
            if el >= seconds or i <= 1:
                break
    return n, el


def cpu_oracle_throughput(model, seconds=12.0, batch=8, threads=None, sweep=True):
    """The reference's CPU path (oracle port, fp32, torch CPU ops on `threads` host threads): latent-steps/s of
    UNet evaluation + DDPM update on a bounded sample of the workload; plus short samples at batch 1 / 16 / 64 and
    the BASELINE config-1 trajectory time (batch 1, 999 steps) extrapolated from the batch-1 rate."""
    import torch
    UO, W, DiffusionOracle = _load_oracle()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    variant = "unet" if model == "unet" else "unetPhosc"
    n, el = _cpu_sample(UO, W, DiffusionOracle, variant, batch, seconds)
    out = dict(value=n * batch / el, unit=UNIT, cores=threads, kind="port",
               sample=f"{n} DDPM steps of the {variant} oracle port at batch {batch} (fp32, torch CPU, {threads} threads, "
                      f"{el:.1f} s), context re-encoded every step as in the reference")
    if sweep:
        by = {}
        for b, sec in ((1, 3.0), (16, 4.0), (64, 6.0)):
            nb, elb = _cpu_sample(UO, W, DiffusionOracle, variant, b, sec * seconds / 12.0)
            by[str(b)] = {"latent_steps_per_sec": round(nb * b / elb, 2), "ms_per_unet_eval": round(1e3 * elb / nb, 2), "steps": nb}
        out["by_batch"] = by
        out["config1_trajectory_s_extrapolated"] = round(999.0 / by["1"]["latent_steps_per_sec"], 1)
        out["config1"] = f"{variant}, batch 1, 999 DDPM steps on {threads} host threads, one UNet call per step (train.py:224-228 makes two)"
    return out


def gpu_eager_reference(model, batch, dev, steps=5):
    """Labelled comparator, NOT the reference arm (SURVEY 8d / BASELINE.md section 4 "also report"): what a user of the reference gets
    on this same B200 today -- the reference's forward as torch-eager ops (the oracle, a functional restatement that matches the
    reference modules bit for bit on the CPU) with the weights on the GPU, once in fp32 with torch's defaults (TF32 convolutions
    through cuDNN, fp32 matmuls) and once under autocast(bfloat16); same batch, context re-encoded every step as the reference
    does, CUDA-event timed after a warm-up.  cuDNN / cuBLAS kernels: none of this repo's code runs here."""
    import torch
    UO, W, DiffusionOracle = _load_oracle()
    variant = "unet" if model == "unet" else "unetPhosc"
    sd = {k: v.to(dev) for k, v in W.make_state_dict(W.load_spec(variant), 1234).items()}
    inp = {k: v.to(dev) for k, v in W.make_inputs(batch, seed=1234).items()}
    phosc = inp["phosc"] if variant != "unet" else None
    t = torch.full((batch,), 500, dtype=torch.long, device=dev)
    res = {"label": "torch-eager oracle forward on the same GPU (cuDNN / cuBLAS); comparator only, not the reference arm",
           "batch": batch, "model": variant}
    ref32 = None
    for name, ctxmgr in (("fp32_torch_defaults", None), ("autocast_bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
        def run():
            with torch.no_grad():
                if ctxmgr is None:
                    return UO.unet_forward(sd, inp["x"], t, inp["context"], inp["y"], phosc=phosc, variant=variant)
                with ctxmgr:
                    return UO.unet_forward(sd, inp["x"], t, inp["context"], inp["y"], phosc=phosc, variant=variant)
        for _ in range(2):
            out = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[name] = {"ms_per_step": round(ms, 3), "latent_steps_per_sec": round(batch / (ms * 1e-3), 1)}
        if ref32 is None:
            ref32 = out.float()
        else:
            res[name]["max_rel_vs_fp32_eager"] = float((out.float() - ref32).abs().max() / ref32.abs().max())
    del sd, inp
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8
    t0 = time.perf_counter()
    cb = cpu_oracle_throughput(args.model, seconds=max(2.0, 3.0 * args.steps), batch=batch, sweep=False)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * batch / cb["value"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.model} DDPM denoise step (UNet eval + update), CPU oracle port of the reference, "
                                  f"bounded sample batch {batch}", "batch_per_step": batch},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(out), flush=True)


def _build_model(cls, variant, dev, W):
    from worddiffusion_b200.unet_base import default_args
    m = cls(args=default_args(dev), **MODEL_KW)
    m.load_state_dict(W.make_state_dict(W.load_spec(variant), 1234), strict=True)
    return m.to(dev).eval()


def sharded_step_leg(model, variant, global_batch, world, rank, dev, steps, warmup, W):
    """K DDPM steps of a GLOBAL batch split over the ranks (contiguous shards, Philox keyed by the global latent index, no
    per-step collective) + the trajectory's all-gather, device-timed, max over ranks.  `strong_scaling`: BASELINE configs[1]'s
    256 latents over N GPUs; `config3`: unetPhosc, 1024 latents over N GPUs (BASELINE configs[2])."""
    import torch
    import torch.distributed as dist
    from worddiffusion_b200.diffusion import Diffusion, all_gather_latents, shard_bounds
    lo, hi = shard_bounds(global_batch, world, rank)
    inp = W.make_inputs(global_batch, seed=777)
    ctx, y = inp["context"][lo:hi].to(dev), inp["y"][lo:hi].to(dev)
    phosc = inp["phosc"][lo:hi].to(dev) if variant != "unet" else None
    diff = Diffusion(noise_steps=1000, device=dev)
    eng = model.engine(dev)
    eng.encode_context(ctx, phosc)
    x = inp["x"][lo:hi].to(dev).clone()
    T = diff.noise_steps

    def step(i, k):
        eng.sampler_step(x, i, y, 1, diff._ddpm_coef[i], philox_seed=1234, sample_offset=lo, step_index=k)

    for k in range(warmup):
        step(T - 1 - k, k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(T - 1 - warmup - k, warmup + k)
    if world > 1:
        all_gather_latents(x, global_batch, world)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"metric": METRIC, "model": variant, "global_batch": global_batch, "batch_per_gpu": hi - lo, "n_gpus": world,
            "scaling": "strong", "steps": steps, "ms_per_step": ms, "value": global_batch / (ms * 1e-3), "unit": UNIT,
            "gpu_launches_per_step": eng.last_launch_count,
            "step_tflops": round(global_batch * GFLOP_PER_LATENT[variant] * 1e9 / (ms * 1e-3) / 1e12, 1)}


def ddim50_sweep_leg(model, world, rank, dev, batches, W):
    """BASELINE configs[4]: DDIM eta = 0, 50 strided steps, unetPhosc2 (10 chars + 769 PHOSC tokens), GLOBAL batch sweep sharded
    over the ranks; whole trajectories (context encoding, device-Philox x_T, 50 fused steps, all-gather), second run timed."""
    import torch
    import torch.distributed as dist
    from worddiffusion_b200.diffusion import Diffusion
    d = Diffusion(noise_steps=1000, device=dev)
    rows = []
    for N in batches:
        if N < world:
            continue
        inp = W.make_inputs(N, seed=1234)
        ctx, y, ph = inp["context"].to(dev), inp["y"].to(dev), inp["phosc"].to(dev)
        ms = None
        for rep in range(2):  # first trajectory = warm-up (plan + arena allocation for this batch)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x = d.sample_latents_sharded(model, ctx, y, phosc=ph, seed=7, ddim_steps=50)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ok = bool(x.shape[0] == N and torch.isfinite(x).all())
        rows.append({"global_batch": N, "n_gpus": world, "trajectory_ms": round(ms, 3), "word_latents_per_sec": round(N / (ms * 1e-3), 2),
                     "latent_steps_per_sec": round(50 * N / (ms * 1e-3), 1), "finite": ok})
        del x, ctx, y, ph
    return {"workload": "unetPhosc2.UNetModelPhosc DDIM eta=0, 50 of 1000 timesteps, ctx 779 tokens; global batch sharded over the "
                        "ranks, one all-gather per trajectory", "scaling": "strong", "rows": rows}


def train_leg(args, world, rank, dev):
    """BASELINE.json configs[3]: noise-prediction training step (train.py:281-294), global batch 224 sharded over the ranks
    (strong scaling: 28 latents per GPU on 8 GPUs), forward + backward + NCCL gradient all-reduce + fused AdamW + EMA."""
    import torch
    import torch.distributed as dist
    import weights as W
    from worddiffusion_b200.diffusion import Diffusion
    from worddiffusion_b200.training import FusedTrainStep, shard_batch
    from worddiffusion_b200.unet import UNetModel, default_args
    gb = args.train_batch
    lo, hi = shard_batch(gb, rank, world)
    m = UNetModel(args=default_args(dev), **MODEL_KW)
    m.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
    m = m.to(dev).train()
    step = FusedTrainStep(m, lr=1e-4)
    diff = Diffusion(device=dev)
    inp = W.make_inputs(gb, seed=4321)
    lat = (torch.randn((gb, 4, 8, 32), generator=torch.Generator().manual_seed(99)) * 0.18215)[lo:hi].to(dev)
    ctx, y = inp["context"][lo:hi].to(dev), inp["y"][lo:hi].to(dev)

    def one():
        t = diff.sample_timesteps(hi - lo).to(dev)
        x_t, noise = diff.noise_images(lat, t)
        return step.step(x_t, t, ctx, y, noise)

    first = float(one())
    for _ in range(3):
        one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.train_steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.train_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    f, b = step.eng.launch_counts
    return {"metric": "train_latents_per_sec", "value": gb / (ms * 1e-3), "unit": "latents/s", "ms_per_step": ms,
            "global_batch": gb, "batch_per_gpu": hi - lo, "scaling": "strong", "steps": args.train_steps,
            "model_tflops": round(3 * (9.153 + 0.039) * 1e9 * gb / (ms * 1e-3) / 1e12, 1),
            "gpu_launches_per_step": f + b + 1, "loss_first": first, "loss_last": float(loss),
            "grad_allreduce": (f"{len(step.buckets)} buckets by backward stage, each launched (async, NCCL) behind the stage that "
                               "finishes it" if world > 1 and len(step.buckets) > 1 else "one all-reduce after the backward pass"
                               if world > 1 else "none (1 GPU)"),
            "workload": "unet.UNetModel noise-prediction training step: forward + backward + gradient all-reduce + AdamW + EMA "
                        "(bf16 tensor-core operands, fp32 master weights / accumulation)"}


def fp32_leg(args, model, diff, inp, dev, variant):
    """BASELINE.json configs[1] asks for the batch-256 sampling step in fp32 as well: the same module with
    ``precision = "fp32"`` (csrc/f32_path.cu: fp32 storage; Linear / 1x1 / 3x3 contractions on the split-TF32 tcgen05 kernel of
    csrc/f32_gemm_tc.cu, the rest fp32 SIMT; 1e-4 of the reference).  Rank 0 only, a few steps."""
    import torch
    B = args.batch
    ctx, y = inp["context"].to(dev), inp["y"].to(dev)
    phosc = inp["phosc"].to(dev) if variant != "unet" else None
    x = inp["x"].to(dev).clone()
    t_probe = torch.full((B,), 500, dtype=torch.long, device=dev)
    with torch.no_grad():   # the same evaluation in both precisions (bf16 engine first: the module is still in bf16 mode)
        e16 = model(x, phosc, timesteps=t_probe, context=ctx, y=y) if variant != "unet" else \
            model(x, None, timesteps=t_probe, context=ctx, y=y)
    model.precision = "fp32"
    try:
        eng = model.engine(dev)
        eng.encode_context(ctx, phosc)
        T = diff.noise_steps

        def step(i, k):
            eng.sampler_step(x, i, y, 1, diff._ddpm_coef[i], philox_seed=1234, step_index=k)

        step(T - 1, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.fp32_steps):
            step(T - 2 - k, 1 + k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.fp32_steps
        with torch.no_grad():
            e32 = model(inp["x"].to(dev), phosc, timesteps=t_probe, context=ctx, y=y) if variant != "unet" else \
                model(inp["x"].to(dev), None, timesteps=t_probe, context=ctx, y=y)
        launches = eng.last_launch_count
        ws = eng.workspace_bytes
    finally:
        model.precision = "bf16"
    tf = B * GFLOP_PER_LATENT[variant] * 1e9 / (ms * 1e-3) / 1e12
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    # ceiling of the split-TF32 route: kind::tf32 runs at half the bf16 rate and every fp32 product costs three MMAs
    tc_ceiling = load_peaks()["tf_burst"] / 2.0 / 3.0
    err = float((e16.double() - e32.double()).abs().max() / e32.double().abs().max())
    return {"metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": args.fp32_steps, "dtype": "f32",
            "step_tflops": round(tf, 2), "frac_of_fp32_ffma_peak": round(tf / peak, 4),
            "fp32_ffma_peak_tflops": round(peak, 1), "split_tf32_ceiling_tflops": round(tc_ceiling, 1),
            "frac_of_split_tf32_ceiling": round(tf / tc_ceiling, 4),
            "tensor_core_route": os.environ.get("WD_F32_TC", "1") != "0", "gpu_launches_per_step": launches + 1,
            "workspace_gb": round(ws / 1e9, 2), "bf16_engine_vs_fp32_mode_max_rel": err,
            "workload": f"{variant} DDPM sampling step in fp32 mode (fp32 storage; contractions as three kind::tf32 tcgen05 MMAs per K step "
                        f"on split operands with fp32 second-level accumulation, norms / attention fp32 SIMT; separate sampler-update "
                        f"kernel), batch {B}; parity 1e-4 vs the reference (tests/test_gpu_zfp32.py)"}


def vae_leg(dev, W, batch=64):
    """the VAE decode at the end of a sampling run (train.py:239-247) through worddiffusion_b200.vae.AutoencoderKL
    (wd_vae_decode, Stable Diffusion v1 decoder shape, random-init weights): latents [batch, 4, 8, 32] -> images [batch, 3, 64, 256]."""
    import torch
    from worddiffusion_b200.vae import AutoencoderKL
    vae = AutoencoderKL()
    spec = [(k, tuple(v.shape)) for k, v in vae.state_dict().items()]
    vae.load_state_dict(W.make_state_dict(spec, seed=77))
    vae.to(dev)
    z = torch.randn(batch, 4, 8, 32, device=dev)
    # builds the engine and sizes the arena for this batch (a smaller warm-up batch left the arena growth -- a multi-GB
    # cudaFree + cudaMalloc whose cost depends on what the earlier legs left allocated: 87 ms one run, 480 ms another -- inside
    # the timed call)
    vae.decode(z, scale=1 / 0.18215, postprocess=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img = vae.decode(z, scale=1 / 0.18215, postprocess=True).sample
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"metric": "vae_decode_images_per_sec", "value": batch / (ms * 1e-3), "unit": "images/s", "batch": batch, "ms": round(ms, 2),
            "image_shape": list(img.shape[1:]), "dtype": "f32", "finite": bool(torch.isfinite(img).all()),
            "workload": "AutoencoderKL decode (SD v1 decoder: 49.5 M parameters, 1.24 TFLOP per 64x64 latent -> 78 GFLOP per 8x32 latent) on the "
                        "fp32 path (split-TF32 tcgen05 contractions where the shapes allow, fp32 SIMT otherwise); parity 1e-5 vs "
                        "oracle/vae_oracle.py (tests/test_gpu_vae.py; parity unpinned: diffusers absent)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import weights as W
    from worddiffusion_b200.diffusion import Diffusion, all_gather_latents
    from worddiffusion_b200.unet import UNetModel, default_args
    from worddiffusion_b200.unetPhosc import UNetModelPhosc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    real_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout must carry exactly one JSON line, and NCCL writes its version banner to fd 1 when the first communicator is
        # created: everything else goes to stderr, the JSON line to the saved descriptor
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(dev))

    B, K, Wm = args.batch, args.steps, args.warmup
    variant = args.model
    cls = UNetModel if variant == "unet" else UNetModelPhosc
    model = cls(args=default_args(dev), **MODEL_KW)
    model.load_state_dict(W.make_state_dict(W.load_spec(variant), 1234), strict=True)
    model = model.to(dev).eval()
    inp = W.make_inputs(B, seed=1234 + rank)
    ctx = inp["context"].to(dev)
    y = inp["y"].to(dev)
    phosc = inp["phosc"].to(dev) if variant != "unet" else None
    diff = Diffusion(noise_steps=1000, device=dev)
    eng = model.engine(dev)
    eng.encode_context(ctx, phosc)      # time-invariant conditioning: once per trajectory
    x = inp["x"].to(dev).clone()
    T = diff.noise_steps

    def step(i, k):
        eng.sampler_step(x, i, y, 1, diff._ddpm_coef[i], philox_seed=1234, sample_offset=rank * B, step_index=k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    for k in range(Wm):
        step(T - 1 - k, k)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    ev0.record()
    for k in range(K):
        step(max(1, T - 1 - Wm - k), Wm + k)
    if world > 1:
        all_gather_latents(x, B * world, world)      # the trajectory's only collective
    ev1.record()
    barrier()
    w1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop(w0, w1)
    # Per-kernel device times (roofline leg): the SAME K steps again with a CUDA event between every launch.  This is a
    # separate pass on purpose: an event between two launches breaks their programmatic dependent launch (the next kernel's
    # prologue no longer overlaps the previous kernel's tail), so the profiled step is slower than the timed one above.
    eng.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for k in range(K):
        step(max(1, T - 1 - Wm - K - k), Wm + K + k)
    pe1.record()
    barrier()
    prof_ms = pe0.elapsed_time(pe1) / K
    n_steps_prof, prof = eng.profile_read()
    eng.set_profiling(False)
    launches = eng.last_launch_count * K
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---------------- end to end through the drop-in nn.Module with host buffers ----------------
    hx = inp["x"].pin_memory()
    ht = torch.full((B,), 500, dtype=torch.long).pin_memory()
    hctx, hy = inp["context"].pin_memory(), inp["y"].pin_memory()
    hph = inp["phosc"].pin_memory() if variant != "unet" else None
    heps = torch.empty((B, 4, 8, 32), dtype=torch.float32).pin_memory()

    def e2e_step():
        dx = hx.to(dev, non_blocking=True)
        dt = ht.to(dev, non_blocking=True)
        dc = hctx.to(dev, non_blocking=True)
        dy = hy.to(dev, non_blocking=True)
        with torch.no_grad():
            if variant == "unet":
                e = model(dx, None, timesteps=dt, context=dc, y=dy)
            else:
                e = model(dx, hph.to(dev, non_blocking=True), timesteps=dt, context=dc, y=dy)
        heps.copy_(e, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads heps on the host after every call

    for _ in range(max(1, min(Wm, 3))):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = hx.numel() * 4 + ht.numel() * 8 + hctx.numel() * 8 + hy.numel() * 8 + (hph.numel() * 8 if hph is not None else 0)
    d2h = heps.numel() * 4
    e2e = {"value": world * B * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms / K,
           "api": "worddiffusion_b200.%s forward(x, timesteps, context, y) with pinned host tensors; context re-encoded "
                  "per call" % ("unet.UNetModel" if variant == "unet" else "unetPhosc.UNetModelPhosc")}

    # training leg (configs[3]): EVERY rank takes part (sharded batch, gradient all-reduce) -- it must run before the
    # non-zero ranks leave
    train = None
    if args.train_steps > 0 and variant == "unet":
        train = train_leg(args, world, rank, dev)

    # sharded sampling legs: every rank takes part.  strong scaling of the headline workload (N = 1: the headline itself),
    # BASELINE configs[2] (unetPhosc, global 1024) and configs[4] (DDIM-50 sweep on unetPhosc2)
    extra = {}
    if variant == "unet" and not args.no_extra_legs:
        try:
            if world > 1:
                extra["strong_scaling"] = sharded_step_leg(model, "unet", B, world, rank, dev, K, Wm, W)
            else:
                extra["strong_scaling"] = {"metric": METRIC, "model": "unet", "global_batch": B, "batch_per_gpu": B, "n_gpus": 1,
                                           "scaling": "strong", "steps": K, "ms_per_step": ms / K, "value": value, "unit": UNIT,
                                           "note": "N = 1: the headline run itself"}
            from worddiffusion_b200.unetPhosc2 import UNetModelPhosc as UNetModelPhosc2
            mp = _build_model(UNetModelPhosc, "unetPhosc", dev, W)
            extra["config3_unetPhosc_b1024"] = sharded_step_leg(mp, "unetPhosc", 1024, world, rank, dev, max(5, K // 2), 3, W)
            del mp
            torch.cuda.empty_cache()
            mp2 = _build_model(UNetModelPhosc2, "unetPhosc", dev, W)
            extra["config5_ddim50_sweep"] = ddim50_sweep_leg(mp2, world, rank, dev, [int(b) for b in args.ddim_batches.split(",")], W)
            del mp2
            torch.cuda.empty_cache()
        except Exception as ex:   # a side leg must not cost the headline line (every rank fails the same way: no hang)
            extra["error"] = f"{type(ex).__name__}: {ex}"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # N = 1 only (like the CPU baseline): at N > 1 the other ranks have already left
    fp32 = None
    if args.fp32_steps > 0 and world == 1:
        try:
            fp32 = fp32_leg(args, model, diff, inp, dev, variant)
        except Exception as ex:   # a side leg must not cost the headline line
            fp32 = {"error": f"{type(ex).__name__}: {ex}"}
            model.precision = "bf16"

    # ---------------- roofline from the per-launch device times ----------------
    peaks = load_peaks()
    cls_t, cls_f, cls_b, cls_n = {}, {}, {}, {}
    for kind, fl, by, msum in prof:
        cls_t[kind] = cls_t.get(kind, 0.0) + msum / max(n_steps_prof, 1)
        cls_f[kind] = cls_f.get(kind, 0.0) + fl
        cls_b[kind] = cls_b.get(kind, 0.0) + by
        cls_n[kind] = cls_n.get(kind, 0) + 1
    total_t = sum(cls_t.values())
    kernels = {}
    for kname in cls_t:
        tms = cls_t[kname]
        kernels[kname] = {"launches_per_step": cls_n[kname], "ms_per_step": round(tms, 4),
                          "share": round(tms / total_t, 4) if total_t else None,
                          "tflops": round(cls_f[kname] / (tms * 1e-3) / 1e12, 2) if tms > 0 else None,
                          "gbs": round(cls_b[kname] / (tms * 1e-3) / 1e9, 1) if tms > 0 else None}
    if args.ops_out:
        with open(args.ops_out, "w") as f:
            json.dump([{"i": i, "kind": k, "gflop": fl / 1e9, "mbytes": by / 1e6, "us": 1e3 * msum / max(n_steps_prof, 1)}
                       for i, (k, fl, by, msum) in enumerate(prof)], f, indent=0)
    g = "gemm_tc"
    ach = cls_f[g] / (cls_t[g] * 1e-3) / 1e12
    roofline = {"kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM: 3x3 convs, 1x1 convs, linears)", "bound": "tensor",
                "achieved": round(ach, 2), "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": round(ach / peaks["tf_burst"], 4),
                "peak_source": f"{peaks['src']} bf16 burst (MEASURED_PEAKS.json bf16_tflops): the timed region is tens of ms and the "
                               "SM clock stays at its maximum (see `clocks`)",
                "peak_sustained": peaks["tf_sust"], "frac_sustained": round(ach / peaks["tf_sust"], 4),
                "flops_per_launch_avg": cls_f[g] / cls_n[g], "launches_per_step": cls_n[g],
                "avg_launch_ms": round(cls_t[g] / cls_n[g], 5), "share_of_step": kernels[g]["share"], "traffic": None,
                "timing": "CUDA events around every launch of the same K steps, run as a second pass right after the timed region "
                          "(events between launches switch off programmatic dependent launch: that pass took %.3f ms/step)" % prof_ms}
    if "tblock" in cls_t and cls_t["tblock"] > 0:
        # second-largest class: the fused transformer block (csrc/tblock.cu).  `tflops` counts the reference's operations
        # (to_q / QK^T / PV / to_out per attention); the kernel executes the algebraically folded form (8x fewer attention FLOPs)
        tb = cls_f["tblock"] / (cls_t["tblock"] * 1e-3) / 1e12
        roofline["fused_block"] = {"kernel": "tblock_unet_kernel (tcgen05 / TMEM / TMA, cta_group::2)", "bound": "tensor",
                                   "achieved_reference_ops": round(tb, 2), "frac": round(tb / peaks["tf_burst"], 4),
                                   "launches_per_step": cls_n["tblock"], "share_of_step": kernels["tblock"]["share"],
                                   "evidence": "profiles/R4q_ncu_tblock.txt, profiles/R2y_tblock_trace_pair1.txt"}
    tr = load_traffic()
    if tr and variant == "unet" and B == tr.get("batch"):
        roofline["traffic"] = tr["gemm_tc_kernel"]["dram_bytes_per_launch"]
        roofline["traffic_source"] = tr["gemm_tc_kernel"]["source"]
        roofline["algorithmic_bytes_per_launch_avg"] = cls_b[g] / cls_n[g]
    step_tf = B * GFLOP_PER_LATENT[variant] * 1e9 / (ms / K * 1e-3) / 1e12

    cb = cpu_oracle_throughput(variant, seconds=args.cpu_seconds) if world == 1 and args.cpu_seconds > 0 else None
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": f"{variant} DDPM sampling step (UNet eval + fused sampler update), batch {B} latents 4x8x32 "
                                  "per GPU, random-init weights (oracle/weights.py seed 1234)",
                      "batch_per_gpu": B, "global_batch": B * world, "noise_steps": 1000,
                      "l2": "inputs larger than L2: %.1f GB activation arena per step vs 126 MB L2" % (eng.workspace_bytes / 1e9)},
           "unet_steps_per_sec": K / (ms * 1e-3), "word_latents_per_sec_999_steps": value / 999.0,
           "step_tflops": round(step_tf, 1), "step_frac_of_bf16_burst": round(step_tf / peaks["tf_burst"], 4),
           "step_frac_of_bf16_sustained": round(step_tf / peaks["tf_sust"], 4),
           "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels}
    if cb is not None:
        out["cpu_baseline"] = cb
    if train is not None:
        out["train_step"] = train
    if fp32 is not None:
        out["fp32_mode"] = fp32
    if world == 1 and args.vae_batch > 0:
        try:
            out["vae_decode"] = vae_leg(dev, W, args.vae_batch)
        except Exception as ex:   # a side leg must not cost the headline line
            out["vae_decode"] = {"error": f"{type(ex).__name__}: {ex}"}
    out.update(extra)
    if world == 1 and args.eager_steps > 0:
        try:
            out["gpu_eager_reference"] = gpu_eager_reference(variant, B, dev, steps=args.eager_steps)
        except Exception as ex:
            out["gpu_eager_reference"] = {"error": f"{type(ex).__name__}: {ex}"}
    if real_stdout is not None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    else:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="unet", choices=["unet", "unetPhosc"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ops-out", default=None, help="write the per-launch device times of one step (JSON) to this path")
    ap.add_argument("--train-steps", type=int, default=10, help="timed steps of the training leg (0: skip it)")
    ap.add_argument("--fp32-steps", type=int, default=3, help="timed steps of the fp32-mode leg on rank 0 (0: skip it)")
    ap.add_argument("--train-batch", type=int, default=224, help="GLOBAL batch of the training leg (BASELINE config 4)")
    ap.add_argument("--vae-batch", type=int, default=64, help="latents of the VAE-decode leg on rank 0 (0: skip it)")
    ap.add_argument("--eager-steps", type=int, default=5, help="timed steps of the torch-eager-on-GPU comparator (0: skip it)")
    ap.add_argument("--ddim-batches", default="1,16,256,1024,4096", help="global batches of the DDIM-50 sweep leg (config 5)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip strong-scaling / config-3 / config-5 legs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
