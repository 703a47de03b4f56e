"""GPU parity of the training step (reference train.py:281-294) through the drop-in nn.Module -> ctypes -> C ABI -> sm_100a
kernels: `loss.backward()` on `worddiffusion_b200.unet.UNetModel` vs
  * the committed gradient signatures of the UNMODIFIED reference model (tests/golden/unet_train.npz), and
  * the CPU oracle's autograd (oracle/train_oracle.py) on fresh seeded inputs, tensor by tensor;
the fused AdamW + EMA kernel vs torch-restated AdamW / EMA; and a short optimisation run (loss must fall).
Tolerances (bf16 tensor-core compute, fp32 accumulation; written per assert): predicted noise 1e-2 max-rel (north_star);
per-parameter gradient relative L2 error 4e-2 (every activation gradient is rounded to bf16 once per layer)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import train_oracle as TO  # noqa: E402
import weights as W  # noqa: E402
from gpu_util import DEV, P, S, relerr  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402
from worddiffusion_b200.training import FusedTrainStep  # noqa: E402
from worddiffusion_b200.unet import UNetModel, default_args  # noqa: E402

SEED = 1234
KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)
GRAD_TOL = 4e-2


def _model():
    m = UNetModel(args=default_args(DEV), **KW)
    sd = W.make_state_dict(W.load_spec("unet"), SEED)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).train(), sd


def _rel_l2(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _backward(m, inp, noise):
    for p in m.parameters():
        p.grad = None
    pred = m(inp["x"].to(DEV), None, timesteps=inp["t"].to(DEV), context=inp["context"].to(DEV), y=inp["y"].to(DEV))
    loss = torch.nn.MSELoss()(noise.to(DEV), pred)
    loss.backward()
    return loss, pred


def test_backward_vs_reference_golden(golden_dir):
    m, sd = _model()
    g = np.load(os.path.join(golden_dir, "unet_train.npz"))
    inp = W.make_inputs(2, seed=SEED)
    loss, _ = _backward(m, inp, torch.from_numpy(g["noise"]))
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-2 * float(g["loss"])
    names = [n for n, _ in m.named_parameters()]
    assert names == list(sd.keys())
    dev = {}
    for i, (n, p) in enumerate(m.named_parameters()):
        if g["grad_norm"][i] < 0:
            assert p.grad is None, f"{n}: the reference gives no gradient to this parameter"
            continue
        assert p.grad is not None, n
        if g["grad_norm"][i] < 1e-6:
            continue  # rounding-noise-only gradient in the reference (key bias under a softmax)
        a, b = TO.signature(p.grad.cpu(), 77 + i)
        # norm, and projection on a random unit-variance direction (differs by at most ~|g - g_ref|)
        dev[n] = max(abs(a - g["grad_norm"][i]), abs(b - g["grad_proj"][i])) / g["grad_norm"][i]
    bad = {n: e for n, e in dev.items() if e >= GRAD_TOL}
    assert not bad, f"{len(bad)} of {len(dev)} gradients beyond {GRAD_TOL}: {sorted(bad.items(), key=lambda kv: -kv[1])[:10]}"
    worst = max(dev.values())
    for key in g.files:
        if key.startswith("grad::"):
            n = key[6:]
            assert _rel_l2(dict(m.named_parameters())[n].grad, torch.from_numpy(g[key])) < GRAD_TOL, n
    print(f"worst gradient-norm deviation vs the reference: {worst:.3e}")


@pytest.mark.parametrize("B", [3, 8])
def test_backward_vs_oracle_autograd(B):
    m, sd = _model()
    inp = W.make_inputs(B, seed=SEED + B)
    noise = torch.randn((B, 4, 8, 32), generator=torch.Generator().manual_seed(5 + B))
    loss, pred = _backward(m, inp, noise)
    ref_loss, ref_eps, ref = TO.unet_loss_and_grads(sd, inp["x"], inp["t"], inp["context"], inp["y"], noise)
    assert relerr(pred.detach(), ref_eps) < 1e-2
    assert abs(float(loss.detach()) - float(ref_loss)) < 1e-2 * float(ref_loss)
    errs = {}
    for n, p in m.named_parameters():
        if ref[n] is None:
            assert p.grad is None, n
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if float(ref[n].norm()) < 1e-6:
            continue
        errs[n] = _rel_l2(p.grad, ref[n])
    bad = {n: e for n, e in errs.items() if e >= GRAD_TOL}
    assert not bad, f"{len(bad)} of {len(errs)} gradients beyond {GRAD_TOL}: {sorted(bad.items(), key=lambda kv: -kv[1])[:8]}"
    print(f"B={B}: {len(errs)} gradients, worst rel-L2 {max(errs.values()):.3e}, median {sorted(errs.values())[len(errs) // 2]:.3e}")


def test_gradients_accumulate_like_autograd():
    """Two backward passes without zero_grad add up (torch semantics the reference loop relies on after zero_grad)."""
    m, _ = _model()
    inp = W.make_inputs(2, seed=SEED)
    noise = torch.randn((2, 4, 8, 32), generator=torch.Generator().manual_seed(1))
    _backward(m, inp, noise)
    g1 = m.out[2].weight.grad.clone()
    pred = m(inp["x"].to(DEV), None, timesteps=inp["t"].to(DEV), context=inp["context"].to(DEV), y=inp["y"].to(DEV))
    torch.nn.MSELoss()(noise.to(DEV), pred).backward()
    assert _rel_l2(m.out[2].weight.grad, 2 * g1) < 1e-3


def test_reference_loop_with_torch_optimizer_and_deepcopy():
    """train.py:403-411,281-294 verbatim: AdamW(model.parameters()), deepcopy EMA model, zero_grad/backward/step."""
    m, _ = _model()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    ema_model = copy.deepcopy(m).eval().requires_grad_(False)
    inp = W.make_inputs(4, seed=SEED)
    noise = torch.randn((4, 4, 8, 32), generator=torch.Generator().manual_seed(2)).to(DEV)
    losses = []
    for _ in range(6):
        pred = m(inp["x"].to(DEV), None, timesteps=inp["t"].to(DEV), context=inp["context"].to(DEV), y=inp["y"].to(DEV))
        loss = torch.nn.MSELoss()(noise, pred)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    with torch.no_grad():  # the deep-copied EMA model still runs (its own inference engine, old weights)
        e = ema_model(inp["x"].to(DEV), None, timesteps=inp["t"].to(DEV), context=inp["context"].to(DEV), y=inp["y"].to(DEV))
    assert torch.isfinite(e).all()


def test_adamw_ema_kernel_vs_restated_update():
    n = 100003
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=gen)
    m0 = torch.zeros(n)
    v0 = torch.zeros(n)
    p, m, v, ema = p0.clone().to(DEV), m0.clone().to(DEV), v0.clone().to(DEV), torch.zeros(n, device=DEV)
    pr, mr, vr, er = p0.clone(), m0.clone(), v0.clone(), torch.zeros(n)
    for step in range(1, 5):
        g = torch.randn(n, generator=gen) * 0.01
        gd = g.to(DEV)
        mode = 1 if step <= 2 else 2      # two warm-up copies, then the moving average
        check(lib().wd_adamw_ema_step(P(p), P(gd), P(m), P(v), P(ema), n, 1e-4, 0.9, 0.999, 1e-8, 0.01, step, 0.995, mode, 1.0,
                                      S()), "adamw")
        TO.adamw_update(pr, g, mr, vr, step)
        TO.ema_update(er, pr, 0 if mode == 1 else 2000)
    torch.cuda.synchronize()
    assert float((p.cpu() - pr).abs().max()) < 1e-6
    assert float((ema.cpu() - er).abs().max()) < 1e-6
    assert relerr(v.cpu(), vr) < 1e-4  # fp32 contraction order (FMA) differs between the kernel and torch


def test_fused_train_step_reduces_loss_and_tracks_ema():
    m, _ = _model()
    step = FusedTrainStep(m, lr=1e-4, step_start_ema=3)
    B = 8
    inp = W.make_inputs(B, seed=SEED)
    noise = torch.randn((B, 4, 8, 32), generator=torch.Generator().manual_seed(4)).to(DEV)
    x, t, c, y = (inp[k].to(DEV) for k in ("x", "t", "context", "y"))
    losses = [float(step.step(x, t, c, y, noise)) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses
    f, b = step.eng.launch_counts
    assert f > 60 and b > 150, (f, b)
    # the module's parameters ARE the flat buffer (updated in place by the kernel); the EMA lags behind them
    w = m.out[2].weight
    assert w.data_ptr() >= step.flat_param.data_ptr() and w.data_ptr() < step.flat_param.data_ptr() + 4 * step.flat_param.numel()
    esd = step.ema_state_dict()
    d = float((esd["out.2.weight"] - w.detach()).abs().max())
    assert 0 < d < 1e-2
    # inference through the same module sees the trained weights
    m.eval()
    with torch.no_grad():
        e = m(x, None, timesteps=t, context=c, y=y)
    assert float(torch.nn.functional.mse_loss(noise, e)) < losses[0]


def test_staged_backward_equals_the_whole_backward():
    """wd_trainer_backward_stages (the backward pass cut at layer boundaries for a bucketed, overlapped gradient
    all-reduce): every parameter's gradient is final after the stage wd_trainer_grad_stage names, the flat buffer is laid out in
    that order, and running the stages bucket by bucket gives the gradients of the single-call backward (the fp32 atomics of
    the weight-gradient epilogues make the two agree to rounding, not bit for bit).  Four rounds: eager, capture, replay."""
    from worddiffusion_b200.training import plan_grad_buckets
    m, _ = _model()
    eng = m.train_engine(DEV)
    eng.bind()
    eng.sync_weights()
    B = 4
    inp = {k: v.to(DEV) for k, v in W.make_inputs(B, seed=SEED + 5).items()}
    d_eps = torch.randn((B, 4, 8, 32), generator=torch.Generator().manual_seed(6)).to(DEV) / 1024
    st = [eng.stage_of[n] for n, _ in eng.live]
    assert eng.n_stages > 10 and st == sorted(st) and min(st) == 0 and max(st) == eng.n_stages - 1
    assert eng.stage_of["out.2.weight"] == 0 and eng.stage_of["time_embed.0.weight"] >= eng.stage_of["input_blocks.1.0.in_layers.2.weight"]
    offs = [eng.offsets[n] for n, _ in eng.live]
    assert offs == sorted(offs)
    buckets = eng.grad_buckets(4)
    assert len(buckets) == 4 and buckets[-1][0] == eng.n_stages and buckets[-1][2] == eng.flat_grad.numel()
    assert buckets == plan_grad_buckets(st, [(p.numel() + 63) // 64 * 64 for _, p in eng.live], eng.n_stages, 4)
    eng.forward(inp["x"], inp["t"], inp["y"], inp["context"])
    eng.backward(d_eps)
    whole = eng.flat_grad.clone()
    for rnd in range(4):
        eng.forward(inp["x"], inp["t"], inp["y"], inp["context"])
        s0 = 0
        for s1, lo, hi in buckets:
            eng.backward_stages(d_eps, s0, s1)
            # what an all-reduce launched here would read: already final
            part = eng.flat_grad[lo:hi].clone()
            assert relerr(part, whole[lo:hi]) < 1e-5, (rnd, s1)
            s0 = s1
        assert relerr(eng.flat_grad, whole) < 1e-5
    # stages out of order are refused
    from worddiffusion_b200._lib import WdError
    eng.forward(inp["x"], inp["t"], inp["y"], inp["context"])
    with pytest.raises(WdError, match="in order"):
        eng.backward_stages(d_eps, 1, 2)


def test_trainer_validates_every_extent():
    """the C side copies batch * C * H * W floats from the pointers it is handed; a wrong latent size / ragged
    conditioning must raise before any kernel runs, not read out of bounds."""
    from worddiffusion_b200._lib import WdError
    m, _ = _model()
    eng = m.train_engine(DEV)
    eng.bind()
    eng.sync_weights()
    inp = {k: v.to(DEV) for k, v in W.make_inputs(2, seed=1).items()}
    with pytest.raises(WdError, match="trainer built for latents"):
        eng.forward(torch.zeros(2, 4, 8, 16, device=DEV), inp["t"], inp["y"], inp["context"])
    with pytest.raises(WdError, match="timesteps"):
        eng.forward(inp["x"], inp["t"][:1], inp["y"], inp["context"])
    with pytest.raises(WdError, match="y must be"):
        eng.forward(inp["x"], inp["t"], inp["y"][:1], inp["context"])
    with pytest.raises(WdError, match="context"):
        eng.forward(inp["x"], inp["t"], inp["y"], inp["context"][:1])
    eng.forward(inp["x"], inp["t"], inp["y"], inp["context"])
    with pytest.raises(WdError, match="d_eps"):
        eng.backward(torch.zeros(1, 4, 8, 32, device=DEV))


def test_training_at_the_reference_default_latent_size():
    """train.Diffusion defaults to img_size = (64, 128) -> 8 x 16 latents (train.py:175): the drop-in forward builds a trainer for
    the latent size it is called with; eps and gradients vs the oracle's autograd."""
    m, sd = _model()
    B = 3
    inp = W.make_inputs(B, seed=SEED + 31, latent=(4, 8, 16))
    noise = torch.randn((B, 4, 8, 16), generator=torch.Generator().manual_seed(9))
    loss, pred = _backward(m, inp, noise)
    assert m._train_engine.latent_hw == (8, 16) and pred.shape == (B, 4, 8, 16)
    ref_loss, ref_eps, ref = TO.unet_loss_and_grads(sd, inp["x"], inp["t"], inp["context"], inp["y"], noise)
    assert relerr(pred.detach(), ref_eps) < 1.5e-2
    errs = {n: _rel_l2(p.grad, ref[n]) for n, p in m.named_parameters() if ref[n] is not None and float(ref[n].norm()) >= 1e-6}
    bad = {n: e for n, e in errs.items() if e >= GRAD_TOL}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]
    # back to 8 x 32: a new trainer, same parameters
    inp2 = W.make_inputs(2, seed=SEED)
    _backward(m, inp2, torch.randn((2, 4, 8, 32), generator=torch.Generator().manual_seed(1)))
    assert m._train_engine.latent_hw == (8, 32)


def test_fused_step_invalidates_both_inference_engines():
    """the AdamW kernel writes the parameters without bumping torch's version counters; after a fused step the bf16
    engine AND the fp32-mode engine must serve the updated weights (eps vs the oracle on the updated state_dict)."""
    import unet_oracle as UO
    m, _ = _model()
    inp = W.make_inputs(2, seed=SEED + 7)
    x, t, c, y = (inp[k].to(DEV) for k in ("x", "t", "context", "y"))
    m.eval()
    m.precision = "fp32"
    with torch.no_grad():
        e_before = m(x, None, timesteps=t, context=c, y=y).clone()   # builds the fp32 engine on the initial weights
    m.train()
    step = FusedTrainStep(m, lr=5e-4, use_ema=False)
    noise = torch.randn((2, 4, 8, 32), generator=torch.Generator().manual_seed(6)).to(DEV)
    for _ in range(3):
        step.step(x, t, c, y, noise)
    m.eval()
    sd_now = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    ref = UO.unet_forward(sd_now, inp["x"], inp["t"], inp["context"], inp["y"], variant="unet")
    with torch.no_grad():
        e32 = m(x, None, timesteps=t, context=c, y=y).clone()
        m.precision = "bf16"
        e16 = m(x, None, timesteps=t, context=c, y=y).clone()
    assert relerr(e_before, ref) > 1e-2, "the steps did not move the weights enough for this test to mean anything"
    assert relerr(e32, ref) < 1e-4
    assert relerr(e16, ref) < 1e-2


def test_data_parallel_training_two_gpus():
    """N = 2 ranks over NCCL (skipped on a single-GPU box): tools/ddp_check.py."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(root, "tools", "ddp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "DDP_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_noise_images_kernel_matches_the_reference_formula():
    """Diffusion.noise_images (train.py:190-194) as one kernel: exact given eps; Philox noise has unit moments and does not depend
    on how the batch is sharded."""
    from worddiffusion_b200.diffusion import Diffusion
    diff = Diffusion(noise_steps=1000, device=DEV)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(64, 4, 8, 32, generator=g).to(DEV)
    t = torch.randint(1, 1000, (64,), generator=g)
    eps = torch.randn(64, 4, 8, 32, generator=g).to(DEV)
    x_t, e = diff.noise_images(x, t, eps=eps)
    ah = diff.alpha_hat.to(DEV)[t.to(DEV)][:, None, None, None]
    want = torch.sqrt(ah) * x + torch.sqrt(1 - ah) * eps
    assert torch.equal(e, eps) and float((x_t - want).abs().max()) < 1e-6
    x_a, e_a = diff.noise_images(x, t, seed=9)
    x_b, e_b = diff.noise_images(x[40:], t[40:], seed=9, sample_offset=40)
    assert torch.equal(e_a[40:], e_b) and torch.equal(x_a[40:], x_b)
    assert abs(float(e_a.mean())) < 2e-2 and abs(float(e_a.std()) - 1.0) < 2e-2
    _, e1 = diff.noise_images(x, t)
    _, e2 = diff.noise_images(x, t)
    assert not torch.equal(e1, e2)          # successive training steps draw fresh noise
    with pytest.raises(IndexError):
        diff.noise_images(x, torch.full((64,), 1000))


def test_mse_loss_and_gradient_kernel():
    from worddiffusion_b200._lib import check, lib
    from gpu_util import P, S
    for n in (5, 4 * 8 * 32 * 3, 224 * 1024 + 17):
        g = torch.Generator().manual_seed(n)
        a = torch.randn(n, generator=g).to(DEV)
        b = torch.randn(n, generator=g).to(DEV)
        ws = torch.zeros(int(lib().wd_mse_workspace_bytes(n)), device=DEV, dtype=torch.uint8)
        d = torch.empty_like(a)
        loss = torch.empty((), device=DEV)
        for _ in range(2):  # the ticket counter re-arms itself
            check(lib().wd_mse_loss_grad(P(a), P(b), P(d), P(loss), P(ws), n, S()), "wd_mse_loss_grad")
        ref = torch.nn.functional.mse_loss(a.double(), b.double())
        assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, float(ref))
        assert float((d - (a - b) * (2.0 / n)).abs().max()) < 1e-7
