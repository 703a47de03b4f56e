"""GPU parity of the fp32 mode (csrc/f32_path.cu, ``model.precision = "fp32"``) through the drop-in nn.Module -> ctypes -> C ABI.
Tolerance: BASELINE.json north_star -- fp32 mode, per-step predicted noise within 1e-4 max relative error (max |err| / max |ref|)
of the reference; the references are the committed outputs of the UNMODIFIED reference modules (tests/golden) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import unet_oracle as UO  # noqa: E402
import weights as W  # noqa: E402
from gpu_util import DEV, P, S, f32, relerr  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402
from worddiffusion_b200.diffusion import Diffusion  # noqa: E402
from worddiffusion_b200.unet import UNetModel, default_args  # noqa: E402
from worddiffusion_b200.unetPhosc import UNetModelPhosc  # noqa: E402

SEED = 1234
TOL_FP32 = 1e-4
KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)


def _model(cls, variant):
    m = cls(args=default_args(DEV), **KW)
    sd = W.make_state_dict(W.load_spec(variant), SEED)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    return m, sd


@pytest.fixture(scope="module")
def unet():
    return _model(UNetModel, "unet")


@pytest.fixture(scope="module")
def phosc():
    return _model(UNetModelPhosc, "unetPhosc")


def _cuda(inp):
    return {k: v.to(DEV) for k, v in inp.items()}


# ---------------------------------------------------------------- single operators
@pytest.mark.parametrize("stride,up,Cin,Cout,B,H,W", [(1, 0, 8, 12, 3, 8, 32), (2, 0, 320, 64, 2, 8, 32), (1, 1, 16, 320, 2, 4, 16),
                                                      (1, 0, 640, 320, 1, 4, 16)])
def test_f32_conv3x3_operator(stride, up, Cin, Cout, B, H, W):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    xi = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    want = F.conv2d(xi.double(), w.double(), b.double(), stride=stride, padding=1).permute(0, 2, 3, 1)
    out = torch.full(tuple(want.shape), float("nan"), device=DEV, dtype=torch.float32)
    xd, wd, bd = f32(x.permute(0, 2, 3, 1)), f32(w), f32(b)  # named: a temporary would be recycled by the next allocation
    check(lib().wd_f32_op_conv3x3(P(xd), P(wd), P(bd), P(out), B, H, W, Cin, Cout, stride, up, S()), "wd_f32_op_conv3x3")
    assert relerr(out, want) < 2e-6


@pytest.mark.parametrize("B,Sq,Skv,heads,d,scale", [(2, 64, 10, 4, 80, 80 ** -0.5), (1, 256, 779, 4, 80, 80 ** -0.5),
                                                    (3, 10, 10, 1, 320, 1.0), (2, 37, 5, 2, 160, 0.3)])
def test_f32_attention_operator(B, Sq, Skv, heads, d, scale):
    g = torch.Generator().manual_seed(4)
    q, k, v = (torch.randn(B, s, heads * d, generator=g) for s in (Sq, Skv, Skv))

    def split(t):
        return t.double().reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)

    attn = (split(q) @ split(k).transpose(-1, -2) * scale).softmax(-1)
    want = (attn @ split(v)).permute(0, 2, 1, 3).reshape(B, Sq, heads * d)
    out = torch.full((B, Sq, heads * d), float("nan"), device=DEV, dtype=torch.float32)
    qd, kd, vd = f32(q), f32(k), f32(v)
    check(lib().wd_f32_op_attention(P(qd), P(kd), P(vd), P(out), B, Sq, Skv, heads, d, scale, S()), "wd_f32_op_attention")
    assert relerr(out, want) < 5e-6


# ---------------------------------------------------------------- whole UNet
def test_unet_fp32_vs_reference_golden(unet, golden_dir):
    m, _ = unet
    g = np.load(os.path.join(golden_dir, "unet_fwd.npz"))
    inp = _cuda(W.make_inputs(2, seed=SEED))
    with torch.no_grad():
        eps = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert eps.shape == (2, 4, 8, 32) and eps.dtype == torch.float32
    assert m._engine_f32.last_launch_count > 100 and m._engine is None  # the fp32 kernels ran, the bf16 engine was never built
    err = relerr(eps, torch.from_numpy(g["eps"]))
    print("unet fp32-mode eps max-rel err vs reference:", err)
    assert err < TOL_FP32


def test_unet_phosc_fp32_vs_reference_golden(phosc, golden_dir):
    m, _ = phosc
    g = np.load(os.path.join(golden_dir, "unetPhosc_fwd.npz"))
    inp = _cuda(W.make_inputs(2, seed=SEED))
    with torch.no_grad():
        eps = m(inp["x"], inp["phosc"], timesteps=inp["t"], context=inp["context"], y=inp["y"])
    err = relerr(eps, torch.from_numpy(g["eps"]))
    print("unetPhosc fp32-mode eps max-rel err vs reference:", err)
    assert err < TOL_FP32


def test_unet_fp32_vs_oracle_fresh_inputs(unet):
    m, sd = unet
    inp = W.make_inputs(5, seed=99)
    ci = _cuda(inp)
    with torch.no_grad():
        eps = m(ci["x"], None, timesteps=ci["t"], context=ci["context"], y=ci["y"])
    ref = UO.unet_forward(sd, inp["x"], inp["t"], inp["context"], inp["y"], variant="unet")
    err = relerr(eps, ref)
    print("unet fp32-mode eps max-rel err vs oracle:", err)
    assert err < TOL_FP32


def test_ddpm_trajectory_fp32_vs_reference_golden(unet, golden_dir):
    """T = 6 trajectory of train.py:217-236 with the reference's own pre-generated noise: per-step eps and final latent at 1e-4."""
    m, _ = unet
    g = np.load(os.path.join(golden_dir, "unet_ddpm_T6.npz"))
    inp = _cuda(W.make_inputs(2, seed=SEED))
    d = Diffusion(noise_steps=6, device=DEV)
    x, trace = d.sample_latents(m, inp["context"], inp["y"], x_T=torch.from_numpy(g["x_T"]),
                                noise=torch.from_numpy(g["noises"]), return_eps_trace=True)
    for k, e in enumerate(trace):
        err = relerr(e, torch.from_numpy(g["eps_steps"][k]))
        print("step", k, "fp32-mode eps err", err)
        assert err < TOL_FP32
    err = relerr(x, torch.from_numpy(g["x_final"]))
    print("fp32-mode final latent err", err)
    assert err < TOL_FP32, "final latent after 5 steps, fp32 mode: 1e-4 of max |x|"


def test_fp32_and_bf16_modes_agree(unet):
    """The two precisions of the same module: the bf16 engine stays within its 1e-2 of the fp32 mode."""
    m, _ = unet
    inp = _cuda(W.make_inputs(4, seed=17))
    with torch.no_grad():
        e32 = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
        m.precision = "bf16"
        try:
            e16 = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
        finally:
            m.precision = "fp32"
    err = relerr(e16, e32)
    print("bf16 engine vs fp32 mode:", err)
    assert 0 < err < 1e-2


def test_fp32_batch_invariance(unet):
    m, _ = unet
    big = _cuda(W.make_inputs(9, seed=21))
    with torch.no_grad():
        e_big = m(big["x"], None, timesteps=big["t"], context=big["context"], y=big["y"])
        idx = torch.tensor([0, 4, 8], device=DEV)
        e_small = m(big["x"][idx], None, timesteps=big["t"][idx], context=big["context"][idx], y=big["y"][idx])
    assert torch.isfinite(e_big).all()
    assert torch.equal(e_big[idx], e_small)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("bf16", 1e-2)])
@pytest.mark.parametrize("B,L", [(1, 10), (1, 7), (3, 4)])
def test_ragged_context_length_and_single_latent(unet, precision, tol, B, L):
    """Edge cases of the seam: one latent, and a context shorter than max_seq_len (unet.py:872 adds pe[:L]) -- both precisions."""
    m, sd = unet
    inp = W.make_inputs(B, seed=31 + L, L=L)
    ci = _cuda(inp)
    m.precision = precision
    try:
        with torch.no_grad():
            eps = m(ci["x"], None, timesteps=ci["t"], context=ci["context"], y=ci["y"])
    finally:
        m.precision = "fp32"
    ref = UO.unet_forward(sd, inp["x"], inp["t"], inp["context"], inp["y"], variant="unet")
    err = relerr(eps, ref)
    print(f"B={B} L={L} {precision}: eps max-rel err vs oracle {err:.3e}")
    assert err < tol


def test_empty_batch_returns_empty(unet):
    m, _ = unet
    z = torch.zeros((0, 4, 8, 32), device=DEV)
    with torch.no_grad():
        out = m(z, None, timesteps=torch.zeros(0, dtype=torch.long, device=DEV),
                context=torch.zeros((0, 10), dtype=torch.long, device=DEV), y=torch.zeros(0, dtype=torch.long, device=DEV))
    assert out.shape == (0, 4, 8, 32)


def test_attention_maps_variant_vs_reference_golden(golden_dir):
    """args.attentionMaps == 1 (unet.py:1336-1364,1645-1836): middle_block1 checkpoint layout, 5-tuple return, maps and context
    against the UNMODIFIED reference's (oracle/make_golden_attnmaps.py)."""
    ren = {"middle_block.0.": "middle_block1.0.0.", "middle_block.1.": "middle_block1.0.1.", "middle_block.2.": "middle_block1.1.0."}
    sd1 = {}
    for k, v in W.make_state_dict(W.load_spec("unet"), SEED).items():
        for old, new in ren.items():
            if k.startswith(old):
                k = new + k[len(old):]
                break
        sd1[k] = v
    m = UNetModel(args=default_args(DEV, attentionMaps=1), **KW)
    m.load_state_dict(sd1, strict=True)
    m = m.to(DEV).eval()
    g = np.load(os.path.join(golden_dir, "unet_attnmaps.npz"))
    g0 = np.load(os.path.join(golden_dir, "unet_fwd.npz"))
    inp = _cuda(W.make_inputs(2, seed=SEED))
    with torch.no_grad():
        out = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert isinstance(out, tuple) and len(out) == 5
    eps, a1, a2, a3, ctx = out
    assert relerr(eps, torch.from_numpy(g0["eps"])) < TOL_FP32
    assert ctx.shape == (2, 10, 320) and relerr(ctx, torch.from_numpy(g["context"])) < TOL_FP32
    for a, key, s in zip((a1, a2, a3), ("attn1", "attn2", "attn3"), g["scales"]):
        s = int(s)
        assert a.shape == (2, 64, 256, 10) and a.dtype == torch.float32
        sub = a[:, ::s, ::s]
        assert torch.equal(sub.repeat_interleave(s, 1).repeat_interleave(s, 2), a)
        err = relerr(sub, torch.from_numpy(g[key]))
        print(key, "max-rel err vs reference:", err)
        assert err < TOL_FP32
    m.precision = "bf16"
    with pytest.raises(NotImplementedError):
        m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])


@pytest.mark.parametrize("M,N,K", [(128, 160, 32), (256, 320, 320), (128, 160, 544), (512, 320, 2880), (256, 640, 5760),
                                   (256, 512, 512), (384, 128, 4608), (20096, 256, 2304)])
def test_f32_tc_gemm_operator(M, N, K):
    """C = A W^T + bias through three kind::tf32 MMAs per K step on pre-split operands, K walked in chunks of 320 whose TMEM
    accumulators are summed in fp32 registers: fp32-class accuracy (vs fp64 torch) whatever K.  (With ONE TMEM accumulation the
    error grew with K -- 1.6e-6 at K = 320, 1.3e-5 at K = 2880, profiles/r03h_* -- because the tensor core's accumulation truncates.)"""
    g = torch.Generator().manual_seed(6)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    want = a.double() @ w.double().t() + b.double()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    ad, wd_, bd = f32(a), f32(w), f32(b)
    check(lib().wd_f32_op_gemm_tc(P(ad), P(wd_), P(bd), P(out), M, N, K, S()), "wd_f32_op_gemm_tc")
    err = relerr(out, want)
    print(f"split-TF32 tcgen05 GEMM {M}x{N}x{K}: max-rel err vs fp64 {err:.3e}")
    assert err < 5e-6


@pytest.mark.parametrize("B,H,W,C1,C2,Cout", [(2, 8, 32, 320, 0, 320), (4, 4, 16, 320, 320, 320), (1, 8, 32, 64, 32, 160),
                                              (2, 8, 16, 640, 0, 640), (8, 4, 4, 32, 0, 160), (1, 16, 64, 512, 0, 512),
                                              (2, 64, 256, 128, 0, 128)])
def test_f32_tc_conv3x3_operator(B, H, W, C1, C2, Cout):
    """3x3 pad-1 convolution as an implicit GEMM on the split-TF32 kernel: shifted 4-D TMA boxes of the split NHWC sources (zero fill =
    padding), channel concatenation of two sources (decoder skip, unet.py:1750), tiles inside one image and tiles over several."""
    g = torch.Generator().manual_seed(B * 100 + C1)
    x1 = torch.randn(B, C1, H, W, generator=g)
    x2 = torch.randn(B, C2, H, W, generator=g) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, generator=g) / (3 * (C1 + C2) ** 0.5)
    b = torch.randn(Cout, generator=g)
    xin = torch.cat([x1, x2], dim=1) if C2 else x1
    want = F.conv2d(xin.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    out = torch.full(tuple(want.shape), float("nan"), device=DEV, dtype=torch.float32)
    x1d = f32(x1.permute(0, 2, 3, 1))
    x2d = f32(x2.permute(0, 2, 3, 1)) if C2 else None
    wd_, bd = f32(w), f32(b)
    check(lib().wd_f32_op_conv3x3_tc(P(x1d), P(x2d), P(wd_), P(bd), P(out), B, H, W, C1, C2, Cout, S()), "wd_f32_op_conv3x3_tc")
    err = relerr(out, want)
    print(f"split-TF32 implicit conv B={B} {H}x{W} {C1}+{C2}->{Cout}: {err:.3e}")
    assert err < 5e-6
