"""GPU parity of the single operators (C ABI `wd_op_*`, the same kernels the engine launches) against plain torch fp32
evaluated on the SAME bf16-rounded operands: the only differences left are accumulation order and the bf16 rounding of
the stored result, so the tolerances are tight (stated per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, P, S, bf, conv3x3, conv3x3_gn_silu, f32, gemm, pack_conv, pack_linear, relerr, sync  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402

BF16_STORE = 2 ** -8   # bf16 has 8 significant bits: storing fp32 -> bf16 costs <= 2^-9 relative, 2^-8 with slack


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (128, 160, 64), (2, 1280, 320), (200, 320, 1280), (1000, 960, 320),
                                   (512, 320, 3520)])
def test_gemm_plain(M, N, K):
    a = bf(torch.randn(M, K, generator=g(1)))
    w = torch.randn(N, K, generator=g(2)) / math.sqrt(K)
    wp = pack_linear(w)
    out = gemm(a, wp)
    sync()
    ref = a.float() @ wp.float().t()
    assert not torch.isnan(out.float()).any()
    assert relerr(out.float(), ref) < BF16_STORE


def test_gemm_epilogues():
    M, N, K = 300, 320, 320
    a = bf(torch.randn(M, K, generator=g(3)))
    w = torch.randn(N, K, generator=g(4)) / math.sqrt(K)
    wp = pack_linear(w)
    bias = f32(torch.randn(N, generator=g(5)))
    res = bf(torch.randn(M, N, generator=g(6)))
    base = a.float() @ wp.float().t()
    o = gemm(a, wp, bias=bias)
    assert relerr(o.float(), base + bias) < BF16_STORE
    o = gemm(a, wp, bias=bias, residual=res)
    assert relerr(o.float(), base + bias + res.float()) < BF16_STORE
    o = gemm(a, wp, bias=bias, silu=True)
    assert relerr(o.float(), F.silu(base + bias)) < BF16_STORE
    o = gemm(a, wp, bias=bias, out_f32=True)
    assert o.dtype == torch.float32 and relerr(o, base + bias) < 1e-5


@pytest.mark.parametrize("M,N,K", [(512, 320, 320), (300, 320, 320), (20000, 320, 320), (256, 320, 1280), (256, 640, 192)])
def test_gemm_fp16_residual_stream(M, N, K):
    """Linear + residual on the fp16 token stream (unet.py:337-345).  K <= 320: the residual is accumulated on the tensor pipe as
    extra K blocks against an identity tile (GemmArgs::res_k); K = 1280: the TMA-prefetched residual epilogue."""
    a = bf(torch.randn(M, K, generator=g(60)))
    w = torch.randn(N, K, generator=g(61)) / math.sqrt(K)
    wp = pack_linear(w)
    bias = f32(torch.randn(N, generator=g(62)))
    res = (torch.randn(M, N, generator=g(63)) * 3).to(device=DEV, dtype=torch.float16)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16)
    check(lib().wd_op_gemm_f16(P(a), P(wp), P(bias), P(res), P(out), M, N, K, S()), "gemm_f16")
    ref = a.float() @ wp.float().t() + bias + res.float()
    assert not torch.isnan(out.float()).any()
    assert relerr(out.float(), ref) < 2 ** -10  # fp16 storage: 11 significant bits


def test_gemm_geglu():
    """GEGLU (unet.py:122-129): proj -> chunk(2) -> a * gelu(gate), exact erf GELU."""
    M, K, inner = 256, 320, 1280
    a = bf(torch.randn(M, K, generator=g(7)))
    w = torch.randn(2 * inner, K, generator=g(8)) / math.sqrt(K)
    b = torch.randn(2 * inner, generator=g(9)) * 0.1
    wp = pack_linear(w, geglu=True)
    bp = torch.empty(2 * inner, device=DEV)
    check(lib().wd_op_pack_vec_geglu(P(f32(b)), P(bp), 2 * inner, S()), "pack_vec")
    o = gemm(a, wp, bias=bp, geglu=True)
    h = a.float() @ bf(w).float().t() + f32(b)
    val, gate = h.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    assert o.shape == (M, inner)
    assert relerr(o.float(), ref) < BF16_STORE


@pytest.mark.parametrize("B,H,W,Cin,stride", [(2, 8, 32, 320, 1), (3, 4, 16, 320, 1), (4, 4, 16, 640, 1), (2, 8, 32, 320, 2),
                                              (1, 8, 32, 64, 1), (5, 4, 16, 320, 1), (90, 8, 32, 320, 1), (90, 8, 32, 640, 1)])
def test_conv3x3(B, H, W, Cin, stride):
    """Implicit-GEMM 3x3 conv (pad 1): TMA zero-fill is the padding; stride 2 is the Downsample op (unet.py:540-551)."""
    Cout = 320
    x = bf(torch.randn(B, H, W, Cin, generator=g(10)))
    w = torch.randn(Cout, Cin, 3, 3, generator=g(11)) / math.sqrt(9 * Cin)
    bias = f32(torch.randn(Cout, generator=g(12)) * 0.1)
    rowbias = f32(torch.randn(B, Cout, generator=g(13)))
    wp = pack_conv(w)
    o = conv3x3(x, wp, bias, rowbias=rowbias, stride=stride)
    sync()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), bf(w).float(), bias, stride=stride, padding=1)
    ref = (ref + rowbias[:, :, None, None]).permute(0, 2, 3, 1)
    assert not torch.isnan(o.float()).any()
    assert relerr(o.float(), ref) < BF16_STORE


@pytest.mark.parametrize("B,H,W,Cin", [(2, 8, 32, 320), (3, 4, 16, 320), (4, 4, 16, 640), (5, 4, 16, 320), (1, 8, 32, 640),
                                       (148, 8, 32, 640), (2, 4, 8, 320), (90, 8, 32, 640), (95, 8, 32, 320)])
def test_conv3x3_groupnorm_silu_in_the_producer(B, H, W, Cin):
    """ResBlock front half (unet.py:657-667 then :592-594): conv3x3 + bias + emb row bias -> GroupNorm32 -> SiLU with the
    normalisation applied by the conv kernel's epilogue (CTA-pair kernel: the 256-row tile holds whole samples).  Covers one
    sample per pair tile (8x32), four / eight samples per tile, ragged last tiles (B = 3, 5) and more tiles than CTA pairs."""
    Cout = 320
    x = bf(torch.randn(B, H, W, Cin, generator=g(40)))
    w = torch.randn(Cout, Cin, 3, 3, generator=g(41)) / math.sqrt(9 * Cin)
    bias = f32(torch.randn(Cout, generator=g(42)) * 0.1)
    rowbias = f32(torch.randn(B, Cout, generator=g(43)))
    gamma = f32(1.0 + 0.2 * torch.randn(Cout, generator=g(44)))
    beta = f32(0.2 * torch.randn(Cout, generator=g(45)))
    o = conv3x3_gn_silu(x, pack_conv(w), bias, rowbias, gamma, beta, 1e-5)
    sync()
    h = F.conv2d(x.float().permute(0, 3, 1, 2), bf(w).float(), bias, padding=1) + rowbias[:, :, None, None]
    ref = F.silu(F.group_norm(h, 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert not torch.isnan(o.float()).any()
    assert relerr(o.float(), ref) < BF16_STORE
    # and it is the same function as the two-kernel form (conv3x3 then the GroupNorm kernel) up to the storage rounding
    h16 = conv3x3(x, pack_conv(w), bias, rowbias=rowbias).reshape(B, H * W, Cout).contiguous()
    two = torch.empty_like(h16)
    check(lib().wd_op_groupnorm(P(h16), P(two), P(gamma), P(beta), B, H * W, Cout, 32, 1e-5, 1, S()), "groupnorm")
    assert relerr(o.float().reshape(B, H * W, Cout), two.float()) < 3 * BF16_STORE


@pytest.mark.parametrize("B", [2, 90])
def test_conv3x3_residual(B):
    """B = 90: 90 pair tiles on 74 CTA pairs -> the 16 tiles of the last round are cut into 160-column halves (tail split)."""
    H, W, C = 8, 32, 320
    x = bf(torch.randn(B, H, W, C, generator=g(14)))
    w = torch.randn(C, C, 3, 3, generator=g(15)) / math.sqrt(9 * C)
    bias = f32(torch.zeros(C))
    res = bf(torch.randn(B, H, W, C, generator=g(16)))
    o = conv3x3(x, pack_conv(w), bias, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), bf(w).float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    assert relerr(o.float(), ref) < BF16_STORE


@pytest.mark.parametrize("B,HW,C,eps,silu", [(3, 256, 320, 1e-5, 1), (2, 64, 320, 1e-6, 0), (2, 256, 640, 1e-5, 1),
                                             (1, 64, 640, 1e-5, 1)])
def test_groupnorm(B, HW, C, eps, silu):
    """GroupNorm32 (+SiLU), fp32 statistics (unet.py:429-431; eps 1e-6 variant unet.py:161-162)."""
    x = bf(torch.randn(B, HW, C, generator=g(20)) * 2 + 0.5)
    gamma = f32(1 + 0.1 * torch.randn(C, generator=g(21)))
    beta = f32(0.1 * torch.randn(C, generator=g(22)))
    out = torch.empty_like(x)
    check(lib().wd_op_groupnorm(P(x), P(out), P(gamma), P(beta), B, HW, C, 32, eps, silu, S()), "groupnorm")
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert relerr(out.float(), ref.permute(0, 2, 1)) < BF16_STORE


def test_layernorm():
    M, C = 1000, 320
    x = bf(torch.randn(M, C, generator=g(23)) * 3 - 1)
    gamma = f32(1 + 0.1 * torch.randn(C, generator=g(24)))
    beta = f32(0.1 * torch.randn(C, generator=g(25)))
    out = torch.empty_like(x)
    check(lib().wd_op_layernorm(P(x), P(out), P(gamma), P(beta), M, C, 1e-5, S()), "layernorm")
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert relerr(out.float(), ref) < BF16_STORE


def _attn_ref(q, k, v, heads):
    B, Sq, C = q.shape
    d = C // heads
    qh, kh, vh = (t.float().reshape(B, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh) * d ** -0.5
    p = sim.softmax(-1)
    return torch.einsum("bhij,bhjd->bhid", p, vh).permute(0, 2, 1, 3).reshape(B, Sq, C), p


@pytest.mark.parametrize("Sq,L", [(256, 10), (64, 10), (256, 16), (64, 1)])
def test_attention_small(Sq, L):
    """Cross-attention over the 10-token character context (unet.py:185-279), incl. the attention-probability output."""
    B, heads, C = 3, 4, 320
    q, k, v = (bf(torch.randn(B, n, C, generator=g(30 + i))) for i, n in enumerate((Sq, L, L)))
    out = torch.empty_like(q)
    probs = torch.empty(B, heads, Sq, L, device=DEV)
    check(lib().wd_op_attention_small(P(q), P(k), P(v), P(out), P(probs), B, Sq, L, heads, 80 ** -0.5, S()), "attn_small")
    ref, p = _attn_ref(q, k, v, heads)
    assert relerr(out.float(), ref) < BF16_STORE
    assert relerr(probs, p) < 1e-4


@pytest.mark.parametrize("B,Sq,L,heads,with_bias", [(3, 256, 10, 4, True), (5, 128, 10, 4, False), (2, 256, 12, 4, True), (2, 256, 16, 4, False),
                                                    (2, 128, 1, 4, False), (150, 256, 10, 4, True)])
def test_q_projection_ctx_attention_fused(B, Sq, L, heads, with_bias):
    """to_q Linear + cross-attention over the short character context in ONE tcgen05 GEMM launch (the attention runs in the
    epilogue on the fp32 accumulator, unet.py:175-207).  Reference: torch fp32 on the same bf16 operands; q stays fp32 in both."""
    C = heads * 80
    a = bf(torch.randn(B * Sq, C, generator=g(50)))
    w = torch.randn(C, C, generator=g(51)) / math.sqrt(C)
    wp = pack_linear(w)
    bias = f32(torch.randn(C, generator=g(52)) * 0.2) if with_bias else None
    kv = bf(torch.randn(B, L, 2 * C, generator=g(53)))
    out = torch.full((B * Sq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_q_ctx_attention(P(a), P(wp), P(bias) if with_bias else None, P(kv), P(out), B, Sq, L, heads, 80 ** -0.5, S()),
          "q_ctx_attention")
    q = a.float() @ wp.float().t()
    if with_bias:
        q = q + bias
    ref, _ = _attn_ref(q.reshape(B, Sq, C), kv[..., :C], kv[..., C:], heads)
    assert not torch.isnan(out.float()).any()
    assert relerr(out.float().reshape(B, Sq, C), ref) < BF16_STORE


@pytest.mark.parametrize("Sq,Skv", [(256, 256), (64, 64), (256, 779), (64, 779), (100, 70), (256, 10), (64, 10), (100, 16),
                                    (64, 1)])
def test_attention_flash(Sq, Skv):
    """Self-attention (256/64 tokens), cross-attention over the 779-token char+PHOSC context (unetPhosc.py:176-196) and over
    the 10-token character context (Skv <= 16: the dedicated short-context kernel).
    P is rounded to bf16 before P.V (as flash kernels do): tolerance 2^-7."""
    B, heads, C = 2, 4, 320
    q, k, v = (bf(torch.randn(B, n, C, generator=g(40 + i))) for i, n in enumerate((Sq, Skv, Skv)))
    out = torch.full_like(q, float("nan"))
    check(lib().wd_op_attention(P(q), C, P(k), P(v), C, P(out), C, B, Sq, Skv, heads, 80 ** -0.5, S()), "attn_flash")
    ref, _ = _attn_ref(q, k, v, heads)
    assert not torch.isnan(out.float()).any()
    assert relerr(out.float(), ref) < 2 ** -7
