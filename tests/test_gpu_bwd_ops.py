"""GPU parity of the single BACKWARD operators (C ABI `wd_op_*_bwd` / `wd_op_wgrad_*`, the kernels the trainer launches)
against torch autograd in fp32 on the SAME bf16-rounded operands.  Activation gradients are stored in bf16 (tolerance
2^-7 of the tensor's max: one bf16 rounding of the result plus fp32 accumulation-order noise); weight gradients are fp32
accumulators of bf16 x bf16 products (tolerance 1e-3 of the max)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, P, S, bf, conv3x3, f32, gemm, relerr, sync  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402

BF16_GRAD = 2 ** -7
F32_ACC = 1e-3


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(512, 320, 320), (7168, 320, 1280), (300, 2560, 320), (28, 1280, 320), (280, 960, 320),
                                   (1000, 320, 128), (4096, 64, 320)])
def test_wgrad_linear(M, N, K):
    x = bf(torch.randn(M, K, generator=g(1)))
    dy = bf(torch.randn(M, N, generator=g(2)))
    dw = torch.zeros(N, K, device=DEV)
    check(lib().wd_op_wgrad_linear(P(x), P(dy), P(dw), M, N, K, S()), "wgrad_linear")
    sync()
    ref = dy.float().t() @ x.float()
    assert relerr(dw, ref) < F32_ACC
    # accumulation semantics: a second call adds
    check(lib().wd_op_wgrad_linear(P(x), P(dy), P(dw), M, N, K, S()), "wgrad_linear")
    sync()
    assert relerr(dw, 2 * ref) < F32_ACC


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride", [(4, 8, 32, 320, 320, 1), (3, 4, 16, 640, 320, 1), (5, 8, 32, 320, 320, 2),
                                                    (2, 8, 32, 320, 64, 1), (28, 8, 32, 320, 320, 1)])
def test_wgrad_conv3x3(B, H, W, Cin, Cout, stride):
    x = bf(torch.randn(B, H, W, Cin, generator=g(3)))
    dy = bf(torch.randn(B, H // stride, W // stride, Cout, generator=g(4)))
    dw = torch.zeros(Cout, Cin, 3, 3, device=DEV)
    check(lib().wd_op_wgrad_conv3x3(P(x), P(dy), P(dw), B, H, W, Cin, Cout, stride, S()), "wgrad_conv3x3")
    sync()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros(Cout, Cin, 3, 3, device=DEV, requires_grad=True)
    y = F.conv2d(xr, w, padding=1, stride=stride)
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert relerr(dw, w.grad) < F32_ACC


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(3, 8, 32, 320, 320), (2, 4, 16, 640, 320)])
def test_dgrad_conv3x3_via_transposed_pack(B, H, W, Cin, Cout):
    """dX of a stride-1 conv = the forward implicit-GEMM kernel run on dY with the flipped / transposed weight pack."""
    w = torch.randn(Cout, Cin, 3, 3, generator=g(5)) / math.sqrt(9 * Cin)
    dy = bf(torch.randn(B, H, W, Cout, generator=g(6)))
    wt = torch.empty(Cin, 9 * Cout, device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_pack_conv3x3_t(P(f32(w)), P(wt), Cout, Cin, S()), "pack_conv3x3_T")
    dx = conv3x3(dy, wt, None)
    sync()
    wb = f32(w).bfloat16().float()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wb, padding=1).permute(0, 2, 3, 1)
    assert relerr(dx.float(), ref) < BF16_GRAD


def test_dgrad_linear_via_transposed_pack():
    M, N, K = 500, 320, 1280
    w = torch.randn(N, K, generator=g(7)) / math.sqrt(K)
    dy = bf(torch.randn(M, N, generator=g(8)))
    wt = torch.empty(K, N, device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_pack_linear_t(P(f32(w)), P(wt), N, K, S()), "pack_linear_T")
    dx = gemm(dy, wt)
    sync()
    ref = dy.float() @ f32(w).bfloat16().float()
    assert relerr(dx.float(), ref) < BF16_GRAD


@pytest.mark.parametrize("B,HW,C,silu,eps", [(3, 256, 320, 1, 1e-5), (2, 64, 320, 0, 1e-6), (2, 256, 640, 1, 1e-5)])
def test_groupnorm_bwd(B, HW, C, silu, eps):
    x = bf(torch.randn(B, HW, C, generator=g(9)) * 1.5 + 0.3)
    dy = bf(torch.randn(B, HW, C, generator=g(10)))
    gamma = f32(1 + 0.1 * torch.randn(C, generator=g(11)))
    beta = f32(0.1 * torch.randn(C, generator=g(12)))
    dx = torch.empty_like(x)
    dgamma = torch.zeros(C, device=DEV)
    dbeta = torch.zeros(C, device=DEV)
    check(lib().wd_op_groupnorm_bwd(P(x), P(dy), P(gamma), P(beta), P(dx), P(dgamma), P(dbeta), B, HW, C, 32, eps, silu, S()),
          "groupnorm_bwd")
    sync()
    xr = x.float().permute(0, 2, 1).reshape(B, C, HW, 1).requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gr, br, eps)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().permute(0, 2, 1).reshape(B, C, HW, 1))
    ref_dx = xr.grad.reshape(B, C, HW).permute(0, 2, 1)
    assert relerr(dx.float(), ref_dx) < BF16_GRAD
    assert relerr(dgamma, gr.grad) < F32_ACC
    assert relerr(dbeta, br.grad) < F32_ACC


@pytest.mark.parametrize("M,C,with_add", [(777, 320, True), (64, 320, False)])
def test_layernorm_bwd(M, C, with_add):
    x = bf(torch.randn(M, C, generator=g(13)) * 2 + 0.5)
    dy = bf(torch.randn(M, C, generator=g(14)))
    add = bf(torch.randn(M, C, generator=g(15))) if with_add else None
    gamma = f32(1 + 0.1 * torch.randn(C, generator=g(16)))
    dx = torch.empty_like(x)
    dgamma = torch.zeros(C, device=DEV)
    dbeta = torch.zeros(C, device=DEV)
    check(lib().wd_op_layernorm_bwd(P(x), P(dy), P(gamma), P(add), P(dx), P(dgamma), P(dbeta), M, C, 1e-5, S()), "layernorm_bwd")
    sync()
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = torch.zeros(C, device=DEV, requires_grad=True)
    F.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy.float())
    ref = xr.grad + (add.float() if with_add else 0)
    assert relerr(dx.float(), ref) < BF16_GRAD
    assert relerr(dgamma, gr.grad) < F32_ACC
    assert relerr(dbeta, br.grad) < F32_ACC


def test_geglu_fwd_bwd():
    M, H = 300, 1280
    p = bf(torch.randn(M, 2 * H, generator=g(17)))
    dout = bf(torch.randn(M, H, generator=g(18)))
    out = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    dp = torch.empty(M, 2 * H, device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_geglu_fwd(P(p), P(out), M, H, S()), "geglu_fwd")
    check(lib().wd_op_geglu_bwd(P(p), P(dout), P(dp), M, H, S()), "geglu_bwd")
    sync()
    pr = p.float().requires_grad_(True)
    a, gate = pr.chunk(2, dim=-1)
    y = a * F.gelu(gate)
    y.backward(dout.float())
    assert relerr(out.float(), y.detach()) < BF16_GRAD
    assert relerr(dp.float(), pr.grad) < BF16_GRAD


@pytest.mark.parametrize("B,Sq,L", [(3, 256, 10), (2, 64, 10), (2, 256, 16), (1, 300, 3)])
def test_attention_small_bwd(B, Sq, L):
    heads, dh = 4, 80
    C = heads * dh
    q = bf(torch.randn(B, Sq, C, generator=g(19)))
    k = bf(torch.randn(B, L, C, generator=g(20)))
    v = bf(torch.randn(B, L, C, generator=g(21)))
    do = bf(torch.randn(B, Sq, C, generator=g(22)))
    dq = torch.empty_like(q)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    scale = dh ** -0.5
    check(lib().wd_op_attention_small_bwd(P(q), P(k), P(v), P(do), P(dq), P(dk), P(dv), B, Sq, L, heads, scale, S()), "attn_bwd")
    sync()
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))

    def split(t):
        return t.reshape(B, t.shape[1], heads, dh).permute(0, 2, 1, 3)

    attn = (torch.einsum("bhid,bhjd->bhij", split(qr), split(kr)) * scale).softmax(-1)
    o = torch.einsum("bhij,bhjd->bhid", attn, split(vr)).permute(0, 2, 1, 3).reshape(B, Sq, C)
    o.backward(do.float())
    assert relerr(dq.float(), qr.grad) < BF16_GRAD
    assert relerr(dk.float(), kr.grad) < BF16_GRAD
    assert relerr(dv.float(), vr.grad) < BF16_GRAD
