"""Helpers of the `-m gpu` parity tests: every kernel is called through the C ABI (ctypes), torch only owns memory."""
import ctypes as C

import torch

from worddiffusion_b200._lib import check, lib

DEV = "cuda:0"


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bf(t):
    return t.to(device=DEV, dtype=torch.bfloat16).contiguous()


def f32(t):
    return t.to(device=DEV, dtype=torch.float32).contiguous()


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def sync():
    torch.cuda.synchronize()


def pack_linear(w, geglu=False):
    N, K = w.shape
    dst = torch.empty(N, K, device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_pack_linear(P(f32(w)), P(dst), N, K, 1 if geglu else 0, S()), "pack_linear")
    return dst


def pack_conv(w):
    Cout, Cin = w.shape[:2]
    dst = torch.empty(Cout, 9 * Cin, device=DEV, dtype=torch.bfloat16)
    check(lib().wd_op_pack_conv3x3(P(f32(w)), P(dst), Cout, Cin, S()), "pack_conv3x3")
    return dst


def gemm(a, w_packed, bias=None, residual=None, silu=False, geglu=False, out_f32=False):
    M, K = a.shape
    N = w_packed.shape[0]
    out = torch.full((M, N // 2 if geglu else N), float("nan"), device=DEV,
                     dtype=torch.float32 if out_f32 else torch.bfloat16)
    check(lib().wd_op_gemm(P(a), P(w_packed), P(bias), P(residual), P(out), M, N, K, int(silu), int(geglu), int(out_f32),
                           S()), "wd_op_gemm")
    return out


def conv3x3(x_nhwc, w_packed, bias, rowbias=None, residual=None, stride=1):
    B, H, W, Cin = x_nhwc.shape
    Cout = w_packed.shape[0]
    out = torch.full((B, H // stride, W // stride, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    rb_ld = rowbias.shape[1] if rowbias is not None else 0
    check(lib().wd_op_conv3x3(P(x_nhwc), P(w_packed), P(bias), P(rowbias), rb_ld, P(residual), P(out), B, H, W, Cin, Cout,
                              stride, S()), "wd_op_conv3x3")
    return out


def conv3x3_gn_silu(x_nhwc, w_packed, bias, rowbias, gamma, beta, eps=1e-5):
    """conv3x3 + bias + row bias -> GroupNorm32 -> SiLU with the normalisation in the conv kernel's epilogue."""
    B, H, W, Cin = x_nhwc.shape
    Cout = w_packed.shape[0]
    out = torch.full((B, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ws = torch.zeros(B * 32 * max(H * W // 32, 1) * 2, device=DEV, dtype=torch.float32)
    rb_ld = rowbias.shape[1] if rowbias is not None else 0
    check(lib().wd_op_conv3x3_gn_silu(P(x_nhwc), P(w_packed), P(bias), P(rowbias), rb_ld, P(gamma), P(beta), float(eps), P(out),
                                      P(ws), B, H, W, Cin, Cout, S()), "wd_op_conv3x3_gn_silu")
    return out
