"""GPU parity of the fused transformer-block kernel (csrc/tblock.cu) as a single operator, through the C ABI: every stage of
the block (proj_in, cross-attention 1, cross-attention 2, GEGLU feed-forward, proj_out + residual) against a plain torch fp32
restatement of reference unet.py:337-345,381-412 on the same 16-bit-rounded inputs.  Tolerance: 2^-7 of max |ref| per stage
(bf16 / fp16 operands, fp32 accumulation; the whole-UNet budget is 1e-2)."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, P, S, relerr  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402

CH, HEADS, DH = 320, 4, 80
TOL = 2.0 ** -7


def _weights(seed):
    g = torch.Generator().manual_seed(seed)

    def lin(n, k, bias=True):
        b = 1.0 / math.sqrt(k)
        w = (torch.rand(n, k, generator=g) * 2 - 1) * b
        return w, ((torch.rand(n, generator=g) * 2 - 1) * 0.05 if bias else None)

    def norm():
        return 1.0 + 0.1 * torch.randn(CH, generator=g), 0.05 * torch.randn(CH, generator=g)
    W = {}
    W["pi_w"], W["pi_b"] = lin(CH, CH)
    W["n2_w"], W["n2_b"] = norm()
    W["n3_w"], W["n3_b"] = norm()
    for a in (1, 2):
        W[f"q{a}"], _ = lin(CH, CH, False)
        W[f"k{a}"], _ = lin(CH, CH, False)
        W[f"v{a}"], _ = lin(CH, CH, False)
        W[f"o{a}_w"], W[f"o{a}_b"] = lin(CH, CH)
    W["ff1_w"], W["ff1_b"] = lin(8 * CH, CH)
    W["ff2_w"], W["ff2_b"] = lin(CH, 4 * CH)
    W["po_w"], W["po_b"] = lin(CH, CH)
    return {k: v.to(DEV).float().contiguous() for k, v in W.items()}


ORDER = ["pi_w", "pi_b", "n2_w", "n2_b", "n3_w", "n3_b", "q1", "k1", "v1", "o1_w", "o1_b", "q2", "k2", "v2", "o2_w", "o2_b",
         "ff1_w", "ff1_b", "ff2_w", "ff2_b", "po_w", "po_b"]


def _plain_ln(x):
    return F.layer_norm(x, (CH,), None, None, 1e-5)


def _reference(W, g, x_in, ctx, B, HW, L, mid=False):
    """-> dict stage -> tensor [B*HW, 320] (fp32).  mid: x_in is the residual stream itself (no proj_in / proj_out)."""
    out = {}
    x = x_in if mid else g @ W["pi_w"].T + W["pi_b"]
    out[1] = _plain_ln(x)
    c = ctx.view(B, L, CH)
    for a in (1, 2):
        h = F.layer_norm(x, (CH,), W["n2_w"], W["n2_b"], 1e-5).view(B, HW, CH)   # unet.py:337-341: norm2 feeds both attentions
        q = (h @ W[f"q{a}"].T).view(B, HW, HEADS, DH).permute(0, 2, 1, 3)
        k = (c @ W[f"k{a}"].T).view(B, L, HEADS, DH).permute(0, 2, 1, 3)
        v = (c @ W[f"v{a}"].T).view(B, L, HEADS, DH).permute(0, 2, 1, 3)
        p = torch.softmax(q @ k.transpose(-1, -2) * DH ** -0.5, dim=-1)
        o = (p @ v).permute(0, 2, 1, 3).reshape(B * HW, CH)
        x = x + o @ W[f"o{a}_w"].T + W[f"o{a}_b"]
        out[1 + a] = _plain_ln(x)
    h = F.layer_norm(x, (CH,), W["n3_w"], W["n3_b"], 1e-5)
    pr = h @ W["ff1_w"].T + W["ff1_b"]
    val, gate = pr.chunk(2, dim=-1)
    x = x + (val * F.gelu(gate)) @ W["ff2_w"].T + W["ff2_b"]
    out[4] = x
    out[5] = x
    out[0] = x @ W["po_w"].T + W["po_b"] + x_in
    return out


def _run(W, g16, x16, ctx16, B, HW, L, stage, want_stats=False, gn=None):
    tensors = [g16, x16, ctx16] + [W[k] for k in ORDER] + (list(gn) if gn is not None else [])
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    out = torch.full((B * HW, CH), float("nan"), device=DEV, dtype=torch.float16)
    stats = torch.zeros((B, 32, HW // 32, 2), device=DEV, dtype=torch.float32) if want_stats else None
    check(lib().wd_op_tblock_unet(arr, len(tensors), B, HW, L, stage, P(out), P(stats), S()), "wd_op_tblock_unet")
    return out, stats


@pytest.mark.parametrize("B,HW,L", [(2, 256, 10), (5, 128, 16), (3, 256, 1), (160, 256, 10), (6, 64, 10), (5, 64, 16), (1, 64, 3), (301, 64, 10)])
def test_fused_block_every_stage(B, HW, L):
    W = _weights(3 + B)
    gen = torch.Generator().manual_seed(100 + B)
    g16 = torch.randn(B * HW, CH, generator=gen).to(DEV).to(torch.bfloat16)
    x16 = (torch.randn(B * HW, CH, generator=gen) * 1.5).to(DEV).to(torch.float16)
    ctx16 = torch.randn(B * L, CH, generator=gen).to(DEV).to(torch.bfloat16)
    ref = _reference(W, g16.float(), x16.float(), ctx16.float(), B, HW, L)
    errs = {}
    for stage in (1, 2, 3, 4, 0):
        out, stats = _run(W, g16, x16, ctx16, B, HW, L, stage, want_stats=(stage == 0))
        assert torch.isfinite(out.float()).all(), f"stage {stage}: non-finite output"
        errs[stage] = relerr(out.float(), ref[stage])
    print(f"B={B} HW={HW} L={L}: stage errors", {k: f"{v:.2e}" for k, v in errs.items()})
    for stage, e in errs.items():
        assert e < TOL, (stage, errs)
    # GroupNorm partial sums of the output: [sample][group of 10 channels][32-row slot][sum, sum of squares]
    o = ref[0].view(B, HW // 32, 32, 32, 10)                 # [b, slot, row, group, c]
    want = torch.stack([o.sum(dim=(2, 4)), (o * o).sum(dim=(2, 4))], dim=-1).permute(0, 2, 1, 3)   # [b, group, slot, 2]
    assert relerr(stats, want) < TOL


def test_fused_block_rows_do_not_mix():
    """Row-local by construction: permuting whole samples permutes the output (tiles of different samples share a CTA over time)."""
    B, HW, L = 6, 256, 10
    W = _weights(11)
    gen = torch.Generator().manual_seed(7)
    g16 = torch.randn(B, HW, CH, generator=gen).to(DEV).to(torch.bfloat16)
    x16 = torch.randn(B, HW, CH, generator=gen).to(DEV).to(torch.float16)
    ctx16 = torch.randn(B, L, CH, generator=gen).to(DEV).to(torch.bfloat16)
    a, _ = _run(W, g16.view(-1, CH), x16.view(-1, CH), ctx16.view(-1, CH), B, HW, L, 0)
    perm = torch.tensor([3, 0, 5, 1, 4, 2], device=DEV)
    b, _ = _run(W, g16[perm].reshape(-1, CH).contiguous(), x16[perm].reshape(-1, CH).contiguous(),
                ctx16[perm].reshape(-1, CH).contiguous(), B, HW, L, 0)
    assert torch.equal(a.view(B, HW, CH)[perm], b.view(B, HW, CH))


@pytest.mark.parametrize("B,HW,L", [(7, 64, 10), (256, 64, 10), (4, 256, 10), (3, 128, 7)])
def test_fused_block_middle_form(B, HW, L):
    """SpatialTransformers whose channel count differs from 320 (the 4 x 16 level, 640 channels) keep proj_in / proj_out as GEMMs:
    the kernel takes the fp16 residual stream and returns it after both attentions and the feed-forward (stage 5)."""
    W = _weights(40 + B)
    gen = torch.Generator().manual_seed(200 + B)
    g16 = torch.zeros(B * HW, CH).to(DEV).to(torch.bfloat16)
    x16 = torch.randn(B * HW, CH, generator=gen).to(DEV).to(torch.float16)
    ctx16 = torch.randn(B * L, CH, generator=gen).to(DEV).to(torch.bfloat16)
    ref = _reference(W, g16.float(), x16.float(), ctx16.float(), B, HW, L, mid=True)
    out, _ = _run(W, g16, x16, ctx16, B, HW, L, 5)
    assert torch.isfinite(out.float()).all()
    e = relerr(out.float(), ref[5])
    print(f"middle form B={B} HW={HW} L={L}: {e:.2e}")
    assert e < TOL


@pytest.mark.parametrize("B,HW,L", [(2, 256, 10), (5, 128, 16), (7, 64, 10), (96, 256, 10)])
def test_fused_block_with_its_input_groupnorm(B, HW, L):
    """SpatialTransformer.norm (GroupNorm32, eps 1e-6, no SiLU: unet.py:388,161-162) inside the kernel: the tile of x_in is normalised in
    shared memory from the producer's partial statistics before proj_in; everything downstream as in the 25-tensor form."""
    W = _weights(60 + B)
    gen = torch.Generator().manual_seed(300 + B)
    x16 = (torch.randn(B * HW, CH, generator=gen) * 1.5 + 0.3).to(DEV).to(torch.float16)
    ctx16 = torch.randn(B * L, CH, generator=gen).to(DEV).to(torch.bfloat16)
    gamma = (1.0 + 0.1 * torch.randn(CH, generator=gen)).to(DEV)
    beta = (0.05 * torch.randn(CH, generator=gen)).to(DEV)
    xf = x16.float().view(B, HW, CH)
    g = F.group_norm(xf.transpose(1, 2), 32, gamma, beta, eps=1e-6).transpose(1, 2).reshape(B * HW, CH)
    g16 = g.to(torch.bfloat16)  # the kernel rounds the normalised operand to bf16 like groupnorm_apply_bulk_kernel does
    ref = _reference(W, g16.float(), x16.float(), ctx16.float(), B, HW, L)
    dummy = torch.zeros_like(g16)
    out, stats = _run(W, dummy, x16, ctx16, B, HW, L, 0, want_stats=True, gn=(gamma, beta))
    e = relerr(out.float(), ref[0])
    s1, _ = _run(W, dummy, x16, ctx16, B, HW, L, 1, gn=(gamma, beta))
    e1 = relerr(s1.float(), ref[1])
    print(f"fused block with input GroupNorm B={B} HW={HW}: out {e:.2e}, after proj_in {e1:.2e}")
    assert e < TOL and e1 < TOL
    a, _ = _run(W, g16, x16, ctx16, B, HW, L, 0)   # the 25-tensor form on the pre-normalised operand
    assert relerr(out.float(), a.float()) < 2.0 ** -9
