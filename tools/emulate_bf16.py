#!/usr/bin/env python
"""Design study (CPU, test infrastructure): where does the bf16 error of the engine come from?
Re-runs the oracle's op graph with bf16 rounding applied at the points where the engine stores bf16
(weights, GN/LN outputs, GEMM outputs, residual stream) and reports max-rel error of eps vs the fp32 oracle for
several storage policies.  Usage: python tools/emulate_bf16.py [unet|unetPhosc] [seed]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet_oracle as UO  # noqa: E402
import weights as W  # noqa: E402


def r(t):
    return t.bfloat16().float()


class Emu:
    def __init__(self, sd, variant, policy):
        self.sd = sd
        self.variant = variant
        self.p = policy  # dict of flags
        self.wcache = {}

    def w(self, k):  # GEMM weights are bf16
        if k not in self.wcache:
            self.wcache[k] = r(self.sd[k])
        return self.wcache[k]

    def res(self, t):  # residual-stream storage
        return t if self.p.get("fp32_residual") else r(t)

    def mid(self, t):  # intermediate consumed only by a norm (h2) -- fp32 if policy says so
        return t if self.p.get("fp32_mid") else r(t)

    def lin(self, pfx, x, bias=True):
        return F.linear(x, self.w(pfx + ".weight"), self.sd[pfx + ".bias"] if bias else None)

    def gn(self, pfx, x, eps, silu):
        y = F.group_norm(x, 32, self.sd[pfx + ".weight"], self.sd[pfx + ".bias"], eps)
        return r(F.silu(y) if silu else y)

    def ln(self, pfx, x):
        return r(F.layer_norm(x, (x.shape[-1],), self.sd[pfx + ".weight"], self.sd[pfx + ".bias"], 1e-5))

    def resblock(self, pfx, x, emb_out):
        sd = self.sd
        a1 = self.gn(pfx + "in_layers.0", x, 1e-5, True)
        h = F.conv2d(a1, self.w(pfx + "in_layers.2.weight"), sd[pfx + "in_layers.2.bias"], padding=1)
        h = self.mid(h + emb_out[:, :, None, None])
        a2 = self.gn(pfx + "out_layers.0", h, 1e-5, True)
        h = F.conv2d(a2, self.w(pfx + "out_layers.3.weight"), sd[pfx + "out_layers.3.bias"], padding=1)
        if pfx + "skip_connection.weight" in sd:
            xs = r(x)  # A operand of the fused 1x1 skip conv is bf16
            x = F.conv2d(xs, self.w(pfx + "skip_connection.weight"), sd[pfx + "skip_connection.bias"])
        return self.res(x + h)

    def attn(self, pfx, xn, ctx, heads):
        q = r(self.lin(pfx + "to_q", xn, False))
        c = xn if ctx is None else ctx
        k = r(self.lin(pfx + "to_k", c, False))
        v = r(self.lin(pfx + "to_v", c, False))
        B, Sq, inner = q.shape
        d = inner // heads
        sp = lambda t: t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)
        sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, Sq, inner)
        return self.lin(pfx + "to_out.0", r(o))

    def st(self, pfx, x, ctx, heads):
        b, c, h, w = x.shape
        g = self.gn(pfx + "norm", x, 1e-6, False)
        t = F.conv2d(g, self.w(pfx + "proj_in.weight"), self.sd[pfx + "proj_in.bias"])
        t = self.res(t.permute(0, 2, 3, 1).reshape(b, h * w, -1))
        tp = pfx + "transformer_blocks.0."
        if self.variant == "unet":
            t = self.res(self.attn(tp + "attn1.", self.ln(tp + "norm2", t), ctx, heads) + t)
        else:
            t = self.res(self.attn(tp + "attn1.", self.ln(tp + "norm1", t), None, heads) + t)
        t = self.res(self.attn(tp + "attn2.", self.ln(tp + "norm2", t), ctx, heads) + t)
        hcat = self.lin(tp + "ff.net.0.proj", self.ln(tp + "norm3", t))
        a, gate = hcat.chunk(2, -1)
        t = r(self.lin(tp + "ff.net.2", r(a * F.gelu(gate))) + t)  # x3 feeds proj_out as a bf16 A operand
        t = t.reshape(b, h, w, -1).permute(0, 3, 1, 2)
        return self.res(F.conv2d(t, self.w(pfx + "proj_out.weight"), self.sd[pfx + "proj_out.bias"]) + x)

    def run_block(self, pfx, h, embs, ctx, heads):
        sd = self.sd
        for i, kind in enumerate(UO._block_layers(sd, pfx)):
            p = f"{pfx}{i}."
            if kind == "res":
                h = self.resblock(p, h, embs[p])
            elif kind == "st":
                h = self.st(p, h, ctx, heads)
            elif kind == "down":
                h = self.res(F.conv2d(r(h), self.w(p + "op.weight"), sd[p + "op.bias"], stride=2, padding=1))
            elif kind == "up":
                h = F.interpolate(r(h), scale_factor=2, mode="nearest")
                h = self.res(F.conv2d(h, self.w(p + "conv.weight"), sd[p + "conv.bias"], padding=1))
            else:  # conv_in: fp32 SIMT kernel
                h = self.res(F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1))
        return h

    def forward(self, x, t, context, y, phosc=None):
        sd = self.sd
        temb = r(UO.timestep_embedding(t, 320))
        h1 = r(F.silu(self.lin("time_embed.0", temb)))
        emb = self.lin("time_embed.2", h1) + F.embedding(y, sd["label_emb.weight"])
        emb_act = r(F.silu(emb))
        embs = {}
        for k in sd:
            if k.endswith("emb_layers.1.weight"):
                p = k[: -len("emb_layers.1.weight")]
                embs[p] = F.linear(emb_act, self.w(k), sd[p + "emb_layers.1.bias"])
        ctx = r(UO.encode_context(sd, context, phosc, variant=self.variant))
        h = x.float()
        hs = []
        i = 0
        while any(f"input_blocks.{i}.0.{s}" in sd for s in ("weight", "in_layers.0.weight", "op.weight")):
            h = self.run_block(f"input_blocks.{i}.", h, embs, ctx, 4)
            hs.append(h)
            i += 1
        h = self.run_block("middle_block.", h, embs, ctx, 4)
        i = 0
        while f"output_blocks.{i}.0.in_layers.0.weight" in sd:
            h = torch.cat([h, hs.pop()], 1)
            h = self.run_block(f"output_blocks.{i}.", h, embs, ctx, 4)
            i += 1
        a = self.gn("out.0", h, 1e-5, True)
        return F.conv2d(a, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "unet"
    seeds = [int(s) for s in sys.argv[2:]] or [1234, 99, 7]
    sd = W.make_state_dict(W.load_spec(variant), 1234)
    for seed in seeds:
        inp = W.make_inputs(2, seed=seed)
        ph = inp["phosc"] if variant != "unet" else None
        with torch.no_grad():
            ref = UO.unet_forward(sd, inp["x"], inp["t"], inp["context"], inp["y"], phosc=ph, variant=variant)
            for name, pol in [("all bf16 (engine today)", {}), ("fp32 mid (h2)", {"fp32_mid": 1}),
                              ("fp32 residual stream", {"fp32_residual": 1}),
                              ("fp32 residual + mid", {"fp32_residual": 1, "fp32_mid": 1})]:
                e = Emu(sd, variant, pol).forward(inp["x"], inp["t"], inp["context"], inp["y"], ph)
                err = float((e - ref).abs().max() / ref.abs().max())
                print(f"{variant} seed {seed:5d} {name:28s} max-rel err {err:.3e}")


if __name__ == "__main__":
    main()
