"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference UNet forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this file.  The product path (``worddiffusion_b200``) never does and has no CPU fallback.

Every function restates one reference symbol (file:line under /root/reference) as a pure function of a
``state_dict`` -- plain ``torch`` fp32 ops on the CPU, NCHW layout, no modules, no autograd.

Parity pinning: the reference ships no golden vectors or tests, so this restatement is pinned
against outputs of the *reference modules themselves*, run in the build container by ``oracle/make_golden.py``
(fixtures in ``tests/golden/``; checked by ``tests/test_oracle_golden.py``).
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000):
    """unet.py:96-116 (unetPhosc.py:89-109)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def positional_encoding(max_seq_len, dim):
    """CharacterEncoder.get_positional_encoding, unet.py:876-882."""
    pe = torch.zeros(max_seq_len, dim)
    for pos in range(max_seq_len):
        for i in range(0, dim, 2):
            pe[pos, i] = math.sin(pos / (10000 ** (i / dim)))
            pe[pos, i + 1] = math.cos(pos / (10000 ** ((i + 1) / dim)))
    return pe


def _lin(sd, pfx, x, bias=True):
    return F.linear(x, sd[pfx + ".weight"], sd[pfx + ".bias"] if bias else None)


def character_encoder(sd, tokens, max_seq_len, always_pe):
    """CharacterEncoder.forward + Word_Attention.forward.
    unet.py:851-874,825-836 (PE always added) ; unetPhosc.py:721-731,698-708 (PE only if len <= max_seq_len)."""
    x = F.embedding(tokens, sd["word_emb.embedding.weight"])
    L, D = x.shape[1], x.shape[2]
    if always_pe or L <= max_seq_len:
        x = x + positional_encoding(max_seq_len, D)[:L, :].to(x.device)
    q = _lin(sd, "word_emb.attention.linear_query", x)
    k = _lin(sd, "word_emb.attention.linear_key", x)
    v = _lin(sd, "word_emb.attention.linear_value", x)
    scores = torch.softmax(q @ k.transpose(-2, -1), dim=-1)  # NOT scaled by 1/sqrt(d)
    return scores @ v


def group_norm(sd, pfx, x, eps):
    """GroupNorm32 unet.py:429-431 (eps 1e-5) / Normalize unet.py:161-162 (eps 1e-6); 32 groups."""
    return F.group_norm(x.float(), 32, sd[pfx + ".weight"], sd[pfx + ".bias"], eps)


def res_block(sd, pfx, x, emb):
    """ResBlock._forward, unet.py:646-671 (use_scale_shift_norm=False, no up/down, dropout 0)."""
    h = F.silu(group_norm(sd, pfx + "in_layers.0", x, 1e-5))
    h = F.conv2d(h, sd[pfx + "in_layers.2.weight"], sd[pfx + "in_layers.2.bias"], padding=1)
    emb_out = _lin(sd, pfx + "emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = F.silu(group_norm(sd, pfx + "out_layers.0", h, 1e-5))
    h = F.conv2d(h, sd[pfx + "out_layers.3.weight"], sd[pfx + "out_layers.3.bias"], padding=1)
    if pfx + "skip_connection.weight" in sd:
        x = F.conv2d(x, sd[pfx + "skip_connection.weight"], sd[pfx + "skip_connection.bias"])
    return x + h


def cross_attention(sd, pfx, x, context, heads):
    """CrossAttention.forward, unet.py:185-279 / unetPhosc.py:176-198.  Returns (out, attn[B, heads, Sq, Skv])."""
    context = x if context is None else context
    q = _lin(sd, pfx + "to_q", x, bias=False)
    k = _lin(sd, pfx + "to_k", context, bias=False)
    v = _lin(sd, pfx + "to_v", context, bias=False)
    B, Sq, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(B, Sq, inner)
    return _lin(sd, pfx + "to_out.0", out), attn


def feed_forward(sd, pfx, x):
    """FeedForward / GEGLU, unet.py:122-149 (exact erf GELU)."""
    h = _lin(sd, pfx + "net.0.proj", x)
    a, gate = h.chunk(2, dim=-1)
    return _lin(sd, pfx + "net.2", a * F.gelu(gate))


def layer_norm(sd, pfx, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + ".weight"], sd[pfx + ".bias"], 1e-5)


def transformer_block(sd, pfx, x, context, heads, variant):
    """unet.py:337-345 (both attentions are cross-attention fed by norm2; norm1 unused)
    vs unetPhosc.py:241-246 (self-attn(norm1), cross-attn(norm2), ff(norm3))."""
    if variant == "unet":
        x1, _ = cross_attention(sd, pfx + "attn1.", layer_norm(sd, pfx + "norm2", x), context, heads)
        x = x1 + x
        x1, attn = cross_attention(sd, pfx + "attn2.", layer_norm(sd, pfx + "norm2", x), context, heads)
        x = x1 + x
    else:
        x1, _ = cross_attention(sd, pfx + "attn1.", layer_norm(sd, pfx + "norm1", x), None, heads)
        x = x1 + x
        x1, attn = cross_attention(sd, pfx + "attn2.", layer_norm(sd, pfx + "norm2", x), context, heads)
        x = x1 + x
    x = feed_forward(sd, pfx + "ff.", layer_norm(sd, pfx + "norm3", x)) + x
    return x, attn


def spatial_transformer(sd, pfx, x, context, heads, variant, depth=1, maps=None):
    """SpatialTransformer.forward, unet.py:381-412 / unetPhosc.py:282-300.  ``maps``: list that receives the attention
    probabilities of the last block's attn2, [B, heads, h, w, Skv] (what the attentionMaps = 1 variant returns, unet.py:402-410)."""
    b, c, h, w = x.shape
    x_in = x
    x = group_norm(sd, pfx + "norm", x, 1e-6)
    x = F.conv2d(x, sd[pfx + "proj_in.weight"], sd[pfx + "proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    for d in range(depth):
        x, attn = transformer_block(sd, f"{pfx}transformer_blocks.{d}.", x, context, heads, variant)
    if maps is not None:
        maps.append(attn.reshape(b, attn.shape[1], h, w, attn.shape[-1]))
    x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[pfx + "proj_out.weight"], sd[pfx + "proj_out.bias"])
    return x + x_in


def _block_layers(sd, pfx):
    """Sub-layer kinds of one TimestepEmbedSequential, recovered from the state_dict keys."""
    idx = 0
    kinds = []
    while True:
        p = f"{pfx}{idx}."
        if p + "in_layers.0.weight" in sd:
            kinds.append("res")
        elif p + "proj_in.weight" in sd:
            kinds.append("st")
        elif p + "op.weight" in sd:
            kinds.append("down")
        elif p + "conv.weight" in sd:
            kinds.append("up")
        elif p + "weight" in sd:
            kinds.append("conv")
        else:
            break
        idx += 1
    return kinds


def _run_block(sd, pfx, h, emb, context, heads, variant, maps=None):
    """TimestepEmbedSequential.forward, unet.py:452-469."""
    for i, kind in enumerate(_block_layers(sd, pfx)):
        p = f"{pfx}{i}."
        if kind == "res":
            h = res_block(sd, p, h, emb)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, heads, variant, maps=maps)
        elif kind == "down":  # Downsample unet.py:549-551
            h = F.conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)
        elif kind == "up":    # Upsample unet.py:490-500
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
        else:
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
    return h


def encode_context(sd, context, phosc=None, *, variant="unet", max_seq_len=10):
    """unet.py:1626-1636 ; unetPhosc.py:1117-1130 (PHOSC tokens embedded by the same encoder and concatenated)."""
    ctx = character_encoder(sd, context.long(), max_seq_len, always_pe=(variant == "unet"))
    if variant != "unet" and phosc is not None:
        ctx_p = character_encoder(sd, phosc.int().long(), max_seq_len, always_pe=False)
        ctx = torch.cat([ctx, ctx_p], dim=1)
    return ctx


ATTENTION_MAP_SCALES = (8, 16, 8)  # unet.py:1787,1791,1795


def canonical_state_dict(sd):
    """attentionMaps = 1 stores the middle block as middle_block1.{0.0, 0.1, 1.0}.* (unet.py:1336-1364): same layers, same order
    as middle_block.{0, 1, 2}.* -- returns ``sd`` under the attentionMaps = 0 names."""
    ren = (("middle_block1.0.0.", "middle_block.0."), ("middle_block1.0.1.", "middle_block.1."),
           ("middle_block1.1.0.", "middle_block.2."))
    out = {}
    for k, v in sd.items():
        for old, new in ren:
            if k.startswith(old):
                k = new + k[len(old):]
                break
        out[k] = v
    return out


def finish_attention_map(attn, scale):
    """unet.py:1786-1797: sum over heads, nearest-upsample (scale, scale): [B, heads, h, w, L] -> [B, scale h, scale w, L]."""
    a = attn.sum(dim=1)
    a = F.interpolate(a.permute(0, 3, 1, 2), scale_factor=(scale, scale), mode="nearest")
    return a.permute(0, 2, 3, 1)


def unet_forward(sd, x, timesteps, context, y, phosc=None, *, variant="unet", model_channels=320, heads=4,
                 max_seq_len=10, add_label_emb=True, ctx_encoded=None, attention_maps=False):
    """UNetModel.forward unet.py:1499-1836 (ocrTraining=0) / UNetModelPhosc.forward unetPhosc.py:1068-1159.
    ``sd``: fp32 CPU state_dict with the reference keys.  ``attention_maps`` (args.attentionMaps = 1, unet.py only): returns
    the reference's 5-tuple (eps, attn1, attn2, attn3, context) -- the maps of the last SpatialTransformer of the input blocks,
    of the middle block and of the output blocks (unet.py:1660,1703,1716: each assignment overwrites the previous one)."""
    sd = canonical_state_dict(sd)
    m_in, m_mid, m_out = ([], [], []) if attention_maps else (None, None, None)
    t_emb = timestep_embedding(timesteps, model_channels)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", t_emb)))
    if add_label_emb and "label_emb.weight" in sd:
        emb = emb + F.embedding(y.long(), sd["label_emb.weight"])
    ctx = ctx_encoded if ctx_encoded is not None else encode_context(sd, context, phosc, variant=variant,
                                                                     max_seq_len=max_seq_len)
    h = x.float()
    hs = []
    i = 0
    while f"input_blocks.{i}.0.weight" in sd or f"input_blocks.{i}.0.in_layers.0.weight" in sd or \
            f"input_blocks.{i}.0.op.weight" in sd:
        h = _run_block(sd, f"input_blocks.{i}.", h, emb, ctx, heads, variant, maps=m_in)
        hs.append(h)
        i += 1
    h = _run_block(sd, "middle_block.", h, emb, ctx, heads, variant, maps=m_mid)
    i = 0
    while f"output_blocks.{i}.0.in_layers.0.weight" in sd:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"output_blocks.{i}.", h, emb, ctx, heads, variant, maps=m_out)
        i += 1
    h = F.silu(group_norm(sd, "out.0", h, 1e-5))
    eps = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    if not attention_maps:
        return eps
    a1, a2, a3 = (finish_attention_map(m[-1], s) for m, s in zip((m_in, m_mid, m_out), ATTENTION_MAP_SCALES))
    return eps, a1, a2, a3, ctx
