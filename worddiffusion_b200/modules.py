"""Parameter containers that reproduce the reference ``state_dict`` layout.

The reference hot path is a ``torch.nn.Module`` whose ``state_dict`` key set is its compatibility contract
(264 keys for ``unet.py``, 246 for ``unetPhosc*.py``, including parameters its forward
never reads).  The classes below hold exactly those parameters, under exactly those names and with the
reference's initialisation (torch defaults + ``zero_module``), but they contain **no arithmetic**: the forward
pass is executed by the sm_100a engine (``engine.py`` -> ``libwd_b200.so``).  Calling one of these holders
directly raises.

Reference constructors mirrored here: ``ResBlock`` unet.py:554-644, ``CrossAttention`` :164-183,
``BasicTransformerBlock`` :305-318, ``SpatialTransformer`` :347-380, ``Upsample`` :472-488, ``Downsample``
:515-547, ``CharacterEncoder`` / ``Word_Attention`` :815-882, ``ResBlockConditional`` :886-1050.
"""
import math

import torch
import torch.nn as nn


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class _Holder(nn.Module):
    """A module that only owns parameters; its arithmetic lives in the CUDA engine."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is a parameter holder of the B200 engine; "
                           "call the enclosing UNetModel instead (there is no eager fallback)")


class ResBlock(_Holder):
    def __init__(self, channels, emb_channels, out_channels=None, conv_skip=False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=0.0),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif conv_skip:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)


class _ConvOp(_Holder):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.op = nn.Conv2d(cin, cout, 3, stride=stride, padding=1)


class ResBlockConditional(ResBlock):
    """``self.res`` of unet.UNetModel (unet.py:1472): constructed, never called (``if 0`` at :1593)."""

    def __init__(self, channels, emb_channels, out_channels):
        super().__init__(channels, emb_channels, out_channels, conv_skip=True)
        self.h_upd = _ConvOp(channels, channels, 2)
        self.x_upd = _ConvOp(channels, channels, 2)
        for n in ("emb_layers", "out_layers", "skip_connection"):  # keep the reference registration order
            self._modules[n] = self._modules.pop(n)


class Downsample(_Holder):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)


class Upsample(_Holder):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)


class CrossAttention(_Holder):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        context_dim = context_dim if context_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)  # dead in the reference forward
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))


class GEGLU(_Holder):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_Holder):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim))


class BasicTransformerBlock(_Holder):
    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head)
        self.attnc = CrossAttention(dim, heads=n_heads, dim_head=d_head)  # dead in the reference forward
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)


class SpatialTransformer(_Holder):
    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None):
        super().__init__()
        inner = n_heads * d_head
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner, in_channels, kernel_size=1))


class Word_Attention(_Holder):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.linear_query = nn.Linear(input_size, hidden_size)
        self.linear_key = nn.Linear(input_size, hidden_size)
        self.linear_value = nn.Linear(input_size, hidden_size)


def character_positional_encoding(max_seq_len, dim):
    """The reference's (non-standard) table, unet.py:876-882: even i -> sin(p / 1e4^(i/d)), odd i+1 -> cos(p / 1e4^((i+1)/d)),
    i.e. the *odd* index itself sits in the exponent.  Plain attribute in the reference (not a buffer, not in the state_dict)."""
    pe = torch.zeros(max_seq_len, dim)
    for pos in range(max_seq_len):
        for i in range(0, dim, 2):
            pe[pos, i] = math.sin(pos / (10000 ** (i / dim)))
            pe[pos, i + 1] = math.cos(pos / (10000 ** ((i + 1) / dim)))
    return pe


class CharacterEncoder(_Holder):
    def __init__(self, input_size, hidden_size, max_seq_len):
        super().__init__()
        self.embedding = nn.Embedding(input_size, hidden_size)
        self.attention = Word_Attention(hidden_size, hidden_size)
        self.embedding_dim = hidden_size
        self.max_seq_len = max_seq_len
        self.positional_encoding = character_positional_encoding(max_seq_len, hidden_size)


class CTCtopC(_Holder):
    """OCR head of args.ocrTraining == 1 (unet.py:1054-1092): parameter / buffer tree only (the arithmetic is wd_f32_ctc_head)."""

    def __init__(self, input_size, head_cfg, nclasses):
        super().__init__()
        hidden_size, num_layers = head_cfg

        def stage(cin):
            return nn.Sequential(nn.Conv2d(cin, hidden_size, kernel_size=(1, 5), stride=(1, 1), padding=(0, 2)),
                                 nn.BatchNorm2d(hidden_size), nn.ReLU(), nn.Dropout(.25))
        self.temporal_i = stage(input_size)
        self.temporal_m = nn.ModuleList([stage(hidden_size) for _ in range(num_layers)])
        self.temporal_o = nn.Conv2d(hidden_size, nclasses, kernel_size=(1, 5), stride=1, padding=(0, 2))
        self.lin1 = nn.Linear(32, 128)
        self.lin2 = nn.Linear(128, 256)


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("parameter holder of the B200 engine; call the enclosing UNetModel")


def add_blocks(model, *, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
               channel_mult, num_heads, num_head_channels, transformer_depth, context_dim, ted, attention_maps_layout=False):
    """Populates ``model`` with input_blocks / middle_block / output_blocks / out (same loops as unet.py:1248-1458).
    ``attention_maps_layout``: args.attentionMaps == 1 registers the middle block as ``middle_block1`` = ModuleList([Res + ST],
    [Res]) (unet.py:1336-1364), which renames 54 state_dict keys."""
    def heads_for(ch):
        if num_head_channels == -1:
            return num_heads, ch // num_heads
        return ch // num_head_channels, num_head_channels

    model.input_blocks = nn.ModuleList(
        [TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
    chans = [model_channels]
    ch, ds = model_channels, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            layers = [ResBlock(ch, ted, out_channels=mult * model_channels)]
            ch = mult * model_channels
            if ds in attention_resolutions:
                h, d = heads_for(ch)
                layers.append(SpatialTransformer(ch, h, d, depth=transformer_depth, context_dim=context_dim))
            model.input_blocks.append(TimestepEmbedSequential(*layers))
            chans.append(ch)
        if level != len(channel_mult) - 1:
            model.input_blocks.append(TimestepEmbedSequential(Downsample(ch, out_channels=ch)))
            chans.append(ch)
            ds *= 2
    h, d = heads_for(ch)
    if attention_maps_layout:
        model.middle_block1 = nn.ModuleList([
            TimestepEmbedSequential(ResBlock(ch, ted),
                                    SpatialTransformer(ch, h, d, depth=transformer_depth, context_dim=context_dim)),
            TimestepEmbedSequential(ResBlock(ch, ted))])
    else:
        model.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted), SpatialTransformer(ch, h, d, depth=transformer_depth, context_dim=context_dim),
            ResBlock(ch, ted))
    model.output_blocks = nn.ModuleList([])
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            layers = [ResBlock(ch + ich, ted, out_channels=model_channels * mult)]
            ch = model_channels * mult
            if ds in attention_resolutions:
                h, d = heads_for(ch)
                layers.append(SpatialTransformer(ch, h, d, depth=transformer_depth, context_dim=context_dim))
            if level and i == num_res_blocks:
                layers.append(Upsample(ch, out_channels=ch))
                ds //= 2
            model.output_blocks.append(TimestepEmbedSequential(*layers))
    model.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(),
                              zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))
