"""VAE decode of the sampling loop's tail.

Reference: ``train.py:239-247`` / ``regenerateFromtrain2.py:624-636``::

    latents = 1 / 0.18215 * x
    image = vae.decode(latents).sample
    image = (image / 2 + 0.5).clamp(0, 1)

with ``vae = AutoencoderKL.from_pretrained(args.stable_dif_path, subfolder="vae")`` (``train.py:415``), diffusers' Stable Diffusion
v1 autoencoder.  :class:`AutoencoderKL` below is the drop-in for the *decode* half of that object: the same constructor defaults,
the same ``post_quant_conv.*`` / ``decoder.*`` parameter tree (so ``load_state_dict`` takes a diffusers checkpoint; encoder /
quant_conv entries are ignored with ``strict=False``), ``decode(z).sample``.  The arithmetic runs in ``csrc/f32_path.cu``
(``wd_vae_decode``: fp32 storage and arithmetic); there is no torch / CPU fallback.
"""
import ctypes as C
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib
from .engine import _ptr, _stream_ptr


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise _lib.WdError("parameter holder: the arithmetic runs in libwd_b200 (wd_vae_decode)")


def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, k, padding=k // 2)


class _Resnet(_Holder):  # diffusers ResnetBlock2D (temb_channels=None)
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = _conv(cin, cout, 3)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = _conv(cout, cout, 3)
        if cin != cout:
            self.conv_shortcut = _conv(cin, cout, 1)


class _Attention(_Holder):  # diffusers Attention(heads = 1, dim_head = channels, bias=True, residual_connection=True)
    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])


class _Upsample(_Holder):
    def __init__(self, ch):
        super().__init__()
        self.conv = _conv(ch, ch, 3)


class _UpBlock(_Holder):
    def __init__(self, cin, cout, n, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(n)])
        if add_upsample:
            self.upsamplers = nn.ModuleList([_Upsample(cout)])


class _MidBlock(_Holder):
    def __init__(self, ch, groups):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(ch, groups)])
        self.resnets = nn.ModuleList([_Resnet(ch, ch, groups), _Resnet(ch, ch, groups)])


class _Decoder(_Holder):
    def __init__(self, latent_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = _conv(latent_channels, rev[0], 3)
        self.mid_block = _MidBlock(rev[0], groups)
        ups, prev = [], rev[0]
        for i, ch in enumerate(rev):
            ups.append(_UpBlock(prev, ch, layers_per_block + 1, groups, add_upsample=i != len(rev) - 1))
            prev = ch
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(groups, rev[-1], eps=1e-6)
        self.conv_out = _conv(rev[-1], out_channels, 3)


class AutoencoderKL(nn.Module):
    """Decode half of diffusers' ``AutoencoderKL`` (defaults = Stable Diffusion v1's ``vae/config.json``)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        if latent_channels != 4 or norm_num_groups != 32:
            raise NotImplementedError("wd_vae_decode takes 4 latent channels and GroupNorm(32)")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                                      layers_per_block=layers_per_block, latent_channels=latent_channels,
                                      norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self.decoder = _Decoder(latent_channels, out_channels, tuple(block_out_channels), layers_per_block, norm_num_groups)
        self.post_quant_conv = _conv(latent_channels, latent_channels, 1)
        self._h = None
        self._sig = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Checkpoints of the full autoencoder also carry ``encoder.*`` / ``quant_conv.*`` (the encode half, outside this path)
        and older diffusers releases name the attention projections query / key / value / proj_attn: both are accepted."""
        ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        sd = {}
        for k, v in state_dict.items():
            if k.startswith("encoder.") or k.startswith("quant_conv."):
                continue
            parts = k.split(".")
            if "attentions" in parts and parts[-2] in ren:
                k = ".".join(parts[:-2] + [ren[parts[-2]], parts[-1]])
                if v.dim() == 4:  # very old checkpoints store the projections as 1x1 convolutions
                    v = v[:, :, 0, 0]
            sd[k] = v
        return super().load_state_dict(sd, strict=strict, **kw)

    def __del__(self):
        try:
            if self._h is not None and self._h.value:
                lib().wd_f32_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _engine(self, device):
        params = [(k, v) for k, v in self.state_dict().items()]
        sig = tuple((k, v.data_ptr(), v._version) for k, v in params) + (str(device),)
        if self._h is not None and sig == self._sig:
            return self._h
        l = lib()
        if self._h is not None:
            l.wd_f32_destroy(self._h)
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(l.wd_vae_create(C.byref(h)), "wd_vae_create")
            sp = _stream_ptr()
            keep = []
            for name, t in params:
                src = t.detach().to(device=device, dtype=torch.float32).contiguous()
                keep.append(src)
                shape = (C.c_int64 * max(src.dim(), 1))(*src.shape)
                check(l.wd_f32_load_param(h, name.encode(), _ptr(src), shape, src.dim(), sp), f"vae load_param({name})")
            torch.cuda.current_stream().synchronize()
        self._h, self._sig = h, sig
        return h

    @torch.no_grad()
    def decode(self, z, return_dict=True, *, scale=1.0, postprocess=False, chunk=32):
        """``vae.decode(latents).sample`` -> fp32 ``[n, 3, 8h, 8w]``.  ``scale`` / ``postprocess`` fold the reference's
        ``1 / 0.18215 *`` and ``(image / 2 + 0.5).clamp(0, 1)`` into the same call (``Diffusion.sampling`` uses them)."""
        if z.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if z.dim() != 4 or z.shape[1] != 4:
            raise _lib.WdError(f"latents must be [n, 4, h, w], got {tuple(z.shape)}")
        zz = z.detach().to(torch.float32).contiguous()
        n, _, h, w = zz.shape
        ups = len(self.config.block_out_channels) - 1
        out = torch.empty((n, self.config.out_channels, h << ups, w << ups), device=z.device, dtype=torch.float32)
        hnd = self._engine(z.device)
        if n:
            with torch.cuda.device(z.device):
                check(lib().wd_vae_decode(hnd, n, _ptr(zz), h, w, float(scale), int(bool(postprocess)), _ptr(out), int(chunk),
                                          _stream_ptr()), "wd_vae_decode")
        return SimpleNamespace(sample=out) if return_dict else (out,)

    def encode(self, *a, **k):
        raise NotImplementedError("the encode half of the autoencoder (train.py:277) is outside the sampling hot path")

    def forward(self, z):
        return self.decode(z).sample
