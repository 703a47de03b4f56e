"""ORACLE / test infrastructure -- never imported by the product package.

CPU restatement (plain torch fp32 functional ops) of the decode half of the autoencoder the reference uses at the end of sampling:

    latents = 1 / 0.18215 * x ; image = vae.decode(latents).sample ; image = (image / 2 + 0.5).clamp(0, 1)
    (/root/reference/train.py:239-247, regenerateFromtrain2.py:624-636; vae = AutoencoderKL.from_pretrained(..., subfolder="vae"),
     train.py:415)

PARITY UNPINNED.  The algorithm lives in a third-party dependency that is absent from /root/reference and from this image:
huggingface `diffusers` (requirements of the reference; `AutoencoderKL`, `models/autoencoders/vae.py::Decoder`,
`models/unets/unet_2d_blocks.py::UNetMidBlock2D / UpDecoderBlock2D`, `models/resnet.py::ResnetBlock2D`,
`models/attention_processor.py::Attention`, `models/upsampling.py::Upsample2D`; the reference pins no version) with the Stable
Diffusion v1 `vae/config.json` (block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4, norm_num_groups 32,
act_fn silu).  No golden vector of it exists in the reference; this file restates the published forward pass:

    z = post_quant_conv(z)                                      1x1 conv
    h = conv_in(z)                                              3x3, pad 1
    h = mid_block(h):   ResnetBlock2D, Attention, ResnetBlock2D
    h = up_blocks[i](h): (layers_per_block + 1) x ResnetBlock2D, then Upsample2D (nearest x2 + 3x3 conv) except in the last block
    h = conv_out(silu(conv_norm_out(h)))
    ResnetBlock2D(x) = shortcut(x) + conv2(silu(norm2(conv1(silu(norm1(x))))))        GroupNorm(32, eps 1e-6), output_scale_factor 1
    Attention(x)     = x + to_out(softmax(q k^T / sqrt(C)) v),  q/k/v = Linear(group_norm(x) as [HW, C]), one head of width C

The state_dict keys are diffusers' (`decoder.up_blocks.2.resnets.0.conv_shortcut.weight`, ...), which the product module
worddiffusion_b200.vae.AutoencoderKL mirrors.
"""
import math

import torch
import torch.nn.functional as F


def _gn(sd, pfx, x):
    return F.group_norm(x, 32, sd[pfx + ".weight"], sd[pfx + ".bias"], eps=1e-6)


def _conv(sd, pfx, x):
    w = sd[pfx + ".weight"]
    return F.conv2d(x, w, sd[pfx + ".bias"], padding=w.shape[-1] // 2)


def _resnet(sd, pfx, x):
    h = _conv(sd, pfx + "conv1", F.silu(_gn(sd, pfx + "norm1", x)))
    h = _conv(sd, pfx + "conv2", F.silu(_gn(sd, pfx + "norm2", h)))
    if pfx + "conv_shortcut.weight" in sd:
        x = _conv(sd, pfx + "conv_shortcut", x)
    return x + h


def _attention(sd, pfx, x):
    b, c, hh, ww = x.shape
    t = _gn(sd, pfx + "group_norm", x).view(b, c, hh * ww).transpose(1, 2)        # [b, HW, C]
    q = F.linear(t, sd[pfx + "to_q.weight"], sd[pfx + "to_q.bias"])
    k = F.linear(t, sd[pfx + "to_k.weight"], sd[pfx + "to_k.bias"])
    v = F.linear(t, sd[pfx + "to_v.weight"], sd[pfx + "to_v.bias"])
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
    o = F.linear(p @ v, sd[pfx + "to_out.0.weight"], sd[pfx + "to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


@torch.no_grad()
def vae_decode(sd, z):
    """sd: diffusers-keyed fp32 state_dict (decode half); z: [n, 4, h, w] -> [n, 3, 8h, 8w]  (vae.decode(z).sample)."""
    h = _conv(sd, "post_quant_conv", z) if "post_quant_conv.weight" in sd else z
    h = _conv(sd, "decoder.conv_in", h)
    h = _resnet(sd, "decoder.mid_block.resnets.0.", h)
    h = _attention(sd, "decoder.mid_block.attentions.0.", h)
    h = _resnet(sd, "decoder.mid_block.resnets.1.", h)
    i = 0
    while f"decoder.up_blocks.{i}.resnets.0.norm1.weight" in sd:
        j = 0
        while f"decoder.up_blocks.{i}.resnets.{j}.norm1.weight" in sd:
            h = _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}.", h)
            j += 1
        if f"decoder.up_blocks.{i}.upsamplers.0.conv.weight" in sd:
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
        i += 1
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h)))


def sampling_tail(sd, x):
    """train.py:239-243: latents -> images in [0, 1]."""
    image = vae_decode(sd, 1 / 0.18215 * x)
    return (image / 2 + 0.5).clamp(0, 1)
