"""GPU parity of the VAE decode (reference train.py:239-247) through the C ABI (wd_vae_decode) against the CPU
restatement oracle/vae_oracle.py of diffusers' AutoencoderKL decoder (parity unpinned: diffusers is not in the image -- see the
oracle's header).  fp32 storage and arithmetic on both sides: tolerance 1e-4 of max |ref| (north_star's fp32-mode figure)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from gpu_util import DEV, relerr  # noqa: E402
from worddiffusion_b200.vae import AutoencoderKL  # noqa: E402

TOL = 1e-4


def _weights(model, seed):
    import weights as W  # oracle/weights.py: deterministic per-key synthetic tensors
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    return W.make_state_dict(spec, seed=seed)


@pytest.mark.parametrize("n,h,w,small", [(2, 8, 32, False), (3, 8, 16, False), (5, 4, 8, True), (1, 8, 32, True)])
def test_vae_decode_matches_the_oracle(n, h, w, small):
    import vae_oracle
    kw = dict(block_out_channels=(64, 128, 128), layers_per_block=1) if small else {}
    m = AutoencoderKL(**kw)
    sd = _weights(m, 77 + n)
    m.load_state_dict(sd, strict=True)
    m.to(DEV)
    z = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(n))
    ref = vae_oracle.vae_decode(sd, z)
    got = m.decode(z.to(DEV), chunk=2).sample
    assert got.shape == ref.shape == (n, 3, h << (len(m.config.block_out_channels) - 1), w << (len(m.config.block_out_channels) - 1))
    e = relerr(got, ref)
    print(f"vae decode n={n} {h}x{w} small={small}: {e:.2e}")
    assert e < TOL
    # the reference's tail: 1 / 0.18215 scaling in front, (image / 2 + 0.5).clamp(0, 1) behind, folded into the same call
    ref2 = vae_oracle.sampling_tail(sd, z)
    got2 = m.decode(z.to(DEV), scale=1 / 0.18215, postprocess=True).sample
    assert float((got2.cpu() - ref2).abs().max()) < 2e-4
    assert float(got2.min()) >= 0.0 and float(got2.max()) <= 1.0


def test_vae_accepts_a_full_autoencoder_checkpoint_with_old_attention_names():
    m = AutoencoderKL(block_out_channels=(64, 128), layers_per_block=1)
    sd = _weights(m, 5)
    old = {}
    ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    for k, v in sd.items():
        for a, b in ren.items():
            if f".attentions.0.{a}." in k:
                k = k.replace(f".{a}.", f".{b}.")
        old[k] = v
    old["encoder.conv_in.weight"] = torch.zeros(64, 3, 3, 3)
    old["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)
    m2 = AutoencoderKL(block_out_channels=(64, 128), layers_per_block=1)
    m2.load_state_dict(old, strict=True)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k])


def test_sampling_returns_images_through_the_vae():
    """Diffusion.sampling(model, vae, ...) end to end (train.py:200-251): [n, 3, 64, 256] in [0, 1]."""
    import weights as W
    from worddiffusion_b200 import unet as wd_unet
    from worddiffusion_b200.diffusion import Diffusion
    from test_gpu_model import KW
    model = wd_unet.UNetModel(args=wd_unet.default_args(DEV), **KW)
    model.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
    model.to(DEV).eval()
    vae = AutoencoderKL(block_out_channels=(64, 128, 128, 128), layers_per_block=1)
    vae.load_state_dict(_weights(vae, 9))
    vae.to(DEV)
    diff = Diffusion(noise_steps=4, img_size=(64, 256), device=DEV)
    y = torch.tensor([1, 2], device=DEV)
    img = diff.sampling(model, vae, 2, "word", y, seed=3)
    assert img.shape == (2, 3, 64, 256) and img.dtype == torch.float32
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 and torch.isfinite(img).all()
