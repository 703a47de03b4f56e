"""GPU parity of the flag variants of unet.UNetModel against outputs of the UNMODIFIED reference built with the
same flags (tests/golden/unet_variants.npz, made by oracle/make_golden_variants.py): the OCR head (args.ocrTraining), the
character-image convolutions (args.charImages), the character-level embedding (args.charLevelEmb), the style-vector context
(args.wrdChrWrStyl) and the style interpolation (args.interpolation + mix_rate).  Tolerances: north_star's 1e-2 (bf16) / 1e-4 (fp32
mode) of max |ref|."""
import json
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import weights as W  # noqa: E402
from gpu_util import DEV, relerr  # noqa: E402
from test_gpu_model import KW, SEED  # noqa: E402
from worddiffusion_b200.unet import UNetModel, default_args  # noqa: E402


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "unet_variants.npz")), np.load(os.path.join(golden_dir, "unet_fwd.npz"))


def _model(rename=None, **flags):
    m = UNetModel(args=default_args(DEV, **flags), **KW)
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(W.variant_state_dict(spec, rename), strict=True)
    return m.to(DEV).eval()


def _inputs():
    inp = W.make_inputs(2, seed=SEED)
    return {k: v.to(DEV) for k, v in inp.items()}


def test_ocr_head_matches_the_reference(gold):
    v, g = gold
    m = _model(W.rename_to_attnmaps, attentionMaps=1, ocrTraining=1)
    i = _inputs()
    with torch.no_grad():
        out = m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"])
    assert len(out) == 5
    assert relerr(out[0], torch.from_numpy(g["eps"])) < 1e-4
    tdec = out[4]
    assert tuple(tdec.shape) == v["ocr_tdec"].shape == (256, 2, KW["vocab_size"] - 2)
    e = relerr(tdec, torch.from_numpy(v["ocr_tdec"]))
    print(f"ocr head tdec: {e:.2e}")
    assert e < 1e-4
    m.train()
    with pytest.raises(NotImplementedError), torch.no_grad():
        m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"])


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_char_images_and_char_level_embedding_leave_eps_unchanged(gold, precision, tol):
    _, g = gold
    i = _inputs()
    ref = torch.from_numpy(g["eps"])
    m = _model(charImages=1)
    m.precision = precision
    imgs = torch.randn(2, 10, 4, 8, 32, device=DEV)
    with torch.no_grad():
        out = m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"], charContextImages=imgs)
        with pytest.raises(AttributeError):  # the reference dereferences charContextImages unconditionally (unet.py:1519)
            m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"])
        with pytest.raises(RuntimeError):    # ... and reshapes it to [max_seq_len * BS, 4, 8, 32]
            m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"], charContextImages=imgs[:, :3])
    assert relerr(out, ref) < tol
    m2 = _model(charLevelEmb=1)
    m2.precision = precision
    with torch.no_grad():
        out2 = m2(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"])
        with pytest.raises(RuntimeError):
            m2(i["x"], None, timesteps=i["t"], context=i["context"][:, :7], y=i["y"])
    assert relerr(out2, ref) < tol


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_style_vector_context_matches_the_reference(gold, precision, tol):
    v, _ = gold
    i = _inputs()
    m = _model(wrdChrWrStyl=1)
    m.precision = precision
    style = torch.from_numpy(v["style_in"]).float().to(DEV)
    with torch.no_grad():
        out = m(i["x"], style, timesteps=i["t"], context=i["context"], y=i["y"])
    e = relerr(out, torch.from_numpy(v["style_eps"]))
    print(f"style-vector context ({precision}): {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_style_interpolation_matches_the_reference(gold, precision, tol):
    v, g = gold
    i = _inputs()
    m = _model(interpolation=True)
    m.precision = precision
    random.seed(7)
    with torch.no_grad():
        out = m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"], mix_rate=0.3)
        plain = m(i["x"], None, timesteps=i["t"], context=i["context"], y=i["y"])
    e = relerr(out, torch.from_numpy(v["mix_eps"]))
    print(f"style interpolation ({precision}): {e:.2e}, writers {v['mix_writers']}")
    assert e < tol
    assert relerr(plain, torch.from_numpy(g["eps"])) < tol
    assert relerr(out, plain) > 10 * tol  # the mix really changed the conditioning


def test_sampling_with_mix_rate_runs_two_evaluations_per_step():
    """Diffusion.sampling(..., mix_rate=) (train.py:221-236 with args.interpolation): two model calls per step, each drawing its
    own writer pair, lerped with cfg_scale -- compared with the same loop spelled out through forward() and torch.lerp."""
    from worddiffusion_b200.diffusion import Diffusion, label_padding
    m = _model(interpolation=True)
    m.precision = "fp32"
    diff = Diffusion(noise_steps=5, img_size=(64, 256), device=DEV)
    y = torch.tensor([4, 9], device=DEV)
    random.seed(11)
    lat = diff.sampling(m, None, 2, "abc", y, mix_rate=0.25, cfg_scale=3, seed=5)
    # the same trajectory by hand
    from worddiffusion_b200.diffusion import philox_normal_latents
    random.seed(11)
    x = philox_normal_latents(2, (4, 8, 32), 5, 0, DEV)
    toks = torch.tensor([label_padding("abc")] * 2, dtype=torch.int64, device=DEV)
    eng = m.engine(DEV, latent_hw=(8, 32))
    with torch.no_grad():
        for i in reversed(range(1, 5)):
            t = torch.full((2,), i, device=DEV, dtype=torch.int64)
            e1 = m(x, None, timesteps=t, context=toks, y=y, mix_rate=0.25)
            e2 = m(x, None, timesteps=t, context=toks, y=y, mix_rate=0.25)
            eps = torch.lerp(e2, e1, 3.0)
            eng.sampler_update(x, eps, 1, diff._ddpm_coef[i], philox_seed=(5 if i > 1 else None), sample_offset=0, step_index=i)
    assert relerr(lat, x / 0.18215) < 1e-5
