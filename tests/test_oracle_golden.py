"""CPU: the oracle (oracle/unet_oracle.py, oracle/diffusion_oracle.py) is pinned against outputs of the UNMODIFIED
reference modules, generated in the build container by oracle/make_golden.py and committed under tests/golden/.
The reference ships no golden vectors of its own."""
import json
import math
import os

import numpy as np
import pytest
import torch

import unet_oracle as UO
import weights as W
from diffusion_oracle import DiffusionOracle

SEED = 1234


def _relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def unet_sd():
    return W.make_state_dict(W.load_spec("unet"), SEED)


@pytest.fixture(scope="module")
def phosc_sd():
    return W.make_state_dict(W.load_spec("unetPhosc"), SEED)


def test_state_dict_spec_sizes():
    assert len(W.load_spec("unet")) == 264
    assert len(W.load_spec("unetPhosc")) == 246


def test_inputs_are_deterministic(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_fwd.npz"))
    inp = W.make_inputs(2, seed=SEED)
    assert np.array_equal(g["x"], inp["x"].numpy())
    assert np.array_equal(g["t"], inp["t"].numpy())
    assert np.array_equal(g["context"], inp["context"].numpy())
    assert np.array_equal(g["y"], inp["y"].numpy())


def test_unet_forward_matches_reference(golden_dir, unet_sd):
    g = np.load(os.path.join(golden_dir, "unet_fwd.npz"))
    inp = W.make_inputs(2, seed=SEED)
    ctx = UO.encode_context(unet_sd, inp["context"], variant="unet")
    assert _relerr(ctx, g["ctx"]) < 1e-5
    eps = UO.unet_forward(unet_sd, inp["x"], inp["t"], inp["context"], inp["y"], variant="unet")
    assert float(np.abs(g["eps"]).max()) > 1e-3      # not the vacuous all-zero output of a fresh reference model
    assert _relerr(eps, g["eps"]) < 1e-4, "fp32 tolerance of BASELINE.json north_star"


def test_unet_phosc_forward_matches_reference(golden_dir, phosc_sd):
    g = np.load(os.path.join(golden_dir, "unetPhosc_fwd.npz"))
    inp = W.make_inputs(2, seed=SEED)
    assert np.array_equal(g["phosc"], inp["phosc"].numpy())
    eps = UO.unet_forward(phosc_sd, inp["x"], inp["t"], inp["context"], inp["y"], phosc=inp["phosc"],
                          variant="unetPhosc")
    assert _relerr(eps, g["eps"]) < 1e-4


def test_unet_phosc2_is_identical_to_unet_phosc(golden_dir):
    with open(os.path.join(golden_dir, "unetPhosc2_same.json")) as f:
        assert json.load(f)["max_abs_diff_vs_unetPhosc"] == 0.0


def test_ddpm_trajectory_matches_reference_loop(golden_dir, unet_sd):
    """Literal transcription of train.py:217-236 (T = 6) driving the reference UNet vs DiffusionOracle + UNet oracle."""
    g = np.load(os.path.join(golden_dir, "unet_ddpm_T6.npz"))
    inp = W.make_inputs(2, seed=SEED)
    d = DiffusionOracle(noise_steps=6)
    ctx = UO.encode_context(unet_sd, inp["context"], variant="unet")
    eps_trace = []

    def eps_fn(x, t):
        e = UO.unet_forward(unet_sd, x, t, inp["context"], inp["y"], variant="unet", ctx_encoded=ctx)
        eps_trace.append(e)
        return e

    x = d.ddpm_sample(eps_fn, torch.from_numpy(g["x_T"]), torch.from_numpy(g["noises"]))
    for k, e in enumerate(eps_trace):
        assert _relerr(e, g["eps_steps"][k]) < 1e-4, f"step {k}"
    assert _relerr(x, g["x_final"]) < 1e-4


def test_schedule_closed_form():
    d = DiffusionOracle(1000)
    assert d.beta.shape == (1000,)
    assert math.isclose(float(d.beta[0]), 1e-4, rel_tol=1e-6) and math.isclose(float(d.beta[-1]), 0.02, rel_tol=1e-6)
    ah = np.cumprod(1.0 - np.linspace(1e-4, 0.02, 1000, dtype=np.float64))
    assert np.allclose(d.alpha_hat.numpy(), ah, rtol=2e-4)


def test_ddim_eta0_properties():
    """DDIM eta=0 (not in the reference): with the true eps the step lands exactly on the x0-consistent point, and the
    last step (t_prev = -1, alpha_hat = 1) returns x0."""
    d = DiffusionOracle(1000)
    ts = d.ddim_timesteps(50)
    assert len(ts) == 50 and ts[0] == 980 and ts[-1] == 0 and all(a - b == 20 for a, b in zip(ts, ts[1:]))
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 4, 8, 32, generator=g)
    eps = torch.randn(2, 4, 8, 32, generator=g)
    t, tp = 500, 480
    xt = d.noise_images(x0, torch.tensor([t, t]), eps)
    xp = d.ddim_step(xt, eps, t, tp)
    want = d.noise_images(x0, torch.tensor([tp, tp]), eps)
    assert _relerr(xp, want) < 1e-5
    assert _relerr(d.ddim_step(xt, eps, t, -1), x0) < 1e-5


def test_timestep_embedding_and_pe():
    e = UO.timestep_embedding(torch.tensor([0, 7]), 320)
    assert e.shape == (2, 320) and torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    pe = UO.positional_encoding(10, 320)
    assert pe[0, 0] == 0 and pe[0, 1] == 1
    assert math.isclose(float(pe[3, 5]), math.cos(3 / 10000 ** (5 / 320)), rel_tol=1e-6)  # odd index in the exponent
