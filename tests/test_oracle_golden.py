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


# ---------------------------------------------------------------------------------------------------------------
# training step (train.py:281-294): oracle autograd / AdamW / EMA vs the UNMODIFIED reference model + torch.optim.AdamW
# ---------------------------------------------------------------------------------------------------------------
def test_train_step_oracle_matches_reference(golden_dir, unet_sd):
    import train_oracle as TO
    g = np.load(os.path.join(golden_dir, "unet_train.npz"))
    inp = W.make_inputs(2, seed=SEED)
    noise = torch.from_numpy(g["noise"])
    loss, _, grads = TO.unet_loss_and_grads(unet_sd, inp["x"], inp["t"], inp["context"], inp["y"], noise)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    names = list(unet_sd.keys())
    n_none = 0
    for i, n in enumerate(names):
        if g["grad_norm"][i] < 0:
            assert grads[n] is None, f"{n}: the reference gives no gradient"
            n_none += 1
            continue
        a, b = TO.signature(grads[n], 77 + i)
        assert abs(a - g["grad_norm"][i]) <= 2e-4 * g["grad_norm"][i] + 1e-9, n
        assert abs(b - g["grad_proj"][i]) <= 2e-4 * g["grad_norm"][i] + 1e-9, n
    assert n_none == 58  # parameters the reference forward never reads
    for key in g.files:
        if key.startswith("grad::"):
            assert _relerr(grads[key[6:]], g[key]) < 2e-4, key
    # AdamW step 1 (lr 1e-4, torch defaults) on those gradients
    for i, n in enumerate(names):
        p = unet_sd[n].clone()
        if grads[n] is None:
            assert g["upd_norm"][i] == 0.0  # torch skips parameters without grad (no weight decay either)
            continue
        if g["grad_norm"][i] < 1e-6:
            continue  # a gradient that is pure rounding noise (e.g. the key bias under a softmax): sign(g) is arbitrary
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        TO.adamw_update(p, grads[n], m, v, 1)
        a, b = TO.signature(p - unet_sd[n], 977 + i)
        # the first Adam step is -lr * sign(g) (+ decay): elements whose gradient is ~0 may flip; compare norms loosely
        assert abs(a - g["upd_norm"][i]) <= 2e-2 * g["upd_norm"][i] + 1e-9, n


def test_ema_oracle_semantics():
    import train_oracle as TO
    p = torch.tensor([1.0, 2.0])
    ema = torch.zeros(2)
    TO.ema_update(ema, p, 0)            # warm-up: EMA.reset_parameters copies (train.py:162-165)
    assert torch.equal(ema, p)
    TO.ema_update(ema, p * 3, 2000)     # afterwards: old * beta + (1 - beta) * new (train.py:156-159)
    assert torch.allclose(ema, p * 0.995 + 0.005 * p * 3)


def test_reduced_call_sampler_matches_reference_loop(golden_dir, unet_sd):
    """regenerateFromtrain2.py:520-618 (fullSampling = 0): stale predicted noise between evaluations, noise-free update.
    Fixture: literal transcription of that loop around the UNMODIFIED reference UNet (oracle/make_golden_reduced.py), T = 12."""
    g = np.load(os.path.join(golden_dir, "unet_reduced_T12.npz"))
    inp = W.make_inputs(2, seed=1234)
    d = DiffusionOracle(noise_steps=12)
    assert [i for i in reversed(range(1, 12)) if d.reduced_call_predicate(i, 12)] == list(g["called"]) == [11, 10, 5]

    def eps_fn(x, t):
        return UO.unet_forward(unet_sd, x, t, inp["context"], inp["y"], variant="unet")
    x, called = d.reduced_call_sample(eps_fn, torch.from_numpy(g["x_T"]))
    assert called == [11, 10, 5]
    assert float((x - torch.from_numpy(g["x_final"])).abs().max()) < 1e-4


def test_phosc_oracle_matches_reference_generators(golden_dir):
    """PHOS / PHOC pyramids of the reference's own generator functions (oracle/make_golden_phosc.py), bit-exact; the bigram part
    the reference never sets stays zero."""
    import phosc_oracle as P
    g = np.load(os.path.join(golden_dir, "phosc_labels.npz"))
    assert g["labels"].shape[1] == 769 and int(g["labels"][:, -100:].sum()) == 0
    for w, lab in zip(g["words"], g["labels"]):
        assert (P.phosc(str(w)) == lab).all(), str(w)
    assert P.segments(7)[1:3] == [(0, 3), (3, 7)] and len(P.segments(3)) == 15
    with pytest.raises(KeyError):
        P.phosc("abc1")


def test_oracle_attention_maps_vs_reference(golden_dir):
    """args.attentionMaps == 1: the oracle's 5-tuple against the reference module's (oracle/make_golden_attnmaps.py).  The fixture
    keeps one sample per nearest-upsampled block; the maps must be constant over each block."""
    import unet_oracle as UO
    import weights as W
    g = np.load(os.path.join(golden_dir, "unet_attnmaps.npz"))
    g0 = np.load(os.path.join(golden_dir, "unet_fwd.npz"))
    sd = W.make_state_dict(W.load_spec("unet"), 1234)
    ren = {"middle_block.0.": "middle_block1.0.0.", "middle_block.1.": "middle_block1.0.1.", "middle_block.2.": "middle_block1.1.0."}
    sd1 = {}
    for k, v in sd.items():   # the layout a checkpoint of the attentionMaps = 1 model has
        for old, new in ren.items():
            if k.startswith(old):
                k = new + k[len(old):]
                break
        sd1[k] = v
    inp = W.make_inputs(2, seed=1234)
    eps, a1, a2, a3, ctx = UO.unet_forward(sd1, inp["x"], inp["t"], inp["context"], inp["y"], attention_maps=True)
    assert float((eps - torch.from_numpy(g0["eps"])).abs().max()) < 1e-5
    assert float((ctx - torch.from_numpy(g["context"])).abs().max()) < 1e-5
    for a, key, s in zip((a1, a2, a3), ("attn1", "attn2", "attn3"), g["scales"]):
        s = int(s)
        assert a.shape == (2, 64, 256, 10)
        sub = a[:, ::s, ::s]
        assert torch.equal(sub.repeat_interleave(s, 1).repeat_interleave(s, 2), a)
        assert float((sub - torch.from_numpy(g[key])).abs().max()) < 1e-5
        assert float((a.sum(-1) - 4.0).abs().max()) < 1e-4    # four heads, each row of probabilities sums to one
