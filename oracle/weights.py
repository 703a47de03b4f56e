"""ORACLE / test infrastructure: deterministic synthetic weights shared by the golden generator, the oracle and
the CUDA parity tests.  No checkpoint ships with the reference, and a fresh reference model outputs exactly 0
(zero_module on every out_layers.3 / proj_out / out.2, unet.py:152-158,620-622,375-379,1457), so every tensor --
including the zero-initialised ones -- is drawn with a torch-default-like scale from a per-key seeded CPU generator.
The values depend only on (seed, key order, shape): any process on any box reproduces them bit for bit."""
import json
import math
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_spec(variant):
    """[(key, shape)] of the reference state_dict (dumped from the reference by oracle/make_golden.py)."""
    name = "unet" if variant == "unet" else "unetPhosc"
    with open(os.path.join(GOLDEN_DIR, f"state_dict_spec_{name}.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


def make_state_dict(spec, seed=1234):
    sd = {}
    for idx, (key, shape) in enumerate(spec):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = (len(shape) == 1 and leaf == "weight")
        if "embedding.weight" in key or key == "label_emb.weight":
            t = torch.randn(shape, generator=g)                      # nn.Embedding default N(0,1)
        elif is_norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)          # norm gains around 1
        elif leaf == "weight":
            fan_in = int(torch.tensor(shape[1:]).prod().item())
            bound = 1.0 / math.sqrt(fan_in)                           # kaiming_uniform(a=sqrt(5)) bound
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:                                                         # biases
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[key] = t.float()
    return sd


def make_inputs(B, seed=1234, L=10, phosc_len=769, vocab=53, num_classes=339, latent=(4, 8, 32), T=1000):
    """Synthetic conditioning as specified in SURVEY.md section 8d."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B,) + tuple(latent), generator=g)
    t = torch.randint(1, T, (B,), generator=g)
    ctx = torch.randint(1, vocab - 1, (B, L), generator=g)
    # realistic tail: words shorter than 10 characters are padded with the PAD id 52 (train.py:74)
    lens = torch.randint(2, L + 1, (B,), generator=g)
    for b in range(B):
        ctx[b, lens[b]:] = vocab - 1
    y = torch.randint(0, num_classes, (B,), generator=g)
    phos = torch.randint(0, 4, (B, 165), generator=g)
    phoc = (torch.rand((B, phosc_len - 165), generator=g) < 0.05).long()
    phosc = torch.cat([phos, phoc], dim=1)
    return dict(x=x, t=t, context=ctx, y=y, phosc=phosc)


def variant_state_dict(spec, rename=None, seed=1234):
    """Fixture weights of a flag variant of unet.UNetModel (oracle/make_golden_variants.py): the keys shared with the plain model
    carry the seed-1234 unet fixture (optionally renamed, e.g. to the attentionMaps = 1 layout), the new ones (auxhead.*,
    conv_layer*.*) come from the same generator with seed 4321; BatchNorm running_var is made positive, num_batches_tracked = 3."""
    base = make_state_dict(load_spec("unet"), seed)
    if rename:
        base = {rename(k): v for k, v in base.items()}
    spec = [(k, tuple(s)) for k, s in spec]
    extra = make_state_dict([(k, s) for k, s in spec if k not in base], seed=4321)
    for k, v in extra.items():
        if k.endswith("running_var"):
            extra[k] = v.abs() + 0.5
        elif k.endswith("num_batches_tracked"):
            extra[k] = torch.tensor(3, dtype=torch.int64)
    sd = dict(base)
    sd.update(extra)
    assert sorted(sd) == sorted(k for k, _ in spec)
    return sd


def rename_to_attnmaps(key):
    """middle_block.{0,1,2}.* (attentionMaps = 0) -> middle_block1.{0.0,0.1,1.0}.* (attentionMaps = 1, unet.py:1336-1364)."""
    for old, new in (("middle_block.0.", "middle_block1.0.0."), ("middle_block.1.", "middle_block1.0.1."),
                     ("middle_block.2.", "middle_block1.1.0.")):
        if key.startswith(old):
            return new + key[len(old):]
    return key
