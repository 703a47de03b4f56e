"""Golden fixtures of the flag variants of unet.UNetModel: runs the UNMODIFIED reference module with
each flag in the build container.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_variants.py

Writes tests/golden/unet_variants.npz and tests/golden/state_dict_spec_unet_variants.json:
  * ocr      args.ocrTraining = 1, attentionMaps = 1 (unet.py:1468,1827-1831): key order + shapes with `auxhead.*`, the 5-tuple's
             last element tdec = auxhead(eps) [256, B, 51] in eval mode (BatchNorm running statistics drawn away from (0, 1))
  * charimg  args.charImages = 1 (unet.py:1217-1223,1517-1539,1625-1627): key order + shapes with conv_layer1-3; eps must equal
             tests/golden/unet_fwd.npz bit for bit (the forward discards the convolutions' output)
  * charlvl  args.charLevelEmb = 1 (unet.py:853-863): eps must equal unet_fwd.npz bit for bit
  * style    args.wrdChrWrStyl = 1 (unet.py:1590-1591,1617-1618): eps with context = wrd_proj(style vectors [B, 10, 4096])
  * mix      args.interpolation = True, mix_rate = 0.3 (unet.py:1558-1572) after random.seed(7): eps and the two drawn writers
Weights: the unet fixture of oracle/weights.py (seed 1234) for the shared keys, the same generator (seed 4321) for the new ones.
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights as W  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
SEED = 1234


def variant_state_dict(model, rename=None):
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    return W.variant_state_dict(spec, rename), spec


def main():
    from make_golden_attnmaps import rename_to_attnmaps
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    inp = W.make_inputs(B, seed=SEED)
    g = np.load(os.path.join(OUT, "unet_fwd.npz"))
    out, specs = {}, {}

    # ---- OCR head ----
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(attentionMaps=1, ocrTraining=1))
    sd, spec = variant_state_dict(m, rename_to_attnmaps)
    specs["ocr"] = [[k, list(s)] for k, s in spec]
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        r = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert len(r) == 5 and np.array_equal(r[0].numpy(), g["eps"])
    out["ocr_tdec"] = r[4].numpy()
    print("tdec", tuple(r[4].shape), float(r[4].abs().max()))

    # ---- character images: dead compute ----
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(charImages=1))
    sd, spec = variant_state_dict(m)
    specs["charimg"] = [[k, list(s)] for k, s in spec]
    m.load_state_dict(sd, strict=True)
    m.eval()
    imgs = torch.randn(B, 10, 4, 8, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        r = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"], charContextImages=imgs)
    assert np.array_equal(r.numpy(), g["eps"]), "charImages = 1 changed eps"

    # ---- character-level embedding: the same lookup ----
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(charLevelEmb=1))
    sd, spec = variant_state_dict(m)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        r = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert np.array_equal(r.numpy(), g["eps"]), "charLevelEmb = 1 changed eps"

    # ---- style vectors as the context ----
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(wrdChrWrStyl=1))
    sd, spec = variant_state_dict(m)
    m.load_state_dict(sd, strict=True)
    m.eval()
    style = torch.randn(B, 10, 4096, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        r = m(inp["x"], style, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    out["style_in"] = style.numpy().astype(np.float16)  # fp16-representable inputs keep the fixture small
    with torch.no_grad():
        r = m(inp["x"], torch.from_numpy(out["style_in"]).float(), timesteps=inp["t"], context=inp["context"], y=inp["y"])
    out["style_eps"] = r.numpy()

    # ---- style interpolation ----
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(interpolation=True))
    sd, spec = variant_state_dict(m)
    m.load_state_dict(sd, strict=True)
    m.eval()
    random.seed(7)
    with torch.no_grad():
        r = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"], mix_rate=0.3)
    random.seed(7)
    s1 = random.randint(0, 338)
    s2 = random.randint(0, 338)
    while s1 == s2:
        s2 = random.randint(0, 338)
    out["mix_eps"] = r.numpy()
    out["mix_writers"] = np.array([s1, s2])
    with torch.no_grad():
        r0 = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert np.array_equal(r0.numpy(), g["eps"]), "interpolation without mix_rate changed eps"

    np.savez_compressed(os.path.join(OUT, "unet_variants.npz"), **out)
    with open(os.path.join(OUT, "state_dict_spec_unet_variants.json"), "w") as f:
        json.dump(specs, f)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
