"""Golden fixture of the attention-map variant (args.attentionMaps = 1, unet.py:1336-1364,1645-1836): runs the UNMODIFIED reference
unet.UNetModel with that flag in the build container.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_attnmaps.py

Writes
  tests/golden/state_dict_spec_unet_attnmaps.json   key order + shapes (middle_block1.{0,1}.* instead of middle_block.*)
  tests/golden/unet_attnmaps.npz                    the 5-tuple (eps, attn1, attn2, attn3, context) of the reference for the
                                                    synthetic weights / inputs of oracle/weights.py (seed 1234, B = 2).  The maps
                                                    are nearest-upsampled by 8 / 16 / 8 in the reference (unet.py:1786-1797):
                                                    the fixture keeps one sample per constant block (``[:, ::s, ::s]``) and the
                                                    generator asserts that nothing is lost by that.
The weights are the attentionMaps = 0 fixture under the renamed keys, so eps must equal tests/golden/unet_fwd.npz bit for bit.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights as W  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
SEED = 1234
SCALES = (8, 16, 8)


def rename_to_attnmaps(key):
    """middle_block.{0,1,2}.* (attentionMaps = 0) -> middle_block1.{0.0,0.1,1.0}.* (attentionMaps = 1)."""
    for old, new in (("middle_block.0.", "middle_block1.0.0."), ("middle_block.1.", "middle_block1.0.1."),
                     ("middle_block.2.", "middle_block1.1.0.")):
        if key.startswith(old):
            return new + key[len(old):]
    return key


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    inp = W.make_inputs(B, seed=SEED)
    m = ref_shims.build_reference_model("unet", args=ref_shims.default_args(attentionMaps=1))
    spec = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    with open(os.path.join(OUT, "state_dict_spec_unet_attnmaps.json"), "w") as f:
        json.dump(spec, f)
    base = W.make_state_dict(W.load_spec("unet"), SEED)
    sd = {rename_to_attnmaps(k): v for k, v in base.items()}
    assert sorted(sd) == sorted(k for k, _ in spec), "renaming does not reproduce the attentionMaps = 1 key set"
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert isinstance(out, tuple) and len(out) == 5
    eps, a1, a2, a3, ctx = out
    g = np.load(os.path.join(OUT, "unet_fwd.npz"))
    assert np.array_equal(eps.numpy(), g["eps"]), "eps of the attentionMaps = 1 model differs from the attentionMaps = 0 golden"
    small = []
    for a, s in zip((a1, a2, a3), SCALES):
        assert a.shape == (B, 64, 256, 10), a.shape
        sub = a[:, ::s, ::s]
        up = sub.repeat_interleave(s, dim=1).repeat_interleave(s, dim=2)
        assert torch.equal(up, a)
        small.append(sub.numpy())
    np.savez_compressed(os.path.join(OUT, "unet_attnmaps.npz"), attn1=small[0], attn2=small[1], attn3=small[2],
                        context=ctx.numpy(), scales=np.array(SCALES))
    print("maps", [s.shape for s in small], "sum over chars (= heads):", float(a1[0, 0, 0].sum()), "ctx", tuple(ctx.shape))


if __name__ == "__main__":
    main()
