"""Generates tests/golden/* by running the UNMODIFIED reference modules (imported from /root/reference under the
shims of oracle/ref_shims.py) in the build container.  The reference cannot travel to the GPU box, the fixtures do.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Fixtures:
  state_dict_spec_{unet,unetPhosc}.json   key order + shapes of the reference state_dict
  unet_fwd.npz / unetPhosc_fwd.npz        eps = model(x, t, context, y[, phosc]) for the synthetic weights/inputs of
                                          oracle/weights.py (seed 1234, B = 2), plus the encoded context
  unetPhosc2_same.json                    max |unetPhosc2 - unetPhosc| on the same inputs (must be 0)
  unet_ddpm_T6.npz                        literal transcription of Diffusion.sampling (train.py:217-236) with
                                          noise_steps = 6 driving the reference unet.UNetModel, fixed noise
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


def dump_spec(model, name):
    spec = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    with open(os.path.join(OUT, f"state_dict_spec_{name}.json"), "w") as f:
        json.dump(spec, f)
    return [(k, tuple(s)) for k, s in spec]


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    inp = W.make_inputs(B, seed=SEED)

    # ---------------- unet.UNetModel ----------------
    m = ref_shims.build_reference_model("unet")
    spec = dump_spec(m, "unet")
    sd = W.make_state_dict(spec, SEED)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        eps = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
        ctx = m.word_emb(inp["context"])
    assert float(eps.abs().max()) > 1e-3, "vacuous parity: reference output is ~0"
    np.savez_compressed(os.path.join(OUT, "unet_fwd.npz"), eps=eps.numpy(), ctx=ctx.numpy(),
                        x=inp["x"].numpy(), t=inp["t"].numpy(), context=inp["context"].numpy(), y=inp["y"].numpy())
    print("unet eps std", float(eps.std()), "absmax", float(eps.abs().max()))

    # literal transcription of train.py:217-236 (one UNet call per step, pre-generated noise), T = 6
    T = 6
    beta = torch.linspace(1e-4, 0.02, T)
    alpha = 1.0 - beta
    alpha_hat = torch.cumprod(alpha, dim=0)
    g = torch.Generator().manual_seed(SEED + 7)
    x = torch.randn((B, 4, 8, 32), generator=g)
    x_T = x.clone()
    noises = torch.randn((T, B, 4, 8, 32), generator=g)
    eps_steps = []
    with torch.no_grad():
        for i in reversed(range(1, T)):
            t = (torch.ones(B) * i).long()
            predicted_noise = m(x, None, timesteps=t, context=inp["context"], y=inp["y"])
            eps_steps.append(predicted_noise.numpy().copy())
            a = alpha[t][:, None, None, None]
            ah = alpha_hat[t][:, None, None, None]
            b = beta[t][:, None, None, None]
            noise = noises[i] if i > 1 else torch.zeros_like(x)
            x = 1 / torch.sqrt(a) * (x - ((1 - a) / (torch.sqrt(1 - ah))) * predicted_noise) + torch.sqrt(b) * noise
    np.savez_compressed(os.path.join(OUT, "unet_ddpm_T6.npz"), x_T=x_T.numpy(), noises=noises.numpy(),
                        x_final=x.numpy(), eps_steps=np.stack(eps_steps))
    del m

    # ---------------- unetPhosc.UNetModelPhosc (ctx = 10 chars + 769 PHOSC tokens) ----------------
    mp = ref_shims.build_reference_model("unetPhosc")
    spec_p = dump_spec(mp, "unetPhosc")
    sd_p = W.make_state_dict(spec_p, SEED)
    mp.load_state_dict(sd_p, strict=True)
    with torch.no_grad():
        eps_p = mp(inp["x"], inp["phosc"], timesteps=inp["t"], context=inp["context"], y=inp["y"])
    assert float(eps_p.abs().max()) > 1e-3
    np.savez_compressed(os.path.join(OUT, "unetPhosc_fwd.npz"), eps=eps_p.numpy(), phosc=inp["phosc"].numpy())
    print("unetPhosc eps std", float(eps_p.std()))
    del mp

    mp2 = ref_shims.build_reference_model("unetPhosc2")
    assert [(k, tuple(v.shape)) for k, v in mp2.state_dict().items()] == spec_p
    mp2.load_state_dict(sd_p, strict=True)
    with torch.no_grad():
        eps_p2 = mp2(inp["x"], inp["phosc"], timesteps=inp["t"], context=inp["context"], y=inp["y"])
    d = float((eps_p2 - eps_p).abs().max())
    with open(os.path.join(OUT, "unetPhosc2_same.json"), "w") as f:
        json.dump({"max_abs_diff_vs_unetPhosc": d}, f)
    print("unetPhosc2 vs unetPhosc max abs diff", d)


if __name__ == "__main__":
    main()
