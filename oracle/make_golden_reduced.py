"""Golden fixture of the reference's reduced-call sampler (regenerateFromtrain2.py:520-618, fullSampling = 0), made by a
literal transcription of that loop driving the UNMODIFIED reference unet.UNetModel (imported under oracle/ref_shims.py):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_reduced.py     ->  tests/golden/unet_reduced_T12.npz

noise_steps = 12: the UNet is evaluated at i = 11, 10, 5 and the stale predicted noise drives the other eight steps."""
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


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B, T = 2, 12
    inp = W.make_inputs(B, seed=SEED)
    m = ref_shims.build_reference_model("unet")
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(W.make_state_dict(spec, SEED), strict=True)
    beta = torch.linspace(1e-4, 0.02, T)
    alpha = 1.0 - beta
    alpha_hat = torch.cumprod(alpha, dim=0)
    g = torch.Generator().manual_seed(SEED + 11)
    x = torch.randn((B, 4, 8, 32), generator=g)
    x_T = x.clone()
    fullSampling = 0
    epoch = 0
    called, eps_steps = [], []
    with torch.no_grad():
        for i in reversed(range(1, T)):
            t = (torch.ones(B) * i).long()
            if fullSampling or ((i % (100) == 0 or i % 5 == 0 or i == T or i == (T - 1) or (epoch > 3 and i % (25) == 0) or
                                 (epoch > 5 and i % (15) == 0) or (epoch > 10 and i % (10) == 0) or epoch > 50 == 0)):
                predicted_noise = m(x, None, timesteps=t, context=inp["context"], y=inp["y"])
                called.append(i)
                eps_steps.append(predicted_noise.numpy().copy())
            a = alpha[t][:, None, None, None]
            ah = alpha_hat[t][:, None, None, None]
            x = 1 / torch.sqrt(a) * (x - ((1 - a) / (torch.sqrt(1 - ah))) * predicted_noise)
    np.savez_compressed(os.path.join(OUT, "unet_reduced_T12.npz"), x_T=x_T.numpy(), x_final=x.numpy(),
                        called=np.array(called), eps_steps=np.stack(eps_steps))
    print("evaluated at", called, "final |x| max", float(x.abs().max()))


if __name__ == "__main__":
    main()
