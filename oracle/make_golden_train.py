"""Generates tests/golden/unet_train.npz by running one training step of the UNMODIFIED reference ``unet.UNetModel``
(train.py:281-294: nn.MSELoss, loss.backward(), torch.optim.AdamW(lr=1e-4).step(), EMA(0.995)) in the build container.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_train.py

Stored (small): the loss, per-parameter gradient signatures (norm + projection on a fixed random direction; -1 norm for
parameters whose grad is None), a few complete small gradients, and per-parameter signatures of the AdamW update."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import train_oracle as TO  # noqa: E402
import weights as W  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
SEED = 1234
FULL = ["out.2.weight", "out.2.bias", "time_embed.0.bias", "input_blocks.0.0.weight", "middle_block.1.norm.weight",
        "output_blocks.3.1.transformer_blocks.0.norm2.weight", "word_emb.attention.linear_query.bias"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    inp = W.make_inputs(B, seed=SEED)
    g = torch.Generator().manual_seed(SEED + 11)
    noise = torch.randn((B, 4, 8, 32), generator=g)
    m = ref_shims.build_reference_model("unet")
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = W.make_state_dict(spec, SEED)
    m.load_state_dict(sd, strict=True)
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)            # train.py:405
    pred = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    loss = torch.nn.MSELoss()(noise, pred)                        # train.py:287
    opt.zero_grad()
    loss.backward()
    names = [n for n, _ in m.named_parameters()]
    assert names == [k for k, _ in spec]
    gnorm, gproj = [], []
    full = {}
    for i, (n, p) in enumerate(m.named_parameters()):
        if p.grad is None:
            gnorm.append(-1.0)
            gproj.append(0.0)
        else:
            a, b = TO.signature(p.grad, 77 + i)
            gnorm.append(a)
            gproj.append(b)
            if n in FULL:
                full["grad::" + n] = p.grad.numpy().copy()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt.step()
    unorm, uproj = [], []
    for i, (n, p) in enumerate(m.named_parameters()):
        a, b = TO.signature(p.detach() - before[n], 977 + i)
        unorm.append(a)
        uproj.append(b)
    np.savez_compressed(os.path.join(OUT, "unet_train.npz"), loss=np.float64(loss.item()), noise=noise.numpy(),
                        grad_norm=np.array(gnorm), grad_proj=np.array(gproj), upd_norm=np.array(unorm), upd_proj=np.array(uproj),
                        **full)
    n_none = sum(1 for v in gnorm if v < 0)
    print("loss", loss.item(), "params", len(names), "without grad", n_none)


if __name__ == "__main__":
    main()
