"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference training step
(train.py:281-294: MSE noise-prediction loss, ``loss.backward()``, ``optim.AdamW(lr=1e-4)``, ``EMA(0.995)``).

Gradients come from torch autograd over the forward restatement ``unet_oracle.unet_forward`` -- the same ATen ops the
reference modules call (all hot-path arithmetic executes inside PyTorch).  ``train.py`` itself cannot be
imported (module-level ``diffusers`` import), so ``EMA`` (train.py:140-170) and the AdamW update rule (torch.optim.AdamW,
the reference's pinned dependency) are restated here.  Pinned by tests/test_oracle_golden.py against gradients of the
UNMODIFIED reference ``unet.UNetModel`` under ``nn.MSELoss`` + ``torch.optim.AdamW`` (oracle/make_golden_train.py ->
tests/golden/unet_train.npz)."""
import math

import torch
import torch.nn.functional as F

import unet_oracle as UO


def unet_loss_and_grads(sd, x_t, t, context, y, noise, variant="unet"):
    """-> (loss, eps, {key: grad or None}).  None = the reference forward never reads that parameter."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    eps = UO.unet_forward(p, x_t, t, context, y, variant=variant)
    loss = F.mse_loss(noise, eps)  # nn.MSELoss()(noise, predicted_noise), train.py:287
    loss.backward()
    return loss.detach(), eps.detach(), {k: (v.grad.detach() if v.grad is not None else None) for k, v in p.items()}


def adamw_update(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """One torch.optim.AdamW step (decoupled weight decay; train.py:405 uses the defaults except lr).  In place."""
    b1, b2 = betas
    p.mul_(1 - lr * weight_decay)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def ema_update(ema, p, ema_step, beta=0.995, step_start_ema=2000):
    """EMA.step_ema (train.py:161-167): `ema_step` = EMA.step BEFORE the call.  Copy during the warm-up, average afterwards."""
    if ema_step < step_start_ema:
        ema.copy_(p)
    else:
        ema.mul_(beta).add_(p, alpha=1 - beta)


def signature(t, seed):
    """Two numbers that pin a tensor without storing it: its L2 norm and its projection on a fixed random direction."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    tf = t.detach().double().reshape(-1)
    return float(tf.norm()), float((tf * r).sum() / math.sqrt(t.numel()))
