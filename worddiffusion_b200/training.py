"""Training step of the B200 path -- the host side of reference ``train.train`` (train.py:255-294):

    t = diffusion.sample_timesteps(n); x_t, noise = diffusion.noise_images(latents, t)
    predicted_noise = model(x_t, ..., timesteps=t, context=text_features, y=s_id)
    loss = mse_loss(noise, predicted_noise); optimizer.zero_grad(); loss.backward(); optimizer.step()
    ema.step_ema(ema_model, model)

Two ways in:
  * drop-in: ``UNetModel.forward`` in training mode with autograd enabled routes through ``_UNetTrainFn``, so the reference
    loop above runs unchanged (``loss.backward()`` fills ``param.grad`` of every parameter the reference's own backward
    reaches; the 58 parameters its forward never reads keep ``grad is None``), with any ``torch.optim`` optimizer;
  * fused: ``FusedTrainStep`` keeps parameters / gradients / AdamW moments / EMA in flat fp32 buffers, runs forward +
    backward through the C ABI, all-reduces the flat gradient over NCCL (data parallel, train.py ``--ddp``) and applies
    AdamW + EMA in one kernel (``wd_adamw_ema_step``).
PyTorch only owns memory, streams and the process group here; every kernel is in libwd_b200.so.  No CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import WdConfig, check, lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class TrainEngine:
    """Handle of one ``wd_trainer`` bound to the parameters of a ``worddiffusion_b200.unet.UNetModel``."""

    def __init__(self, module, device, latent_hw=(8, 32)):
        self.device = torch.device(device)
        self.latent_hw = (int(latent_hw[0]), int(latent_hw[1]))
        if self.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 trains on a CUDA (sm_100a) device only; there is no CPU path")
        self.module = module
        cfg = WdConfig()
        cfg.variant = module.VARIANT
        cfg.in_channels, cfg.model_channels, cfg.out_channels = module.in_channels, module.model_channels, module.out_channels
        cfg.num_res_blocks = module.num_res_blocks
        cfg.n_channel_mult = len(module.channel_mult)
        for i, m in enumerate(module.channel_mult):
            cfg.channel_mult[i] = int(m)
        ar = sorted(set(int(a) for a in module.attention_resolutions))
        cfg.n_attention_resolutions = len(ar)
        for i, a in enumerate(ar):
            cfg.attention_resolutions[i] = a
        cfg.num_heads, cfg.num_head_channels = module.num_heads, module.num_head_channels
        cfg.transformer_depth = module.transformer_depth
        cfg.context_dim, cfg.vocab_size = module.context_dim, module.vocab_size
        cfg.num_classes = module.num_classes or 0
        cfg.max_seq_len = module.max_seq_len
        cfg.latent_h, cfg.latent_w = self.latent_hw
        cfg.add_label_emb = 1 if module._add_label_emb() else 0
        cfg.phosc_len = module._phosc_len()
        self.cfg = cfg
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().wd_trainer_create(C.byref(cfg), C.byref(self._h)), "wd_trainer_create")
        self._bound_sig = None
        self._weights_sig = None
        self.live = []        # [(name, param)] parameters that receive a gradient
        self.flat_grad = None
        self.grad_views = {}
        self._hold = None

    def __deepcopy__(self, memo):
        return None  # copy.deepcopy(model) (train.py:410, the EMA model) gets its own engine lazily

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().wd_trainer_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def bind(self, flat_grad=None):
        """(Re)bind the module's parameters and a flat fp32 gradient buffer (one slice per live parameter)."""
        named = [(n, p) for n, p in self.module.named_parameters()]
        sig = tuple((n, p.data_ptr()) for n, p in named)
        if sig == self._bound_sig and (flat_grad is None or flat_grad is self.flat_grad):
            return
        l = lib()
        for n, p in named:
            if p.device != self.device or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.WdError(f"parameter {n} must be a contiguous fp32 tensor on {self.device}")
        # first pass: which parameters are live (the engine ignores the ones the reference forward never reads)
        offsets, total = {}, 0
        probe = torch.zeros(1, device=self.device)
        live = []
        for n, p in named:
            shape = (C.c_int64 * max(p.dim(), 1))(*p.shape)
            rc = check(l.wd_trainer_bind_param(self._h, n.encode(), _ptr(p.data), _ptr(probe), shape, p.dim()), f"bind({n})")
            if rc == _lib.WD_OK:
                offsets[n] = total
                total += (p.numel() + 63) // 64 * 64  # 256-byte aligned slices
                live.append((n, p))
        # lay the slices out in the order the backward pass finishes them (wd_trainer_grad_stage): a bucket of the gradient
        # all-reduce is then one contiguous range that can leave while the later stages still run (plan_grad_buckets)
        nst = C.c_int(0)
        check(l.wd_trainer_num_grad_stages(self._h, C.byref(nst)), "wd_trainer_num_grad_stages")
        self.n_stages = nst.value
        stage_of = {}
        for n, _ in live:
            st = C.c_int(0)
            check(l.wd_trainer_grad_stage(self._h, n.encode(), C.byref(st)), f"grad_stage({n})")
            stage_of[n] = st.value
        live.sort(key=lambda np_: stage_of[np_[0]])  # stable: state_dict order inside a stage
        offsets, total = {}, 0
        for n, p in live:
            offsets[n] = total
            total += (p.numel() + 63) // 64 * 64  # 256-byte aligned slices
        self.stage_of = stage_of
        if flat_grad is None:
            flat_grad = torch.zeros(total, device=self.device, dtype=torch.float32)
        elif flat_grad.numel() != total:
            raise _lib.WdError(f"flat gradient buffer must hold {total} elements")
        self.flat_grad = flat_grad
        self.grad_views = {}
        for n, p in live:
            gv = flat_grad[offsets[n]:offsets[n] + p.numel()].view(p.shape)
            self.grad_views[n] = gv
            shape = (C.c_int64 * max(p.dim(), 1))(*p.shape)
            check(l.wd_trainer_bind_param(self._h, n.encode(), _ptr(p.data), _ptr(gv), shape, p.dim()), f"bind({n})")
        self.live = live
        self.offsets = offsets
        self.total = total
        pe = self.module.word_emb.positional_encoding.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            check(l.wd_trainer_set_pos_encoding(self._h, _ptr(pe), _stream_ptr()), "set_pos_encoding")
            torch.cuda.current_stream().synchronize()
        self._bound_sig = sig
        self._weights_sig = None

    def sync_weights(self, force=False):
        """Refresh the bf16 tensor-core packs from the fp32 parameters (after an optimizer step)."""
        sig = tuple((p.data_ptr(), p._version) for _, p in self.live)
        if force or sig != self._weights_sig:
            with torch.cuda.device(self.device):
                check(lib().wd_trainer_sync_weights(self._h, _stream_ptr()), "wd_trainer_sync_weights")
            self._weights_sig = sig

    # ------------------------------------------------------------------ step
    def forward(self, x, timesteps, y, context):
        # the C side copies batch * C * H * W floats from / to these pointers: every extent is checked here
        want = (self.module.in_channels,) + self.latent_hw
        if x.dim() != 4 or tuple(x.shape[1:]) != want:
            raise _lib.WdError(f"trainer built for latents [B, {want[0]}, {want[1]}, {want[2]}], got x {tuple(x.shape)} "
                               "(module.train_engine(device, latent_hw) builds one for another size)")
        B = x.shape[0]
        if B < 1:
            raise _lib.WdError("empty training batch")
        if timesteps.dim() != 1 or timesteps.shape[0] != B:
            raise _lib.WdError(f"timesteps must be [{B}], got {tuple(timesteps.shape)}")
        if context is None or context.dim() != 2 or context.shape[0] != B:
            raise _lib.WdError(f"context must be [{B}, L] token ids, got {None if context is None else tuple(context.shape)}")
        if self.cfg.add_label_emb:
            if y is None or y.dim() != 1 or y.shape[0] != B:
                raise _lib.WdError(f"y must be [{B}], got {None if y is None else tuple(y.shape)}")
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        t = timesteps.to(device=self.device, dtype=torch.int64).contiguous()
        ctx = context.to(device=self.device, dtype=torch.int64).contiguous()
        yy = y.to(device=self.device, dtype=torch.int64).contiguous() if y is not None else None
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            check(lib().wd_trainer_forward(self._h, B, _ptr(x), _ptr(t), _ptr(yy), _ptr(ctx), ctx.shape[1], _ptr(out),
                                           _stream_ptr()), "wd_trainer_forward")
        self._hold = (x, t, ctx, yy)
        self._hold_shape = tuple(x.shape)
        return out

    def grad_buckets(self, n_buckets):
        """[(stage_end, lo, hi)]: once the backward stages [.., stage_end) have run, flat_grad[lo:hi] is final."""
        stages = [self.stage_of[n] for n, _ in self.live]
        sizes = [(p.numel() + 63) // 64 * 64 for _, p in self.live]
        return plan_grad_buckets(stages, sizes, self.n_stages, n_buckets)

    def backward_stages(self, d_eps, stage_begin, stage_end, zero=True):
        """Stages [stage_begin, stage_end) of the backward pass of the last forward (wd_trainer_backward_stages); the call
        that starts at stage 0 zeroes the flat gradient buffer (``zero``) and reads ``d_eps``."""
        if self._hold is None:
            raise _lib.WdError("backward called without a forward")
        d = None
        if stage_begin == 0:
            if tuple(d_eps.shape) != self._hold_shape:
                raise _lib.WdError(f"d_eps must have the shape of the forward's x {self._hold_shape}, got {tuple(d_eps.shape)}")
            d = d_eps.to(device=self.device, dtype=torch.float32).contiguous()
            if zero:
                self.flat_grad.zero_()
        with torch.cuda.device(self.device):
            check(lib().wd_trainer_backward_stages(self._h, _ptr(d), int(stage_begin), int(stage_end), _stream_ptr()),
                  "wd_trainer_backward_stages")
        if stage_end == self.n_stages:
            self._hold = None

    def backward(self, d_eps, zero=True):
        """Accumulates the parameter gradients of the last forward into the flat gradient buffer."""
        if self._hold is None:
            raise _lib.WdError("backward called without a forward")
        _, _, ctx, yy = self._hold
        if tuple(d_eps.shape) != self._hold_shape:
            raise _lib.WdError(f"d_eps must have the shape of the forward's x {self._hold_shape}, got {tuple(d_eps.shape)}")
        d = d_eps.to(device=self.device, dtype=torch.float32).contiguous()
        if zero:
            self.flat_grad.zero_()
        with torch.cuda.device(self.device):
            check(lib().wd_trainer_backward(self._h, _ptr(d), _ptr(yy), _ptr(ctx), _stream_ptr()), "wd_trainer_backward")
        self._hold = None

    @property
    def launch_counts(self):
        f, b = C.c_int(0), C.c_int(0)
        check(lib().wd_trainer_launch_counts(self._h, C.byref(f), C.byref(b)), "wd_trainer_launch_counts")
        return f.value, b.value

    @property
    def workspace_bytes(self):
        return lib().wd_trainer_workspace_bytes(self._h)


class _UNetTrainFn(torch.autograd.Function):
    """eps = UNetModel(x, timesteps, context, y) with a hand-written backward (no autograd graph inside)."""

    @staticmethod
    def forward(ctx, module, x, timesteps, context, y, *params):
        eng = module.train_engine(x.device, x.shape[2:])
        eng.bind()
        eng.sync_weights()
        ctx.eng = eng
        ctx.names = [n for n, _ in module.named_parameters()]
        return eng.forward(x, timesteps, y, context)

    @staticmethod
    def backward(ctx, d_eps):
        eng = ctx.eng
        eng.backward(d_eps)
        # fresh tensors: autograd may keep (or add into) what it is handed, the flat buffer is reused next step
        grads = [eng.grad_views[n].clone() if n in eng.grad_views else None for n in ctx.names]
        return (None, None, None, None, None, *grads)


def unet_train_forward(module, x, timesteps, context, y):
    params = [p for _, p in module.named_parameters()]
    return _UNetTrainFn.apply(module, x, timesteps, context, y, *params)


def plan_grad_buckets(stages, sizes, n_stages, n_buckets):
    """Cuts the flat gradient buffer into at most ``n_buckets`` contiguous ranges of roughly equal size at backward-stage
    boundaries.  ``stages[i]`` / ``sizes[i]``: final stage and (padded) element count of slice i, in buffer order, stages
    non-decreasing.  Returns ``[(stage_end, lo, hi)]`` with stage_end strictly increasing and the last one == n_stages:
    after stages [.., stage_end) the elements [lo, hi) are final (torch DDP's gradient buckets, built from the engine's own
    backward order instead of autograd hooks)."""
    if any(b < a for a, b in zip(stages, stages[1:])):
        raise ValueError("slices must be ordered by the stage that finishes them")
    total = sum(sizes)
    n_buckets = max(1, int(n_buckets))
    buckets, lo, pos, k = [], 0, 0, 1
    for i, (st, sz) in enumerate(zip(stages, sizes)):
        pos += sz
        last_of_stage = i + 1 == len(stages) or stages[i + 1] != st
        if last_of_stage and k < n_buckets and pos >= total * k / n_buckets and i + 1 < len(stages):
            buckets.append((st + 1, lo, pos))
            lo = pos
            while pos >= total * k / n_buckets:
                k += 1
    buckets.append((n_stages, lo, total))
    return buckets


def allreduce_sum_(flat, group=None):
    """Data-parallel gradient exchange (train.py --ddp): in-place SUM all-reduce of the flat gradient buffer.
    Returns the world size (the optimizer kernel divides by it: DDP's gradient averaging)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    ws = dist.get_world_size(group)
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return ws


def shard_batch(n, rank, world):
    """Contiguous slice [lo, hi) of a global batch of n samples owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FusedTrainStep:
    """Noise-prediction step with flat parameter storage and a fused AdamW + EMA update (train.py:281-294,140-170,405).

    ``step(x_t, t, context, y, noise)`` returns the MSE loss (device scalar).  With an initialised process group the flat
    gradient is sum-all-reduced over NCCL and averaged (DDP semantics) before the update."""

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, ema_beta=0.995,
                 step_start_ema=2000, process_group=None, use_ema=True, latent_hw=(8, 32), grad_buckets=None):
        p0 = next(module.parameters())
        self.module = module
        self.device = p0.device
        self.eng = module.train_engine(self.device, latent_hw)
        self.eng.bind()
        eng = self.eng
        # move the live parameters into one flat fp32 buffer (same slices as the gradient buffer) and re-point them
        self.flat_param = torch.empty(eng.total, device=self.device, dtype=torch.float32)
        self.flat_param.zero_()
        for n, p in eng.live:
            sl = self.flat_param[eng.offsets[n]:eng.offsets[n] + p.numel()].view(p.shape)
            sl.copy_(p.data)
            p.data = sl
        eng.bind(eng.flat_grad)  # pointers changed
        self.m = torch.zeros_like(self.flat_param)
        self.v = torch.zeros_like(self.flat_param)
        self.ema = self.flat_param.clone() if use_ema else None
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.ema_beta, self.step_start_ema = ema_beta, step_start_ema
        self.t = 0
        self.pg = process_group
        # gradient exchange: `grad_buckets` contiguous ranges of the flat buffer, each all-reduced (async, NCCL's own stream)
        # as soon as the backward stage that finishes it has been launched, under the remaining stages
        if grad_buckets is None:
            import os
            grad_buckets = int(os.environ.get("WD_GRAD_BUCKETS", "4"))  # 1: one all-reduce after the whole backward pass
        self.buckets = eng.grad_buckets(grad_buckets)
        eng.sync_weights(force=True)

    def world_size(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.pg)
        return 1

    def step(self, x_t, timesteps, context, y, noise):
        eng = self.eng
        eps = eng.forward(x_t, timesteps, y, context)
        # nn.MSELoss (train.py:287) and its gradient in one kernel: loss = mean((eps - noise)^2), d_eps = 2 (eps - noise) / n
        noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        n_el = eps.numel()
        if getattr(self, "_mse_ws_n", None) != n_el:
            self._mse_ws = torch.zeros(int(lib().wd_mse_workspace_bytes(n_el)), device=self.device, dtype=torch.uint8)
            self._mse_ws_n = n_el
        d_eps = torch.empty_like(eps)
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(lib().wd_mse_loss_grad(_ptr(eps), _ptr(noise), _ptr(d_eps), _ptr(loss), _ptr(self._mse_ws), n_el, _stream_ptr()),
                  "wd_mse_loss_grad")
        ws = self.world_size()
        if ws > 1 and len(self.buckets) > 1:
            import torch.distributed as dist
            works, s0 = [], 0
            for s1, lo, hi in self.buckets:
                eng.backward_stages(d_eps, s0, s1)
                # the collective is ordered after the launches above (NCCL's stream waits for the current stream here) and
                # runs beside the stages launched next; the optimizer kernel waits for all of them
                works.append(dist.all_reduce(eng.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
                s0 = s1
            for w in works:
                w.wait()
        else:
            eng.backward(d_eps)
            ws = allreduce_sum_(eng.flat_grad, self.pg)
        self.t += 1
        # EMA.step_ema (train.py:161-167): copy during the warm-up, moving average afterwards
        ema_mode = 0 if self.ema is None else (1 if self.t <= self.step_start_ema else 2)
        with torch.cuda.device(self.device):
            check(lib().wd_adamw_ema_step(_ptr(self.flat_param), _ptr(eng.flat_grad), _ptr(self.m), _ptr(self.v),
                                          _ptr(self.ema), self.flat_param.numel(), self.lr, self.betas[0], self.betas[1],
                                          self.eps, self.weight_decay, self.t, self.ema_beta, ema_mode, 1.0 / ws,
                                          _stream_ptr()), "wd_adamw_ema_step")
        eng.sync_weights(force=True)
        # the kernel wrote the parameters behind torch's back (no _version bump): both inference engines (bf16 and fp32 mode)
        # re-pack on their next call
        self.module.invalidate_engine()
        return loss

    def ema_state_dict(self):
        """state_dict of the EMA model (live parameters from the EMA buffer, the rest copied from the module)."""
        sd = {k: v.detach().clone() for k, v in self.module.state_dict().items()}
        if self.ema is not None:
            for n, p in self.eng.live:
                sd[n] = self.ema[self.eng.offsets[n]:self.eng.offsets[n] + p.numel()].view(p.shape).clone()
        return sd
