"""Python handle of the C++ engine (``csrc/engine.cu``): owns one ``wd_engine``, keeps its packed weights in sync
with the ``nn.Module`` parameters and forwards the hot-path calls through ctypes.  PyTorch is only used for device
memory and streams here."""
import ctypes as C

import torch

from . import _lib
from ._lib import WdConfig, check, lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def make_config(*, variant, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
                num_heads, num_head_channels, transformer_depth, context_dim, vocab_size, num_classes, max_seq_len, latent_hw,
                add_label_emb, phosc_len):
    """``wd_config`` (include/wd_b200.h) from the reference constructor arguments (unet.py:1126-1156)."""
    cfg = WdConfig()
    cfg.variant = variant
    cfg.in_channels, cfg.model_channels, cfg.out_channels = in_channels, model_channels, out_channels
    cfg.num_res_blocks = num_res_blocks
    cfg.n_channel_mult = len(channel_mult)
    for i, m in enumerate(channel_mult):
        cfg.channel_mult[i] = int(m)
    ar = sorted(set(int(a) for a in attention_resolutions))
    cfg.n_attention_resolutions = len(ar)
    for i, a in enumerate(ar):
        cfg.attention_resolutions[i] = a
    cfg.num_heads, cfg.num_head_channels = num_heads, num_head_channels
    cfg.transformer_depth = transformer_depth
    cfg.context_dim, cfg.vocab_size = context_dim, vocab_size
    cfg.num_classes = num_classes or 0
    cfg.max_seq_len = max_seq_len
    cfg.latent_h, cfg.latent_w = latent_hw
    cfg.add_label_emb = 1 if add_label_emb else 0
    cfg.phosc_len = phosc_len
    return cfg


class HotPathEngine:
    def __init__(self, *, variant, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 channel_mult, num_heads, num_head_channels, transformer_depth, context_dim, vocab_size, num_classes,
                 max_seq_len, latent_hw, add_label_emb, phosc_len, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        cfg = make_config(variant=variant, in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                          num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions,
                          channel_mult=channel_mult, num_heads=num_heads, num_head_channels=num_head_channels,
                          transformer_depth=transformer_depth, context_dim=context_dim, vocab_size=vocab_size,
                          num_classes=num_classes, max_seq_len=max_seq_len, latent_hw=latent_hw,
                          add_label_emb=add_label_emb, phosc_len=phosc_len)
        self.cfg = cfg
        self.latent_hw = tuple(latent_hw)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().wd_engine_create(C.byref(cfg), C.byref(self._h)), "wd_engine_create")
        self._weights_sig = None
        self._ctx_key = None

    def __deepcopy__(self, memo):
        return None  # copy.deepcopy(model) (train.py:410, the EMA model) gets its own engine lazily

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().wd_engine_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    @staticmethod
    def weights_signature(named_params):
        return tuple((p.data_ptr(), p._version) for _, p in named_params)

    def load_state(self, named_tensors, pos_encoding):
        """named_tensors: iterable of (state_dict key, tensor).  Repacks every tensor the forward reads."""
        l = lib()
        with torch.cuda.device(self.device):
            sp = _stream_ptr()
            keep = []
            for name, t in named_tensors:
                src = t.detach()
                if src.device != self.device or src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(src)
                shape = (C.c_int64 * max(src.dim(), 1))(*src.shape)
                check(l.wd_engine_load_param(self._h, name.encode(), _ptr(src), shape, src.dim(), sp),
                      f"load_param({name})")
            pe = pos_encoding.to(device=self.device, dtype=torch.float32).contiguous()
            check(l.wd_engine_set_pos_encoding(self._h, _ptr(pe), sp), "set_pos_encoding")
            missing = l.wd_engine_finalize_params(self._h, sp)
            if missing != 0:
                msg = l.wd_last_error()
                raise _lib.WdError(f"state_dict incomplete for the engine: {msg.decode() if msg else missing}")
            torch.cuda.current_stream().synchronize()  # staging copies in `keep` may be freed after this
        self._ctx_key = None

    # ------------------------------------------------------------------ hot path
    def encode_context(self, context, phosc=None):
        """context: int64 [B, L] token ids; phosc: [B, 769] (any numeric dtype) or None."""
        B, L = context.shape
        # The conditioning is time-invariant: a sampling loop passes the SAME tensor objects at every step (train.py:221-236),
        # and re-encoding them is pure waste.  The key holds the tensors themselves (their storage cannot be recycled while
        # cached) and their in-place version counters; a new tensor object, even with equal contents, is encoded again.
        key = (context, context._version, phosc, None if phosc is None else phosc._version, B)
        old = self._ctx_key
        if (isinstance(old, tuple) and old[0] is context and old[1] == key[1] and old[2] is phosc and old[3] == key[3]
                and old[4] == B and not getattr(self, "_ctx_dirty", False)):
            return
        ctx = context.to(device=self.device, dtype=torch.int64).contiguous()
        ph = None
        if self.cfg.phosc_len > 0:
            if phosc is None:
                raise _lib.WdError("this model needs phoscLabels")
            ph = phosc.to(self.device).int().contiguous()  # reference: phoscLabels.int() (unetPhosc.py:1124)
            if ph.shape != (B, self.cfg.phosc_len):
                raise _lib.WdError(f"phoscLabels must be [{B}, {self.cfg.phosc_len}], got {tuple(ph.shape)}")
        with torch.cuda.device(self.device):
            check(lib().wd_encode_context(self._h, B, _ptr(ctx), L, _ptr(ph), _stream_ptr()), "wd_encode_context")
        self._ctx_hold = (ctx, ph)
        self._ctx_key = key
        self._ctx_dirty = False

    def set_context(self, ctx_dense):
        """Dense fp32 context [B, L, context_dim] in place of tokens (args.wrdChrWrStyl == 1, unet.py:1617-1618)."""
        ctx = ctx_dense.to(device=self.device, dtype=torch.float32).contiguous()
        B, L, D = ctx.shape
        if D != self.cfg.context_dim:
            raise _lib.WdError(f"dense context must be [B, L, {self.cfg.context_dim}], got {tuple(ctx.shape)}")
        with torch.cuda.device(self.device):
            check(lib().wd_set_context(self._h, B, _ptr(ctx), L, _stream_ptr()), "wd_set_context")
        self._ctx_hold = (ctx, None)
        self._ctx_key = None

    def set_label_mix(self, row, s1, s2, mix):
        with torch.cuda.device(self.device):
            check(lib().wd_engine_set_label_mix(self._h, int(row), int(s1), int(s2), float(mix), _stream_ptr()),
                  "wd_engine_set_label_mix")

    def unet_eval(self, x, timesteps, y, out=None):
        B = x.shape[0]
        if out is None:
            out = torch.empty_like(x)
        t_ptr, t_scalar = C.c_void_p(0), 0
        if isinstance(timesteps, int):
            t_scalar = timesteps
        else:
            timesteps = timesteps.to(device=self.device, dtype=torch.int64).contiguous()
            t_ptr = _ptr(timesteps)
        with torch.cuda.device(self.device):
            check(lib().wd_unet_eval(self._h, B, _ptr(x), t_ptr, t_scalar, _ptr(y), _ptr(out), _stream_ptr()),
                  "wd_unet_eval")
        return out

    def sampler_step(self, x, t, y, mode, coef, noise=None, philox_seed=None, sample_offset=0, step_index=0,
                     eps_out=None):
        B = x.shape[0]
        c4 = (C.c_float * 4)(*[float(v) for v in coef])
        use_philox = 1 if (noise is None and philox_seed is not None) else 0
        with torch.cuda.device(self.device):
            check(lib().wd_sampler_step(self._h, B, _ptr(x), int(t), _ptr(y), mode, c4, _ptr(noise), use_philox,
                                        int(philox_seed or 0), int(sample_offset), int(step_index), _ptr(eps_out),
                                        _stream_ptr()), "wd_sampler_step")
        return x

    def sampler_update(self, x, eps, mode, coef, noise=None, philox_seed=None, sample_offset=0, step_index=0):
        """x <- update(x, eps) with a predicted noise the caller kept (reduced-call sampling); no UNet evaluation."""
        B = x.shape[0]
        c4 = (C.c_float * 4)(*[float(v) for v in coef])
        use_philox = 1 if (noise is None and philox_seed is not None) else 0
        with torch.cuda.device(self.device):
            check(lib().wd_sampler_update(_ptr(x), _ptr(eps), B, x[0].numel(), mode, c4, _ptr(noise), use_philox,
                                          int(philox_seed or 0), int(sample_offset), int(step_index), _stream_ptr()),
                  "wd_sampler_update")
        return x

    OP_KINDS = ("timestep_embed", "gemm_tc", "groupnorm", "layernorm", "attn_small", "attn_flash", "conv_in",
                "groupnorm_stats", "upsample", "tblock", "out_head")

    def set_profiling(self, on):
        check(lib().wd_engine_set_profiling(self._h, 1 if on else 0), "wd_engine_set_profiling")

    def profile_read(self):
        """-> (n_steps, [(kind_name, flops, bytes, ms_sum)] per launch of the step plan)."""
        cap = 1024
        kinds, fl, by = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        ms, ns = (C.c_float * cap)(), C.c_int(0)
        n = check(lib().wd_engine_profile_read(self._h, cap, kinds, fl, by, ms, C.byref(ns)), "wd_engine_profile_read")
        return ns.value, [(self.OP_KINDS[kinds[i]], fl[i], by[i], ms[i]) for i in range(n)]

    def reserve(self, batch):
        self._ctx_key = None
        with torch.cuda.device(self.device):
            check(lib().wd_engine_reserve(self._h, batch), "wd_engine_reserve")

    @property
    def last_launch_count(self):
        return lib().wd_engine_last_launch_count(self._h)

    @property
    def workspace_bytes(self):
        return lib().wd_engine_workspace_bytes(self._h)

    @property
    def weight_bytes(self):
        return lib().wd_engine_weight_bytes(self._h)


class F32Engine:
    """fp32 mode of the hot path (``csrc/f32_path.cu``): the same UNet with fp32 storage and FFMA arithmetic, for the
    north_star's 1e-4 tolerance.  Same interface as :class:`HotPathEngine` (``Diffusion`` drives either); the sampler update
    is the separate ``wd_sampler_update`` kernel instead of the output conv's epilogue."""

    def __init__(self, cfg, latent_hw, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        self.cfg = cfg
        self.latent_hw = tuple(latent_hw)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().wd_f32_create(C.byref(cfg), C.byref(self._h)), "wd_f32_create")
        self._ctx_key = None

    def __deepcopy__(self, memo):
        return None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().wd_f32_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state(self, named_tensors, pos_encoding):
        l = lib()
        with torch.cuda.device(self.device):
            sp = _stream_ptr()
            keep = []
            for name, t in named_tensors:
                src = t.detach()
                if src.device != self.device or src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(src)
                shape = (C.c_int64 * max(src.dim(), 1))(*src.shape)
                check(l.wd_f32_load_param(self._h, name.encode(), _ptr(src), shape, src.dim(), sp), f"f32 load_param({name})")
            pe = pos_encoding.to(device=self.device, dtype=torch.float32).contiguous()
            check(l.wd_f32_set_pos_encoding(self._h, _ptr(pe), sp), "wd_f32_set_pos_encoding")
            torch.cuda.current_stream().synchronize()
        self._ctx_key = None

    def encode_context(self, context, phosc=None):
        B, L = context.shape
        key = (context, context._version, phosc, None if phosc is None else phosc._version, B)
        old = self._ctx_key
        if (isinstance(old, tuple) and old[0] is context and old[1] == key[1] and old[2] is phosc and old[3] == key[3]
                and old[4] == B):
            return
        ctx = context.to(device=self.device, dtype=torch.int64).contiguous()
        ph = None
        if self.cfg.phosc_len > 0:
            if phosc is None:
                raise _lib.WdError("this model needs phoscLabels")
            ph = phosc.to(self.device).int().contiguous()
            if ph.shape != (B, self.cfg.phosc_len):
                raise _lib.WdError(f"phoscLabels must be [{B}, {self.cfg.phosc_len}], got {tuple(ph.shape)}")
        with torch.cuda.device(self.device):
            check(lib().wd_f32_encode_context(self._h, B, _ptr(ctx), L, _ptr(ph), _stream_ptr()), "wd_f32_encode_context")
        self._ctx_key = key

    def set_context(self, ctx_dense):
        ctx = ctx_dense.to(device=self.device, dtype=torch.float32).contiguous()
        B, L, D = ctx.shape
        if D != self.cfg.context_dim:
            raise _lib.WdError(f"dense context must be [B, L, {self.cfg.context_dim}], got {tuple(ctx.shape)}")
        with torch.cuda.device(self.device):
            check(lib().wd_f32_set_context(self._h, B, _ptr(ctx), L, _stream_ptr()), "wd_f32_set_context")
        self._ctx_key = None

    def set_label_mix(self, row, s1, s2, mix):
        with torch.cuda.device(self.device):
            check(lib().wd_f32_set_label_mix(self._h, int(row), int(s1), int(s2), float(mix), _stream_ptr()),
                  "wd_f32_set_label_mix")

    def ctc_head(self, eps):
        """tdec = auxhead(eps) (CTCtopC, unet.py:1054-1092) in eval mode -> fp32 [256, B, nclasses]."""
        B, Cc, H, W = eps.shape
        ncls = self.cfg.vocab_size - 2
        out = torch.empty((256, B, ncls), device=self.device, dtype=torch.float32)
        e = eps.to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            check(lib().wd_f32_ctc_head(self._h, B, _ptr(e), Cc, H, W, _ptr(out), _stream_ptr()), "wd_f32_ctc_head")
        return out

    def unet_eval(self, x, timesteps, y, out=None):
        B = x.shape[0]
        if out is None:
            out = torch.empty_like(x)
        t_ptr, t_scalar = C.c_void_p(0), 0
        if isinstance(timesteps, int):
            t_scalar = timesteps
        else:
            timesteps = timesteps.to(device=self.device, dtype=torch.int64).contiguous()
            t_ptr = _ptr(timesteps)
        with torch.cuda.device(self.device):
            check(lib().wd_f32_unet_eval(self._h, B, _ptr(x), t_ptr, t_scalar, _ptr(y), _ptr(out), _stream_ptr()),
                  "wd_f32_unet_eval")
        return out

    def unet_eval_maps(self, x, timesteps, y, scales):
        """args.attentionMaps == 1 (unet.py:1645-1836): -> (eps, [attn1, attn2, attn3], context) with attn_i fp32
        [B, H_i * scales[i], W_i * scales[i], L] and context fp32 [B, L, context_dim]."""
        B = x.shape[0]
        out = torch.empty_like(x)
        t_ptr, t_scalar = C.c_void_p(0), 0
        if isinstance(timesteps, int):
            t_scalar = timesteps
        else:
            timesteps = timesteps.to(device=self.device, dtype=torch.int64).contiguous()
            t_ptr = _ptr(timesteps)
        l = lib()
        with torch.cuda.device(self.device):
            sp = _stream_ptr()
            check(l.wd_f32_unet_eval_maps(self._h, B, _ptr(x), t_ptr, t_scalar, _ptr(y), _ptr(out), sp), "wd_f32_unet_eval_maps")
            maps = []
            L = 0
            for which, sc in enumerate(scales):
                H, W, Lc = C.c_int(0), C.c_int(0), C.c_int(0)
                check(l.wd_f32_read_attention_map(self._h, which, int(sc), C.c_void_p(0), C.byref(H), C.byref(W), C.byref(Lc), sp),
                      "wd_f32_read_attention_map")
                m = torch.empty((B, H.value * sc, W.value * sc, Lc.value), device=self.device, dtype=torch.float32)
                check(l.wd_f32_read_attention_map(self._h, which, int(sc), _ptr(m), None, None, None, sp),
                      "wd_f32_read_attention_map")
                maps.append(m)
                L = Lc.value
            ctx = torch.empty((B, L, self.cfg.context_dim), device=self.device, dtype=torch.float32)
            check(l.wd_f32_read_context(self._h, _ptr(ctx), ctx.numel() * 4, sp), "wd_f32_read_context")
        return out, maps, ctx

    def sampler_step(self, x, t, y, mode, coef, noise=None, philox_seed=None, sample_offset=0, step_index=0,
                     eps_out=None):
        eps = self.unet_eval(x, int(t), y, out=eps_out)
        return self.sampler_update(x, eps, mode, coef, noise=noise, philox_seed=philox_seed, sample_offset=sample_offset,
                                   step_index=step_index)

    sampler_update = HotPathEngine.sampler_update

    def reserve(self, batch):
        self._ctx_key = None

    @property
    def last_launch_count(self):
        return lib().wd_f32_last_launch_count(self._h)

    @property
    def workspace_bytes(self):
        return lib().wd_f32_workspace_bytes(self._h)
