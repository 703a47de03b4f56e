"""Shared shell of the drop-in UNet modules: reference constructor arguments in, reference ``state_dict`` layout
out, forward pass delegated to the sm_100a engine."""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib
from .engine import F32Engine, HotPathEngine, make_config
from .modules import CharacterEncoder, add_blocks


def default_args(device="cuda:0", **over):
    """The ``args`` namespace the reference constructors read (unet.py:1209-1217,1336,1468 / unetPhosc.py:1120)."""
    ns = SimpleNamespace(device=device, interpolation=False, charLevelEmb=0, charImages=0, attentionMaps=0,
                         ocrTraining=0, imgConditioned=0, wrdChrWrStyl=0, phosc=1, phos=0)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


class UNetBase(nn.Module):
    VARIANT = None
    PHOSC_LEN = 769  # 165 PHOS + 604 PHOC entries (ResPhoSCNetZSL/modules/utils/phos_generator.py:70-78, phoc_generator.py:78-90)

    def _init_common(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                     attention_resolutions, dropout, channel_mult, conv_resample, dims, num_classes, use_checkpoint,
                     use_fp16, num_heads, num_head_channels, num_heads_upsample, use_scale_shift_norm,
                     resblock_updown, use_new_attention_order, use_spatial_transformer, transformer_depth,
                     context_dim, vocab_size, n_embed, legacy, args, max_seq_len):
        # ---- what the sm_100a path implements; everything else fails loudly (no silent fallback) ----
        unsupported = []
        if dims != 2: unsupported.append("dims != 2")
        if not use_spatial_transformer: unsupported.append("use_spatial_transformer=False (AttentionBlock)")
        if use_scale_shift_norm: unsupported.append("use_scale_shift_norm")
        if resblock_updown: unsupported.append("resblock_updown")
        if not conv_resample: unsupported.append("conv_resample=False")
        if use_fp16: unsupported.append("use_fp16")
        if n_embed is not None: unsupported.append("n_embed (id_predictor head)")
        if legacy: unsupported.append("legacy")
        if dropout: unsupported.append("dropout > 0")
        if isinstance(context_dim, (list, tuple)): unsupported.append("per-level context_dim list")
        if context_dim is None: unsupported.append("context_dim=None")
        if args is None: unsupported.append("args=None")
        if unsupported:
            raise NotImplementedError("worddiffusion_b200 does not implement: " + ", ".join(unsupported))
        if num_heads == -1 and num_head_channels == -1:
            raise AssertionError("Either num_heads or num_head_channels has to be set")
        self.args = args
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.predict_codebook_ids = False
        self.transformer_depth = transformer_depth
        self.context_dim = context_dim
        self.vocab_size = vocab_size
        self.max_seq_len = max_seq_len
        self.interpolation = args.interpolation
        if self.interpolation and self.VARIANT != _lib.VARIANT_UNET:
            raise NotImplementedError("worddiffusion_b200 implements args.interpolation (random style mixing) for unet.UNetModel only")
        self._engine = None
        self._engine_sig = None
        self._train_engine = None
        self._engine_f32 = None
        self._engine_f32_sig = None
        # "bf16": the tcgen05 engine (bf16 operands, fp32 accumulation; 1e-2 of the reference).  "fp32": fp32 storage and FFMA
        # arithmetic (1e-4 of the reference, north_star's fp32 mode).  Not a constructor argument: the reference has none.
        self.precision = "bf16"

    def _build_tree(self, extra_before_blocks=None):
        mc = self.model_channels
        ted = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.word_emb = CharacterEncoder(self.vocab_size, self.context_dim, self.max_seq_len)
        if extra_before_blocks is not None:
            extra_before_blocks()
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(self.num_classes, ted)
        add_blocks(self, in_channels=self.in_channels, model_channels=mc, out_channels=self.out_channels,
                   num_res_blocks=self.num_res_blocks, attention_resolutions=self.attention_resolutions,
                   channel_mult=self.channel_mult, num_heads=self.num_heads,
                   num_head_channels=self.num_head_channels, transformer_depth=self.transformer_depth,
                   context_dim=self.context_dim, ted=ted, attention_maps_layout=self._attention_maps())

    # ------------------------------------------------------------------ engine plumbing
    def _add_label_emb(self):
        return self.num_classes is not None

    def _phosc_len(self):
        return 0

    def _attention_maps(self):
        return False

    _MIDDLE_RENAMES = (("middle_block1.0.0.", "middle_block.0."), ("middle_block1.0.1.", "middle_block.1."),
                       ("middle_block1.1.0.", "middle_block.2."))

    def _engine_state_items(self):
        """state_dict entries under the names the engines know: the attentionMaps = 1 layout (middle_block1.*, unet.py:1336-1364)
        holds the same layers in the same order as middle_block.{0,1,2}.*."""
        for k, v in self.state_dict().items():
            for old, new in self._MIDDLE_RENAMES:
                if k.startswith(old):
                    k = new + k[len(old):]
                    break
            if k == "label_emb.weight" and self._engine_classes() != self.num_classes:
                v = torch.cat([v, v.new_zeros((1, v.shape[1]))], dim=0)  # scratch row of the style interpolation
            yield k, v

    def _engine_classes(self):
        """args.interpolation: the engines get one scratch class after the model's own (wd_engine_set_label_mix writes it)."""
        if self.num_classes is not None and getattr(self, "interpolation", False):
            return self.num_classes + 1
        return self.num_classes

    def engine(self, device=None, latent_hw=None):
        """The B200 engine bound to this module's parameters (created / re-synchronised lazily)."""
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 has no CPU path: move the model and its inputs to a CUDA (B200) device")
        latent_hw = tuple(latent_hw) if latent_hw is not None else (8, 32)
        if self.precision == "fp32":
            return self._f32_engine(device, latent_hw)
        if self.precision != "bf16":
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        if self._engine is None or self._engine.device != device or self._engine.latent_hw != latent_hw:
            self._engine = HotPathEngine(
                variant=self.VARIANT, in_channels=self.in_channels, model_channels=self.model_channels,
                out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult,
                num_heads=self.num_heads, num_head_channels=self.num_head_channels,
                transformer_depth=self.transformer_depth, context_dim=self.context_dim, vocab_size=self.vocab_size,
                num_classes=self._engine_classes(), max_seq_len=self.max_seq_len, latent_hw=latent_hw,
                add_label_emb=self._add_label_emb(), phosc_len=self._phosc_len(), device=device)
            self._engine_sig = None
        sig = self._weights_signature()
        if sig != self._engine_sig:
            self._engine.load_state(self._engine_state_items(), self.word_emb.positional_encoding)
            self._engine_sig = sig
        return self._engine

    def _f32_engine(self, device, latent_hw):
        if self._engine_f32 is None or self._engine_f32.device != device or self._engine_f32.latent_hw != latent_hw:
            cfg = make_config(
                variant=self.VARIANT, in_channels=self.in_channels, model_channels=self.model_channels,
                out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult,
                num_heads=self.num_heads, num_head_channels=self.num_head_channels,
                transformer_depth=self.transformer_depth, context_dim=self.context_dim, vocab_size=self.vocab_size,
                num_classes=self._engine_classes(), max_seq_len=self.max_seq_len, latent_hw=latent_hw,
                add_label_emb=self._add_label_emb(), phosc_len=self._phosc_len())
            self._engine_f32 = F32Engine(cfg, latent_hw, device)
            self._engine_f32_sig = None
        sig = self._weights_signature()
        if sig != self._engine_f32_sig:
            self._engine_f32.load_state(self._engine_state_items(), self.word_emb.positional_encoding)
            self._engine_f32_sig = sig
        return self._engine_f32

    def _weights_signature(self):
        """(data_ptr, version) of every parameter: re-pack the engine's bf16 weights when any of them changed.  Walking the
        module tree costs ~0.5 ms per call (264 parameters), more than the host side of a whole denoising step, so the walk is
        cached as (owner module's ``_parameters`` dict, name, Parameter) triples and every call only re-checks that each dict
        still holds the SAME Parameter object (~20 us): a Parameter replaced by hand (``m.out[2].bias = nn.Parameter(...)``)
        is picked up by the very next forward.  `_apply` (.to / .cuda / .float) and `load_state_dict` drop the cache; a
        sub-MODULE swapped after construction needs `invalidate_engine()`."""
        cache = getattr(self, "_param_cache", None)
        if cache is not None:
            for owner, name, p in cache:
                if owner.get(name) is not p:
                    cache = None
                    break
        if cache is None:
            cache = []
            for mod in self.modules():
                for name, p in mod._parameters.items():
                    if p is not None:
                        cache.append((mod._parameters, name, p))
            self._param_cache = cache
        return HotPathEngine.weights_signature((None, p) for _, _, p in cache)

    def invalidate_engine(self):
        self._param_cache = None
        self._engine_sig = None
        self._engine_f32_sig = None

    def _apply(self, fn, *args, **kwargs):
        self._param_cache = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._param_cache = None
        return super().load_state_dict(*args, **kwargs)

    def train_engine(self, device=None, latent_hw=None):
        """The B200 training engine (forward + hand-written backward) bound to this module's parameters.  One trainer serves
        one latent size (its plans and activation arena are laid out for it): a different ``latent_hw`` builds a new one."""
        from .training import TrainEngine
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 has no CPU path: move the model and its inputs to a CUDA (B200) device")
        latent_hw = tuple(int(v) for v in latent_hw) if latent_hw is not None else (
            self._train_engine.latent_hw if self._train_engine is not None else (8, 32))
        if self._train_engine is None or self._train_engine.device != device or self._train_engine.latent_hw != latent_hw:
            self._train_engine = TrainEngine(self, device, latent_hw)
        return self._train_engine

    def _run(self, x, timesteps, context, y, phosc, dense_context=None, label_mix=None):
        if context is None:
            raise NotImplementedError("worddiffusion_b200 needs the character context (context=None is not implemented)")
        if x.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 has no CPU path: inputs must live on a CUDA (B200) device")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise AssertionError(f"x must be [B, {self.in_channels}, H, W], got {tuple(x.shape)}")
        if x.shape[0] == 0:   # an empty batch passes through every reference layer: [0, out_channels, H, W]
            return x.new_empty((0, self.out_channels) + tuple(x.shape[2:]))
        eng = self.engine(x.device, latent_hw=x.shape[2:])
        xin = x.to(torch.float32).contiguous()
        if y is not None:
            y = y.to(device=x.device, dtype=torch.int64).contiguous()
        if label_mix is not None:
            eng.set_label_mix(self.num_classes, *label_mix)
        if dense_context is not None:
            eng.set_context(dense_context)
        else:
            eng.encode_context(context, phosc)
        out = eng.unet_eval(xin, timesteps, y)
        return out.type(x.dtype)
