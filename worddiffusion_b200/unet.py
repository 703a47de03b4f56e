"""Drop-in for the reference ``unet.UNetModel`` (reference unet.py:1096-1836): same constructor, same ``forward``
signature, same ``state_dict`` keys (264), B200 engine underneath."""
import torch
import torch.nn as nn

from ._lib import VARIANT_UNET
from .modules import ResBlockConditional
from .unet_base import UNetBase, default_args  # noqa: F401


class UNetModel(UNetBase):
    VARIANT = VARIANT_UNET

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=True, transformer_depth=1, context_dim=768, vocab_size=256, n_embed=None,
                 legacy=False, args=None, max_seq_len=20):
        super().__init__()
        self._init_common(image_size, in_channels, model_channels, out_channels, num_res_blocks,
                          attention_resolutions, dropout, channel_mult, conv_resample, dims, num_classes,
                          use_checkpoint, use_fp16, num_heads, num_head_channels, num_heads_upsample,
                          use_scale_shift_norm, resblock_updown, use_new_attention_order, use_spatial_transformer,
                          transformer_depth, context_dim, vocab_size, n_embed, legacy, args, max_seq_len)
        bad = [f for f in ("charLevelEmb", "charImages", "ocrTraining", "wrdChrWrStyl") if getattr(args, f, 0)]
        if bad:
            # these flags change the state_dict key set and the return arity (unet.py:1468,1217-1223);
            # SURVEY.md section 8f ranks them as "next"
            raise NotImplementedError("worddiffusion_b200.unet.UNetModel does not implement args." + ", args.".join(bad))
        if getattr(args, "attentionMaps", 0) == 1:
            # unet.py:1336-1364,1645-1836: middle_block1 key layout, forward returns (eps, attn1, attn2, attn3, context).  The maps
            # are the attention probabilities, which only the fp32 path materialises (the bf16 engine keeps them in registers).
            self.precision = "fp32"

        def extras():
            self.wrd_proj = nn.Linear(4096, 320)  # unet.py:1243, only read when args.wrdChrWrStyl == 1

        self._build_tree(extras)
        # unet.py:1472 -- constructed, never called (gated by `if 0` at :1593); kept for state_dict parity
        self.res = ResBlockConditional(32, 1280, 320)

    def _attention_maps(self):
        return getattr(self.args, "attentionMaps", 0) == 1

    def _add_label_emb(self):
        # unet.py:1578-1581: the writer-style embedding is skipped when args.imgConditioned == 1
        return self.num_classes is not None and getattr(self.args, "imgConditioned", 0) != 1

    def forward(self, x, wrdChrWrStyl=None, original_images=None, timesteps=None, context=None, y=None,
                charContextImages=None, original_context=None, or_images=None, mix_rate=None, **kwargs):
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
        if self._attention_maps():
            return self._run_attention_maps(x, timesteps, context, y)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # train.py:285: the noise-prediction step; gradients come from the hand-written backward (training.py)
            from .training import unet_train_forward
            if context is None:
                raise NotImplementedError("worddiffusion_b200 needs the character context (context=None is not implemented)")
            return unet_train_forward(self, x, timesteps, context, y).type(x.dtype)
        return self._run(x, timesteps, context, y, None)

    ATTENTION_MAP_SCALES = (8, 16, 8)  # unet.py:1787,1791,1795

    def _run_attention_maps(self, x, timesteps, context, y):
        """args.attentionMaps == 1 (unet.py:1645-1836): returns the reference's 5-tuple ``(eps, attn1, attn2, attn3, context)`` --
        attn_i = head-summed attn2 probabilities of the last SpatialTransformer of the input blocks / the middle block / the
        output blocks, nearest-upsampled by 8 / 16 / 8 to [B, 64, 256, 10]; ``context`` = the encoded characters [B, 10, 320]."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("worddiffusion_b200: training with args.attentionMaps == 1 is not implemented")
        if self.precision != "fp32":
            raise NotImplementedError("worddiffusion_b200: attention maps come from the fp32 path (model.precision = 'fp32')")
        if context is None:
            raise NotImplementedError("worddiffusion_b200 needs the character context (context=None is not implemented)")
        if x.device.type != "cuda":
            from . import _lib
            raise _lib.WdError("worddiffusion_b200 has no CPU path: inputs must live on a CUDA (B200) device")
        eng = self.engine(x.device, latent_hw=x.shape[2:])
        xin = x.to(torch.float32).contiguous()
        if y is not None:
            y = y.to(device=x.device, dtype=torch.int64).contiguous()
        eng.encode_context(context, None)
        eps, maps, ctx = eng.unet_eval_maps(xin, timesteps, y, self.ATTENTION_MAP_SCALES)
        return eps.type(x.dtype), maps[0], maps[1], maps[2], ctx
