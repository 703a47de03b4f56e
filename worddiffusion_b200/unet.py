"""Drop-in for the reference ``unet.UNetModel`` (reference unet.py:1096-1836): same constructor, same ``forward``
signature, same ``state_dict`` keys (264), B200 engine underneath."""
import torch
import torch.nn as nn

import random

from . import _lib
from ._lib import VARIANT_UNET, check, lib
from .engine import _ptr, _stream_ptr
from .modules import CTCtopC, ResBlockConditional
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
        # SURVEY.md section 8f rank 4 -- the flag variants of unet.py:
        #   charLevelEmb (unet.py:853-863): the same embedding lookup through a flatten / view(BS, 10, 320) round trip
        #   charImages   (unet.py:1217-1223,1517-1539,1625-1627): three extra convolutions whose result the forward discards
        #   ocrTraining  (unet.py:1468,1829): the CTCtopC head on the predicted noise, returned with the attention maps
        #   wrdChrWrStyl (unet.py:1590-1591,1617-1618): context = wrd_proj(style vectors) instead of the encoded characters
        #   interpolation (unet.py:1558-1577): the label embedding of two random writers mixed by mix_rate
        self.interpolation = bool(getattr(args, "interpolation", False))
        if getattr(args, "attentionMaps", 0) == 1:
            # unet.py:1336-1364,1645-1836: middle_block1 key layout, forward returns (eps, attn1, attn2, attn3, context).  The maps
            # are the attention probabilities, which only the fp32 path materialises (the bf16 engine keeps them in registers).
            self.precision = "fp32"

        def extras():
            if getattr(args, "charImages", 0) == 1:  # unet.py:1217-1223
                self.conv_layer1 = nn.Conv2d(4, 16, kernel_size=(4, 16))
                self.conv_layer2 = nn.Conv2d(16, 160, kernel_size=(4, 12))
                self.conv_layer3 = nn.Conv2d(160, 320, kernel_size=(2, 6))
            self.wrd_proj = nn.Linear(4096, 320)  # unet.py:1243, only read when args.wrdChrWrStyl == 1

        self._build_tree(extras)
        if getattr(args, "ocrTraining", 0) == 1:
            self.auxhead = CTCtopC(4, (256, 3), vocab_size - 2)  # unet.py:1468
        # unet.py:1472 -- constructed, never called (gated by `if 0` at :1593); kept for state_dict parity
        self.res = ResBlockConditional(32, 1280, 320)

    def _attention_maps(self):
        return getattr(self.args, "attentionMaps", 0) == 1

    def _add_label_emb(self):
        # unet.py:1578-1581: the writer-style embedding is skipped when args.imgConditioned == 1
        return self.num_classes is not None and getattr(self.args, "imgConditioned", 0) != 1

    def forward(self, x, wrdChrWrStyl=None, original_images=None, timesteps=None, context=None, y=None,
                charContextImages=None, original_context=None, or_images=None, mix_rate=None, **kwargs):
        args = self.args
        if getattr(args, "charImages", 0) == 1:
            # unet.py:1517-1539: conv_layer1-3 run on the character images, and unet.py:1625-1627 then drops their output
            # (`context = context #+ output`): nothing of it reaches the result.  Only the reference's input contract is kept.
            bs = charContextImages.size(0)
            charContextImages.reshape(self.max_seq_len * bs, 4, 8, 32)
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
        variant = (getattr(args, "wrdChrWrStyl", 0) == 1 or (self.interpolation and mix_rate is not None))
        training = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if getattr(args, "charLevelEmb", 0) and context is not None:
            # unet.py:853-863: flatten -> embedding -> view(BS, 10, 320): the plain lookup when the context is [BS, 10] and the
            # embedding 320 wide, a RuntimeError from .view() otherwise
            if context.shape[1] != 10 or self.context_dim != 320:
                raise RuntimeError(f"shape '[{context.shape[0] * context.shape[1] // 10}, 10, 320]' is invalid for input of size "
                                   f"{context.numel() * self.context_dim}")
        if training and (variant or getattr(args, "ocrTraining", 0) == 1 and self._attention_maps()):
            raise NotImplementedError("worddiffusion_b200: the training step covers the default unet.UNetModel path only "
                                      "(no wrdChrWrStyl / interpolation / OCR-head gradients)")
        dense = None
        if getattr(args, "wrdChrWrStyl", 0) == 1:
            proj = self._wrd_proj(wrdChrWrStyl)  # unet.py:1590-1591
            if getattr(args, "imgConditioned", 0) == 0 and context is not None:
                dense = proj                       # unet.py:1617-1618: context = wrdChrWrStyl
        y_eff = y
        mix = None
        if self.interpolation and mix_rate is not None:
            # unet.py:1558-1572: two distinct random writers out of the first 339, drawn with the `random` module per call
            s1 = random.randint(0, 338)
            s2 = random.randint(0, 338)
            while s1 == s2:
                s2 = random.randint(0, 338)
            if max(s1, s2) >= self.num_classes:
                raise IndexError("index out of range in self")  # nn.Embedding's error for a table shorter than 339 rows
            mix = (s1, s2, float(mix_rate))
            y_eff = torch.full((x.shape[0],), self.num_classes, device=x.device, dtype=torch.int64)  # the scratch row
        if self._attention_maps():
            return self._run_attention_maps(x, timesteps, context, y_eff, dense, mix)
        if training:
            # train.py:285: the noise-prediction step; gradients come from the hand-written backward (training.py)
            from .training import unet_train_forward
            if context is None:
                raise NotImplementedError("worddiffusion_b200 needs the character context (context=None is not implemented)")
            return unet_train_forward(self, x, timesteps, context, y).type(x.dtype)
        if dense is None and mix is None:
            return self._run(x, timesteps, context, y, None)
        return self._run(x, timesteps, context, y_eff, None, dense_context=dense, label_mix=mix)

    def _wrd_proj(self, style):
        """wrd_proj(wrdChrWrStyl): Linear(4096, 320) in fp32 on the device (wd_f32_op_linear)."""
        if style.device.type != "cuda":
            raise _lib.WdError("worddiffusion_b200 has no CPU path: inputs must live on a CUDA (B200) device")
        s2 = style.detach().to(torch.float32).contiguous()
        K = s2.shape[-1]
        w, b = self.wrd_proj.weight, self.wrd_proj.bias
        if K != w.shape[1]:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({s2.numel() // K}x{K} and {w.shape[1]}x{w.shape[0]})")
        out = torch.empty(s2.shape[:-1] + (w.shape[0],), device=s2.device, dtype=torch.float32)
        M = s2.numel() // K
        with torch.cuda.device(s2.device):
            check(lib().wd_f32_op_linear(_ptr(s2), _ptr(w.detach().to(s2.device).float().contiguous()),
                                         _ptr(b.detach().to(s2.device).float().contiguous()), _ptr(out), M, w.shape[0], K,
                                         _stream_ptr()), "wd_f32_op_linear")
            torch.cuda.current_stream().synchronize()  # the temporaries above may be freed
        return out

    ATTENTION_MAP_SCALES = (8, 16, 8)  # unet.py:1787,1791,1795

    def _run_attention_maps(self, x, timesteps, context, y, dense=None, mix=None):
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
        if mix is not None:
            eng.set_label_mix(self.num_classes, *mix)
        if dense is not None:
            eng.set_context(dense)
        else:
            eng.encode_context(context, None)
        eps, maps, ctx = eng.unet_eval_maps(xin, timesteps, y, self.ATTENTION_MAP_SCALES)
        if getattr(self.args, "ocrTraining", 0) == 1:
            # unet.py:1827-1831: (h, attn1, attn2, attn3, tdec), tdec = auxhead(h)
            if self.training:
                raise NotImplementedError("worddiffusion_b200: the OCR head runs in eval mode (BatchNorm running statistics, no "
                                          "Dropout); call model.eval() as the reference's sampling does (train.py:203)")
            return eps.type(x.dtype), maps[0], maps[1], maps[2], eng.ctc_head(eps)
        return eps.type(x.dtype), maps[0], maps[1], maps[2], ctx
