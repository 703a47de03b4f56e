"""Drop-in for the reference ``unetPhosc.UNetModelPhosc`` (reference unetPhosc.py:751-1159): standard transformer
block (self-attention, then cross-attention over chars + PHOSC tokens), 246 ``state_dict`` keys."""
import torch

from ._lib import VARIANT_PHOSC
from .unet_base import UNetBase, default_args  # noqa: F401


class UNetModelPhosc(UNetBase):
    VARIANT = VARIANT_PHOSC
    STRICT_Y = False

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
        self._build_tree()

    def _phosc_len(self):
        # unetPhosc.py:1120-1130: PHOSC tokens are embedded and concatenated when args.phosc or args.phos is set
        return self.PHOSC_LEN if (getattr(self.args, "phosc", 0) == 1 or getattr(self.args, "phos", 0) == 1) else 0

    def forward(self, x, phoscLabels=None, timesteps=None, context=None, y=None, mix_rate=None, **kwargs):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if self.num_classes is not None:
            if self.STRICT_Y:
                assert y.shape == (x.shape[0],)       # unetPhosc2.py:1122
            elif y.shape[0] != x.shape[0]:
                y = y[:x.shape[0]]                    # unetPhosc.py:1089-1090
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the training step is built for unet.UNetModel, the model train.py:403 trains; running the inference
            # engine here would hand the caller a tensor without grad_fn and a loss.backward() that silently trains nothing
            raise NotImplementedError(
                "worddiffusion_b200: training UNetModelPhosc is not implemented (no backward for the 256-token self-attention "
                "and the 779-token cross-attention); call model.eval() or wrap the call in torch.no_grad() for inference")
        return self._run(x, timesteps, context, y, phoscLabels if self._phosc_len() else None)
