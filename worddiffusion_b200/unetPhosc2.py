"""Drop-in for the reference ``unetPhosc2.UNetModelPhosc`` (reference unetPhosc2.py:1100-1210): numerically identical to
``unetPhosc``; the forward takes no ``mix_rate`` / ``**kwargs`` and asserts ``y.shape`` strictly (:1122).  The
reference's per-forward ``logging.info`` calls (:1181-1200) are not reproduced."""
from .unetPhosc import UNetModelPhosc as _Base
from .unet_base import default_args  # noqa: F401


class UNetModelPhosc(_Base):
    STRICT_Y = True

    def forward(self, x, phoscLabels=None, timesteps=None, context=None, y=None):
        return super().forward(x, phoscLabels, timesteps=timesteps, context=context, y=y)
