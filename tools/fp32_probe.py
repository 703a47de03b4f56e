"""One fp32-mode UNet evaluation at the benchmark batch (for `ncu -k regex:f32_gemm_kernel`); prints its device time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import weights as W  # noqa: E402
from worddiffusion_b200.unet import UNetModel, default_args  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
kw = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 1),
          channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320, vocab_size=53, max_seq_len=10)
m = UNetModel(args=default_args(dev), **kw)
m.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
m = m.to(dev).eval()
m.precision = "fp32"
inp = {k: v.to(dev) for k, v in W.make_inputs(B, seed=1234).items()}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    eps = m(inp["x"], None, timesteps=inp["t"], context=inp["context"], y=inp["y"])
    e1.record()
torch.cuda.synchronize()
print(f"fp32 mode, batch {B}: {e0.elapsed_time(e1):.2f} ms (first call: includes the arena allocation), "
      f"{m._engine_f32.last_launch_count} launches, finite={bool(torch.isfinite(eps).all())}")
