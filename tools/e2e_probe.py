import sys, os, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import weights as W
from worddiffusion_b200.unet import UNetModel, default_args
dev = "cuda:0"
KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 1),
          channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320, vocab_size=53, max_seq_len=10)
m = UNetModel(args=default_args(dev), **KW)
m.load_state_dict(W.make_state_dict(W.load_spec("unet"), 1234), strict=True)
m = m.to(dev).eval()
B = 256
inp = W.make_inputs(B, seed=1)
x, t, ctx, y = inp["x"].to(dev), torch.full((B,), 500, device=dev), inp["context"].to(dev), inp["y"].to(dev)
eng = m.engine(dev)
def ev(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    return e0.elapsed_time(e1) / n * 1e3, (w1 - w0) / n * 1e6
eng.encode_context(ctx, None)
out = torch.empty_like(x)
print("unet_eval only          (gpu us, wall us):", ev(lambda: eng.unet_eval(x, t, y, out)))
def enc():
    eng._ctx_key = None
    eng.encode_context(ctx, None)
print("encode_context only     :", ev(enc))
with torch.no_grad():
    print("forward same ctx object :", ev(lambda: m(x, None, timesteps=t, context=ctx, y=y)))
    print("forward new ctx object  :", ev(lambda: m(x, None, timesteps=t, context=ctx.clone(), y=y)))
hx, ht, hc, hy = (v.cpu().pin_memory() for v in (x, t, ctx, y))
heps = torch.empty_like(hx).pin_memory()
def e2e():
    dx = hx.to(dev, non_blocking=True); dt = ht.to(dev, non_blocking=True); dc = hc.to(dev, non_blocking=True); dy = hy.to(dev, non_blocking=True)
    with torch.no_grad():
        e = m(dx, None, timesteps=dt, context=dc, y=dy)
    heps.copy_(e, non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("e2e step                :", ev(e2e))
def host_only():
    m._weights_signature()
t0 = time.perf_counter()
for _ in range(200): host_only()
print("weights_signature host us:", (time.perf_counter() - t0) / 200 * 1e6)
