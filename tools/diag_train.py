"""GPU diagnostic: training-forward / inference-forward / oracle agreement and per-parameter gradient errors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import train_oracle as TO
import weights as W
from worddiffusion_b200.unet import UNetModel, default_args
DEV = "cuda:0"
KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
          attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
          vocab_size=53, max_seq_len=10)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
seed = 1234 + B
m = UNetModel(args=default_args(DEV), **KW)
sd = W.make_state_dict(W.load_spec("unet"), 1234)
zero = [z for z in os.environ.get("DIAG_ZERO", "").split(",") if z]
for k in sd:
    if any(z in k for z in zero):
        sd[k] = torch.zeros_like(sd[k])
m.load_state_dict(sd, strict=True)
m = m.to(DEV)
inp = W.make_inputs(B, seed=seed)
noise = torch.randn((B, 4, 8, 32), generator=torch.Generator().manual_seed(5 + B))
args = dict(timesteps=inp["t"].to(DEV), context=inp["context"].to(DEV), y=inp["y"].to(DEV))
m.eval()
with torch.no_grad():
    e_inf = m(inp["x"].to(DEV), None, **args).cpu()
m.train()
pred = m(inp["x"].to(DEV), None, **args)
loss = torch.nn.MSELoss()(noise.to(DEV), pred)
loss.backward()
e_tr = pred.detach().cpu()
ref_loss, ref_eps, ref = TO.unet_loss_and_grads(sd, inp["x"], inp["t"], inp["context"], inp["y"], noise)
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
print("t", inp["t"].tolist(), "y", inp["y"].tolist())
print("inference vs oracle", rel(e_inf, ref_eps), " train-fwd vs oracle", rel(e_tr, ref_eps), " train vs inference", rel(e_tr, e_inf))
for b in range(B):
    print(f"  sample {b}: inf {rel(e_inf[b], ref_eps[b]):.3e} train {rel(e_tr[b], ref_eps[b]):.3e}")
print("loss", float(loss), float(ref_loss))
errs = []
for n, p in m.named_parameters():
    if ref[n] is None or float(ref[n].norm()) < 1e-6: continue
    g = p.grad.cpu().double().reshape(-1); r = ref[n].double().reshape(-1)
    errs.append((float((g - r).norm() / r.norm()), float(g.norm() / r.norm()), n))
errs.sort(reverse=True)
for e in errs[:40]: print("%.3e  ratio %.4f  %s" % e)
print("median", errs[len(errs)//2][0], "n", len(errs))

# ---- time-embedding chain vs fp32 torch ----
import ctypes as C
import torch.nn.functional as F
import unet_oracle as UO
from worddiffusion_b200._lib import lib, check
eng = m._train_engine
pred = m(inp["x"].to(DEV), None, **args)  # fresh forward (the backward above consumed the plan state)
def rd(name, shape, dtype):
    t = torch.empty(shape, device=DEV, dtype=dtype)
    check(lib().wd_trainer_read_tensor(eng._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel() * t.element_size(),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), name)
    torch.cuda.synchronize()
    return t.float().cpu()
temb = UO.timestep_embedding(inp["t"], 320)
h1p = F.linear(temb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
h1 = F.silu(h1p)
embp = F.linear(h1, sd["time_embed.2.weight"], sd["time_embed.2.bias"]) + F.embedding(inp["y"], sd["label_emb.weight"])
emb_act = F.silu(embp)
print("temb", rel(rd("temb", (B, 320), torch.bfloat16), temb))
print("h1p", rel(rd("h1p", (B, 1280), torch.bfloat16), h1p))
print("h1", rel(rd("h1", (B, 1280), torch.bfloat16), h1))
print("embp", rel(rd("embp", (B, 1280), torch.bfloat16), embp))
print("emb_act", rel(rd("emb_act", (B, 1280), torch.bfloat16), emb_act))
eo = rd("emb_out", (B, 2560), torch.float32)
names = [k[:-len("emb_layers.1.weight")] for k in sd if k.endswith("emb_layers.1.weight") and not k.startswith("res.")]
print("emb_layers order", names)
for i, pfx in enumerate(names):
    ref_o = F.linear(emb_act, sd[pfx + "emb_layers.1.weight"], sd[pfx + "emb_layers.1.bias"])
    print("  emb_out", pfx, rel(eo[:, 320 * i:320 * (i + 1)], ref_o))
