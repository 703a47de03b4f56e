"""attention operator determinism: the same (sample, head, query tile) must give the same bits whatever the batch around it."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, P, S, bf, sync
from worddiffusion_b200._lib import check, lib
g = torch.Generator().manual_seed(0)
for (Sq, Skv) in [(256, 779), (256, 256), (64, 779)]:
    B = 128
    q = bf(torch.randn(B, Sq, 320, generator=g)); k = bf(torch.randn(B, Skv, 320, generator=g)); v = bf(torch.randn(B, Skv, 320, generator=g))
    def run(qq, kk, vv):
        o = torch.empty_like(qq)
        check(lib().wd_op_attention(P(qq), 320, P(kk), P(vv), 320, P(o), 320, qq.shape[0], Sq, Skv, 4, 1 / math.sqrt(80), S()), "attn")
        sync(); return o
    big = run(q, k, v)
    big2 = run(q, k, v)
    idx = torch.tensor([0, 65, 127], device=DEV)
    small = run(q[idx].contiguous(), k[idx].contiguous(), v[idx].contiguous())
    d = (big[idx].float() - small.float()).abs()
    print(Sq, Skv, "big vs big:", int((big != big2).sum()), " big vs small: n_diff", int((d > 0).sum()), "max", float(d.max()),
          "rows differing per sample", [(int((d[i] > 0).any(dim=1).sum())) for i in range(3)])
    if int((d > 0).sum()):
        i = 1; rows = (d[i] > 0).any(dim=1).nonzero().flatten()[:10].tolist(); cols = (d[i] > 0).any(dim=0).nonzero().flatten()[:10].tolist()
        print("  sample 65 rows", rows, "cols", cols)
