#!/usr/bin/env python
"""Phase timeline of the fused transformer-block kernel (csrc/tblock.cu), CTA 0, from the clock64 stamps recorded under
WD_TBLOCK_TRACE=1.  Usage: WD_TBLOCK_PAIR=0|1 python tools/tblock_trace.py [batch]   (cycles at ~1.9 GHz: 1900 = 1 us)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["WD_TBLOCK_TRACE"] = "1"
from gpu_util import DEV  # noqa: E402
from worddiffusion_b200._lib import LIB_PATH  # noqa: E402
import test_gpu_tblock as T  # noqa: E402

MMA = {0: "tile start", 1: "g landed + X free", 2: "proj_in issued", 3: "A ready (attn1)", 4: "S1 issued", 5: "P1 ready", 6: "PN1 issued",
       7: "A ready (attn2)", 8: "S2 issued", 9: "P2 ready", 10: "PN2 issued", 11: "A ready (ff)", 12: "ff issued", 13: "A ready (proj_out)",
       14: "proj_out issued"}
EPI = {0: "tile start", 1: "proj_in acc", 2: "LN copy done", 3: "P1 written", 4: "attn1 acc", 5: "LN copy done", 6: "P2 written",
       7: "attn2 acc", 8: "LN copy done", 9: "ff chunks done", 10: "ff acc", 11: "x3 copy done", 12: "proj_out acc", 13: "out tile formed",
       14: "stored"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    HW, L = 256, 10
    W = T._weights(5)
    gen = torch.Generator().manual_seed(1)
    g16 = torch.randn(B * HW, T.CH, generator=gen).to(DEV).to(torch.bfloat16)
    x16 = torch.randn(B * HW, T.CH, generator=gen).to(DEV).to(torch.float16)
    ctx16 = torch.randn(B * L, T.CH, generator=gen).to(DEV).to(torch.bfloat16)
    raw = C.CDLL(LIB_PATH)
    buf = (C.c_ulonglong * (2 * 8 * 64))()
    for _ in range(2):
        T._run(W, g16, x16, ctx16, B, HW, L, 0)
    raw.wdx_tblock_trace_read(buf, 1)
    T._run(W, g16, x16, ctx16, B, HW, L, 0)
    raw.wdx_tblock_trace_read(buf, 0)
    print(f"# WD_TBLOCK_PAIR={os.environ.get('WD_TBLOCK_PAIR', '1')} batch {B}: CTA 0")
    base = buf[0]
    for it in range(8):
        m = [buf[(0 * 8 + it) * 64 + k] for k in range(64)]
        e = [buf[(1 * 8 + it) * 64 + k] for k in range(64)]
        if not m[0]:
            break
        print(f"## tile {it}: start {m[0] - base} .. proj_out issued {m[14] - base} ; epilogue stored {e[14] - base}  (tile total {e[14] - m[0]})")
        print("  MMA thread: " + " | ".join(f"{MMA[k]} +{m[k] - m[k - 1]}" for k in range(1, 15) if m[k] and m[k - 1]))
        print("  epilogue  : " + " | ".join(f"{EPI[k]} +{e[k] - e[k - 1]}" for k in range(1, 15) if e[k] and e[k - 1]))
        print("  ff chunks (MMA issue done, since A ready): " + " ".join(str(m[16 + c] - m[11]) for c in range(8) if m[16 + c]))
        print("  ff chunks (epilogue: acc full / operand written, since MMA's A ready): " +
              " ".join(f"{e[16 + 2 * c] - m[11]}/{e[17 + 2 * c] - m[11]}" for c in range(8) if e[16 + 2 * c]))


if __name__ == "__main__":
    main()
