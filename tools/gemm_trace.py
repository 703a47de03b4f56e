#!/usr/bin/env python
"""Epilogue phase timeline of gemm_tc_kernel (CTA 0, leader warp of column half 0): run with WD_GEMM_DBG=1024.
Prints, per tile of that CTA, the clock64 deltas between the stamps (SM cycles; 1.9 GHz -> 1900 cycles = 1 us):
 0 tile start | 1 bias/row-bias vector ready | 2 LayerNorm row statistics read | 3 accumulator complete (tmem_full) |
 4 TMEM drained + released | 5 staging buffer free | 6 arithmetic + staging writes done | 7 fence + barrier | 8 TMA store issued."""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("WD_GEMM_DBG", "1024")
from gpu_util import DEV, P, S, bf, f32, pack_linear  # noqa: E402
from worddiffusion_b200._lib import LIB_PATH, check, lib  # noqa: E402


def trace(name, fn):
    raw = C.CDLL(LIB_PATH)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    raw.wdx_gemm_trace_clear()
    fn()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (12 * 64))()
    raw.wdx_gemm_trace_read(buf, 12 * 64)
    ct = (C.c_ulonglong * (4 * 160))()
    raw.wdx_gemm_cta_times_read(ct)
    rows = [[ct[i * 4 + k] for k in range(4)] for i in range(160) if ct[i * 4]]
    # the launch just traced, timed with events inside a back-to-back train of the same launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"## {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (events, 20 back-to-back)")
    if rows:
        t0 = min(r[0] for r in rows)
        print(f"   CTAs {len(rows)}: entry spread {max(r[0] for r in rows) - t0} ns | prologue done (max) {max(r[1] for r in rows) - t0} ns | "
              f"first exit {min(r[3] for r in rows) - t0} ns | last exit {max(r[3] for r in rows) - t0} ns | CTA0 {[x - t0 for x in rows[0]]}")
    prev_end = None
    for it in range(64):
        st = [buf[it * 12 + k] for k in range(9)]
        if st[0] == 0:
            break
        d = [st[k + 1] - st[k] if st[k + 1] and st[k] else 0 for k in range(8)]
        gap = st[0] - prev_end if prev_end else 0
        prev_end = st[8] or st[7]
        print(f"tile {it:2d}: gap {gap:6d} | vec {d[0]:5d} ln {d[1]:5d} tmem_full {d[2]:6d} drain {d[3]:5d} stg_free {d[4]:5d} math {d[5]:5d} "
              f"fence+bar {d[6]:5d} store {d[7]:5d} | total {st[8] - st[0] if st[8] else 0:6d}")


def main():
    M = 65536
    a = bf(torch.randn(M, 320))
    w = pack_linear(torch.randn(320, 320) / math.sqrt(320))
    b = f32(torch.randn(320) * 0.1)
    o = torch.empty(M, 320, device=DEV, dtype=torch.bfloat16)
    trace("lin320 (bf16 out, bias)", lambda: check(lib().wd_op_gemm(P(a), P(w), P(b), P(None), P(o), M, 320, 320, 0, 0, 0, S()), "g"))
    r16 = torch.randn(M, 320, device=DEV).half()
    o16 = torch.empty(M, 320, device=DEV, dtype=torch.float16)
    trace("lin320 + fp16 residual (res_k)", lambda: check(lib().wd_op_gemm_f16(P(a), P(w), P(b), P(r16), P(o16), M, 320, 320, S()), "g"))
    wg = pack_linear(torch.randn(2560, 320) / math.sqrt(320), geglu=True)
    bg = f32(torch.randn(2560) * 0.1)
    og = torch.empty(M, 1280, device=DEV, dtype=torch.bfloat16)
    trace("geglu 320 -> 2560", lambda: check(lib().wd_op_gemm(P(a), P(wg), P(bg), P(None), P(og), M, 2560, 320, 0, 1, 0, S()), "g"))
    kv = bf(torch.randn(M // 256, 10, 640))
    trace("to_q + ctx attention", lambda: check(lib().wd_op_q_ctx_attention(P(a), P(w), P(b), P(kv), P(o), M // 256, 256, 10, 4, 80 ** -0.5, S()), "g"))


if __name__ == "__main__":
    if os.environ.get("WD_TRACE_STREAM", "0") == "1":  # a non-default stream (PDL on the legacy NULL stream?)
        with torch.cuda.stream(torch.cuda.Stream()):
            main()
    else:
        main()
