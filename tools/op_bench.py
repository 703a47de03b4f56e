#!/usr/bin/env python
"""GPU micro-benchmark of single hot-path operators through the C ABI (wd_op_*), CUDA-event timed.
Each case rotates over enough buffer sets to exceed the 126 MB L2 ("cold") and also reports the same-buffer ("hot") time.
    python tools/op_bench.py [case ...]        (env WD_GEMM_DBG=<flags> selects kernel experiment switches)"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, P, S, bf, f32, pack_conv, pack_linear  # noqa: E402
from worddiffusion_b200._lib import check, lib  # noqa: E402


def timeit(fns, iters=20):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def gemm_case(M, N, K, residual=False, geglu=False, bias=True, nsets=6):
    w = pack_linear(torch.randn(N, K) / math.sqrt(K), geglu=geglu)
    b = f32(torch.randn(N) * 0.1) if bias else None
    oc = N // 2 if geglu else N
    sets = []
    for _ in range(nsets):
        a = bf(torch.randn(M, K))
        r = bf(torch.randn(M, oc)) if residual else None
        o = torch.empty(M, oc, device=DEV, dtype=torch.bfloat16)
        sets.append((a, r, o))

    def mk(a, r, o):
        return lambda: check(lib().wd_op_gemm(P(a), P(w), P(b), P(r), P(o), M, N, K, 0, int(geglu), 0, S()), "gemm")
    fns = [mk(*s) for s in sets]
    cold, hot = timeit(fns), timeit(fns[:1])
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + M * oc * (2 if residual else 1))
    return cold, hot, fl, by


def conv_case(B, H, W, Cin, Cout=320, residual=False, nsets=6):
    wp = pack_conv(torch.randn(Cout, Cin, 3, 3) / math.sqrt(9 * Cin))
    b = f32(torch.randn(Cout) * 0.1)
    sets = []
    for _ in range(nsets):
        x = bf(torch.randn(B, H, W, Cin))
        r = bf(torch.randn(B, H, W, Cout)) if residual else None
        o = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.bfloat16)
        sets.append((x, r, o))

    def mk(x, r, o):
        return lambda: check(lib().wd_op_conv3x3(P(x), P(wp), P(b), P(None), 0, P(r), P(o), B, H, W, Cin, Cout, 1, S()), "conv")
    fns = [mk(*s) for s in sets]
    cold, hot = timeit(fns), timeit(fns[:1])
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    by = 2.0 * B * H * W * (Cin + Cout * (2 if residual else 1))
    return cold, hot, fl, by


CASES = {
    "lin320": lambda: gemm_case(65536, 320, 320),
    "lin320_res": lambda: gemm_case(65536, 320, 320, residual=True),
    "qkv960": lambda: gemm_case(65536, 960, 320, bias=False),
    "geglu": lambda: gemm_case(65536, 2560, 320, geglu=True, nsets=4),
    "ffout_res": lambda: gemm_case(65536, 320, 1280, residual=True, nsets=4),
    "conv8x32": lambda: conv_case(256, 8, 32, 320),
    "conv8x32_res": lambda: conv_case(256, 8, 32, 320, residual=True),
    "conv8x32_640": lambda: conv_case(256, 8, 32, 640),
    "conv4x16": lambda: conv_case(256, 4, 16, 320),
    "lin320_4x16": lambda: gemm_case(16384, 320, 320),
}


def main():
    names = sys.argv[1:] or list(CASES)
    print(f"# WD_GEMM_DBG={os.environ.get('WD_GEMM_DBG', '0')}")
    print(f"{'case':14s} {'cold_us':>8s} {'hot_us':>8s} {'TF/s cold':>10s} {'GB/s cold':>10s}")
    for n in names:
        cold, hot, fl, by = CASES[n]()
        print(f"{n:14s} {cold:8.1f} {hot:8.1f} {fl / cold / 1e6:10.1f} {by / cold / 1e3:10.1f}", flush=True)


if __name__ == "__main__":
    main()
