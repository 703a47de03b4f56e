#!/usr/bin/env python
"""conv3x3 (+GroupNorm+SiLU in the producer) at the benchmark shapes, for ncu / CUDA-event timing of the pair kernel's
epilogue.  Usage: op_gnfuse.py [iters]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gpu_util import bf, conv3x3, conv3x3_gn_silu, f32, pack_conv, sync  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator().manual_seed(0)
for (B, H, W, Cin) in [(256, 4, 16, 320), (256, 4, 16, 640), (256, 8, 32, 640)]:
    x = bf(torch.randn(B, H, W, Cin, generator=g))
    wp = pack_conv(torch.randn(320, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin))
    bias = f32(torch.zeros(320))
    rb = f32(torch.randn(B, 320, generator=g))
    gamma, beta = f32(torch.ones(320)), f32(torch.zeros(320))
    for name, fn in (("plain", lambda: conv3x3(x, wp, bias, rowbias=rb)), ("gn", lambda: conv3x3_gn_silu(x, wp, bias, rb, gamma, beta))):
        for _ in range(2):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        sync()
        print(f"B={B} {H}x{W} Cin={Cin} {name}: {1e3 * e0.elapsed_time(e1) / iters:.1f} us (incl. output alloc fill)")
