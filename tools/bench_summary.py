#!/usr/bin/env python
"""Prints the headline numbers and the per-kernel-class table of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"ms/step {d['ms_per_step']:.3f}  value {d['value']:.0f} {d['unit']}  step {d['step_tflops']} TF/s "
      f"({d['step_frac_of_bf16_sustained']:.3f} of sustained)  gemm frac {d['roofline']['frac']}  e2e {d['e2e']['value']:.0f}"
      f"  clocks {d['clocks']}")
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"  {k:16s} n={v['launches_per_step']:3d}  {v['ms_per_step']:8.4f} ms  share {v['share']:.3f}  "
          f"{v['tflops']} TF/s  {v['gbs']} GB/s")
