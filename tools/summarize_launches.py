#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of ONE DDPM step of
bench.py (the launches from the first kernel of a step -- the time-embedding table lookup, or the sinusoid kernel in builds
without the table -- up to the next one).  Usage: summarize_launches.py launches.csv [step]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("wd::", "").replace("void ", ""),
              float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]) for r in rows]
    starts = [i for i, n in enumerate(names) if "emb_from_table" in n[0]]
    if len(starts) < which + 2:
        starts = [i for i, n in enumerate(names) if "timestep_embed" in n[0]]
    lo, hi = starts[which], starts[which + 1]
    step = names[lo:hi]
    agg = collections.OrderedDict()
    for n, t, g, b in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: step {which} of the run = launches {lo}..{hi - 1} ({hi - lo} kernels, {total / 1e6:.3f} ms summed, "
          "cold-cache serialised ncu times: compare SHARES)")
    print(f"{'kernel':52s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:52s} {c:8d} {t / 1e3:10.1f} {t / total:7.3f}")
    print("\n# every launch of the step, in order")
    for n, t, g, b in step:
        print(f"{n:52s} grid={g:18s} block={b:14s} {t / 1e3:9.1f} us")


if __name__ == "__main__":
    main()
