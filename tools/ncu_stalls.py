#!/usr/bin/env python
"""Top warp-stall sites of one launch of an `ncu --set full --import-source on` report, grouped by SASS instruction
(the `--page source` view is SASS-level).  Usage: ncu_stalls.py report.ncu-rep <launch index> [top N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("#", rows[0][1][:140])
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[col["# Samples"]])
        except ValueError:
            continue
        total += n
        data.append((n, r))
    print(f"# total samples {total}")
    order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(order):
        n, r = data[i]
        st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {100.0 * n / max(total, 1):5.1f}%  {r[col['Source']][:90]:90s} " + " ".join(f"{c}={v}" for v, c in st if v))


if __name__ == "__main__":
    main()
