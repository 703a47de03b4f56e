#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: per launch duration, SM clock, tensor-pipe activity, DRAM bytes
(the `roofline.traffic` figure), L2 / L1 throughput, registers, shared memory, and the top warp-stall sites.
Usage: summarize_ncu.py report.ncu-rep [max_launches]"""
import csv
import io
import subprocess
import sys


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.avg.per_second", "GHz"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
            ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smemB"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]
    print(f"# {rep}: {len(data)} launches (ncu --set full --clock-control none; cold-cache, serialised)")
    print("# " + " | ".join(f"{n}[{units[col[m]]}]" for m, n in want if m in col))
    for d in data[:maxn]:
        name = d[col["Kernel Name"]] if "Kernel Name" in col else ""
        name = name.split("(")[0].replace("wd::", "")[:48]
        vals = []
        for m, n in want:
            if m in col:
                v = d[col[m]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(v)
        print(f"{name:48s} " + " | ".join(vals))


if __name__ == "__main__":
    main()
