#!/usr/bin/env python
"""Per-kernel summary of an ncu report:  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.md"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__inst_executed.sum", "warp insts")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full --clock-control none, report `{rep}` (cold-cache, serialised launches: shares, not absolutes)\n")
    print("| # | kernel | grid | block | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|---|---|" + "---|" * len(COLS))
    for k, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("ilsm::", "")
        cells = []
        for m, _ in COLS:
            if m not in col:
                cells.append("")
                continue
            v, u = r[col[m]], units[col[m]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(v + (" " + u if u in ("ms", "us", "ns", "Mbyte", "Kbyte", "Gbyte", "byte") else ""))
        print(f"| {k} | {name} | {r[col['Grid Size']]} | {r[col['Block Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
