#!/usr/bin/env python3
"""Print the handful of ncu metrics we track from a .ncu-rep (needs `ncu` on PATH, no GPU).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("sm__cycles_elapsed.max", "cyc"),
    ("smsp__cycles_active.avg", "cyc_act"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    path = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), short) for m, short in WANT if m in hdr]
    print("| kernel | " + " | ".join(s for _, s in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "")[:44]
        vals = []
        for i, s in cols:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            u = units[i]
            vals.append(f"{v} {u}".strip() if s in ("time", "dram_rd", "dram_wr") else v)
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
