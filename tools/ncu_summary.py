#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch with
duration, DRAM traffic, occupancy, issue utilisation, cache hit rates, pipe utilisation and the top stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.csv
"""
import csv
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    cols = [c for c in COLS if c in hdr]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "block", "grid"] + ["%s [%s]" % (c, units[hdr.index(c)]) for c in cols] + ["top stalls (warps per issue)"])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        st = sorted(((float(r[hdr.index(k)]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for k in stall), reverse=True)[:5]
        w.writerow([name, r[hdr.index("Block Size")], r[hdr.index("Grid Size")]] + [r[hdr.index(c)] for c in cols] +
                   ["; ".join("%s=%.2f" % (n, v) for v, n in st)])


if __name__ == "__main__":
    main()
