#!/usr/bin/env python
"""Key numbers of every kernel in an .ncu-rep (`ncu --set full`): duration, DRAM bytes, instructions,
occupancy, issue rate and the warp-stall reasons (warps stalled per issue-active cycle).
Usage: scripts/ncu_rep_summary.py <report.ncu-rep> [kernel-substring]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want not in name:
            continue
        print(name[:100])
        for k in KEYS:
            if k in hdr:
                print("    %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = sorted(((float(r[i] or 0), h) for h, i in stall), reverse=True)[:6]
        print("    stalls (warps per issue-active cycle): " + ", ".join(
            "%s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for v, h in st))


if __name__ == "__main__":
    main()
