#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` log: one line per kernel launch with every metric.
Usage: scripts/ncu_csv_summary.py <log.csv> [kernel-substring]"""
import csv
import re
import sys


def main():
    rows = {}
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for r in csv.reader(open(sys.argv[1])):
        if len(r) < 15 or r[0] == "ID":
            continue
        k = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        if want in k:
            rows.setdefault((int(r[0]), k), {})[r[12]] = float(r[14].replace(",", ""))
    for (i, k), m in sorted(rows.items()):
        ms = m.get("gpu__time_duration.sum", 0) / 1e6
        print("%4d %-28s %8.3f ms " % (i, k, ms) + " ".join(
            "%s=%.4g" % (a.replace("l1tex__", "").replace(".sum", "").replace(".avg.pct_of_peak_sustained_active", "%"), b)
            for a, b in sorted(m.items()) if a != "gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
