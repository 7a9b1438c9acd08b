#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file F ...).
usage: launch_summary.py launches.csv [first_id [last_id]]   (ids select a window, e.g. the timed steps)"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    tot = OrderedDict()
    for r in rows[1:]:
        if not (lo <= int(r[ii]) <= hi):
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)[:76]
        t, n = tot.get(name, (0.0, 0))
        tot[name] = (t + float(r[vi].replace(",", "")) / 1e6, n + 1)
    s = sum(t for t, _ in tot.values())
    print("# %s  launches %d..%d: %d kernels, %.3f ms" % (path, lo, min(hi, len(rows) - 2), sum(n for _, n in tot.values()), s))
    for name, (t, n) in tot.items():
        print("%-78s %5d x %9.3f ms %6.1f%%" % (name, n, t, 100.0 * t / s))


if __name__ == "__main__":
    main()
