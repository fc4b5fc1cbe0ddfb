#!/usr/bin/env python
"""Summarise an ncu launch list (csv of gpu__time_duration.sum) for one step of bench.py."""
import collections
import csv
import re
import sys


def main(path, detail=False, which=1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if r["Kernel Name"].startswith("prep_kernel")]
    step = rows[idx[which]:idx[which + 1]]
    tot = sum(float(r["Metric Value"]) for r in step)
    print("%s: %d launches in one step, sum of kernel times %.3f ms" % (path, len(step), tot / 1e6))
    agg = collections.OrderedDict()
    for r in step:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:60]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("   %-60s n=%4d  %9.1f us  %5.1f%%  avg %8.2f us" % (k, c, v / 1e3, 100 * v / tot, v / 1e3 / c))
    if detail:
        for r in step:
            n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("gemm_f64_kernel", "gemm")[:40]
            print("%-42s grid=%-14s %8.1f us" % (n, r["Grid Size"], float(r["Metric Value"]) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], "--detail" in sys.argv)
