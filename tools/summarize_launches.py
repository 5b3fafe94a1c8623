#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot:.3f} ms (cold-cache, serialised)")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:70]:70s} n={n:4d} {v:9.3f} ms {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
