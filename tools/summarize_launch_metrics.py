#!/usr/bin/env python
"""Per-kernel summary of an ncu CSV with several metrics per launch (gpu__time_duration.sum, dram bytes, tensor pipe %):
launch count, time, share of the step, DRAM TB/s, time-weighted tensor-pipe activity."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, mi, vi, ui, idi, gi = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID", "Grid Size"))
    by_id = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = by_id.setdefault(int(r[idi]), {"name": r[ki].split("(")[0].replace("void ", ""), "grid": r[gi]})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
        d[r[mi]] = v
    return by_id


def main(path, top=40):
    by_id = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in by_id.values():
        a = agg[d["name"]]
        t = d.get("gpu__time_duration.sum", 0.0)
        a[0] += 1
        a[1] += t
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(by_id)} launches, {tot:.3f} ms (cold-cache, serialised)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{k[:66]:66s} n={a[0]:5d} {a[1]:8.3f} ms {100 * a[1] / tot:5.1f}%  dram {a[2] / max(a[1], 1e-9) / 1e9:6.2f} TB/s  tensor {a[3] / max(a[1], 1e-9):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
