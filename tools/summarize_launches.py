#!/usr/bin/env python3
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, average, share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> microseconds
        name = re.sub(r"^void\s+", "", r["Kernel Name"])
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, v))
    agg = OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / 1e3:.3f} | {t / c:.2f} | {t / total:.3f} |")
    print(f"\nTotal {total / 1e3:.3f} ms over {len(rows)} launches")


if __name__ == "__main__":
    main(sys.argv[1])
