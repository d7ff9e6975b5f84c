#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, average).
Usage: summarize_launches.py <launches.csv> [header line ...] > profiles/<name>.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows:
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    name = r[ki].split("(")[0] if r[ki].startswith("ecne::") else r[ki][:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
for h in sys.argv[2:]:
    print("# " + h)
print(f"# total {total / 1e6:.3f} ms over {len(rows)} launches")
print("kernel,launches,total_us,share_pct,avg_us")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{name}\",{n},{ns / 1e3:.1f},{100 * ns / total:.1f},{ns / 1e3 / n:.2f}")
