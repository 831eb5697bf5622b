#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel count / total / share of the step."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    rows.append((name, ns))
tot = sum(ns for _, ns in rows) or 1.0
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    agg[name][0] += 1
    agg[name][1] += ns
print("kernel,launches,total_ms,share_pct,avg_us")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{ns / 1e6:.3f},{100 * ns / tot:.2f},{ns / n / 1e3:.1f}")
print(f"TOTAL,{len(rows)},{tot / 1e6:.3f},100.00,")
