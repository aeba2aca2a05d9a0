#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (profiles/r*_launches.txt).

  python tools/ncu_launches.py gpurun_out/launches.csv "command that was profiled" > profiles/r1_launches.txt
"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
    key = (r["Kernel Name"], r.get("Grid Size", ""))
    agg.setdefault(key, []).append(us)
total = sum(sum(v) for v in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# (cold-cache, serialised per-launch times: compare shares, not absolutes)")
print("# kernel | grid | launches | total us | share | min us | max us")
for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:140]} | {g} | {len(v)} | {sum(v):.1f} | {100 * sum(v) / total:.1f}% | {min(v):.1f} | {max(v):.1f}")
