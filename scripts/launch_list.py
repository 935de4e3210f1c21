"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name / grid / block the
number of launches, mean and total duration.   python scripts/launch_list.py launches.csv"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((r["Kernel Name"][:48], r["Grid Size"], r["Block Size"], us))
agg = OrderedDict()
for k, g, b, us in rows:
    a = agg.setdefault((k, g, b), [0, 0.0])
    a[0] += 1
    a[1] += us
print("| kernel | grid | block | launches | mean us | total ms |\n|---|---|---|---|---|---|")
for (k, g, b), (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {g} | {b} | {n} | {tot / n:.2f} | {tot / 1e3:.3f} |")
