#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: summarize_launches.py in.csv [header text]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# per-launch times are cold-cache / serialised under ncu: compare SHARES, not absolutes")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} n={n:4d} total={t:10.1f}us avg={t / n:8.1f}us share={100 * t / tot:5.1f}%")
print(f"total {tot:.1f} us")
