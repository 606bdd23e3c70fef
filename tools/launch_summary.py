#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
d = collections.defaultdict(list)
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    d[name].append(v)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':58s} {'n':>5s} {'avg us':>9s} {'total ms':>9s} {'share':>6s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:58]:58s} {len(v):5d} {sum(v)/len(v):9.1f} {sum(v)/1000:9.2f} {100*sum(v)/tot:5.1f}%")
