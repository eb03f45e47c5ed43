#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

src = sys.argv[1]
lines = [l for l in open(src) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
    tot[name][0] += 1
    tot[name][1] += v
T = sum(v[1] for v in tot.values())
print(f"# {src}: {sum(v[0] for v in tot.values())} launches, {T / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:10.3f} ms {100 * v[1] / T:6.2f}%  n={v[0]:5d}  avg={v[1] / v[0] / 1e3:9.1f} us  {k[:120]}")
