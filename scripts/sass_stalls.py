#!/usr/bin/env python
"""Warp-stall sample totals by reason from `ncu --page source --csv`.  usage: sass_stalls.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); seen = set()
ia = hdr.index("Address")
for r in rows[2:]:
    if len(r) <= max(cols) or r[ia] in seen: continue
    seen.add(r[ia])
    for c in cols:
        try: tot[hdr[c]] += int(r[c])
        except ValueError: pass
S = sum(tot.values())
for k, v in tot.most_common(): print(f"{k:28s} {v:8d} {100*v/S:6.2f}%")
