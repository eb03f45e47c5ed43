#!/usr/bin/env python
"""Hottest SASS lines (warp-stall samples) with their top stall reasons from `ncu --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
def num(x):
    try: return int(x)
    except Exception: return 0
seen, allrows = set(), []
for r in rows[2:]:
    if len(r) <= iex or r[ia] in seen or r[ia] == "Address": continue
    seen.add(r[ia]); allrows.append(r)
tot = sum(num(r[ismp]) for r in allrows)
for i, r in enumerate(allrows):
    n = num(r[ismp])
    if tot and n / tot > thr:
        top = sorted(((num(r[c]), h[6:]) for h, c in cols.items()), reverse=True)[:2]
        prev = allrows[i - 1][isrc].strip()[:46] if i else ""
        print(f"{100*n/tot:5.1f}%  {r[isrc].strip()[:66]:66s} {top} | prev: {prev}")
