#!/usr/bin/env python
"""Opcode mix weighted by executed count from `ncu --page source --csv` (SASS view).
usage: sass_mix.py file.csv [n_scores]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = collections.Counter(); smp = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] if not op.startswith("MUFU") else op
    try: n = int(r[iex])
    except: continue
    tot[op] += n
    try: smp[op] += int(r[ismp])
    except: pass
N = sum(tot.values()); S = sum(smp.values())
scale = float(sys.argv[2]) / 32 if len(sys.argv) > 2 else None
print("total warp-instr", N, "samples", S)
for op, n in tot.most_common(40):
    print(f"{op:14s} {n:12d} {100*n/N:6.2f}%  samples {100*smp[op]/max(S,1):6.2f}%" + (f"  per-score {n/scale:6.3f}" if scale else ""))
