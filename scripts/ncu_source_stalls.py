#!/usr/bin/env python
"""Stall-reason totals and the hottest instructions of a kernel from an ncu source page:
ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME | python scripts/ncu_source_stalls.py [top]"""
import collections
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
recs = []
seen = set()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit():
        continue
    if r[0] in seen:      # the page repeats per launch of the same kernel: keep the first
        continue
    seen.add(r[0])
    n = int(r[ix["# Samples"]])
    per = {c: int(r[ix[c]] or 0) for c in stall_cols}
    for c, v in per.items():
        tot[c] += v
    recs.append((n, r[ix["Source"]].strip(), per, int(r[ix["Instructions Executed"]] or 0)))
s = sum(tot.values())
print("stall totals:", ", ".join(f"{c[6:]} {100.0*v/s:.1f}%" for c, v in tot.most_common(10)))
recs.sort(key=lambda x: -x[0])
allsamp = sum(r[0] for r in recs)
for n, src, per, ex in recs[:top]:
    why = ", ".join(f"{c[6:]} {v}" for c, v in sorted(per.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100.0*n/allsamp:5.2f}%  ex {ex:9d}  {src[:70]:70s} {why}")
