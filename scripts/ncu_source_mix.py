#!/usr/bin/env python
"""Executed-instruction mix and stall samples per opcode from an ncu source page:
ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME | python scripts/ncu_source_mix.py [units]
`units` = number of warp-level work items (e.g. warp-faces) to normalise by."""
import collections
import csv
import re
import sys

units = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter()
st = collections.Counter()
tot_ex = tot_s = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    op = src.split()[0].split(".")[0] if src else "?"
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    ex[op] += n
    st[op] += s
    tot_ex += n
    tot_s += s
print(f"total warp instr {tot_ex}  per unit {tot_ex/units:.1f}; samples {tot_s}")
fp64 = sum(v for k, v in ex.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe per unit {fp64/units:.1f}")
for k, v in ex.most_common(28):
    print(f"  {k:8s} {v/units:8.1f} /unit   samples {100.0*st[k]/max(tot_s,1):5.1f}%")
