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
by_count = collections.Counter()       # executed count of an instruction -> (how many instructions, total executed)
by_count_n = collections.Counter()
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
    by_count[n] += n
    by_count_n[n] += 1
    tot_ex += n
    tot_s += s
print(f"total warp instr {tot_ex}  per unit {tot_ex/units:.1f}; samples {tot_s}")
fp64 = sum(v for k, v in ex.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe per unit {fp64/units:.1f}")
for k, v in ex.most_common(28):
    print(f"  {k:8s} {v/units:8.1f} /unit   samples {100.0*st[k]/max(tot_s,1):5.1f}%")
# where the instructions are executed: static instructions grouped by how often each one ran (hot loop = the units'
# count; rarer groups are prologue, row set-up, the boundary-cell path)
print("executed-count groups (count per instruction: static instructions, share of all executed):")
for n, tot in sorted(by_count.items(), key=lambda kv: -kv[1])[:14]:
    print(f"  {n:12d} x {by_count_n[n]:5d} instr  = {tot/units:8.1f} /unit  {100.0*tot/max(tot_ex,1):5.1f}%")
