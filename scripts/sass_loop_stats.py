#!/usr/bin/env python
"""Instruction mix of the hot loop of a kernel: scripts/sass_loop_stats.py <sass dump> <function substring>
(`cuobjdump -sass lib.so > dump`).  The hot loop = the largest backward-branch span of the function."""
import collections
import re
import sys


def main(path, pat):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "Function :" in l and pat in l)
    end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
    ins = []
    for l in lines[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    best = None
    loops = []
    for addr, txt in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?(0x[0-9a-f]+)\)?", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                loops.append((tgt, addr))
                if best is None or addr - tgt > best[1] - best[0]:
                    best = (tgt, addr)
    # --inner: the SMALLEST loop that still holds at least half of the FP64 instructions of the largest one
    if "--inner" in sys.argv and best is not None:
        def nfp(lo, hi):
            return sum(1 for a, t in ins if lo <= a <= hi and re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
                       in ("DFMA", "DMUL", "DADD"))
        ref = nfp(*best)
        cands = [l for l in loops if nfp(*l) * 2 >= ref]
        best = min(cands, key=lambda l: l[1] - l[0])
    print(lines[start].strip(), "total instr", len(ins))
    if best is None:
        print("no loop")
        return
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    cnt = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        cnt[t.split()[0].split(".")[0]] += 1
    fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"loop {best[0]:#x}..{best[1]:#x}: {len(body)} instr, FP64-pipe {fp64}, MUFU {cnt['MUFU']}")
    print("  " + ", ".join(f"{k} {v}" for k, v in cnt.most_common(30)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
