#!/usr/bin/env python3
"""SASS opcode histogram (thread instructions executed) per kernel from an ncu source page:
    ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv ; python scripts/ncu_sass_hist.py sass.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, hdr, H = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] in ("Kernel Name", "Function Name") and len(r) > 1:
        kern = r[1]; continue
    if "Source" in r and "Instructions Executed" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        try:
            src = r[hdr.index("Source")]
            n = int(float(r[hdr.index("Thread Instructions Executed")] or 0)) if "Thread Instructions Executed" in hdr else int(float(r[hdr.index("Instructions Executed")] or 0))
        except ValueError:
            continue
        t = src.split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        H.setdefault(kern, collections.Counter())[op] += n
for k, c in H.items():
    tot = sum(c.values())
    print(f"{k}\n{tot} thread instructions (all captured launches of this kernel)")
    for op, v in c.most_common(top):
        print(f"  {op:10s} {100 * v / tot:5.1f}%")
