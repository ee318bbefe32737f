#!/usr/bin/env python3
"""Per-CUDA-source-line stall samples / executed instructions of every kernel in an ncu source page:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [top]
(what profiles/*_hot_lines.txt hold; inlined callees are counted at the call site too)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, cur, fname, hdr = {}, None, "", None
for r in rows:
    if not r:
        continue
    if r[0] in ("Kernel Name", "Function Name") and len(r) > 1:
        cur = r[1]; continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] and r[0].isdigit() and hdr and "# Samples" in hdr:
        iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try:
            out.setdefault(cur, []).append((int(float(r[iS] or 0)), int(float(r[iE] or 0)), fname, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
for k, o in out.items():
    ts, te = sum(x[0] for x in o), sum(x[1] for x in o)
    print(f"{k}\nstall samples {ts}, executed instructions {te}")
    for s, e, fn, l, src in sorted(o, reverse=True)[:top]:
        print(f"{100 * s / max(ts, 1):5.2f}% smp {100 * e / max(te, 1):5.2f}% inst  {fn}:{l}  {src}")
