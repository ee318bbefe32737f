#!/usr/bin/env python3
"""Per-CUDA-source-line executed instructions / samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
out = []
fname = ""
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] and r[0].isdigit() and hdr:
        iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try: out.append((int(float(r[iE] or 0)), int(float(r[iS] or 0)), fname, int(r[0]), r[1].strip()[:90]))
        except ValueError: pass
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total inst", tot, "samples", ts)
for e, s, f, l, src in sorted(out, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100*e/tot:5.2f}% inst {100*s/max(ts,1):5.2f}% smp  {f}:{l}  {src}")
