#!/usr/bin/env python3
"""Histogram of executed SASS opcodes / stall samples from `ncu --page source --csv`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iSamp, iExec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iExec or r[iS] == 'Source' or r[0] == 'Kernel Name': continue
    t = r[iS].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    n = int(float(r[iExec] or 0)); s = int(float(r[iSamp] or 0))
    ops[op] += n; samp[op] += s; tot += n
print("total warp-instructions executed:", tot, " static instructions:", len(rows) - 2)
S = sum(samp.values())
for op, n in ops.most_common(28):
    print(f"{op:10s} {n:12d} {100*n/tot:6.2f}%   samples {100*samp[op]/max(S,1):6.2f}%")
