#!/usr/bin/env python3
"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): python scripts/launch_shares.py file.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
agg, cnt = {}, {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"].split("(")[0]
    agg[k] = agg.get(k, 0.0) + float(d["Metric Value"]); cnt[k] = cnt.get(k, 0) + 1
tot = sum(agg.values())
print(f"{len(rows)-1} launches, {tot/1e6:.3f} ms in kernels (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{v/1e6:10.3f} ms {100*v/tot:5.1f}%  x{cnt[k]:<4d} {k}")
