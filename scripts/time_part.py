#!/usr/bin/env python3
"""Kernel time of the Jacobian pass on ONE part of a partitioned mesh, on one GPU: the other parts are host-only contexts
that only take part in the structure exchange.  python scripts/time_part.py [cells] [px py pz] [rank]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goal_b200
from goal_b200.partition import block_part
from goal_b200.synthetic import MATERIAL, fields
c = int(sys.argv[1]) if len(sys.argv) > 1 else 128
grid = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 1, 1)
me = int(sys.argv[5]) if len(sys.argv) > 5 else 0
world = grid[0] * grid[1] * grid[2]
parts = [block_part(c, grid, r) for r in range(world)]
A = [goal_b200.Assembler(p["coords"], p["conn"], "J2", [MATERIAL], device=0 if r == me else -1, partition=p) for r, p in enumerate(parts)]
for r, a in enumerate(A):
    for pi in range(a.num_peers):
        q = int(parts[r]["peer_rank"][pi]); pj = list(parts[q]["peer_rank"]).index(r)
        A[q].struct_unpack(pj, a.struct_pack(pi))
for a in A:
    a.struct_finalize()
a, p = A[me], parts[me]
f = fields(p["coords"], len(p["conn"]), node_gid=p["node_gid"], elem_gid=p["elem_gid"])
a.set_solution(f["u"], f["p"]); a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
for kv in filter(None, os.environ.get("GX_OPTS", "").split(",")):
    k, v = kv.split("="); a.set_option(k, int(v))
ts = []
for i in range(8):
    a.jacobian(goal_b200.PRIMAL, save=True, out=False); t = a.last_timing(); ts.append(t["assemble_ms"])
print(json.dumps({"cells": c, "grid": grid, "rank": me, "elements": a.ne, "nnz_ghost": a.nnz, "jacobian_ms": float(np.median(ts[2:])), "launches": t["launches"]}))
