#!/usr/bin/env python3
"""Per-repetition stage times of one pass (spread, not just the median): python scripts/time_reps.py [cells] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goal_b200
from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
co, cn = kuhn_cube(cells); f = fields(co, len(cn))
a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])
for kv in filter(None, os.environ.get("GX_OPTS", "").split(",")):
    k, v = kv.split("="); a.set_option(k, int(v))
a.set_solution(f["u"], f["p"]); a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
def run(fn):
    out = []
    for _ in range(reps):
        fn(); s = a.last_stage_timing(); out.append(round(s["element_ms"], 3))
    return out
print(json.dumps({"opts": os.environ.get("GX_OPTS", ""),
                  "adjoint_nosave_A": run(lambda: a.jacobian(goal_b200.ADJOINT, save=False, out=False)),
                  "primal_nosave_A": run(lambda: a.jacobian(goal_b200.PRIMAL, save=False, out=False)),
                  "primal_save_A": run(lambda: a.jacobian(goal_b200.PRIMAL, save=True, out=False)),
                  "adjoint_nosave_A_again": run(lambda: a.jacobian(goal_b200.ADJOINT, save=False, out=False))}))
