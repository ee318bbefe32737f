#!/usr/bin/env python3
"""Device time of every pass of the path on one GPU: python scripts/time_passes.py [cells] [model]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goal_b200
from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = sys.argv[2] if len(sys.argv) > 2 else "J2"
t0 = time.time(); co, cn = kuhn_cube(cells); f = fields(co, len(cn)); t1 = time.time()
a = goal_b200.Assembler(co, cn, model, [MATERIAL]); t2 = time.time()
for kv in filter(None, os.environ.get("GX_OPTS", "").split(",")):  # library options k=v[,k=v]
    k, v = kv.split("="); a.set_option(k, int(v))
a.set_solution(f["u"], f["p"])
if model == "J2":
    a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
out = {"cells": cells, "elements": a.ne, "nodes": a.nn, "nnz": a.nnz, "mesh_s": t1 - t0, "gx_create_s": t2 - t1, "colours": a.num_colors}
stages = {}
def timed(fn, reps=5, name=None):
    fn(); ts = []; st = []
    for _ in range(reps):
        fn(); t = a.last_timing(); ts.append(t["zero_ms"] + t["assemble_ms"]); s = a.last_stage_timing(); st.append((s["element_ms"], s["gather_ms"]))
    if name: stages[name] = [round(float(np.median([x[i] for x in st])), 4) for i in (0, 1)]
    return float(np.median(ts))
out["stages"] = stages  # per pass: [element kernel ms, gather kernel(s) ms]
out["jacobian_primal_save_ms"] = timed(lambda: a.jacobian(goal_b200.PRIMAL, save=True, out=False), name="jacobian_primal_save")
if len(sys.argv) > 3 and sys.argv[3] == "jac":  # Jacobian pass only
    out["jacobian_primal_save_Melem_s"] = a.ne / out["jacobian_primal_save_ms"] / 1e3
    print(json.dumps(out)); sys.exit(0)
out["jacobian_adjoint_nosave_ms"] = timed(lambda: a.jacobian(goal_b200.ADJOINT, save=False, out=False), name="jacobian_adjoint_nosave")
out["residual_save_ms"] = timed(lambda: a.residual(save=True, out=False), name="residual_save")
L = a.L
import ctypes as C
zu, zp, zc = [np.ascontiguousarray(f[k]) for k in ("zu_diff", "zp_diff", "zp_coarse")]
out["localize_ms"] = timed(lambda: a.localize(zu, zp, zc), name="localize")
t = time.time(); R = a.localize(zu, zp, zc).reshape(-1, 4).copy(); eta, _, b = a.element_error(R[:, :3], R[:, 3]); out["localize+element_error_wall_s"] = time.time() - t
for k in list(out):
    if k.endswith("_ms"): out[k.replace("_ms", "_Melem_s")] = a.ne / out[k] / 1e3
print(json.dumps(out))
