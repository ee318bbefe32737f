#!/usr/bin/env python3
"""Operation counts of the kernel's element arithmetic (counting scalar, tests/hostcheck hc_count_ops)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostcheck"), "-s"])
L = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]) * 0.1
mat = np.array([1000.0, 0.25, 100.0, 10.0, 1.0])
Fp = (np.eye(3) + 1e-5 * np.arange(9).reshape(3, 3)).reshape(-1).copy()
p = np.array([0.1, -0.2, 0.3, 0.05])
rows = []
for model, name in ((0, "neohookean"), (1, "J2")):
    for amp, tag in ((1e-5, "elastic"), (3e-3, "plastic")):
        if model == 0 and tag == "plastic":
            continue
        u = amp * np.array([[0, 0, 0], [1, 0.2, 0], [0.3, -1, 0.1], [0, 0.4, 0.7]])
        for what, wname in ((0, "core+save"), (1, "one incidence (row-owner lane)"), (2, "whole element (16 blocks)")):
            out = (C.c_long * 4)(); pl = C.c_int()
            rc = L.hc_count_ops(model, what, 1, dp(x), dp(u), dp(p), dp(mat), dp(Fp), C.c_double(0.001), out, C.byref(pl))
            assert rc == 0
            assert (pl.value == 1) == (tag == "plastic" and model == 1), (name, tag, pl.value)
            add, mul, dv, sp = list(out)
            rows.append((name, tag, wname, add, mul, dv, sp, add + mul + dv + sp))
print(f"{'model':11s} {'branch':8s} {'scope':32s} {'add':>6s} {'mul':>6s} {'div':>4s} {'spec':>5s} {'flops':>6s}")
for r in rows:
    print(f"{r[0]:11s} {r[1]:8s} {r[2]:32s} {r[3]:6d} {r[4]:6d} {r[5]:4d} {r[6]:5d} {r[7]:6d}")
