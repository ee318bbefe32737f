"""The C++ host mirror of the reference's classes (goal_b200/host/goal_gx.hpp: Disc, SolInfo, States,
Primal::compute_resid / compute_jacob, NestedAdjoint::localize) driven from a C++ program, checked
against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from goal_b200.synthetic import MATERIAL, kuhn_cube

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_cpp_host_mirror_matches_oracle(model):
    from oracle.oracle import PRIMAL, Oracle
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "goal_b200", "host"), "-s"])
    n = 6
    out = subprocess.run([os.path.join(ROOT, "goal_b200", "gx_selftest"), model, str(n)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    def grab(pat):
        return float(re.search(pat, out.stdout).group(1))
    assert "fail() ok: unknown state: no_such_state" in out.stdout
    co, cn = kuhn_cube(n)
    tp = 6.283185307179586
    u = np.stack([0.004 * co[:, 0] + 2e-3 * np.sin(tp * co[:, 1]), -0.0016 * co[:, 1] + 2e-3 * np.sin(tp * co[:, 2]),
                  -0.0016 * co[:, 2] + 2e-3 * np.sin(tp * co[:, 0])], -1)
    p = np.cos(3.0 * co[:, 0] + 2.0 * co[:, 1] - co[:, 2])
    o = Oracle(co, cn, model, [MATERIAL])
    o.set_solution(u, p)
    R, A = o.jacobian(PRIMAL, save=True)
    assert f"nodes {o.nn} elems {o.ne} nnz {o.nnz}" in out.stdout
    assert abs(grab(r"jacob \|R\|\^2 (\S+)") - (R * R).sum()) < 1e-11 * (R * R).sum()
    assert abs(grab(r"jacob \|R\|\^2 \S+ \|A\|\^2 (\S+)") - (A * A).sum()) < 1e-11 * (A * A).sum()
    Rr = o.residual(save=True)
    assert abs(grab(r"resid \|R\|\^2 (\S+)") - (Rr * Rr).sum()) < 1e-11 * (Rr * Rr).sum()
    s = o.state("sigma")
    assert abs(grab(r"sigma \|s\|\^2 (\S+)") - (s * s).sum()) < 1e-9 * (s * s).sum()
    o.update_states()  # the self-test calls States::update() before localising
    Rz = o.residual(save=False)
    assert abs(grab(r"localize\(z=1\) \|R\|\^2 (\S+)") - (Rz * Rz).sum()) < 1e-11 * (Rz * Rz).sum()
    # the steps either side of the assembly through the C++ mirror: Functional, set_tbcs, add_soln, get_iso_target_size
    from oracle import driver
    o.set_solution(u, p)
    Rr = o.residual(save=True)
    Jv, dM = o.functional("avg vm", with_dMdu=True)
    assert abs(grab(r"functional avg_vm (\S+)") - Jv) < 1e-11 * abs(Jv)
    assert abs(grab(r"\|dMdu\|\^2 (\S+)") - (dM * dM).sum()) < 1e-11 * (dM * dM).sum()
    Jk = o.functional("max vm", rho=0.05)
    assert abs(grab(r"max_vm (\S+)") - Jk) < 1e-11 * abs(Jk)
    t = cn[0]
    sides = [[t[1], t[2], t[3]], [t[0], t[3], t[2]], [t[0], t[1], t[3]]]
    Rt = Rr.copy()
    for row, v in driver._traction_rhs(co, sides, (0.3, -1.0, 0.25)):
        Rt[row] += v
    assert abs(grab(r"resid\+tbcs \|R\|\^2 (\S+)") - (Rt * Rt).sum()) < 1e-11 * (Rt * Rt).sum()
    du = 1e-4 * np.sin(0.37 * np.arange(4 * o.nn)).reshape(-1, 4)
    o.set_solution(u + du[:, :3], p + du[:, 3])
    Ru = o.residual(save=False)
    assert abs(grab(r"resid\(u\+du\) \|R\|\^2 (\S+)") - (Ru * Ru).sum()) < 1e-11 * (Ru * Ru).sum()
    eta = 1e-5 * (1.0 + (np.arange(o.ne) % 7))
    vs, _ = o.size_field(eta, 2 * o.ne)
    assert abs(grab(r"size field sum (\S+)") - vs.sum()) < 1e-12 * vs.sum()
