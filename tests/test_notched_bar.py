"""BASELINE.json configs[2]: J2 elastoplastic multi-step load history on a notched tet mesh (history state carried
through the assembly).

The reference ships the geometry generator (example/primal/notch/notch.cpp:7-20: the unit plate with a quarter-circle
notch of radius 0.2 at the origin corner, extruded thin) and the 2D load cases example/restart/notch2D_*.yaml
(u_x = 0 on xmin, u_y = 0 on ymin, u_x = 0.01 t on xmax, J2 with the usual material, Newton tolerance 1e-8); the
load-step loop is Solver::solve (src/main_primal.cpp:68-90) with States::update (src/goal_states.cpp:130-141) after
every step.  Here: the same plate as a Kuhn box of N x N x 1 cells with the cells inside the notch radius removed,
the same boundary conditions plus u_z = 0 on zmin, and an 8-step history that loads into the plastic range, unloads
(elements that yielded go back to the elastic branch with Fp_old != I) and reloads past the previous maximum.

Every call the reference's driver makes on the assembler -- compute_jacob and compute_resid of every Newton iterate
with save_state = true, States::update at the end of every step -- is recorded from the oracle's run and replayed on
the other implementation with ITS OWN carried history; R (and the CRS values) must agree to 1e-12 at every call and
sigma, eqps, Fp to 1e-10 at the end of every step."""
import ctypes as C

import numpy as np
import pytest

from conftest import relerr
from goal_b200.synthetic import MATERIAL, kuhn_block

LOAD = [0.004, 0.008, 0.012, 0.016, 0.010, 0.004, 0.012, 0.020]  # u_x on xmax at the end of step k+1


def notched_plate(n=10, radius=0.2):
    """Kuhn box [0,1] x [0,1] x [0,1/n] (n x n x 1 cells) without the cells whose centre lies inside the notch."""
    co, cn = kuhn_block(n, n, 1, (0, 0, 0), n)
    cen = co[cn].mean(axis=1)
    cell = (np.arange(len(cn)) // 6)
    ccen = np.zeros((cell.max() + 1, 3))
    np.add.at(ccen, cell, cen / 6.0)
    keep_cell = np.hypot(ccen[:, 0], ccen[:, 1]) > radius
    cn = cn[keep_cell[cell]]
    used = np.unique(cn)
    new = -np.ones(len(co), dtype=np.int64); new[used] = np.arange(len(used))
    return np.ascontiguousarray(co[used]), np.ascontiguousarray(new[cn].astype(np.int32))


def _bcs(co):
    amp = lambda t: float(np.interp(t, np.arange(len(LOAD) + 1), [0.0] + LOAD))
    z = lambda t: 0.0
    sel = lambda m: np.nonzero(m)[0]
    return [(0, sel(co[:, 0] < 1e-12), z), (1, sel(co[:, 1] < 1e-12), z), (2, sel(co[:, 2] < 1e-12), z),
            (0, sel(co[:, 0] > 1 - 1e-12), amp)]


class Recorder:
    """Forwards the driver's calls to an assembler with the Oracle interface and keeps what went in and came out."""

    def __init__(self, asm):
        self.asm, self.log, self.u, self.p = asm, [], None, None
        self.rowptr, self.colind = asm.rowptr, asm.colind

    def set_solution(self, u, p):
        self.u, self.p = np.array(u, copy=True), np.array(p, copy=True)
        self.asm.set_solution(u, p)

    def jacobian(self, save=True, **kw):
        R, V = self.asm.jacobian(1, save=save)
        self.log.append(("jac", self.u, self.p, np.array(R, copy=True), np.array(V, copy=True), self.asm.plastic_count()))
        return R, V

    def residual(self, save=True, **kw):
        R = self.asm.residual(save=save)
        self.log.append(("res", self.u, self.p, np.array(R, copy=True), None, self.asm.plastic_count()))
        return R

    def update_states(self):
        st = {k: np.array(self.asm.state(k), copy=True) for k in ("sigma", "eqps", "Fp")}
        self.log.append(("update", st))
        self.asm.update_states()

    def plastic_count(self):
        return self.asm.plastic_count()

    def avg_disp(self):
        return self.asm.avg_disp()

    def state(self, k):
        return self.asm.state(k)


@pytest.fixture(scope="module")
def history():
    from oracle import driver
    from oracle.oracle import Oracle
    co, cn = notched_plate()
    rec = Recorder(Oracle(co, cn, "J2", [MATERIAL]))
    r = driver.run_primal(rec, co, _bcs(co), (), num_steps=len(LOAD), max_iters=8)
    return co, cn, rec.log, r


def test_history_exercises_both_branches(history):
    co, cn, log, r = history
    ne = len(cn)
    assert len(cn) == 6 * 100 - 6 * 3 and max(r["newton"]) <= 8  # three cells have their centre inside the notch radius
    pl = r["plastic"]  # plastic elements at the end of every step
    assert pl[0] == 0                      # step 1 is elastic everywhere
    assert 0 < pl[3] < ne                  # at the first maximum the notch root has yielded, the far field has not
    assert pl[4] == 0 and pl[5] == 0       # unloading is elastic -- with Fp_old != I in the elements that yielded
    assert pl[7] > pl[3]                   # reloading past the previous maximum spreads the plastic zone
    ups = [e[1] for e in log if e[0] == "update"]
    eq = np.array([u["eqps"] for u in ups])
    assert np.all(np.diff(eq, axis=0) >= -1e-15) and eq[5].max() == eq[3].max() > 0  # eqps never decreases; frozen while unloading
    dFp = np.abs(ups[5]["Fp"] - np.eye(3).reshape(-1)).max(axis=1)
    # Fp moved wherever the element yielded -- and in more elements: an element that is plastic at an intermediate Newton
    # iterate and elastic at the converged one keeps the iterate's Fp, because the elastic branch does not write Fp
    # (goal_J2.cpp:135-136; SURVEY.md 8a trap 1).  The replay tests below must reproduce exactly that.
    assert np.all(dFp[eq[3] > 0] > 1e-6) and (dFp > 1e-6).sum() > (eq[3] > 0).sum()


def _replay(log, jac, res, update, states, check_values=True):
    n_calls = 0
    for entry in log:
        if entry[0] == "update":
            st = states()
            assert relerr(st["sigma"], entry[1]["sigma"]) < 1e-10
            assert np.abs(st["eqps"] - entry[1]["eqps"]).max() < 1e-10
            assert np.abs(st["Fp"] - entry[1]["Fp"]).max() < 1e-10
            update()
            continue
        kind, u, p, R, V, npl = entry
        if kind == "jac":
            Rg, Vg, plg = jac(u, p)
            if check_values:
                assert relerr(Vg, V) < 1e-12
        else:
            Rg, plg = res(u, p)
        assert relerr(Rg, R) < 1e-12 and plg == npl, (kind, n_calls)
        n_calls += 1
    return n_calls


def test_history_replayed_by_the_device_element_code_on_the_cpu(history, hostcheck):
    """The CUDA path's element code (host build, coloured schedule) through the whole recorded history."""
    co, cn, log, r = history
    nn, ne = len(co), len(cn)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    mat = np.array(MATERIAL)
    st = dict(sigma=np.zeros((ne, 9)), eqps=np.zeros(ne), eqps_old=np.zeros(ne), Fp=np.tile(np.eye(3).reshape(-1), (ne, 1)),
              Fp_old=np.tile(np.eye(3).reshape(-1), (ne, 1)))
    nnz, ncol = C.c_int64(0), C.c_int32(0)
    nul = C.POINTER(C.c_double)()
    hostcheck.hc_assemble(1, 0, 0, nn, ne, ip(cn), dp(co), dp(mat), nul, nul, nul, nul, nul, nul, nul, nul, nul, nul, C.byref(nnz), None, None,
                          C.byref(ncol), C.byref(C.c_int64(0)))

    def run(pass_, u, p):
        R, V, npl = np.zeros(4 * nn), np.zeros(nnz.value), C.c_int64(0)
        rc = hostcheck.hc_assemble(1, pass_, 1, nn, ne, ip(cn), dp(co), dp(mat), dp(np.ascontiguousarray(u)), dp(np.ascontiguousarray(p)), nul,
                                   dp(st["sigma"]), dp(st["eqps"]), dp(st["eqps_old"]), dp(st["Fp"]), dp(st["Fp_old"]), dp(R), dp(V),
                                   C.byref(nnz), None, None, C.byref(ncol), C.byref(npl))
        assert rc == 0
        return R, V, npl.value

    def update():
        st["eqps_old"][:] = st["eqps"]; st["Fp_old"][:] = st["Fp"]

    n = _replay(log, lambda u, p: run(1, u, p), lambda u, p: (lambda o: (o[0], o[2]))(run(0, u, p)), update,
                lambda: dict(sigma=st["sigma"], eqps=st["eqps"], Fp=st["Fp"]))
    assert n == len([e for e in log if e[0] != "update"]) >= 2 * len(LOAD)


@pytest.mark.gpu
def test_history_replayed_on_the_device(history):
    import goal_b200
    co, cn, log, r = history
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])

    def jac(u, p):
        a.set_solution(u, p)
        R, V = a.jacobian(goal_b200.PRIMAL, save=True)
        return R.copy(), V.copy(), a.plastic_count()

    def res(u, p):
        a.set_solution(u, p)
        return a.residual(save=True).copy(), a.plastic_count()

    n = _replay(log, jac, res, a.update_states, lambda: {k: a.get_state(k) for k in ("sigma", "eqps", "Fp")})
    assert n >= 2 * len(LOAD)
    a.close()
    # and the whole load history solved with every assembly / boundary step on the device: same functional, same Newton counts
    from oracle import driver
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])
    g = driver.run_primal(a, co, _bcs(co), (), num_steps=len(LOAD), max_iters=8, device_bcs=True)
    assert g["newton"] == r["newton"] and g["plastic"] == r["plastic"]
    assert np.abs(np.array(g["J"]) - np.array(r["J"])).max() < 1e-10 * np.abs(r["J"]).max()
    a.close()
