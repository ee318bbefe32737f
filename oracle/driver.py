"""Load-stepping Newton driver that mirrors the reference's GoalPrimal loop.

TEST INFRASTRUCTURE.  It reproduces, around any assembler object with the Oracle
interface (oracle.Oracle, or the CUDA path's goal_b200.Assembler), the steps the
reference performs around `assemble`:

  Solver::solve          src/main_primal.cpp:68-90     load-step loop, states->update()
  Primal::solve          src/goal_primal.cpp:111-137   Newton loop
  Primal::compute_jacob  src/goal_primal.cpp:92-109    zero, assemble, tbcs, gather, jac dbcs
  Primal::compute_resid  src/goal_primal.cpp:75-90     zero, assemble, tbcs, gather, resid dbcs
  set_jac_dbcs / set_resid_dbcs   src/goal_dbcs.cpp:39-99
  set_tbcs (apply_bc)    src/goal_tbcs.cpp:29-71        R[row] -= T_d * N_n * w * dv  (= T_d * area / 3)
  Functional (avg disp)  src/goal_functional.cpp:62-70, src/goal_avg_disp.cpp:17-21

The linear solve (Belos GMRES + MueLu in the reference, src/goal_linear_solve.cpp)
is replaced by a sparse direct solve; the reference's linear tolerance is 1e-10.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _traction_rhs(coords, sides, traction):
    """ghost-R contribution of one traction BC: list of (row, value)."""
    out = []
    for tri in sides:
        x = coords[np.asarray(tri)]
        area = 0.5 * np.linalg.norm(np.cross(x[1] - x[0], x[2] - x[0]))
        for n in tri:
            for d in range(3):
                out.append((4 * n + d, -traction[d] * (1.0 / 3.0) * 0.5 * (2.0 * area)))
    return out


def _inward_rhs(coords, sides, scale, center):
    """set_ibcs (src/goal_ibcs.cpp:33-83): T = scale (x_c - center) at the side centroid; list of (row, value)."""
    out = []
    for tri in sides:
        x = coords[np.asarray(tri)]
        area = 0.5 * np.linalg.norm(np.cross(x[1] - x[0], x[2] - x[0]))
        T = (x.sum(0) / 3.0 - np.asarray(center)) * scale
        for n in tri:
            for d in range(3):
                out.append((4 * n + d, -T[d] * (1.0 / 3.0) * 0.5 * (2.0 * area)))
    return out


def _bforce_rhs(coords, conn, b):
    """BForce (src/goal_bforce.cpp:58-68) as a ghost-R increment: R[(n,i)] -= b_e[i] N_n w dv, N_n = 1/4, w dv = vol_e."""
    x = coords[conn]
    vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6.0
    inc = np.zeros(4 * len(coords))
    for n in range(4):
        for i in range(3):
            np.add.at(inc, 4 * conn[:, n] + i, -b[:, i] * 0.25 * vol)
    return inc


def run_primal(asm, coords, dbcs, tbcs=(), num_steps=3, dt=1.0, tol=1e-8, max_iters=5, log=None, device_bcs=False,
               bforce=None, conn=None):
    """dbcs: [(eq, node_ids, g(t))]; tbcs: [(side_tris, T(t) -> 3-vector)]; bforce: b(centroids [Ne,3], t) -> [Ne,3]
    (needs conn), the body force of `mechanics: body force` (src/goal_mechanics.cpp:57-59, goal_bforce.cpp).
    device_bcs: apply the traction terms and the Dirichlet rows with asm.apply_tbcs / asm.apply_dbcs (the CUDA path's
    gx_apply_tbcs / gx_apply_dbcs) on the device-resident result instead of on the host.

    Returns dict(J=[per-step functional], newton=[iterations], plastic=[count at end of step]).
    """
    nn = len(coords)
    u = np.zeros((nn, 3))
    p = np.zeros(nn)
    rowptr, colind = asm.rowptr, asm.colind
    out = dict(J=[], newton=[], plastic=[])
    t_old, t_now = 0.0, dt

    def apply_tbcs(R, t):
        for sides, T in tbcs:
            for row, v in _traction_rhs(coords, sides, T(t)):
                R[row] += v
        if bforce is not None:
            R += _bforce_rhs(coords, conn, bvals(t))

    cen = None if bforce is None else coords[np.asarray(conn)].mean(axis=1)
    bvals = lambda t: np.ascontiguousarray(bforce(cen, t))

    def dbc_rows(t):
        for eq, nodes, g in dbcs:
            for n in nodes:
                yield 4 * n + eq, (u[n, eq] if eq < 3 else p[n]) - g(t)

    for step in range(num_steps):
        it, converged = 1, False
        while it <= max_iters and not converged:
            asm.set_solution(u, p)
            if device_bcs:
                asm.jacobian(save=True, out=False)
                for sides, T in tbcs:
                    asm.apply_tbcs(sides, T(t_now))
                if bforce is not None:
                    asm.apply_bforce(bvals(t_now))
                rows_g = [(4 * n + eq, g(t_now)) for eq, nodes, g in dbcs for n in nodes]
                asm.apply_dbcs([r for r, _ in rows_g], [v for _, v in rows_g], True)
                R, vals = asm.fetch()
            else:
                R, vals = asm.jacobian(save=True)
                R = np.array(R, copy=True)
                vals = np.array(vals, copy=True)
                apply_tbcs(R, t_now)
                for row, r in dbc_rows(t_now):
                    R[row] = r
                    vals[rowptr[row]:rowptr[row + 1]] = 0.0
                    k = rowptr[row] + np.searchsorted(colind[rowptr[row]:rowptr[row + 1]], row)
                    vals[k] = 1.0
            A = sp.csr_matrix((vals, colind, rowptr), shape=(4 * nn, 4 * nn))
            du = spla.spsolve(A.tocsc(), -R).reshape(nn, 4)
            u += du[:, :3]
            p += du[:, 3]
            asm.set_solution(u, p)
            if device_bcs:
                asm.residual(save=True, out=False)
                for sides, T in tbcs:
                    asm.apply_tbcs(sides, T(t_now))
                if bforce is not None:
                    asm.apply_bforce(bvals(t_now))
                rows_g = [(4 * n + eq, g(t_now)) for eq, nodes, g in dbcs for n in nodes]
                asm.apply_dbcs([r for r, _ in rows_g], [v for _, v in rows_g], False)
                R = asm.fetch(values=False)[0]
            else:
                R = np.array(asm.residual(save=True), copy=True)
                apply_tbcs(R, t_now)
                for row, r in dbc_rows(t_now):
                    R[row] = r
            nrm = np.linalg.norm(R)
            if log:
                log(f"step {step + 1} newton {it} ||R|| = {nrm:.3e}")
            converged = nrm < tol
            it += 1
        if not converged:
            raise RuntimeError(f"newton's method failed in {max_iters} iterations")
        out["newton"].append(it - 1)
        out["plastic"].append(asm.plastic_count())
        out["J"].append(asm.avg_disp())
        try:  # the old states this step was solved with (what an error estimate at the end of the step evaluates with)
            get = asm.get_state if hasattr(asm, "get_state") else asm.state
            out["states_old"] = {k: np.array(get(k), copy=True) for k in ("Fp_old", "eqps_old")}
        except Exception:
            out["states_old"] = {}
        asm.update_states()
        t_old, t_now = t_now, t_now + dt
    out["u"], out["p"] = u, p
    return out


def _set_state(asm, name, arr):
    if hasattr(asm, "set_state"):
        asm.set_state(name, arr)
    else:
        asm.state(name)[:] = arr


def run_nested_cycle(asm, nested, u_base, p_base, states_base, dbcs, qoi="avg disp", qoi_kw=None, t_now=1.0, device_bcs=False):
    """One adjoint error-estimation cycle on a nested mesh, the sequence of NestedAdjoint::run
    (src/goal_nested_adjoint.cpp:236-247) around any assembler with the Oracle interface:

      primal state on the nested mesh        P1 fields interpolated, element states inherited from the parent
      compute_adjoint  (:163-180)            transposed Jacobian + dMdu (FADT chain, save=false), jac dbcs
      solve            (:197-215)            dRdu^T z = dMdu; z_fine, z_coarse = set_coarse(z_fine), z_diff; e = -(R . z)
      localize         (:217-234)            error chain weighted with z_diff / z_p-coarse, resid dbcs -> u_error, p_error
      sum_contribs, compute_error, set_error (src/goal_error.cpp:7-56, goal_nested.cpp:395-412)

    asm: assembler built on nested["coords"], nested["tets"].  dbcs: [(eq, node_ids on the nested mesh, g(t))].
    device_bcs: Dirichlet rows through asm.apply_dbcs (gx_apply_dbcs) instead of on the host."""
    from goal_b200.nested import prolong, set_coarse
    nn = len(nested["coords"])
    u, p = prolong(u_base, nested), prolong(p_base, nested)
    asm.set_solution(u, p)
    for name, arr in (states_base or {}).items():
        _set_state(asm, name, np.ascontiguousarray(np.asarray(arr)[nested["parent"]]))
    rows = np.array([4 * n + eq for eq, nodes, g in dbcs for n in nodes], dtype=np.int64)
    gval = np.array([g(t_now) for eq, nodes, g in dbcs for n in nodes])
    sol = np.concatenate([u, p[:, None]], 1).reshape(-1)
    rowptr, colind = asm.rowptr, asm.colind
    kw = dict(qoi_kw or {})
    if device_bcs:
        asm.jacobian(2, save=False, out=False)  # ADJOINT
        J, _ = asm.functional(qoi, with_dMdu=True, **kw)
        asm.apply_dbcs(rows, gval, True)
        R, AT = asm.fetch()
        dMdu = asm.fetch_dMdu()
    else:
        R, AT = [np.array(x, copy=True) for x in asm.jacobian(2, save=False)]
        J, dMdu = asm.functional(qoi, with_dMdu=True, **kw)
        for row, g in zip(rows, gval):  # set_jac_dbcs (src/goal_dbcs.cpp:60-97)
            R[row] = sol[row] - g
            dMdu[row] = 0.0
            AT[rowptr[row]:rowptr[row + 1]] = 0.0
            AT[rowptr[row] + np.searchsorted(colind[rowptr[row]:rowptr[row + 1]], row)] = 1.0
    A = sp.csr_matrix((AT, colind, rowptr), shape=(4 * nn, 4 * nn))
    z = spla.spsolve(A.tocsc(), dMdu)
    zf = z.reshape(nn, 4)
    zc = set_coarse(zf, nested)       # apf::copyData + Nested::set_coarse
    zd = zf - zc                       # NestedAdjoint::subtract
    e_est = -float(R @ z)
    if device_bcs:
        asm.localize(zd[:, :3], zd[:, 3], zc[:, 3])
        asm.apply_dbcs(rows, gval, False)
        Re = asm.fetch(values=False)[0]
    else:
        Re = np.array(asm.localize(zd[:, :3], zd[:, 3], zc[:, 3]), copy=True)
        Re[rows] = sol[rows] - gval    # set_resid_dbcs
    err = Re.reshape(nn, 4)
    n_base = int(nested["parent"].max()) + 1
    eta, eta_parent, bound = asm.element_error(err[:, :3].copy(), err[:, 3].copy(), nested["parent"], n_base)
    return dict(J=J, z=z, e_est=e_est, eta=eta, eta_parent=eta_parent, bound=bound, u=u, p=p)


# The reference's three 3D regression inputs (example/primal/*.yaml), all on
# test/mesh/cube with E=1000, nu=0.25, K=100, Y=10, c0=1, 3 load steps of 1.0.
GOLDEN = {
    # name: (model, golden J, yaml)
    "neohookean_uniaxial_3D": ("neohookean", 2.541285341193943e-03, "example/primal/neohookean_uniaxial_3D.yaml:34-36"),
    "J2_uniaxial_3D": ("J2", 1.073955612775838e-03, "example/primal/J2_uniaxial_3D.yaml:36-38"),
    "J2_traction_3D": ("J2", 2.512233163668167e-04, "example/primal/J2_traction_3D.yaml:37-39"),
}
# survey-time per-step values: intermediate known answers
GOLDEN_STEPS = {
    "neohookean_uniaxial_3D": [8.379491376688085e-04, 1.685073358995205e-03, 2.541285341193988e-03],
    "J2_uniaxial_3D": [8.379491377532382e-04, 9.454391681366886e-04, 1.073955612775823e-03],
    "J2_traction_3D": [8.346891869851834e-05, 1.672096850639751e-04, 2.512233163668262e-04],
}


def golden_case(name, fixture):
    """(dbcs, tbcs) of one reference regression input on the cube fixture."""
    ns, ss = fixture["node_sets"], fixture["side_sets"]
    zero = lambda t: 0.0
    dbcs = [(0, ns["xmin"], zero), (1, ns["ymin"], zero), (2, ns["zmin"], zero)]
    tbcs = []
    if name.endswith("uniaxial_3D"):
        dbcs.append((0, ns["xmax"], lambda t: 0.01 * t))
    else:
        tbcs.append((ss["ymax"], lambda t: (0.0, 1.0 * t, 0.0)))
    return dbcs, tbcs
