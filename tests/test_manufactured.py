"""BASELINE.json configs[1]: manufactured-solution convergence on uniformly refined tet cubes.

The reference's case is example/manufactured/elastic_primal_2D.yaml: Dirichlet data u = (x^2, y^2) on the whole
boundary, the body force that makes it the exact solution (`body force: elastic squared`, src/goal_bforce.cpp:51-68),
functional "avg disp".  `elastic` and 2D are outside the path; this is its 3D analogue on the path's own
model: mixed u/p neo-Hookean WITH stabilization on Kuhn cubes N = 4, 8, 16, exact solution a smooth finite-strain
field u*, p* = kappa/2 (J* - 1/J*), body force b = -Div P(u*, p*) (evaluated at the element integration points, like
the reference does), u = u* on the whole boundary.  P1/P1 with the pressure-Laplacian stabilization (tau = c0 h^2 / 2 mu,
goal_stabilization.cpp:72) approaches O(h^2) in the L2 norm of u from below -- the stabilization's natural boundary
condition on p costs part of the rate on coarse meshes: measured 1.33 (N 4 -> 8) and 1.75 (8 -> 16) with the oracle.

CPU: the oracle driver (go_apply_bforce restatement checked against a numpy one);  GPU: the same Newton history with
every assembly / boundary / body-force step on the device -- the two must agree to solver precision, and converge."""
import numpy as np
import pytest

from goal_b200.synthetic import MATERIAL, kuhn_cube

AMP = 0.05
MAT = (1000.0, 0.25, 100.0, 10.0, 1.0)


def u_exact(x):
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    return AMP * np.stack([np.sin(1.3 * X + 0.7 * Y) * np.cos(0.9 * Z), X * Y + np.sin(1.1 * Z) * 0.5, np.cos(X - 0.8 * Y + 0.6 * Z)], -1)


def grad_u_exact(x, h=1e-6):
    g = np.zeros(x.shape[:-1] + (3, 3))
    for j in range(3):
        d = np.zeros(3); d[j] = h
        g[..., :, j] = (u_exact(x + d) - u_exact(x - d)) / (2 * h)
    return g


def first_pk(x):
    """P(X) of the exact solution: sigma = mu J^(-5/3) dev(F F^T) + p* I, p* = kappa/2 (J - 1/J) (goal_neohookean.cpp:60-72
    with Mixed at the exact pressure, goal_mixed.cpp:34-46), P = J sigma F^-T (goal_neohookean.cpp:80-86)."""
    E, nu = MAT[0], MAT[1]
    kappa, mu = E / (3 * (1 - 2 * nu)), E / (2 * (1 + nu))
    F = np.eye(3) + grad_u_exact(x)
    J = np.linalg.det(F)
    b = F @ np.swapaxes(F, -1, -2)
    devb = b - (np.trace(b, axis1=-2, axis2=-1) / 3)[..., None, None] * np.eye(3)
    sig = mu * (J ** (-5.0 / 3.0))[..., None, None] * devb + (0.5 * kappa * (J - 1 / J))[..., None, None] * np.eye(3)
    return J[..., None, None] * sig @ np.swapaxes(np.linalg.inv(F), -1, -2)


def body_force(x, t=1.0, h=1e-4):
    """b = -Div P, (Div P)_i = dP_ij / dX_j (central differences of the analytic P)"""
    div = np.zeros(x.shape)
    for j in range(3):
        d = np.zeros(3); d[j] = h
        div += (first_pk(x + d)[..., :, j] - first_pk(x - d)[..., :, j]) / (2 * h)
    return -div


def _solve(asm, co, cn, device_bcs=False):
    from oracle import driver
    n1 = round(len(co) ** (1 / 3)) - 1
    on_b = np.nonzero(np.any((co < 1e-12) | (co > 1 - 1e-12), axis=1))[0]
    ue = u_exact(co)
    dbcs = [(eq, on_b[k:k + 1], (lambda t, v=ue[on_b[k], eq]: v)) for eq in range(3) for k in range(len(on_b))]
    # one dbc entry per (eq, node) carries that node's own boundary value
    r = driver.run_primal(asm, co, dbcs, (), num_steps=1, max_iters=10, bforce=lambda c, t: body_force(c), conn=cn, device_bcs=device_bcs)
    err = r["u"] - ue
    x = co[cn]
    vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6.0
    l2 = np.sqrt((vol[:, None] * (err[cn] ** 2).sum(-1)).sum() / 4.0)  # nodal (lumped) L2 norm of the error
    return dict(l2=l2, J=r["J"][-1], newton=r["newton"][-1], u=r["u"], p=r["p"], n=n1)


def test_oracle_bforce_matches_numpy(cube):
    from oracle import driver
    from oracle.oracle import Oracle
    co, cn = cube["coords"], cube["tets"]
    o = Oracle(co, cn, "neohookean", [MATERIAL])
    b = np.random.RandomState(0).randn(len(cn), 3)
    R = o.apply_bforce(b, np.zeros(4 * len(co)))
    Rn = driver._bforce_rhs(co, cn, b)
    assert np.abs(R - Rn).max() < 1e-14 * np.abs(Rn).max()
    zu = np.random.RandomState(1).randn(len(co), 3)
    Rz = o.apply_bforce(b, np.zeros(4 * len(co)), zu_diff=zu)
    assert np.abs(o.apply_bforce(b, np.zeros(4 * len(co)), zu_diff=np.ones((len(co), 3))) - R).max() < 1e-15  # z = 1: the plain weights
    x = co[cn]; vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6.0
    want = -(b * zu[cn].mean(1) * vol[:, None]).sum(0)  # partition of unity: sum over nodes = -int b_i z_i
    assert np.allclose(Rz.reshape(-1, 4)[:, :3].sum(0), want, rtol=1e-12, atol=1e-15)


def test_manufactured_convergence_oracle():
    """N = 4, 8 (16 is left to the GPU test: the sparse direct solves there cost half a minute)."""
    from oracle.oracle import Oracle
    out = []
    for n in (4, 8):
        co, cn = kuhn_cube(n)
        out.append(_solve(Oracle(co, cn, "neohookean", [MAT]), co, cn))
    assert out[0]["newton"] <= 6 and out[1]["newton"] <= 6
    rate = np.log2(out[0]["l2"] / out[1]["l2"])
    assert rate > 1.2, (rate, [o["l2"] for o in out])
    Jex = _functional_exact()
    assert abs(out[1]["J"] - Jex) < abs(out[0]["J"] - Jex)


def _functional_exact(n=48):
    """avg disp of the exact solution: int sum_i u_i dV / 3 (goal_avg_disp.cpp:17-21), midpoint rule on a fine grid"""
    g = (np.arange(n) + 0.5) / n
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return u_exact(np.stack([X, Y, Z], -1)).sum(-1).mean() / 3.0


@pytest.mark.gpu
def test_manufactured_convergence_on_the_device():
    import goal_b200
    from oracle.oracle import Oracle
    out = []
    for n in (4, 8, 16):
        co, cn = kuhn_cube(n)
        a = goal_b200.Assembler(co, cn, "neohookean", [MAT])
        g = _solve(a, co, cn, device_bcs=True)
        a.close()
        if n <= 8:  # the oracle's Newton history on the same mesh: same solution to solver precision
            o = _solve(Oracle(co, cn, "neohookean", [MAT]), co, cn)
            assert g["newton"] == o["newton"]
            assert np.abs(g["u"] - o["u"]).max() < 1e-9 * np.abs(o["u"]).max() and abs(g["J"] - o["J"]) < 1e-10 * abs(o["J"])
        out.append(g)
    r1, r2 = np.log2(out[0]["l2"] / out[1]["l2"]), np.log2(out[1]["l2"] / out[2]["l2"])
    assert r1 > 1.2 and r2 > 1.6 and r2 > r1, (r1, r2, [o["l2"] for o in out])
    Jex = _functional_exact()
    eJ = [abs(o["J"] - Jex) for o in out]
    assert eJ[2] < eJ[1] < eJ[0] and np.log2(eJ[1] / eJ[2]) > 1.0, eJ
