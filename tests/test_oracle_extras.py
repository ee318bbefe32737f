"""More pinning of the oracle (test infrastructure) where the reference has no golden of its own:
  * 50-digit mpmath restatement of the two stress updates (src/goal_neohookean.cpp:60-72, src/goal_J2.cpp:72-143,
    src/goal_mixed.cpp:34-46) on single elements -- SURVEY.md 8(c) "oracle acceptance";
  * the functionals' FADT derivative against central differences of their value (src/goal_qoi.cpp:63-76);
  * identities of the traction terms and the size field."""
import numpy as np
import pytest

from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
from oracle.oracle import Oracle

mp = pytest.importorskip("mpmath")


def _mp_state(x, u, p, model, Fp_old, eqps_old):
    """(sigma after Mixed, eqps, Fp or None, plastic) of one linear tet at 50 digits, formulas of the reference."""
    mp.mp.dps = 50
    E, nu, K, Y, _ = [mp.mpf(v) for v in MATERIAL]
    kappa, mu = E / (3 * (1 - 2 * nu)), E / (2 * (1 + nu))
    X = mp.matrix(x.tolist()); U = mp.matrix(u.tolist())
    Jg = mp.matrix(3, 3)
    for i in range(3):
        for j in range(3):
            Jg[i, j] = X[i + 1, j] - X[0, j]
    dN = mp.matrix([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    G = dN * mp.inverse(Jg).T  # G[n, j] = sum_k Jinv[j, k] dN[n, k]
    F = mp.eye(3)
    for i in range(3):
        for j in range(3):
            F[i, j] += sum(U[n, i] * G[n, j] for n in range(4))
    J = mp.det(F)
    pv = sum(mp.mpf(float(v)) for v in p) / 4
    I = mp.eye(3)
    dev = lambda A: A - (A[0, 0] + A[1, 1] + A[2, 2]) / 3 * I
    pr = kappa * (J - 1 / J) / 2
    plastic, eqps, Fp = False, mp.mpf(eqps_old), None
    if model == "neohookean":
        sigma = mu * J ** (mp.mpf(-5) / 3) * dev(F * F.T) + pr * I
    else:
        Fpo = mp.matrix(Fp_old.reshape(3, 3).tolist())
        Fpi = mp.inverse(Fpo)
        be = J ** (mp.mpf(-2) / 3) * F * (Fpi * Fpi.T) * F.T
        s = mu * dev(be)
        mubar = mu * (be[0, 0] + be[1, 1] + be[2, 2]) / 3
        smag = mp.sqrt(sum(s[i, j] ** 2 for i in range(3) for j in range(3)))
        sq23 = mp.sqrt(mp.mpf(2) / 3)
        f = smag - sq23 * (Y + K * eqps)
        if f > mp.mpf("1e-12"):
            plastic = True
            dgam = f / (2 * mubar + 2 * K / 3)  # fixed point of the Newton loop for linear hardening
            N = s / smag
            s = s - 2 * mubar * dgam * N
            eqps = eqps + sq23 * dgam
            Fp = mp.expm(dgam * N) * Fpo
        sigma = s / J + pr * I
    pbar = (sigma[0, 0] + sigma[1, 1] + sigma[2, 2]) / 3
    sigma = sigma + (pv - pbar) * I
    tof = lambda A: np.array([[float(A[i, j]) for j in range(3)] for i in range(3)]).reshape(-1)
    return tof(sigma), float(eqps), None if Fp is None else tof(Fp), plastic


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_stress_update_against_mpmath(model):
    rng = np.random.RandomState(11)
    seen = set()
    for trial in range(12):
        x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]) * 0.1 + 0.01 * rng.randn(4, 3)
        if np.linalg.det(x[1:] - x[0]) < 0:
            x[[1, 2]] = x[[2, 1]]
        u = (0.0002, 0.003)[trial % 2] * rng.randn(4, 3)
        p = rng.randn(4)
        Fpo = (np.eye(3) + 1e-3 * rng.randn(3, 3)).reshape(-1)
        eqo = 0.01 * rng.rand()
        o = Oracle(x, np.array([[0, 1, 2, 3]], dtype=np.int32), model, [MATERIAL])
        o.set_solution(u, p)
        if model == "J2":
            o.state("Fp_old")[:] = Fpo
            o.state("eqps_old")[:] = eqo
            o.state("Fp")[:] = 7.0
        o.residual(save=True)
        sig, eq, Fp, plastic = _mp_state(x, u, p, model, Fpo, eqo)
        seen.add(plastic)
        assert np.abs(o.state("sigma")[0] - sig).max() < 1e-13 * np.abs(sig).max()
        if model == "J2":
            assert plastic == (o.plastic_count() == 1)
            assert abs(o.state("eqps")[0] - eq) < 1e-15
            if plastic:
                assert np.abs(o.state("Fp")[0] - Fp).max() < 1e-14
            else:
                assert np.all(o.state("Fp")[0] == 7.0)  # elastic branch leaves Fp alone (goal_J2.cpp:135-136)
    assert seen == ({False, True} if model == "J2" else {False})


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_functional_derivatives_match_central_differences(model):
    co, cn = kuhn_cube(3)
    f = fields(co, len(cn), strain=0.02)
    es = (np.arange(len(cn)) % 2).astype(np.int32)
    o = Oracle(co, cn, model, [MATERIAL, MATERIAL], elem_set=es)
    o.set_solution(f["u"], f["p"])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    rng = np.random.RandomState(0)
    du, dp, h = rng.randn(*f["u"].shape), rng.randn(*f["p"].shape), 1e-6
    for t in ("avg disp", "avg disp subdomain", "avg vm", "point wise"):
        kw = dict(elem_set=1, point=(5, 1))
        J, d = o.functional(t, with_dMdu=True, **kw)
        assert J == o.functional(t, **kw)  # ST and FADT chains give the same value
        o.set_solution(f["u"] + h * du, f["p"] + h * dp); Jp = o.functional(t, **kw)
        o.set_solution(f["u"] - h * du, f["p"] - h * dp); Jm = o.functional(t, **kw)
        o.set_solution(f["u"], f["p"])
        dd = d.reshape(-1, 4)
        lin = (dd[:, :3] * du).sum() + (dd[:, 3] * dp).sum()
        assert abs((Jp - Jm) / (2 * h) - lin) < 2e-6 * max(abs(lin), 1e-3), t
    # "max vm": value = max + log(scale)/rho from the SAVED stress (goal_ks_vm.cpp:36-87, 102-105), a smooth bound of the max
    o.residual(save=True)
    sg = o.state("sigma").reshape(-1, 3, 3)
    dv = sg - np.trace(sg, axis1=1, axis2=2)[:, None, None] / 3 * np.eye(3)
    vm = np.sqrt(1.5 * (dv * dv).sum((1, 2)))
    x = co[cn]
    vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6
    for rho in (0.05, 2.0):
        Jk = o.functional("max vm", rho=rho)
        want = vm.max() + np.log((np.exp(rho * (vm - vm.max())) * vol).sum()) / rho
        assert abs(Jk - want) < 1e-12 * abs(want)


def test_traction_and_size_field_identities(cube):
    from oracle import driver
    co = cube["coords"]
    sides = cube["side_sets"]["ymax"]
    T = (0.3, -1.0, 0.25)
    rows = driver._traction_rhs(co, sides, T)
    tot = np.zeros(3)
    for row, v in rows:
        tot[row % 4] += v
    area = sum(0.5 * np.linalg.norm(np.cross(co[t[1]] - co[t[0]], co[t[2]] - co[t[0]])) for t in sides)
    assert abs(area - 1.0) < 1e-14 and np.abs(tot + np.array(T) * area).max() < 1e-14  # sum_n N_n = 1
    # inward traction about the face centroid: resultant force vanishes on the flat unit face
    tot = np.zeros(3)
    for row, v in driver._inward_rhs(co, sides, 2.5, (0.5, 1.0, 0.5)):
        tot[row % 4] += v
    assert np.abs(tot).max() < 1e-14
    # size field: uniform indicators and target = number of elements leave the mesh size unchanged
    o = Oracle(co, cube["tets"], "neohookean", [MATERIAL])
    eta = np.full(o.ne, 3e-4)
    v, G = o.size_field(eta, o.ne)
    x = co[cube["tets"]]
    h = np.sqrt(sum(((x[:, i] - x[:, j]) ** 2).sum(1) for i in range(4) for j in range(i + 1, 4)) / 6)
    cnt = np.bincount(cube["tets"].reshape(-1), minlength=o.nn)
    want = np.bincount(cube["tets"].reshape(-1), weights=np.repeat(h, 4), minlength=o.nn) / cnt
    assert abs(G - o.ne * 3e-4 ** 1.2) < 1e-12 * G and np.abs(v - want).max() < 1e-13


# ---------------------------------------------------------------------------
# A16/A17 (error localisation) against an INDEPENDENT restatement.  The reference holds no value for this chain, so
# the oracle's C++ restatement is checked against a second one written here in numpy straight from the reference's
# text -- different code, different data layout, vectorised over elements:
#   DisplacementAdjoint / PressureAdjoint weights   goal_displacement_adjoint.cpp:37-53, goal_pressure_adjoint.cpp:38-49
#       w_n^i = z_i N_n,  d_j w_n^i = (grad z)_ij N_n + z_i d_j N_n   (z, grad z at the integration point)
#   MResidual  R_u[n][i] += P_ij d_j w_n^i w dv                        goal_mresidual.cpp:26-32   (w = u_z_diff)
#   PResidual  R_p[n] += (p/kappa - (J - 1/J)/2) w_n w dv              goal_presidual.cpp:54-59   (w = p_z_diff)
#   Stabilization R_p[n] += tau J Cinv_ij d_i p d_j w_n w dv           goal_stabilization.cpp:59-81 (w = p_z_COARSE,
#       build_error hands "pwc" to Stabilization: goal_mechanics.cpp:214)
#   neo-Hookean + Mixed                                                goal_neohookean.cpp:60-86, goal_mixed.cpp:34-46
# ---------------------------------------------------------------------------
def _numpy_localize_neohookean(co, cn, u, p, zu, zp, zpc, mat=MATERIAL):
    E, nu, _, _, c0 = mat
    kappa, mu = E / (3 * (1 - 2 * nu)), E / (2 * (1 + nu))
    x = co[cn]                                               # [e, n, 3]
    Jg = x[:, 1:] - x[:, :1]                                 # rows x1-x0, x2-x0, x3-x0
    dv = np.linalg.det(Jg)
    Ji = np.linalg.inv(Jg)                                   # d xi / d x
    G = np.concatenate([-Ji.sum(axis=2, keepdims=True).transpose(0, 2, 1), Ji.transpose(0, 2, 1)], axis=1)  # [e, n, 3] = grad N_n
    N = 0.25
    wdv = dv / 6.0                                           # order-1 rule: weight 1/6 at the centroid
    ue, pe, zue, zpe, zce = u[cn], p[cn], zu[cn], zp[cn], zpc[cn]
    F = np.eye(3)[None] + np.einsum("eni,enj->eij", ue, G)
    J = np.linalg.det(F)
    Finv = np.linalg.inv(F)
    b = F @ F.transpose(0, 2, 1)
    devb = b - np.trace(b, axis1=1, axis2=2)[:, None, None] / 3 * np.eye(3)
    sigma = mu * J[:, None, None] ** (-5.0 / 3.0) * devb + 0.5 * kappa * (J - 1 / J)[:, None, None] * np.eye(3)
    pq = pe.sum(1) * N
    sigma = sigma + (pq - np.trace(sigma, axis1=1, axis2=2) / 3)[:, None, None] * np.eye(3)   # Mixed
    P = J[:, None, None] * sigma @ Finv.transpose(0, 2, 1)
    z = zue.sum(1) * N                                       # [e, 3]
    gz = np.einsum("eni,enj->eij", zue, G)                   # (grad z)_ij = d_j z_i
    # MResidual with the adjoint-weighted test functions
    Ru = (np.einsum("eij,eij->ei", P, gz)[:, None, :] * N + np.einsum("eij,ei,enj->eni", P, z, G)) * wdv[:, None, None]
    # PResidual with p_z_diff
    zs = zpe.sum(1) * N
    Rp = ((pq / kappa - 0.5 * (J - 1 / J)) * zs * wdv)[:, None] * N * np.ones((1, 4))
    # Stabilization with p_z_coarse
    h2 = sum(((x[:, a] - x[:, b]) ** 2).sum(1) for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))) / 6.0
    tau = 0.5 * c0 * h2 / mu
    Cinv = np.linalg.inv(F.transpose(0, 2, 1) @ F)
    gp = np.einsum("en,enj->ej", pe, G)
    zc = zce.sum(1) * N
    gzc = np.einsum("en,enj->ej", zce, G)
    gw = gzc[:, None, :] * N + zc[:, None, None] * G         # d_j of the n-th coarse-weighted test function
    Rp = Rp + (tau * J * wdv)[:, None] * np.einsum("eij,ei,enj->en", Cinv, gp, gw)
    R = np.zeros((len(co), 4))
    np.add.at(R[:, :3], cn, Ru)
    np.add.at(R[:, 3], cn, Rp)
    return R.reshape(-1)


def test_error_localisation_against_independent_numpy_restatement(cube):
    co, cn = cube["coords"], cube["tets"]
    f = fields(co, len(cn), strain=0.01)
    o = Oracle(co, cn, "neohookean", [MATERIAL])
    o.set_solution(f["u"], f["p"])
    Ro = o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"])
    Rn = _numpy_localize_neohookean(co, cn, f["u"], f["p"], f["zu_diff"], f["zp_diff"], f["zp_coarse"])
    assert np.abs(Ro - Rn).max() < 1e-12 * np.abs(Rn).max()
    # per equation too: the pressure rows are orders of magnitude below the momentum rows
    for eq in range(4):
        assert np.abs(Ro[eq::4] - Rn[eq::4]).max() < 1e-11 * np.abs(Rn[eq::4]).max()
    # the check has teeth: the same chain with Stabilization weighted by p_z_diff instead of p_z_coarse (pw for pwc,
    # goal_mechanics.cpp:214) or PResidual by p_z_coarse is far outside the tolerance
    Rw = _numpy_localize_neohookean(co, cn, f["u"], f["p"], f["zu_diff"], f["zp_diff"], f["zp_diff"])
    assert np.abs(Ro[3::4] - Rw[3::4]).max() > 1e-3 * np.abs(Rn[3::4]).max()
    Rw = _numpy_localize_neohookean(co, cn, f["u"], f["p"], f["zu_diff"], f["zp_coarse"], f["zp_coarse"])
    assert np.abs(Ro[3::4] - Rw[3::4]).max() > 1e-3 * np.abs(Rn[3::4]).max()
    # partition of unity: the nodal sum equals the directly integrated weighted residual
    tot = Rn.sum()
    assert abs(Ro.sum() - tot) < 1e-12 * np.abs(Rn).sum()
