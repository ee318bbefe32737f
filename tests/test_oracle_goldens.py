"""Pins the oracle: the reference's golden functional values, structural known answers,
finite differences and the error-localisation identities of SURVEY.md 8(c)."""
import numpy as np
import pytest

from conftest import relerr
from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
from oracle import driver
from oracle.oracle import ADJOINT, PRIMAL, Oracle


@pytest.mark.parametrize("name", list(driver.GOLDEN))
def test_reference_golden_J(cube, name):
    """example/primal/*_3D.yaml regression values, reference tolerance 1e-8 (src/goal_regression.cpp:14-23);
    we hold the oracle to 1e-13 absolute (13 significant digits)."""
    model, J_gold, _ = driver.GOLDEN[name]
    o = Oracle(cube["coords"], cube["tets"], model, [MATERIAL])
    dbcs, tbcs = driver.golden_case(name, cube)
    r = driver.run_primal(o, cube["coords"], dbcs, tbcs)
    assert abs(r["J"][-1] - J_gold) < 1e-13
    for got, want in zip(r["J"], driver.GOLDEN_STEPS[name]):
        assert abs(got - want) < 1e-13
    assert r["newton"] == ([2, 2, 2] if name == "J2_traction_3D" else [3, 3, 3])
    if name == "J2_uniaxial_3D":
        assert r["plastic"] == [0, 132, 132]  # the only golden that pins the plastic branch
    if name == "J2_traction_3D":
        assert r["plastic"] == [0, 0, 0]


def test_graph_known_answers(cube):
    """nnz = 16 (Nn + 2 N_edges): cube fixture 8,176; Kuhn cube N_edges = 3N(N+1)^2 + 3N^2(N+1) + N^3."""
    o = Oracle(cube["coords"], cube["tets"], "neohookean", [MATERIAL])
    assert o.nnz == 8176 and len(o.rowptr) == 205
    for n in (3, 5):
        co, cn = kuhn_cube(n)
        edges = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
        ok = Oracle(co, cn, "neohookean", [MATERIAL])
        assert ok.nnz == 16 * ((n + 1) ** 3 + 2 * edges)
        for r in range(0, 4 * ok.nn, 7):
            row = ok.colind[ok.rowptr[r]:ok.rowptr[r + 1]]
            assert np.all(np.diff(row) > 0)


def test_kuhn_mesh_is_valid():
    co, cn = kuhn_cube(4)
    x = co[cn]
    vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6
    assert vol.min() > 0 and abs(vol.sum() - 1) < 1e-13
    assert len(cn) == 6 * 64 and len(co) == 125


@pytest.mark.parametrize("model,scale", [("neohookean", 1.0), ("J2", 1.0), ("J2", 0.05)])
def test_jacobian_matches_finite_differences(model, scale):
    """scale 1.0 puts every J2 element on the plastic branch, 0.05 every element on the elastic one
    (a mixed state is not differentiable across the yield surface, so FD is checked per branch)."""
    co, cn = kuhn_cube(3)
    f = fields(co, len(cn), strain=0.02)
    f["u"] = scale * f["u"]
    o = Oracle(co, cn, model, [MATERIAL])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    o.set_solution(f["u"], f["p"])
    R, vals = o.jacobian(PRIMAL, save=False)
    if model == "J2":
        assert o.plastic_count() == (o.ne if scale == 1.0 else 0)
    A = o.csr(vals.copy())
    _, valsT = o.jacobian(ADJOINT, save=False)
    assert abs(o.csr(valsT) - A.T).max() < 1e-12 * abs(A).max()  # scatter_adjoint is the transpose
    assert relerr(o.residual(save=False), R) < 1e-14
    d = np.random.RandomState(0).randn(4 * o.nn)
    h = 1e-6
    du, dpp = d.reshape(-1, 4)[:, :3], d.reshape(-1, 4)[:, 3]
    o.set_solution(f["u"] + h * du, f["p"] + h * dpp)
    Rp = o.residual(save=False)
    o.set_solution(f["u"] - h * du, f["p"] - h * dpp)
    Rm = o.residual(save=False)
    assert relerr((Rp - Rm) / (2 * h), A @ d) < 1e-6


def test_j2_state_semantics():
    """Parity traps 1-3 of SURVEY.md 8(a): Fp untouched on the elastic branch, eqps carried, save=false writes nothing."""
    co, cn = kuhn_cube(3)
    f = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, "J2", [MATERIAL])
    o.state("Fp_old")[:] = f["Fp_old"]
    o.state("eqps_old")[:] = f["eqps_old"]
    o.state("Fp")[:] = 7.0
    o.set_solution(f["u"], f["p"])
    o.residual(save=False)
    assert np.all(o.state("Fp") == 7.0) and np.all(o.state("sigma") == 0.0)
    o.residual(save=True)
    npl = o.plastic_count()
    assert 0 < npl < o.ne
    untouched = np.all(o.state("Fp") == 7.0, axis=1)
    assert untouched.sum() == o.ne - npl
    assert np.array_equal(o.state("eqps")[untouched], f["eqps_old"][untouched])
    assert np.all(o.state("eqps")[~untouched] > f["eqps_old"][~untouched])
    o.update_states()
    assert np.array_equal(o.state("Fp_old"), o.state("Fp"))


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_error_localisation_identities(model):
    """goal_error has no reference golden (parity unpinned); these follow from sum_n N_n = 1."""
    co, cn = kuhn_cube(3)
    f = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, model, [MATERIAL])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    o.set_solution(f["u"], f["p"])
    # (1) z == 1: the weights reduce to the shape functions -> the ordinary ST residual
    one3, one = np.ones((o.nn, 3)), np.ones(o.nn)
    assert relerr(o.localize(one3, one, one), o.residual(save=False)) < 1e-13
    # (2) the localisation is linear in z
    R1 = o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"]).copy()
    R2 = o.localize(2 * f["zu_diff"], 2 * f["zp_diff"], 2 * f["zp_coarse"])
    assert relerr(R2, 2 * R1) < 1e-13
    # (3) element indicator / bound / parent sum definitions (src/goal_error.cpp:7-56)
    Rn = R1.reshape(-1, 4)
    parent = (np.arange(o.ne) // 6).astype(np.int32)
    eta, etap, bound = o.element_error(Rn[:, :3], Rn[:, 3], parent, o.ne // 6)
    want = np.abs(0.25 * Rn[cn].sum(axis=(1, 2)))
    assert relerr(eta, want) < 1e-13
    assert relerr(etap, eta.reshape(-1, 6).sum(1)) < 1e-14
    assert abs(bound - np.abs(Rn.sum(1)).sum()) < 1e-12 * bound
