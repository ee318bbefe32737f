"""CPU checks of the CUDA path's own code (no GPU needed): the closed-form per-seed
tangent of goal_b200/csrc/element_math.cuh and the launch/scatter logic of gx_kernels.cuh +
gx_setup.cpp are compiled for the host by tests/hostcheck and compared with the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import blockerr, relerr
from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
from oracle.oracle import ADJOINT, PRIMAL, Oracle

dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_element_math_matches_fad_oracle(hostcheck, model):
    rng = np.random.RandomState(1)
    mat = np.array(MATERIAL)
    seen = set()
    for trial in range(120):
        x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]) * 0.1 + 0.01 * rng.randn(4, 3)
        if np.linalg.det(x[1:] - x[0]) < 0:
            x[[1, 2]] = x[[2, 1]]
        u = (0.0001, 0.0005, 0.003)[trial % 3] * rng.randn(4, 3)
        p = rng.randn(4)
        Fpo = (np.eye(3) + 1e-3 * rng.randn(3, 3)).reshape(-1)
        eqo = 0.01 * rng.rand()
        o = Oracle(x, np.array([[0, 1, 2, 3]], dtype=np.int32), model, [MATERIAL])
        o.set_solution(u, p)
        if model == "J2":
            o.state("Fp_old")[:] = Fpo
            o.state("eqps_old")[:] = eqo
            o.state("Fp")[:] = 7.0
        R, vals = o.jacobian(PRIMAL, save=True)
        Ko = o.csr(vals).toarray()
        K, Rh, sig, Fp = np.zeros(256), np.zeros(16), np.zeros(9), np.full(9, 7.0)
        eq, wf, pl = C.c_double(0), C.c_int(0), C.c_int(0)
        rc = hostcheck.hc_element(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                  1, dp(K), dp(Rh), dp(sig), C.byref(eq), dp(Fp), C.byref(wf), C.byref(pl))
        assert rc == 0
        seen.add(pl.value)
        assert relerr(K.reshape(16, 16), Ko) < 1e-12
        assert relerr(Rh, R) < 1e-12
        Kw = np.zeros(256)  # same with r_m rebuilt from w_m (column_node_w: what the tangent records carry)
        rc = hostcheck.hc_element(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                  2, dp(Kw), dp(np.zeros(16)), dp(np.zeros(9)), C.byref(C.c_double(0)), dp(np.zeros(9)),
                                  C.byref(C.c_int(0)), C.byref(C.c_int(0)))
        assert rc == 0 and relerr(Kw.reshape(16, 16), Ko) < 1e-12
        for mode in (2 | 4, 2 | 4 | 8):  # + jacobian_block_add (plain and transposed): the patch gather's inner loop
            Ka = np.zeros(256)
            rc = hostcheck.hc_element(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                      mode, dp(Ka), dp(np.zeros(16)), dp(np.zeros(9)), C.byref(C.c_double(0)), dp(np.zeros(9)),
                                      C.byref(C.c_int(0)), C.byref(C.c_int(0)))
            assert rc == 0 and relerr(Ka.reshape(16, 16), Ko) < 1e-12
        for mode in (16, 16 | 8):  # the tangent record + pair / diagonal contributions (tangent_record.cuh), plain and transposed
            Kt, Rt = np.zeros(256), np.zeros(16)
            rc = hostcheck.hc_element(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                      mode, dp(Kt), dp(Rt), dp(np.zeros(9)), C.byref(C.c_double(0)), dp(np.zeros(9)),
                                      C.byref(C.c_int(0)), C.byref(C.c_int(0)))
            assert rc == 0 and relerr(Kt.reshape(16, 16), Ko) < 1e-12 and relerr(Rt, R) < 1e-12
            assert blockerr(Kt.reshape(16, 16), Ko) < 1e-10
        assert relerr(sig, o.state("sigma")[0]) < 1e-10
        if model == "J2":
            assert pl.value == o.plastic_count()
            assert abs(eq.value - o.state("eqps")[0]) < 1e-12
            assert np.abs(Fp - o.state("Fp")[0]).max() < 1e-10  # elastic: both still 7.0
        zu, zp, zpc = 1e-2 * rng.randn(4, 3), 1e-2 * rng.randn(4), 1e-2 * rng.randn(4)
        Re, Rh2 = o.localize(zu, zp, zpc), np.zeros(16)
        hostcheck.hc_error_residual(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                    dp(zu), dp(zp), dp(zpc), dp(Rh2))
        assert relerr(Rh2, Re) < 1e-12
    assert seen == ({0, 1} if model == "J2" else {0})


def test_j2_barely_yielding_element_matches_fad_oracle(hostcheck):
    """goal_J2.cpp:108-121: for an element that barely yields (1e-12 < f < ~1e-9) the reference's Newton loop passes
    its |R|/Y < 1e-11 test after ONE iteration, X = f / (2 mubar), and FAD differentiates that iterate -- an O(1)
    difference in the consistent tangent against X = f / (2 mubar + 2K/3).  The closed form must follow it."""
    rng = np.random.RandomState(7)
    mat = np.array(MATERIAL)
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]) * 0.1 + 0.01 * rng.randn(4, 3)
    u = 0.003 * rng.randn(4, 3)
    p = rng.randn(4)
    Fpo = (np.eye(3) + 1e-3 * rng.randn(3, 3)).reshape(-1)
    o = Oracle(x, np.array([[0, 1, 2, 3]], dtype=np.int32), "J2", [MATERIAL])
    o.set_solution(u, p)
    o.state("Fp_old")[:] = Fpo

    def plastic(eq):
        o.state("eqps_old")[:] = eq
        o.residual(save=False)
        return o.plastic_count()

    lo, hi = 0.0, 10.0  # f = |s| - sqrt(2/3)(Y + K eqps_old) falls with eqps_old: find where yielding stops
    assert plastic(lo) == 1 and plastic(hi) == 0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if plastic(mid) else (lo, mid)
    for f_target, first_iter in ((2e-10, True), (1e-7, False)):
        eqo = lo - f_target / (np.sqrt(2.0 / 3.0) * MATERIAL[2])
        o.state("eqps_old")[:] = eqo
        R, vals = o.jacobian(PRIMAL, save=False)
        assert o.plastic_count() == 1
        Ko = o.csr(vals).toarray()
        K, pl = np.zeros(256), C.c_int(0)
        rc = hostcheck.hc_element(1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo), 16, dp(K), dp(np.zeros(16)), dp(np.zeros(9)),
                                  C.byref(C.c_double(0)), dp(np.zeros(9)), C.byref(C.c_int(0)), C.byref(pl))
        assert rc == 0 and pl.value == 1
        assert relerr(K.reshape(16, 16), Ko) < 1e-12, (f_target, first_iter)
    # and the two iterates really differ in the tangent: the same element with the window missed is not within 1e-12
    assert relerr(K.reshape(16, 16), Ko) < 1e-12


def test_j2_return_map_failure_code(hostcheck):
    """non-finite plastic increment -> ERR_J2_RETURN_MAP (3), the reference's fail("J2: return mapping failed")"""
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    u = 1e-3 * np.arange(12.0).reshape(4, 3)
    p, mat, Fpo = np.zeros(4), np.array(MATERIAL), np.eye(3).reshape(-1).copy()
    args = lambda eq: (1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eq), 1, dp(np.zeros(256)), dp(np.zeros(16)), dp(np.zeros(9)),
                       C.byref(C.c_double(0)), dp(np.zeros(9)), C.byref(C.c_int(0)), C.byref(C.c_int(0)))
    assert hostcheck.hc_element(*args(0.0)) == 0
    assert hostcheck.hc_element(*args(-np.inf)) == 3


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_von_mises_derivative_matches_fad_oracle(hostcheck, model):
    """element_von_mises (closed-form d vm / d u) against the oracle's FADT evaluation of AvgVM
    (src/goal_avg_vm.cpp:43-61, goal_von_mises.cpp:6-18) on single elements, elastic and plastic."""
    rng = np.random.RandomState(5)
    mat = np.array(MATERIAL)
    seen = set()
    for trial in range(60):
        x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]) * 0.1 + 0.01 * rng.randn(4, 3)
        if np.linalg.det(x[1:] - x[0]) < 0:
            x[[1, 2]] = x[[2, 1]]
        u = (0.0001, 0.0005, 0.003)[trial % 3] * rng.randn(4, 3)
        p = rng.randn(4)
        Fpo = (np.eye(3) + 1e-3 * rng.randn(3, 3)).reshape(-1)
        eqo = 0.01 * rng.rand()
        o = Oracle(x, np.array([[0, 1, 2, 3]], dtype=np.int32), model, [MATERIAL])
        o.set_solution(u, p)
        if model == "J2":
            o.state("Fp_old")[:] = Fpo
            o.state("eqps_old")[:] = eqo
        Jo, do = o.functional("avg vm", with_dMdu=True)
        seen.add(o.plastic_count() > 0)
        out, dvm = np.zeros(2), np.zeros(12)
        rc = hostcheck.hc_von_mises(0 if model == "neohookean" else 1, dp(x), dp(u), dp(p), dp(mat), dp(Fpo), C.c_double(eqo),
                                    dp(out), dp(dvm))
        assert rc == 0
        assert abs(out[0] * out[1] - Jo) < 1e-12 * abs(Jo)
        do = do.reshape(4, 4)
        assert np.all(do[:, 3] == 0.0)  # von Mises does not see the pressure
        assert relerr(dvm.reshape(4, 3), do[:, :3]) < 1e-11
    assert seen == ({False, True} if model == "J2" else {False})


def test_expm_matches_scipy(hostcheck):
    import scipy.linalg
    rng = np.random.RandomState(2)
    for sc in (1e-4, 1e-3, 1e-2, 0.1, 1.0, 10.0):
        A = sc * rng.randn(3, 3)
        o = np.zeros(9)
        hostcheck.hc_expm3(dp(A.reshape(-1).copy()), dp(o))
        E = scipy.linalg.expm(A)
        assert np.abs(o.reshape(3, 3) - E).max() < 5e-14 * np.abs(E).max()


def _emulate(hostcheck, model, pass_, save, co, cn, f, z5=None):
    nn, ne = len(co), len(cn)
    nnz, ncol, pl = C.c_int64(0), C.c_int32(0), C.c_int64(0)
    mat = np.array(MATERIAL)
    nul = None
    rc = hostcheck.hc_assemble(model, 0, 0, nn, ne, ip(cn), dp(co), dp(mat), nul, nul, nul, nul, nul, nul, nul, nul, nul, nul,
                               C.byref(nnz), nul, nul, C.byref(ncol), C.byref(pl))
    assert rc == 0
    rowptr, colind = np.zeros(4 * nn + 1, dtype=np.int64), np.zeros(nnz.value, dtype=np.int32)
    sig, eq, eqo = np.zeros((ne, 9)), np.zeros(ne), f["eqps_old"].copy()
    Fp, Fpo = np.full((ne, 9), 7.0), f["Fp_old"].copy()
    R, vals = np.zeros(4 * nn), np.zeros(nnz.value)
    u, p = np.ascontiguousarray(f["u"]), np.ascontiguousarray(f["p"])
    rc = hostcheck.hc_assemble(model, pass_, save, nn, ne, ip(cn), dp(co), dp(mat), dp(u), dp(p), dp(z5), dp(sig), dp(eq),
                               dp(eqo), dp(Fp), dp(Fpo), dp(R), dp(vals), C.byref(nnz), lp(rowptr), ip(colind),
                               C.byref(ncol), C.byref(pl))
    assert rc == 0
    return dict(R=R, vals=vals, rowptr=rowptr, colind=colind, sigma=sig, eqps=eq, Fp=Fp, ncolors=ncol.value, plastic=pl.value)


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("mesh", ["cube", "kuhn5"])
def test_whole_mesh_emulation_matches_oracle(hostcheck, cube, model, mesh):
    co, cn = (cube["coords"], cube["tets"]) if mesh == "cube" else kuhn_cube(5)
    mi = 0 if model == "neohookean" else 1
    f = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, model, [MATERIAL])
    o.set_solution(f["u"], f["p"])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
        o.state("Fp")[:] = 7.0
    Ro, Ao = o.jacobian(PRIMAL, save=True)
    h = _emulate(hostcheck, mi, 1, 1, co, cn, f)
    assert np.array_equal(h["rowptr"], o.rowptr) and np.array_equal(h["colind"], o.colind)  # graph bit-exact
    assert h["plastic"] == o.plastic_count()
    assert relerr(h["R"], Ro) < 1e-12 and relerr(h["vals"], Ao) < 1e-12
    assert relerr(h["sigma"], o.state("sigma")) < 1e-10
    if model == "J2":
        assert np.abs(h["eqps"] - o.state("eqps")).max() < 1e-12
        assert np.abs(h["Fp"] - o.state("Fp")).max() < 1e-10
    _, At = o.jacobian(ADJOINT, save=False)
    h2 = _emulate(hostcheck, mi, 2, 0, co, cn, f)
    assert relerr(h2["vals"], At) < 1e-12 and np.all(h2["sigma"] == 0.0)
    assert relerr(_emulate(hostcheck, mi, 0, 0, co, cn, f)["R"], o.residual(save=False)) < 1e-12
    z5 = np.concatenate([f["zu_diff"], f["zp_diff"][:, None], f["zp_coarse"][:, None]], 1).copy()
    Re = o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"])
    assert relerr(_emulate(hostcheck, mi, 3, 0, co, cn, f, z5=z5)["R"], Re) < 1e-12


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("mesh", ["cube", "kuhn5"])
def test_patch_gather_replay_matches_oracle(hostcheck, cube, model, mesh):
    """The default GPU schedule of the Jacobian pass, replayed on the CPU: the patch schedule built by the product's
    host code and interpreted like patch_pair_kernel (record slots from the bulk-copy runs, work items, partial
    sums, one writer per block) with the device's element math -> the oracle's operator, primal and transposed."""
    co, cn = (cube["coords"], cube["tets"]) if mesh == "cube" else kuhn_cube(5)
    co, cn = np.ascontiguousarray(co, dtype=np.float64), np.ascontiguousarray(cn, dtype=np.int32)
    f = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, model, [MATERIAL])
    o.set_solution(f["u"], f["p"])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    mat = np.array(MATERIAL)
    u, p = np.ascontiguousarray(f["u"]), np.ascontiguousarray(f["p"])
    eqo, Fpo = np.ascontiguousarray(f["eqps_old"]), np.ascontiguousarray(f["Fp_old"])
    for mode, tr in ((PRIMAL, 0), (ADJOINT, 1)):
        Ro, Ao = o.jacobian(mode, save=False)
        R, A, npch = np.zeros(4 * len(co)), np.zeros(o.nnz), C.c_int32(0)
        rc = hostcheck.hc_patch_gather(0 if model == "neohookean" else 1, tr, len(co), len(cn), ip(cn), dp(co), dp(mat), dp(u), dp(p),
                                       dp(eqo), dp(Fpo), dp(R), dp(A), C.byref(npch))
        assert rc == 0 and npch.value >= 1
        assert relerr(A, Ao) < 1e-12 and relerr(R, Ro) < 1e-12
        be = blockerr(o.csr(A).toarray(), o.csr(Ao).toarray())
        print(f"per-block relative error {model} {mesh} mode {tr}: {be:.2e}")
        assert be < 1e-11  # every 4x4 block on its own scale (relerr is norm-wise)


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("mesh", ["cube", "kuhn5"])
def test_block_reduced_residual_replay_matches_oracle(hostcheck, cube, model, mesh):
    """The default GPU schedule of the residual and error-localisation passes, replayed on the CPU: the element lines
    from the device's element code, summed through the block schedule built by the product's host code and read like
    elem_residual_block_kernel + node_partial_sum_kernel do (slots, pre-swizzled chunk entries, partial sums in block
    order, one writer per entry) -> the oracle's residual (Primal::compute_resid, src/goal_primal.cpp:75-90) and its
    localised error residual (NestedAdjoint::localize, src/goal_nested_adjoint.cpp:217-234)."""
    co, cn = (cube["coords"], cube["tets"]) if mesh == "cube" else kuhn_cube(5)
    co, cn = np.ascontiguousarray(co, dtype=np.float64), np.ascontiguousarray(cn, dtype=np.int32)
    f = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, model, [MATERIAL])
    o.set_solution(f["u"], f["p"])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    mat = np.array(MATERIAL)
    u, p = np.ascontiguousarray(f["u"]), np.ascontiguousarray(f["p"])
    eqo, Fpo = np.ascontiguousarray(f["eqps_old"]), np.ascontiguousarray(f["Fp_old"])
    z5 = np.ascontiguousarray(np.concatenate([f["zu_diff"], f["zp_diff"][:, None], f["zp_coarse"][:, None]], axis=1))
    m = 0 if model == "neohookean" else 1
    hostcheck.hc_residual_blocks.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 8
    for z, want in ((None, o.residual(save=False).copy()), (z5, o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"]).copy())):
        R = np.zeros(4 * len(co))
        rc = hostcheck.hc_residual_blocks(m, len(co), len(cn), ip(cn), dp(co), dp(mat), dp(u), dp(p), dp(z) if z is not None else None,
                                          dp(eqo), dp(Fpo), dp(R))
        assert rc == 0
        assert relerr(R, want) < 1e-12
        comp = np.abs(want.reshape(-1, 4)).max(axis=0)  # u rows and p rows on their own scales
        assert (np.abs(R - want).reshape(-1, 4).max(axis=0) < 1e-12 * comp).all()
