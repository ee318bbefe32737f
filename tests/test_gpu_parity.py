"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same
seeded inputs; the reference's goldens through the GPU path; size-independent properties at
BASELINE.json's 1M-tet size.  Bars (north_star / BASELINE.md 5): graph and scatter map bit-exact;
R, Jacobian values, error indicators 1e-12 relative (norm-wise); J2 state 1e-10."""
import numpy as np
import pytest

from conftest import blockerr, relerr
from goal_b200.synthetic import MATERIAL, fields, kuhn_block, kuhn_cube

pytestmark = pytest.mark.gpu


def _pair(co, cn, model, f, elem_set=None, mats=(MATERIAL,)):
    import goal_b200
    from oracle.oracle import Oracle
    a = goal_b200.Assembler(co, cn, model, list(mats), elem_set=elem_set)
    o = Oracle(co, cn, model, list(mats), elem_set=elem_set)
    for s in (a, o):
        s.set_solution(f["u"], f["p"])
    if model == "J2":
        a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
        a.set_state("Fp", np.full((a.ne, 9), 7.0))
        o.state("Fp_old")[:] = f["Fp_old"]; o.state("eqps_old")[:] = f["eqps_old"]; o.state("Fp")[:] = 7.0
    return a, o


def _mesh(cube, name):
    return (cube["coords"], cube["tets"]) if name == "cube" else kuhn_cube(int(name[4:]))


@pytest.mark.parametrize("mesh", ["cube", "kuhn8"])
def test_graph_and_scatter_map_bit_exact(cube, mesh):
    import goal_b200
    from oracle.oracle import Oracle
    co, cn = _mesh(cube, mesh)
    a = goal_b200.Assembler(co, cn, "neohookean", [MATERIAL])
    o = Oracle(co, cn, "neohookean", [MATERIAL])
    assert a.nnz == o.nnz
    assert np.array_equal(a.rowptr, o.rowptr) and np.array_equal(a.colind, o.colind)
    if mesh == "cube":
        assert a.nnz == 8176
    # scatter map: value index of ((n,i),(m,k)) = rowptr[4 a_n + i] + 4 bpos + k must hold column 4 a_m + k
    b = a.scatter_map().reshape(-1, 4, 4)
    for i in range(4):
        pos = o.rowptr[4 * cn[:, :, None] + i] + 4 * b.astype(np.int64)
        for k in range(4):
            assert np.array_equal(o.colind[pos + k], np.broadcast_to(4 * cn[:, None, :] + k, pos.shape))


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("mesh", ["cube", "kuhn8"])
def test_jacobian_residual_state_parity(cube, model, mesh):
    import goal_b200
    co, cn = _mesh(cube, mesh)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, model, f)
    R, A = a.jacobian(goal_b200.PRIMAL, save=True)
    assert a.last_timing()["launches"] == 2  # element records + patch gather
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=True)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    assert relerr(a.get_state("sigma"), o.state("sigma")) < 1e-10
    if model == "J2":
        assert a.plastic_count() == o.plastic_count() and 0 < a.plastic_count() < a.ne
        assert np.abs(a.get_state("eqps") - o.state("eqps")).max() < 1e-10
        Fp = a.get_state("Fp")
        assert np.abs(Fp - o.state("Fp")).max() < 1e-10
        assert np.all(Fp[np.all(o.state("Fp") == 7.0, axis=1)] == 7.0)  # elastic branch leaves Fp untouched
    # adjoint mode scatters the transpose (src/goal_displacement.cpp:196-214); save=false writes no state
    sig_before = a.get_state("sigma")
    At = a.jacobian(goal_b200.ADJOINT, save=False)[1].copy()
    _, Ato = o.jacobian(goal_b200.ADJOINT, save=False)
    assert relerr(At, Ato) < 1e-12
    assert np.array_equal(a.get_state("sigma"), sig_before)
    Ap = a.jacobian(goal_b200.PRIMAL, save=False)[1].copy()
    assert abs(a.csr(At) - a.csr(Ap).T).max() < 1e-12 * np.abs(Ap).max()  # same sums, different order
    # residual-only pass (Primal::compute_resid)
    assert relerr(a.residual(save=False), o.residual(save=False)) < 1e-12
    # States::update
    a.update_states(); o.update_states()
    if model == "J2":
        assert np.abs(a.get_state("Fp_old") - o.state("Fp_old")).max() < 1e-10
        assert np.abs(a.get_state("eqps_old") - o.state("eqps_old")).max() < 1e-10
    a.close()


@pytest.mark.parametrize("opts", [dict(), dict(kernel=1), dict(residual_kernel=1)])
@pytest.mark.parametrize("mesh", ["cube", "kuhn7"])
def test_kernel_variants_parity(cube, mesh, opts):
    """Both schedules (owner-computes: element records + patch pairs / gather form = default; coloured elements) against
    the oracle, primal and adjoint, incl. the fixture whose nodes have up to 56 incident elements."""
    import goal_b200
    co, cn = _mesh(cube, mesh)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, "J2", f)
    for k, v in opts.items():
        a.set_option(k, v)
    R, A = a.jacobian(goal_b200.PRIMAL, save=True)
    if not opts:
        assert a.last_timing()["launches"] == 2  # element records + patch pairs
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=True)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    assert blockerr(a.csr(A).toarray(), o.csr(Ao).toarray()) < 1e-12  # every 4x4 block on its own scale
    assert relerr(a.get_state("sigma"), o.state("sigma")) < 1e-10
    assert np.abs(a.get_state("Fp") - o.state("Fp")).max() < 1e-10 and a.plastic_count() == o.plastic_count()
    At = a.jacobian(goal_b200.ADJOINT, save=False)[1].copy()
    assert relerr(At, o.jacobian(goal_b200.ADJOINT, save=False)[1]) < 1e-12
    R1, A1 = [x.copy() for x in a.jacobian(goal_b200.PRIMAL, save=False)]
    R2, A2 = a.jacobian(goal_b200.PRIMAL, save=False)
    assert np.array_equal(R1, R2) and np.array_equal(A1, A2)  # bit-reproducible
    # residual and error-localisation passes under the same option (kernel=1: coloured; residual_kernel=1: element
    # lines + node gather; default: block-reduced)
    assert relerr(a.residual(save=False), o.residual(save=False)) < 1e-12
    if not opts:
        assert a.last_timing()["launches"] == 2  # element blocks + partial sums of the shared nodes
    assert relerr(a.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"]),
                  o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"])) < 1e-12
    a.close()


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("form", [0, 1])
def test_residual_pass_with_state_save(model, form):
    """Primal::compute_resid (src/goal_primal.cpp:75-90) = the residual pass that also saves the state: R, sigma, eqps and
    Fp against the oracle under both forms of the pass, on a mesh of 16 element blocks (ragged last block), elastic and
    plastic elements mixed, twice (bit-identical), and after a shuffle of the element order (no locality: every node is
    finished by the partial-sum kernel)."""
    import goal_b200
    co, cn = kuhn_block(7, 6, 8)
    for shuffle in (False, True):
        if shuffle:
            cn = cn[np.random.default_rng(5).permutation(len(cn))]
        f = fields(co, len(cn), strain=0.004)
        a, o = _pair(co, cn, model, f)
        a.set_option("residual_kernel", form)
        R1 = a.residual(save=True).copy()
        Ro = o.residual(save=True)
        assert relerr(R1, Ro) < 1e-12
        nodal = np.abs(Ro.reshape(-1, 4)).max(axis=0)  # per component (u rows and p rows differ in scale)
        assert (np.abs(R1 - Ro).reshape(-1, 4).max(axis=0) < 1e-12 * nodal).all()
        assert relerr(a.get_state("sigma"), o.state("sigma")) < 1e-10
        if model == "J2":
            assert 0 < a.plastic_count() < a.ne and a.plastic_count() == o.plastic_count()
            assert np.abs(a.get_state("Fp") - o.state("Fp")).max() < 1e-10
            assert np.abs(a.get_state("eqps") - o.state("eqps")).max() < 1e-10
        assert np.array_equal(a.residual(save=True), R1)
        zu, zp, zc = f["zu_diff"], f["zp_diff"], f["zp_coarse"]
        assert relerr(a.localize(zu, zp, zc), o.localize(zu, zp, zc)) < 1e-12
        a.close()


def _fan(k=9):
    """k x k grid in the plane z = 0, every triangle joined to one apex: the apex has 2 (k-1)^2 incident elements."""
    g = np.linspace(0.0, 1.0, k)
    X, Y = np.meshgrid(g, g, indexing="ij")
    co = np.concatenate([np.stack([X.ravel(), Y.ravel(), np.zeros(k * k)], axis=1), [[0.5, 0.5, 1.0]]])
    apex = k * k
    cn = []
    for i in range(k - 1):
        for j in range(k - 1):
            a, b, c, d = i * k + j, (i + 1) * k + j, (i + 1) * k + j + 1, i * k + j + 1
            cn += [[a, b, c, apex], [a, c, d, apex]]
    return co, np.array(cn, dtype=np.int32)


def test_patch_schedule_fallback():
    """A node with more incident elements than a patch stages (242 > 176): the default Jacobian pass must fall back
    to the coloured schedule and still match the oracle."""
    import goal_b200
    co, cn = _fan(12)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, "J2", f)
    R, A = a.jacobian(goal_b200.PRIMAL, save=True)
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=True)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    assert a.last_timing()["launches"] >= 242  # one launch per colour (the apex alone forces 242): not the patch schedule
    At = a.jacobian(goal_b200.ADJOINT, save=False)[1].copy()
    assert relerr(At, o.jacobian(goal_b200.ADJOINT, save=False)[1]) < 1e-12
    a.close()


def test_multiple_elem_sets(cube):
    import goal_b200
    co, cn = kuhn_cube(6)
    es = (np.arange(len(cn)) % 3).astype(np.int32)
    mats = [MATERIAL, (2000.0, 0.3, 50.0, 20.0, 0.5), (500.0, 0.2, 200.0, 5.0, 2.0)]
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, "J2", f, elem_set=es, mats=mats)
    R, A = a.jacobian(goal_b200.PRIMAL, save=True)
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=True)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12 and a.plastic_count() == o.plastic_count()
    a.close()


@pytest.mark.parametrize("model", ["neohookean", "J2"])
def test_error_localisation_parity(cube, model):
    co, cn = kuhn_cube(6)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, model, f)
    R = a.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"]).copy()
    Ro = o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"])
    assert relerr(R, Ro) < 1e-12
    one3, one = np.ones((a.nn, 3)), np.ones(a.nn)
    assert relerr(a.localize(one3, one, one).copy(), a.residual(save=False)) < 1e-12  # z == 1 identity
    Rn = Ro.reshape(-1, 4)
    parent = (np.arange(a.ne) // 6).astype(np.int32)
    eta, etap, bound = a.element_error(Rn[:, :3], Rn[:, 3], parent, a.ne // 6)
    eo, epo, bo = o.element_error(Rn[:, :3], Rn[:, 3], parent, a.ne // 6)
    assert relerr(eta, eo) < 1e-12 and relerr(etap, epo) < 1e-12 and abs(bound - bo) < 1e-12 * bo
    a.close()


def test_functional_and_device_dbcs(cube):
    """'next' rows: avg-disp functional + dMdu (src/goal_avg_disp.cpp, goal_qoi.cpp:63-76) and Dirichlet rows
    (src/goal_dbcs.cpp:39-99) on the device against the oracle / the host construction."""
    import goal_b200
    co, cn = kuhn_cube(6)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, "J2", f)
    J, d = a.avg_disp(with_dMdu=True)
    Jo, do = o.avg_disp(with_dMdu=True)
    assert abs(J - Jo) < 1e-13 * abs(Jo) and relerr(d, do) < 1e-13
    R, A = [x.copy() for x in a.jacobian(goal_b200.PRIMAL, save=False)]
    rows = np.array([4 * n + eq for eq in (0, 2, 3) for n in range(0, a.nn, 7)], dtype=np.int32)
    g = np.linspace(-1, 1, len(rows))
    a.apply_dbcs(rows, g, True)
    R2, A2 = a.fetch()
    sol = np.concatenate([f["u"], f["p"][:, None]], 1).reshape(-1)
    for k, row in enumerate(rows):
        lo, hi = a.rowptr[row], a.rowptr[row + 1]
        want = (a.colind[lo:hi] == row).astype(float)
        assert np.array_equal(A2[lo:hi], want) and R2[row] == sol[row] - g[k]
        A[lo:hi] = want; R[row] = sol[row] - g[k]
    assert np.array_equal(A, A2) and np.array_equal(R, R2)  # every other row untouched
    a.close()


@pytest.mark.parametrize("model", ["neohookean", "J2"])
@pytest.mark.parametrize("mesh", ["cube", "kuhn6"])
def test_all_functionals_match_fad_oracle(cube, model, mesh):
    """every functional of Mechanics::build_functional (src/goal_mechanics.cpp:149-167) and
    its dMdu (QoI<FADT>::scatter, src/goal_qoi.cpp:63-76) on the device against the oracle's ST / FADT evaluation:
    avg disp, avg disp subdomain, avg vm (closed-form von Mises derivative), max vm (KS), point wise."""
    co, cn = _mesh(cube, mesh)
    f = fields(co, len(cn), strain=0.004)
    es = (np.arange(len(cn)) % 3 == 1).astype(np.int32)
    a, o = _pair(co, cn, model, f, elem_set=es, mats=(MATERIAL, MATERIAL))
    a.residual(save=True); o.residual(save=True)  # "max vm" reads the saved sigma state
    for t in ("avg disp", "avg disp subdomain", "avg vm", "max vm", "point wise"):
        kw = dict(elem_set=1, rho=0.05, point=(len(co) // 2, 1))
        J, d = a.functional(t, with_dMdu=True, **kw)
        Jo, do = o.functional(t, with_dMdu=True, **kw)
        assert abs(J - Jo) <= 1e-12 * abs(Jo), t
        assert relerr(d, do) < 1e-12, t
        assert np.array_equal(d, a.fetch_dMdu())
        assert abs(a.functional(t, **kw) - o.functional(t, **kw)) <= 1e-12 * abs(Jo), t  # ST chain: value only
        J2, d2 = a.functional(t, with_dMdu=True, **kw)
        assert J2 == J and np.array_equal(d, d2), t  # deterministic
    # the dedicated avg-disp entry point agrees with the generic one
    assert abs(a.avg_disp() - a.functional("avg disp")) < 1e-15
    a.close()


def test_side_set_boundary_terms(cube):
    """set_tbcs / set_ibcs (src/goal_tbcs.cpp:29-71, goal_ibcs.cpp:41-83) on the device-resident
    ghost R against the host construction of oracle/driver.py, on the reference fixture's ymax side set (16 faces)
    and on every boundary face of a Kuhn cube (sides sharing nodes; per-side traction values)."""
    from oracle import driver
    co, cn = cube["coords"], cube["tets"]
    f = fields(co, len(cn), strain=0.004)
    a, _ = _pair(co, cn, "neohookean", f)
    sides = np.array(cube["side_sets"]["ymax"], dtype=np.int32)
    R0 = a.residual(save=False).copy()
    a.apply_tbcs(sides, (0.3, -1.0, 0.25))
    want = R0.copy()
    for row, v in driver._traction_rhs(co, sides, (0.3, -1.0, 0.25)):
        want[row] += v
    got = a.fetch(values=False)[0]
    assert relerr(got - R0, want - R0) < 1e-13 and relerr(got, want) < 1e-14
    a.apply_ibcs(sides, 2.5, (0.5, 0.25, 0.75))
    for row, v in driver._inward_rhs(co, sides, 2.5, (0.5, 0.25, 0.75)):
        want[row] += v
    got2 = a.fetch(values=False)[0]
    assert relerr(got2 - got, want - got) < 1e-13
    a.close()
    # all boundary faces of a Kuhn cube, a different traction on every side
    co, cn = kuhn_cube(5)
    faces = np.concatenate([cn[:, [1, 2, 3]], cn[:, [0, 3, 2]], cn[:, [0, 1, 3]], cn[:, [0, 2, 1]]])
    key = np.sort(faces, 1)
    _, idx, cnt = np.unique(key, axis=0, return_index=True, return_counts=True)
    bf = np.ascontiguousarray(faces[idx[cnt == 1]], dtype=np.int32)
    assert len(bf) == 6 * 2 * 25
    f = fields(co, len(cn), strain=0.004)
    a, _ = _pair(co, cn, "J2", f)
    R0 = a.residual(save=False).copy()
    T = np.random.RandomState(3).randn(len(bf), 3)
    a.apply_tbcs(bf, T)
    want = R0.copy()
    for s_, tri in enumerate(bf):
        for row, v in driver._traction_rhs(co, [tri], T[s_]):
            want[row] += v
    got = a.fetch(values=False)[0]
    assert relerr(got - R0, want - R0) < 1e-13
    a.apply_tbcs(bf, T)  # repeated application is deterministic
    a2, _ = _pair(co, cn, "J2", f)
    a2.residual(save=False); a2.apply_tbcs(bf, T); a2.apply_tbcs(bf, T)
    assert np.array_equal(a.fetch(values=False)[0], a2.fetch(values=False)[0])
    a.close(); a2.close()


@pytest.mark.parametrize("mesh", ["cube", "kuhn6"])
def test_solution_update_and_size_field(cube, mesh):
    """Disc::add_soln (src/goal_disc.cpp:398-422) on the device-resident solution and
    get_iso_target_size (src/goal_size_field.cpp:39-150) from the element indicators, against the oracle."""
    co, cn = _mesh(cube, mesh)
    f = fields(co, len(cn), strain=0.004)
    a, o = _pair(co, cn, "neohookean", f)
    rng = np.random.RandomState(4)
    du = 1e-3 * rng.randn(len(co), 4)
    a.add_solution(du)
    u, p = a.get_solution()
    assert np.array_equal(u, f["u"] + du[:, :3]) and np.array_equal(p, f["p"] + du[:, 3])
    o.set_solution(u, p)
    assert relerr(a.residual(save=False), o.residual(save=False)) < 1e-12  # the kernels see the updated fields
    eta = np.abs(rng.randn(len(cn))) * 10.0 ** rng.uniform(-8, -2, len(cn))  # spans both clamps
    target = 3 * len(cn)
    v, G = a.size_field(eta, target)
    vo, Go = o.size_field(eta, target)
    assert abs(G - Go) < 1e-13 * Go and relerr(v, vo) < 1e-13
    x = co[cn]
    h = np.sqrt(sum(((x[:, i] - x[:, j]) ** 2).sum(1) for i in range(4) for j in range(i + 1, 4)) / 6)
    assert v.min() >= 0.25 * h.min() * (1 - 1e-12) and v.max() <= 2.0 * h.max() * (1 + 1e-12)
    # partitioned form: sums and counts, with the reduced G passed in
    s2, G2, c2 = a.size_field(eta, target, G=G, counts=True)
    assert G2 == G and relerr(s2 / c2, v) < 1e-15
    assert np.array_equal(c2, np.bincount(cn.reshape(-1), minlength=len(co)))
    a.close()


def test_stabilization_off(cube):
    """mechanics: stabilization: false (src/goal_mechanics.cpp:55-56, 140-143): the Stabilization evaluator is left out.
    All its terms carry tau = c0 h^2 / (2 mu), so the oracle with c0 = 0 is the reference chain without it."""
    import goal_b200
    from oracle.oracle import Oracle
    co, cn = kuhn_cube(5)
    f = fields(co, len(cn), strain=0.004)
    nostab = list(MATERIAL[:4]) + [0.0]
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL], flags=1)
    o = Oracle(co, cn, "J2", [nostab])
    for s_ in (a, o):
        s_.set_solution(f["u"], f["p"])
    a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
    o.state("Fp_old")[:] = f["Fp_old"]; o.state("eqps_old")[:] = f["eqps_old"]
    R, A = a.jacobian(goal_b200.PRIMAL, save=False)
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=False)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    zu, zp, zc = f["zu_diff"], f["zp_diff"], f["zp_coarse"]
    assert relerr(a.localize(zu, zp, zc), o.localize(zu, zp, zc)) < 1e-12
    # and it is not the stabilized chain: the pressure residual differs (4e-4 norm-wise on this input)
    o2 = Oracle(co, cn, "J2", [MATERIAL])
    o2.set_solution(f["u"], f["p"]); o2.state("Fp_old")[:] = f["Fp_old"]; o2.state("eqps_old")[:] = f["eqps_old"]
    assert relerr(R, o2.jacobian(goal_b200.PRIMAL, save=False)[0]) > 1e-5
    a.close()


def test_bitwise_determinism(cube):
    """No atomics on the data path: repeated passes give identical bits."""
    import goal_b200
    co, cn = kuhn_cube(10)
    f = fields(co, len(cn), strain=0.004)
    a, _ = _pair(co, cn, "J2", f)
    R1, A1 = [x.copy() for x in a.jacobian(goal_b200.PRIMAL, save=True)]
    for _ in range(3):
        R2, A2 = a.jacobian(goal_b200.PRIMAL, save=True)
        assert np.array_equal(R1, R2) and np.array_equal(A1, A2)
    a.close()


@pytest.mark.parametrize("name", ["neohookean_uniaxial_3D", "J2_uniaxial_3D", "J2_traction_3D"])
def test_reference_goldens_through_gpu_path(cube, name):
    """The reference's regression values (example/primal/*_3D.yaml) with every residual and
    Jacobian coming from the CUDA path; Newton/BC/solve steps as in oracle/driver.py."""
    import goal_b200
    from oracle import driver
    model, J_gold, _ = driver.GOLDEN[name]
    a = goal_b200.Assembler(cube["coords"], cube["tets"], model, [MATERIAL])
    co, cn = cube["coords"], cube["tets"]

    class WithFunctional:  # avg disp (src/goal_avg_disp.cpp:17-21) evaluated on the host from the solution
        def __getattr__(self, k):
            return getattr(a, k)

        def set_solution(self, u, p):
            self.u = u.copy()
            a.set_solution(u, p)

        def avg_disp(self):
            x = co[cn]
            vol = np.linalg.det(x[:, 1:] - x[:, :1]) / 6
            return float((self.u[cn].mean(1).sum(1) * vol / 3).sum())

    dbcs, tbcs = driver.golden_case(name, cube)
    r = driver.run_primal(WithFunctional(), co, dbcs, tbcs)
    # same run with the traction terms (gx_apply_tbcs), the Dirichlet rows (gx_apply_dbcs) and the functional
    # (gx_functional_avg_disp) on the device
    a2 = goal_b200.Assembler(cube["coords"], cube["tets"], model, [MATERIAL])
    r2 = driver.run_primal(a2, co, dbcs, tbcs, device_bcs=True)
    assert np.abs(np.array(r2["J"]) - np.array(r["J"])).max() < 1e-13 and r2["newton"] == r["newton"]
    a2.close()
    assert abs(r["J"][-1] - J_gold) < 1e-12
    for got, want in zip(r["J"], driver.GOLDEN_STEPS[name]):
        assert abs(got - want) < 1e-12
    assert r["plastic"] == {"neohookean_uniaxial_3D": [0, 0, 0], "J2_uniaxial_3D": [0, 132, 132], "J2_traction_3D": [0, 0, 0]}[name]
    a.close()


def test_error_codes(cube):
    import goal_b200
    co, cn = kuhn_cube(2)
    bad = cn.copy()
    bad[5, [1, 2]] = bad[5, [2, 1]]  # negative volume
    a = goal_b200.Assembler(co, bad, "neohookean", [MATERIAL])
    with pytest.raises(goal_b200.GxError) as e:
        a.residual()
    assert e.value.status == 3 and "element 5" in str(e.value)
    a.close()
    a = goal_b200.Assembler(co, cn, "neohookean", [MATERIAL])
    u = np.zeros((len(co), 3)); u[:, 0] = -1.5 * co[:, 0]  # det F < 0
    a.set_solution(u, np.zeros(len(co)))
    with pytest.raises(goal_b200.GxError) as e:
        a.jacobian()
    assert e.value.status == 4
    with pytest.raises(goal_b200.GxError):
        a.get_state("Fp")  # neo-Hookean registers only sigma (goal_mechanics.cpp:87-95)
    a.close()


def test_error_bound_on_more_than_262144_nodes():
    """sum_contribs (goal_error.cpp:37-56) on a mesh whose node count exceeds what 1023 partial blocks of 256 cover in
    one sweep (the 1024th partial used to be dropped): bound == numpy sum, indicators == the definition."""
    import goal_b200
    co, cn = kuhn_cube(64)
    assert len(co) == 65 ** 3 > 262144
    a = goal_b200.Assembler(co, cn, "neohookean", [MATERIAL])
    rng = np.random.RandomState(3)
    ue, pe = rng.randn(len(co), 3), rng.randn(len(co))
    eta, _, bound = a.element_error(ue, pe)
    want = np.abs(ue.sum(1) + pe).sum()
    assert abs(bound - want) < 1e-12 * want
    e4 = np.concatenate([ue, pe[:, None]], 1)
    assert np.abs(eta - np.abs(0.25 * e4[cn].sum(axis=(1, 2)))).max() < 1e-12 * np.abs(eta).max()
    a.close()


def test_isolated_nodes_are_accepted():
    """Nodes that belong to no element (the last node, and one in the middle): no blocks, zero R, and the rest
    of the operator unchanged -- through the patch schedule, whose diagonal flag must not be set for them."""
    import goal_b200
    co, cn = kuhn_cube(4)
    mid = 40
    co2 = np.concatenate([co[:mid], [[9.0, 9.0, 9.0]], co[mid:], [[7.0, 7.0, 7.0]]])
    cn2 = np.where(cn >= mid, cn + 1, cn).astype(np.int32)
    f = fields(co, len(cn), strain=0.004)
    f2 = dict(f)
    f2["u"] = np.concatenate([f["u"][:mid], [[0.0, 0.0, 0.0]], f["u"][mid:], [[0.0, 0.0, 0.0]]])
    f2["p"] = np.concatenate([f["p"][:mid], [0.0], f["p"][mid:], [0.0]])
    a, o = _pair(co2, cn2, "J2", f2)
    for _ in range(2):
        R, A = a.jacobian(goal_b200.PRIMAL, save=False)
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=False)
    assert np.array_equal(a.rowptr, o.rowptr) and relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    assert np.all(R.reshape(-1, 4)[[mid, len(co2) - 1]] == 0.0)
    assert relerr(a.residual(save=False), o.residual(save=False)) < 1e-12
    a.close()


def test_j2_return_map_failure():
    """goal_J2.cpp:119-120: the reference fail()s when the Newton loop on X does not converge -- with linear hardening
    that only happens on non-finite numbers.  A state that makes the yield function infinite must come back as
    GX_ERR_J2_RETURN_MAP with the element id, and the failed pass must leave no usable result (the reference aborts)."""
    import goal_b200
    co, cn = kuhn_cube(3)
    f = fields(co, len(cn), strain=0.004)
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])
    a.set_solution(f["u"], f["p"])
    bad = f["eqps_old"].copy()
    bad[17] = -np.inf  # f = |s| - sqrt(2/3)(Y + K eqps_old) = +inf
    a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", bad)
    for call in (lambda: a.jacobian(goal_b200.PRIMAL, save=False), lambda: a.residual(save=False)):
        with pytest.raises(goal_b200.GxError) as e:
            call()
        assert e.value.status == 5 and "element 17" in str(e.value)
        with pytest.raises(goal_b200.GxError):
            a.fetch()  # no result after a failed pass
    a.set_state("eqps_old", f["eqps_old"])
    a.jacobian(goal_b200.PRIMAL, save=False)  # the context recovers
    a.close()


def _block_ids(nrow):
    """block index of every CRS value (dof rows of node a hold its blocks side by side: entry i*(4 nb) + 4 j + k)."""
    ids = np.empty(16 * int(nrow[-1]), dtype=np.int64)
    nb_all = np.diff(nrow)
    for nb in np.unique(nb_all):
        rows = np.nonzero(nb_all == nb)[0]
        if nb == 0:
            continue
        pat = np.tile(np.repeat(np.arange(nb), 4), 4)[None, :] + nrow[rows][:, None]
        pos = (16 * nrow[rows])[:, None] + np.arange(16 * nb)[None, :]
        ids[pos.reshape(-1)] = pat.reshape(-1)
    return ids


def blockerr_crs(A, B, nrow):
    """max over 4x4 blocks of max|A_blk - B_blk| / max|B_blk| for CRS value arrays of the node-blocked graph."""
    ids = _block_ids(nrow)
    nblk = int(nrow[-1])
    num = np.zeros(nblk); den = np.zeros(nblk)
    np.maximum.at(num, ids, np.abs(A - B))
    np.maximum.at(den, ids, np.abs(B))
    ok = den > 0
    return float((num[ok] / den[ok]).max())


def test_full_size_oracle_parity_1M():
    """N=55 -> 998,250 tets (BASELINE.json configs[4], the 1M mesh) with the BENCH fields (2 % strain, 99.6 % of the
    elements plastic): the whole pass against the oracle -- R, every CRS value (norm-wise and per 4x4 block), sigma,
    eqps, Fp -- primal with state save, and the transposed operator."""
    import goal_b200
    from oracle.oracle import Oracle
    n = 55
    co, cn = kuhn_cube(n)
    f = fields(co, len(cn))
    a, o = _pair(co, cn, "J2", f)
    assert a.ne == 998250
    R, A = a.jacobian(goal_b200.PRIMAL, save=True)
    Ro, Ao = o.jacobian(goal_b200.PRIMAL, save=True)
    assert a.plastic_count() == o.plastic_count() and a.plastic_count() > 0.99 * a.ne
    assert np.array_equal(a.rowptr, o.rowptr) and np.array_equal(a.colind, o.colind)
    assert relerr(R, Ro) < 1e-12 and relerr(A, Ao) < 1e-12
    nrow, _ = a.node_graph()
    be = blockerr_crs(A, Ao, nrow)
    assert be < 1e-11, be  # every 4x4 block on its own scale
    assert relerr(a.get_state("sigma"), o.state("sigma")) < 1e-10
    assert np.abs(a.get_state("eqps") - o.state("eqps")).max() < 1e-10
    assert np.abs(a.get_state("Fp") - o.state("Fp")).max() < 1e-10
    At = a.jacobian(goal_b200.ADJOINT, save=False)[1]
    assert relerr(At, o.jacobian(goal_b200.ADJOINT, save=False)[1]) < 1e-12
    assert relerr(a.residual(save=False), o.residual(save=False)) < 1e-12
    assert relerr(a.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"]),
                  o.localize(f["zu_diff"], f["zp_diff"], f["zp_coarse"])) < 1e-12
    a.close()


def test_full_size_properties_1M():
    """N=55 -> 998,250 tets: size-independent properties on top of the oracle comparison above -- determinism,
    transpose, translation invariance / K*(rigid translation)=0, and directional finite differences."""
    import goal_b200
    from oracle.oracle import Oracle
    n = 55
    co, cn = kuhn_cube(n)
    f = fields(co, len(cn))
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])
    edges = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    assert a.ne == 998250 and a.nnz == 16 * ((n + 1) ** 3 + 2 * edges) == 40954336
    a.set_solution(f["u"], f["p"]); a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
    R, A = [x.copy() for x in a.jacobian(goal_b200.PRIMAL, save=False)]
    R2, A2 = a.jacobian(goal_b200.PRIMAL, save=False)
    assert np.array_equal(R, R2) and np.array_equal(A, A2)
    K = a.csr(A)
    At = a.jacobian(goal_b200.ADJOINT, save=False)[1].copy()
    assert abs(a.csr(At) - K.T).max() < 1e-12 * np.abs(A).max()
    # rigid translation leaves F unchanged: residual invariant and K t = 0
    t = np.zeros((a.nn, 4)); t[:, :3] = (0.3, -0.2, 0.1)
    assert np.abs(K @ t.reshape(-1)).max() < 1e-9 * np.abs(A).max()
    a.set_solution(f["u"] + t[:, :3], f["p"])
    assert relerr(a.residual(save=False).copy(), R) < 1e-11
    # directional FD; an element that crosses the yield surface inside +-h is not differentiable,
    # so allow a vanishing fraction of rows to miss the bar
    assert a.plastic_count() > 0.99 * a.ne
    d = np.random.RandomState(0).randn(a.nn, 4) * 1e-3
    h = 1e-4
    a.set_solution(f["u"] + h * d[:, :3], f["p"] + h * d[:, 3]); Rp = a.residual(save=False).copy()
    a.set_solution(f["u"] - h * d[:, :3], f["p"] - h * d[:, 3]); Rm = a.residual(save=False).copy()
    fd, an = (Rp - Rm) / (2 * h), K @ d.reshape(-1)
    assert (np.abs(fd - an) > 1e-5 * np.abs(an).max()).mean() < 1e-4
    a.close()
