"""Maximum-size edge case: a mesh whose CRS operator has more than 2^31 nonzeros (64-bit value offsets,
SURVEY.md 8a: nnz = 4.05e9 at 100M tets).  208^3 cells -> 53,990,912 tets, nnz = 2.19e9.  The oracle cannot
assemble that, so rows of sampled nodes -- including the very last ones, whose offsets exceed 2^31 -- are read
straight from device memory and compared with the oracle run on the 2x2x2-cell neighbourhood of each node."""
import ctypes as C

import numpy as np
import pytest

from conftest import relerr
from goal_b200.synthetic import MATERIAL, fields, kuhn_block, kuhn_cube

pytestmark = pytest.mark.gpu


def _dev_read(ptr, offset_doubles, n):
    import torch  # noqa: F401  (loads libcudart)
    rt = C.CDLL("libcudart.so.12")
    out = np.empty(n)
    rc = rt.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr + 8 * int(offset_doubles)), C.c_size_t(8 * n), 2)
    assert rc == 0
    return out


def test_operator_beyond_2_31_nonzeros():
    import psutil
    import torch
    import goal_b200
    from oracle.oracle import PRIMAL, Oracle
    if torch.cuda.mem_get_info()[0] < 90e9 or psutil.virtual_memory().available < 60e9:
        pytest.skip("needs ~80 GB of device and ~50 GB of host memory")
    n = 208
    co, cn = kuhn_cube(n)
    f = fields(co, len(cn))
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL])
    edges = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    assert a.ne == 6 * n ** 3 and a.nnz == 16 * ((n + 1) ** 3 + 2 * edges) and a.nnz > 2 ** 31
    a.set_solution(f["u"], f["p"]); a.set_state("Fp_old", f["Fp_old"]); a.set_state("eqps_old", f["eqps_old"])
    a.jacobian(goal_b200.PRIMAL, save=False, out=False)
    assert a.plastic_count() > 0.99 * a.ne
    nrow, ncol = a.node_graph()
    Rdev, Vdev = a.result_dev()
    s = n + 1
    gid = lambda i, j, k: i + s * (j + s * k)
    samples = [(1, 1, 1), (n // 2, n // 3, n // 5), (n - 1, n - 1, n - 1), (n - 1, 2, n - 1), (3, n - 1, n - 2)]
    for (i, j, k) in samples:
        node = gid(i, j, k)
        assert 16 * nrow[node] > (2 ** 31 if k >= n - 2 else -1)
        nb = int(nrow[node + 1] - nrow[node])
        assert nb == 15  # interior node of a Kuhn mesh: 14 neighbours + itself
        got = _dev_read(Vdev, 16 * nrow[node], 16 * nb).reshape(4, nb, 4)
        gotR = _dev_read(Rdev, 4 * node, 4)
        # oracle on the 2x2x2 cells around the node, same fields (they are functions of global ids / coordinates)
        lco, lcn = kuhn_block(2, 2, 2, (i - 1, j - 1, k - 1), n)
        li, lj, lk = np.meshgrid(np.arange(3), np.arange(3), np.arange(3), indexing="ij")
        lg = np.array([gid(i - 1 + (q % 3), j - 1 + ((q // 3) % 3), k - 1 + q // 9) for q in range(27)])
        cells = np.array([(i - 1 + (c % 2)) + n * ((j - 1 + ((c // 2) % 2)) + n * (k - 1 + c // 4)) for c in range(8)])
        eg = (6 * cells[:, None] + np.arange(6)[None, :]).reshape(-1)
        assert np.allclose(lco, co[lg])
        o = Oracle(lco, lcn, "J2", [MATERIAL])
        o.set_solution(f["u"][lg], f["p"][lg])
        o.state("Fp_old")[:] = f["Fp_old"][eg]; o.state("eqps_old")[:] = f["eqps_old"][eg]
        Ro, Vo = o.jacobian(PRIMAL, save=False)
        Ao = o.csr(Vo).toarray()
        c13 = 13  # the centre node of the 3x3x3 block
        l_of_g = {int(g): q for q, g in enumerate(lg)}
        for b in range(nb):
            lq = l_of_g[int(ncol[nrow[node] + b])]
            want = Ao[4 * c13:4 * c13 + 4, 4 * lq:4 * lq + 4]
            assert np.abs(got[:, b, :] - want).max() < 1e-12 * np.abs(Ao).max()
        assert relerr(gotR, Ro[4 * c13:4 * c13 + 4]) < 1e-11
    a.close()
