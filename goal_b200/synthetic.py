"""Synthetic structured tet meshes and fields.

Host-side numpy generators shared by the tests and bench.py.  The mesh is the
Kuhn/Freudenthal subdivision of an N^3-cell unit cube: for each permutation pi
of the axes the tet (v0, v0+e_pi1, v0+e_pi1+e_pi2, v0+e_1+e_2+e_3); two
vertices are swapped on odd permutations so det J > 0.  Node id
i + (N+1)(j + (N+1)k); element id 6*cell + perm.
"""
import itertools

import numpy as np

# reference material set used by every example (example/primal/J2_uniaxial_3D.yaml:18-22)
MATERIAL = (1000.0, 0.25, 100.0, 10.0, 1.0)  # E, nu, K, Y, c0

_PERMS = list(itertools.permutations(range(3)))


def _parity(p):
    return sum(p[i] > p[j] for i in range(3) for j in range(i + 1, 3)) % 2


def kuhn_block(nx, ny, nz, origin=(0, 0, 0), global_cells=None):
    """Tets of an nx*ny*nz block of cells.  Returns (coords [Nn,3], conn [Ne,4] int32).

    `origin` (cell offset) and `global_cells` (N of the full cube) place the block
    inside a larger cube so that partitioned runs see the same geometry.
    """
    n = global_cells if global_cells is not None else max(nx, ny, nz)
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    coords = np.stack([(i + origin[0]) / n, (j + origin[1]) / n, (k + origin[2]) / n], -1).reshape(-1, 3)
    ck, cj, ci = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (ci + sx * cj + sy * ck).reshape(-1)
    stride = np.array([1, sx, sy])
    conn = np.empty((base.size, 6, 4), dtype=np.int64)
    for q, perm in enumerate(_PERMS):
        v = [np.zeros_like(base)]
        for ax in perm:
            v.append(v[-1] + stride[ax])
        v = [base + t for t in v]
        if _parity(perm):
            v[1], v[2] = v[2], v[1]
        conn[:, q, :] = np.stack(v, -1)
    conn = conn.reshape(-1, 4).astype(np.int32)
    return np.ascontiguousarray(coords), np.ascontiguousarray(conn)


def kuhn_cube(n):
    return kuhn_block(n, n, n, (0, 0, 0), n)


def fields(coords, n_elems, seed_nodal=20260417, seed_elem=20260418, node_gid=None, elem_gid=None, strain=0.02):
    """u, p, Fp_old, eqps_old and adjoint fields of BASELINE.md 4.

    The random part is a counter-based hash of the global node / element id so a
    partitioned mesh sees exactly the values of the serial one.
    """
    x = coords
    nn = len(x)
    ngid = np.arange(nn, dtype=np.uint64) if node_gid is None else node_gid.astype(np.uint64)
    egid = np.arange(n_elems, dtype=np.uint64) if elem_gid is None else elem_gid.astype(np.uint64)

    def uni(ids, seed, stream, ncomp):
        # splitmix64 on (id, component, stream, seed) -> U[0,1)
        with np.errstate(over="ignore"):
            k = (ids[:, None] * np.uint64(ncomp) + np.arange(ncomp, dtype=np.uint64)[None, :])
            z = k + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    sym = lambda ids, seed, stream, nc: 2.0 * uni(ids, seed, stream, nc) - 1.0
    tp = 2.0 * np.pi
    u = strain * np.stack([x[:, 0], -0.4 * x[:, 1], -0.4 * x[:, 2]], -1)
    u += 2e-3 * np.stack([np.sin(tp * x[:, 1]), np.sin(tp * x[:, 2]), np.sin(tp * x[:, 0])], -1)
    u += 1e-4 * sym(ngid, seed_nodal, 1, 3)
    p = sym(ngid, seed_nodal, 2, 1)[:, 0]
    Fp_old = np.tile(np.eye(3).reshape(1, 9), (n_elems, 1)) + 1e-3 * sym(egid, seed_elem, 3, 9)
    eqps_old = 0.01 * uni(egid, seed_elem, 4, 1)[:, 0]
    zu = 1e-2 * sym(ngid, seed_nodal, 5, 3)
    zp = 1e-2 * sym(ngid, seed_nodal, 6, 1)[:, 0]
    zpc = 1e-2 * sym(ngid, seed_nodal, 7, 1)[:, 0]
    return dict(u=u, p=p, Fp_old=Fp_old, eqps_old=eqps_old, zu_diff=zu, zp_diff=zp, zp_coarse=zpc)
