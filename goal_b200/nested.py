"""Nested (uniformly refined) meshes without MeshAdapt -- the bookkeeping of src/goal_nested.cpp in flat arrays.

The reference builds the mesh for its adjoint solve by copying the base mesh and refining it with ma::adapt
(Nested::copy_mesh / refine_uniform, src/goal_nested.cpp:57-61, 118-131), remembers for every nested element the base
element it came from (number_elems, :41-55, the "elems" field that follows the refinement), for every new vertex
the two old vertices of the edge it split (the "nvt" tag, :63-75, :133-143) and uses them in

    Nested::set_coarse   new-vertex value = mean of the two old-vertex values          (:365-393)
    Nested::set_error    base element error = sum over its nested elements            (:395-412)

`refine_uniform` produces exactly these arrays for the regular 1:8 subdivision of a tet (4 corner tets + the inner
octahedron cut along the 1-3 / 0-2 mid-edge diagonal).  Which inner diagonal MeshAdapt picks is not reproduced; it
changes the nested elements, not the meaning of the arrays.
"""
import numpy as np


def refine_uniform(coords, tets):
    """-> dict(coords [Nn',3], tets [8 Ne,4] positively oriented, parent [8 Ne], new_vtx [(new, old0, old1)], n_old)"""
    coords = np.asarray(coords, dtype=np.float64)
    tets = np.asarray(tets, dtype=np.int64)
    nn = len(coords)
    pairs = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
    e = np.stack([np.sort(tets[:, list(p)], axis=1) for p in pairs], axis=1)  # [Ne, 6, 2]
    key = e[:, :, 0] * nn + e[:, :, 1]
    uniq, inv = np.unique(key.reshape(-1), return_inverse=True)
    mid = nn + inv.reshape(-1, 6)  # id of the vertex on each of the 6 edges
    old0, old1 = uniq // nn, uniq % nn
    co = np.concatenate([coords, 0.5 * (coords[old0] + coords[old1])])
    v = tets
    m01, m12, m02, m03, m13, m23 = (mid[:, k] for k in range(6))
    kids = [
        (v[:, 0], m01, m02, m03), (m01, v[:, 1], m12, m13), (m02, m12, v[:, 2], m23), (m03, m13, m23, v[:, 3]),
        (m01, m02, m03, m13), (m01, m02, m13, m12), (m02, m03, m13, m23), (m02, m12, m23, m13),  # octahedron, diagonal m02-m13
    ]
    t = np.stack([np.stack(k, axis=1) for k in kids], axis=1).reshape(-1, 4)  # element 8 e + k
    x = co[t]
    neg = np.linalg.det(x[:, 1:] - x[:, :1]) < 0
    t[neg, 1], t[neg, 2] = t[neg, 2].copy(), t[neg, 1].copy()
    return dict(coords=co, tets=t.astype(np.int32), parent=np.repeat(np.arange(len(tets)), 8).astype(np.int32),
                new_vtx=np.stack([nn + np.arange(len(uniq)), old0, old1], axis=1).astype(np.int32), n_old=nn)


def set_coarse(field, nested):
    """Nested::set_coarse (src/goal_nested.cpp:365-393): overwrite the new vertices' values with the mean of the two old
    vertices of their edge.  field: [Nn'] or [Nn', k] on the nested mesh (old vertices keep their ids)."""
    f = np.array(field, dtype=np.float64, copy=True)
    nv = nested["new_vtx"]
    f[nv[:, 0]] = (f[nv[:, 1]] + f[nv[:, 2]]) * 0.5
    return f


def prolong(field, nested):
    """Base-mesh vertex field -> nested mesh (linear interpolation = what the solution transfer of the refinement
    gives a P1 field): old vertices keep their value, new ones get the edge mean."""
    base = np.asarray(field, dtype=np.float64)
    f = np.zeros((len(nested["coords"]),) + base.shape[1:])
    f[:nested["n_old"]] = base
    return set_coarse(f, nested)
