"""the C++ reader of the reference's mesh inputs (goal_b200/host/goal_mesh_io.hpp: .smb parts,
.dmg model, assoc file) -- no PUMI.  (1) a Kuhn cube is ENCODED here in the .smb / .dmg / assoc formats and read
back; (2) where the reference tree is present (the build container) its own cube fixtures are read and must equal
the committed JSON fixture that the Python decoder tests/golden/make_cube_fixture.py produced."""
import itertools
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from goal_b200.synthetic import kuhn_cube

REF = "/root/reference/test/mesh/cube"


def _dump(*args):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "goal_b200", "host"), "-s", "../gx_meshdump"])
    r = subprocess.run([os.path.join(ROOT, "goal_b200", "gx_meshdump"), *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


def _encode_cube(tmp, n):
    """unit Kuhn cube as cube.smb / cube.dmg / cube.txt; model: 8 vertices, 12 edges, 6 faces, 1 region."""
    co, cn = kuhn_cube(n)
    edges, tris = {}, {}
    e_id = lambda a, b: edges.setdefault((min(a, b), max(a, b)), len(edges))
    t2e, tet2t = [], []
    for t in cn:
        ft = []
        for f in ((0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)):
            v = [int(t[i]) for i in f]
            key = tuple(sorted(v))
            if key not in tris:
                tris[key] = len(tris)
                t2e.append([e_id(v[0], v[1]), e_id(v[1], v[2]), e_id(v[2], v[0])])
            ft.append(tris[key])
        tet2t.append(ft)
    e2v = [k for k, _ in sorted(edges.items(), key=lambda kv: kv[1])]
    tri_keys = [k for k, _ in sorted(tris.items(), key=lambda kv: kv[1])]
    # model tags: vertices 1..8, edges 11..22, faces 31..36 (xmin xmax ymin ymax zmin zmax), region 41
    corners = list(itertools.product((0.0, 1.0), repeat=3))
    vtag = {c: 1 + i for i, c in enumerate(corners)}
    medges = [(a, b) for a, b in itertools.combinations(corners, 2) if sum(x != y for x, y in zip(a, b)) == 1]
    etag = {e: 11 + i for i, e in enumerate(medges)}
    ftag = {(ax, val): 31 + 2 * ax + int(val) for ax in range(3) for val in (0.0, 1.0)}

    def classify(pts):
        fixed = [(ax, v) for ax in range(3) for v in (0.0, 1.0) if all(abs(p[ax] - v) < 1e-12 for p in pts)]
        if len(fixed) == 0:
            return 41, 3
        if len(fixed) == 1:
            return ftag[fixed[0]], 2
        if len(fixed) == 2:
            for (a, b), t in etag.items():
                if all(a[ax] == v and b[ax] == v for ax, v in fixed):
                    return t, 1
        return vtag[tuple(float(pts[0][ax]) for ax in range(3))], 0

    out = struct.pack(">12I", 0, 5, 3, 1, len(co), len(e2v), len(tri_keys), 0, 0, 0, 0, len(cn))
    out += np.array(e2v, dtype=">u4").tobytes() + np.array(t2e, dtype=">u4").tobytes() + np.array(tet2t, dtype=">u4").tobytes()
    out += co.astype(">f8").tobytes() + np.zeros((len(co), 2), dtype=">f8").tobytes()
    out += struct.pack(">I", 0)  # no remotes
    for ents in ([[i] for i in range(len(co))], e2v, tri_keys, cn.tolist()):
        out += np.array([classify(co[list(e)]) for e in ents], dtype=">u4").tobytes()
    open(os.path.join(tmp, "cube.smb"), "wb").write(out)
    with open(os.path.join(tmp, "cube.dmg"), "w") as f:
        f.write("1 6 12 8\n0 0 0\n1 1 1\n")
        for c, t in vtag.items():
            f.write(f"{t} {c[0]} {c[1]} {c[2]}\n")
        for (a, b), t in etag.items():
            f.write(f"{t} {vtag[a]} {vtag[b]}\n")
        for (ax, v), t in ftag.items():
            fe = [et for (a, b), et in etag.items() if a[ax] == v and b[ax] == v]
            f.write(f"{t} 1\n {len(fe)}\n" + "".join(f"  {e} 1\n" for e in fe))
        f.write("41 1\n 6\n" + "".join(f"  {t} 1\n" for t in ftag.values()))
    with open(os.path.join(tmp, "cube.txt"), "w") as f:
        f.write(f"node set xmin 1\n2 {ftag[(0, 0.0)]}\nnode set edge 1\n1 {etag[medges[0]]}\n")
        f.write(f"side set ymax 1\n2 {ftag[(1, 1.0)]}\nelem set cube 1\n3 41\n")
    return co, cn, medges[0]


def test_encoded_cube_round_trip(tmp_path):
    n = 3
    co, cn, medge = _encode_cube(str(tmp_path), n)
    d = _dump(*(str(tmp_path / f) for f in ("cube.smb", "cube.dmg", "cube.txt")))
    assert d["dim"] == 3 and d["nparts"] == 1 and d["peers"] == [] and d["remotes"] == []
    assert np.array_equal(np.array(d["coords"]).reshape(-1, 3), co)
    tets = np.array(d["tets"]).reshape(-1, 4)
    assert np.array_equal(np.sort(tets, 1), np.sort(cn, 1))  # same vertex sets, element order kept
    x = co[tets]
    assert (np.linalg.det(x[:, 1:] - x[:, :1]) > 0).all()  # oriented
    assert d["node_sets"]["xmin"] == np.nonzero(co[:, 0] == 0.0)[0].tolist()
    on_edge = [i for i in range(len(co)) if all(co[i][ax] == medge[0][ax] for ax in range(3) if medge[0][ax] == medge[1][ax])]
    assert d["node_sets"]["edge"] == on_edge and len(on_edge) == n + 1
    sides = np.array(d["side_sets"]["ymax"]).reshape(-1, 3)
    assert len(sides) == 2 * n * n and (co[sides][:, :, 1] == 1.0).all()
    assert d["elem_sets"]["cube"] == list(range(len(cn)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_cube_files_match_committed_fixture(cube):
    d = _dump(f"{REF}/cube-serial0.smb", f"{REF}/cube.dmg", f"{REF}/cube.txt")
    assert np.array_equal(np.array(d["coords"]).reshape(-1, 3), cube["coords"])
    assert np.array_equal(np.array(d["tets"]).reshape(-1, 4), cube["tets"])
    assert d["node_sets"] == cube["node_sets"] and d["elem_sets"] == cube["elem_sets"]
    assert {k: np.array(v).reshape(-1, 3).tolist() for k, v in d["side_sets"].items()} == cube["side_sets"]
    for p, part in enumerate(cube["parts"]):
        dp = _dump(f"{REF}/cube{p}.smb")
        assert dp["nparts"] == 4
        assert np.array_equal(np.array(dp["coords"]).reshape(-1, 3), np.array(part["coords"]))
        assert np.array_equal(np.array(dp["tets"]).reshape(-1, 4), np.array(part["tets"]))
        assert {str(q): r for q, r in zip(dp["peers"], dp["remotes"])} == part["remotes"]
