#!/usr/bin/env python3
"""Decode the reference's PUMI cube fixtures into a plain JSON fixture.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_cube_fixture.py

Inputs (reference data files, read-only):
    /root/reference/test/mesh/cube/cube-serial0.smb   51 verts / 132 tets, 1 part
    /root/reference/test/mesh/cube/cube{0..3}.smb     the same mesh split in 4 parts
    /root/reference/test/mesh/cube/cube.dmg           geometric model (text)
    /root/reference/test/mesh/cube/cube.txt           set association file

Output: tests/golden/cube_fixture.json with
    serial: coords, tets (positively oriented), node_sets{name: [vertex ids]},
            side_sets{name: [[v0,v1,v2], ...]}, elem_sets{name: [tet ids]}
    parts[4]: coords, tets, remotes{peer: [local vertex ids]} (same order on
              both sides of a part boundary), serial_vertex (local -> serial id)

.smb layout (SCOREC/core MDS "version 5", big-endian), as decoded for SURVEY §8(c):
    u32 magic=0, version, dim, nparts; u32 counts[8] (vert, edge, tri, quad, hex,
    prism, pyramid, tet); edge->vert (2 u32), tri->edge (3), tet->tri (4);
    coords nv*3 f64; params nv*2 f64; remotes: np, peers[np], counts[np], then the
    per-peer local vertex lists; (model_tag, model_dim) per entity.
The tet vertex *sets* come from the downward adjacencies; the canonical MDS
vertex order is not needed (a permutation of an element's nodes changes nothing
but round-off), so each tet is just oriented to positive volume.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/test/mesh/cube"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cube_fixture.json")


def decode_smb(path):
    b = open(path, "rb").read()
    hdr = np.frombuffer(b, dtype=">u4", count=12).astype(np.int64)
    off = 48
    magic, ver, dim, nparts = hdr[:4]
    nv, ne, nt, nq, nh, npz, npy, ntet = hdr[4:]
    assert magic == 0 and ver == 5 and dim == 3 and nq == nh == npz == npy == 0

    def u32(n):
        nonlocal off
        a = np.frombuffer(b, dtype=">u4", count=n, offset=off).astype(np.int64)
        off += 4 * n
        return a

    def f64(n):
        nonlocal off
        a = np.frombuffer(b, dtype=">f8", count=n, offset=off).astype(np.float64)
        off += 8 * n
        return a

    e2v = u32(ne * 2).reshape(ne, 2)
    t2e = u32(nt * 3).reshape(nt, 3)
    tet2t = u32(ntet * 4).reshape(ntet, 4)
    xyz = f64(nv * 3).reshape(nv, 3)
    f64(nv * 2)  # parametric coords, unused
    npeers = int(u32(1)[0])
    peers = u32(npeers)
    counts = u32(npeers)
    remotes = {int(p): u32(int(c)).tolist() for p, c in zip(peers, counts)}
    cls = {}
    for name, n in (("vert", nv), ("edge", ne), ("tri", nt), ("tet", ntet)):
        cls[name] = u32(2 * n).reshape(n, 2)  # (model tag, model dim)

    def tri_verts(t):
        e0, e1, _ = t2e[t]
        a, c = set(e2v[e0].tolist()), set(e2v[e1].tolist())
        v1 = (a & c).pop()
        return [(a - {v1}).pop(), v1, (c - {v1}).pop()]

    tris = [tri_verts(t) for t in range(nt)]
    tets = []
    for k in range(ntet):
        tv = list(tris[tet2t[k][0]])
        extra = set(tris[tet2t[k][1]]) - set(tv)
        assert len(extra) == 1
        tv.append(extra.pop())
        x = xyz[tv]
        if np.linalg.det(x[1:] - x[0]) < 0:
            tv[1], tv[2] = tv[2], tv[1]
        tets.append(tv)
    return dict(xyz=xyz, tets=np.array(tets), tris=np.array(tris), cls=cls,
                remotes=remotes, nparts=int(nparts))


def parse_dmg(path):
    tok = open(path).read().split()
    it = iter(tok)
    nreg, nface, nedge, nvert = (int(next(it)) for _ in range(4))
    for _ in range(6):
        next(it)  # bounding box
    for _ in range(nvert):
        for _ in range(4):
            next(it)
    edges = {}
    for _ in range(nedge):
        tag, a, b = int(next(it)), int(next(it)), int(next(it))
        edges[tag] = (a, b)
    faces = {}
    for _ in range(nface):
        tag, nloops = int(next(it)), int(next(it))
        fe = []
        for _ in range(nloops):
            n = int(next(it))
            for _ in range(n):
                fe.append(int(next(it)))
                next(it)  # direction
        faces[tag] = fe
    return edges, faces


def parse_assoc(path):
    sets = {"node set": {}, "side set": {}, "elem set": {}}
    lines = [l.strip() for l in open(path) if l.strip()]
    i = 0
    while i < len(lines):
        kind = lines[i][:8]
        name, n = lines[i][8:].split()
        ents = []
        for j in range(int(n)):
            d, t = lines[i + 1 + j].split()
            ents.append((int(d), int(t)))
        sets[kind][name] = ents
        i += 1 + int(n)
    return sets


def main():
    edges, faces = parse_dmg(f"{REF}/cube.dmg")
    assoc = parse_assoc(f"{REF}/cube.txt")

    def closure(dim, tag):
        """model entities (dim, tag) in the closure of a model face."""
        out = {(dim, tag)}
        if dim == 2:
            for e in faces[tag]:
                out.add((1, e))
                out.update((0, v) for v in edges[e])
        return out

    m = decode_smb(f"{REF}/cube-serial0.smb")
    nv = len(m["xyz"])
    serial = dict(coords=m["xyz"].tolist(), tets=m["tets"].tolist(),
                  node_sets={}, side_sets={}, elem_sets={})
    vcls = [(int(d), int(t)) for t, d in m["cls"]["vert"]]
    for name, ents in assoc["node set"].items():
        cl = set()
        for d, t in ents:
            cl |= closure(d, t)
        serial["node_sets"][name] = [v for v in range(nv) if vcls[v] in cl]
    for name, ents in assoc["side set"].items():
        want = set(ents)
        serial["side_sets"][name] = [m["tris"][t].tolist() for t in range(len(m["tris"]))
                                     if (int(m["cls"]["tri"][t][1]), int(m["cls"]["tri"][t][0])) in want]
    for name, ents in assoc["elem set"].items():
        want = set(ents)
        serial["elem_sets"][name] = [e for e in range(len(m["tets"]))
                                     if (int(m["cls"]["tet"][e][1]), int(m["cls"]["tet"][e][0])) in want]

    # sanity: what SURVEY §8(c) recorded
    vol = sum(np.linalg.det(m["xyz"][t[1:]] - m["xyz"][t[0]]) / 6 for t in m["tets"])
    assert nv == 51 and len(m["tets"]) == 132 and abs(vol - 1.0) < 1e-14
    for nm, ax, val in (("xmin", 0, 0.0), ("xmax", 0, 1.0), ("ymin", 1, 0.0),
                        ("zmin", 2, 0.0), ("zmax", 2, 1.0)):
        geo = [v for v in range(nv) if abs(m["xyz"][v][ax] - val) < 1e-12]
        assert geo == serial["node_sets"][nm] and len(geo) == 13, nm

    key = {tuple(np.round(x, 12)): i for i, x in enumerate(m["xyz"])}
    parts = []
    for p in range(4):
        mp = decode_smb(f"{REF}/cube{p}.smb")
        assert mp["nparts"] == 4
        parts.append(dict(coords=mp["xyz"].tolist(), tets=mp["tets"].tolist(),
                          remotes={str(k): v for k, v in sorted(mp["remotes"].items())},
                          serial_vertex=[key[tuple(np.round(x, 12))] for x in mp["xyz"]]))
    assert sum(len(p["tets"]) for p in parts) == 132
    json.dump(dict(serial=serial, parts=parts), open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
