"""Mesh parts for multi-GPU runs, in the form PUMI hands them to the reference.

The reference is element-partitioned: parts own disjoint element sets, vertices
on part boundaries are shared copies listed per neighbouring part in an order both sides agree on
(the .smb "remotes"), there are no ghost element layers, and every shared vertex has one owning
part.  A part description here is the dict gx_desc wants:

    coords [Nn,3], conn [Ne,4]      part-local ("ghost"/overlap) numbering
    node_gid [Nn] int64             global node id        (apf::makeGlobal, src/goal_disc.cpp:274)
    elem_gid [Ne] int64             global element id     (only used to seed synthetic fields)
    node_owner [Nn] int32           owning rank           (lowest rank holding the vertex)
    peer_rank [P], peer_offset [P+1], peer_nodes [...]    shared-vertex lists, ascending global id
    rank, n_ranks
"""
import numpy as np

from .synthetic import kuhn_block


def _finish(rank, n_ranks, coords, conn, node_gid, elem_gid, holders):
    """holders: dict peer_rank -> sorted local node ids shared with that peer (ascending gid)."""
    nn = len(coords)
    owner = np.full(nn, rank, dtype=np.int32)
    for q, nodes in holders.items():
        if q < rank:
            owner[nodes] = np.minimum(owner[nodes], q)
    peers = sorted(holders)
    off = np.zeros(len(peers) + 1, dtype=np.int32)
    for i, q in enumerate(peers):
        off[i + 1] = off[i] + len(holders[q])
    pn = np.concatenate([holders[q] for q in peers]).astype(np.int32) if peers else np.zeros(0, np.int32)
    return dict(rank=rank, n_ranks=n_ranks, coords=coords, conn=conn, node_gid=node_gid.astype(np.int64),
                elem_gid=elem_gid.astype(np.int64), node_owner=owner, peer_rank=np.array(peers, dtype=np.int32),
                peer_offset=off, peer_nodes=pn)


def block_part(c, grid, rank):
    """Part `rank` of a Kuhn box split into px*py*pz equal blocks.  c = cells per block edge: one int (cubic blocks of
    c^3 cells, "weak scaling": every block sees the same element size h = 1/c) or (cx, cy, cz) (the blocks of a fixed
    global mesh, "strong scaling").

    Rank r sits at block (r % Px, (r // Px) % Py, r // (Px*Py)).  Coordinates are in units of the x block edge."""
    px, py, pz = grid
    n_ranks = px * py * pz
    cx, cy, cz = (c, c, c) if np.isscalar(c) else c
    b = (rank % px, (rank // px) % py, rank // (px * py))
    g = (px * cx, py * cy, pz * cz)  # global cells per axis
    coords, conn = kuhn_block(cx, cy, cz, (b[0] * cx, b[1] * cy, b[2] * cz), cx)
    k, j, i = np.meshgrid(np.arange(cz + 1), np.arange(cy + 1), np.arange(cx + 1), indexing="ij")
    gi, gj, gk = i + b[0] * cx, j + b[1] * cy, k + b[2] * cz
    node_gid = (gi + (g[0] + 1) * (gj + (g[1] + 1) * gk)).reshape(-1).astype(np.int64)
    ck, cj, ci = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
    cell_gid = ((ci + b[0] * cx) + g[0] * ((cj + b[1] * cy) + g[1] * (ck + b[2] * cz))).reshape(-1).astype(np.int64)
    elem_gid = (6 * cell_gid[:, None] + np.arange(6)[None, :]).reshape(-1)
    li, lj, lk = i.reshape(-1), j.reshape(-1), k.reshape(-1)
    holders = {}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                nb = (b[0] + dx, b[1] + dy, b[2] + dz)
                if not (0 <= nb[0] < px and 0 <= nb[1] < py and 0 <= nb[2] < pz):
                    continue
                q = nb[0] + px * (nb[1] + py * nb[2])
                m = np.ones(len(li), dtype=bool)
                for d, l, cd in ((dx, li, cx), (dy, lj, cy), (dz, lk, cz)):
                    if d == -1:
                        m &= l == 0
                    elif d == 1:
                        m &= l == cd
                ids = np.nonzero(m)[0]
                holders[q] = ids[np.argsort(node_gid[ids], kind="stable")]
    return _finish(rank, n_ranks, coords, conn, node_gid, elem_gid, holders)


def fixture_parts(parts):
    """The reference's 4-part cube fixture (tests/golden/cube_fixture.json 'parts') as gx_desc parts.

    `remotes` are PUMI's per-peer shared-vertex lists (same order on both sides); global ids are the
    serial fixture's vertex ids."""
    out = []
    n = len(parts)
    for r, p in enumerate(parts):
        gid = np.array(p["serial_vertex"], dtype=np.int64)
        holders = {int(q): np.array(v, dtype=np.int64) for q, v in p["remotes"].items()}
        out.append(_finish(r, n, np.array(p["coords"]), np.array(p["tets"], dtype=np.int32), gid,
                           np.arange(len(p["tets"]), dtype=np.int64), holders))
    # elem_gid is not recoverable from the part files; make it unique across parts
    base = 0
    for p in out:
        p["elem_gid"] = p["elem_gid"] + base
        base += len(p["conn"])
    return out


def serial_from_parts(parts):
    """Glue parts back into one mesh (global node numbering) -- what a 1-rank run would assemble."""
    ngl = int(max(p["node_gid"].max() for p in parts)) + 1
    coords = np.zeros((ngl, 3))
    conn = []
    for p in parts:
        coords[p["node_gid"]] = p["coords"]
        conn.append(p["node_gid"][p["conn"]])
    return coords, np.concatenate(conn).astype(np.int32)
