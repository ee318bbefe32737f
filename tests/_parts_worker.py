"""Worker of tests/test_parts_gloo.py: one mesh part per process, gloo backend, CPU only.

Runs the product's real partition code (libgoal_b200.so host-only contexts: structure exchange,
extended rows, exchange plan, owned graph) and emulates the value exchange in numpy on
oracle-assembled per-part arrays, then checks every owned row against a serial oracle assembly."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    case, model = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import goal_b200
    from goal_b200.partition import block_part, fixture_parts, serial_from_parts
    from goal_b200.synthetic import MATERIAL, fields
    from oracle.oracle import PRIMAL, Oracle

    if case == "fixture":
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "cube_fixture.json")))
        parts = fixture_parts(fx["parts"])
    else:
        grid = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
        parts = [block_part(3, grid, r) for r in range(world)]
    assert len(parts) == world
    me = parts[rank]
    a = goal_b200.Assembler(me["coords"], me["conn"], model, [MATERIAL], device=-1, partition=me)
    a.exchange_structure(dist)

    # per-part assembly in ghost layout by the oracle, on fields seeded by global ids
    co_s, cn_s = serial_from_parts(parts)
    fs = fields(co_s, len(cn_s), strain=0.004)
    f = fields(me["coords"], len(me["conn"]), node_gid=me["node_gid"], elem_gid=me["elem_gid"], strain=0.004)
    # the synthetic u depends on coordinates; take every nodal field from the serial arrays instead
    for k in ("u", "p"):
        f[k] = fs[k][me["node_gid"]]
    o = Oracle(me["coords"], me["conn"], model, [MATERIAL])
    o.set_solution(f["u"], f["p"])
    if model == "J2":
        o.state("Fp_old")[:] = f["Fp_old"]
        o.state("eqps_old")[:] = f["eqps_old"]
    R, V = o.jacobian(PRIMAL, save=False)
    assert np.array_equal(a.rowptr, o.rowptr) and np.array_equal(a.colind, o.colind)

    # extended rows: ghost row + phantom blocks
    g = a.owned_graph()
    nb_g = (o.rowptr[1::4] - o.rowptr[0:-1:4]) // 4  # ghost blocks per node
    # emulate gx_pack_interface / gx_unpack_add_interface with the library's plan
    ext = {}  # owned node -> [4, 4*nblk_x] array
    for s, node in enumerate(g["nodes"]):
        nbx = (g["rowptr"][4 * s + 1] - g["rowptr"][4 * s]) // 4
        row = np.zeros((4, 4 * nbx))
        for i in range(4):
            row[i, :4 * nb_g[node]] = V[o.rowptr[4 * node + i]:o.rowptr[4 * node + i + 1]]
        ext[node] = row
    Rn = R.reshape(-1, 4).copy()
    plans = [a.exchange_plan(p) for p in range(a.num_peers)]
    ops, inbox = [], {}
    outbox = []
    for pl in plans:
        payload = [Rn[pl["send_nodes"]].reshape(-1)]
        for node in pl["send_nodes"]:
            payload.append(np.concatenate([V[o.rowptr[4 * node + i]:o.rowptr[4 * node + i + 1]] for i in range(4)]))
        t = torch.from_numpy(np.concatenate(payload)) if len(pl["send_nodes"]) else None
        outbox.append(t)
        n_in = 4 * len(pl["recv_nodes"]) + 16 * int(pl["recv_cnt"].sum())
        inbox[pl["rank"]] = torch.zeros(n_in, dtype=torch.float64)
        if t is not None:
            ops.append(dist.P2POp(dist.isend, t, pl["rank"]))
        if n_in:
            ops.append(dist.P2POp(dist.irecv, inbox[pl["rank"]], pl["rank"]))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for pl in plans:  # ascending peer rank, like gx_reduce_interfaces
        buf = inbox[pl["rank"]].numpy()
        nr = len(pl["recv_nodes"])
        Rn[pl["recv_nodes"]] += buf[:4 * nr].reshape(nr, 4)
        pos, mpos = 4 * nr, 0
        for s, node in enumerate(pl["recv_nodes"]):
            nb = int(pl["recv_cnt"][s])
            blk = buf[pos:pos + 16 * nb].reshape(4, nb, 4)
            m = pl["recv_map"][mpos:mpos + nb]
            for j in range(nb):
                ext[node][:, 4 * m[j]:4 * m[j] + 4] += blk[:, j, :]
            pos += 16 * nb
            mpos += nb

    # serial truth
    os_ = Oracle(co_s, cn_s, model, [MATERIAL])
    os_.set_solution(fs["u"], fs["p"])
    if model == "J2":
        os_.state("Fp_old")[:] = fs["Fp_old"][np.concatenate([p["elem_gid"] for p in parts])] if case != "fixture" else fs["Fp_old"]
        os_.state("eqps_old")[:] = fs["eqps_old"][np.concatenate([p["elem_gid"] for p in parts])] if case != "fixture" else fs["eqps_old"]
    Rs, Vs = os_.jacobian(PRIMAL, save=False)
    scale = np.abs(Vs).max()
    worst = 0.0
    for s, node in enumerate(g["nodes"]):
        G = int(me["node_gid"][node])
        assert np.abs(Rn[node] - Rs[4 * G:4 * G + 4]).max() < 1e-12 * np.abs(Rs).max()
        for i in range(4):
            cols = g["col_gid"][g["rowptr"][4 * s + i]:g["rowptr"][4 * s + i + 1]]
            ref = dict(zip(os_.colind[os_.rowptr[4 * G + i]:os_.rowptr[4 * G + i + 1]].tolist(),
                           Vs[os_.rowptr[4 * G + i]:os_.rowptr[4 * G + i + 1]].tolist()))
            got = ext[node][i]
            assert sorted(cols.tolist()) == sorted(ref)
            worst = max(worst, max(abs(got[k] - ref[int(c)]) for k, c in enumerate(cols)))
    assert worst < 1e-12 * scale, worst
    # ---- the owned matrix in Tpetra's local layout (gx_owned_tpetra_graph): the restated makeColMap rule
    tp = a.owned_tpetra_graph()
    no = len(g["nodes"])
    owned_gids = me["node_gid"][g["nodes"]]
    assert tp["n_owned"] == no and np.array_equal(tp["colmap"][:no], owned_gids)  # (1) the owned dofs, in owned_map order
    pairs = [None] * world
    dist.all_gather_object(pairs, (me["node_gid"].tolist(), me["node_owner"].tolist()))
    owner_of = {}
    for gids, owners in pairs:
        owner_of.update(zip(gids, owners))
    rem = tp["colmap"][no:].tolist()
    assert len(set(rem)) == len(rem) and not set(rem) & set(owned_gids.tolist())
    assert all(owner_of[x] != rank for x in rem)
    assert rem == sorted(rem, key=lambda x: (owner_of[x], x))                     # (2) remotes by owning rank, then by global id
    assert np.array_equal(tp["rowptr"], g["rowptr"])
    for s in range(4 * no):
        ci = tp["colind"][tp["rowptr"][s]:tp["rowptr"][s + 1]]
        assert np.all(np.diff(ci) > 0)                                           # (3) local column indices ascending in every row
        gdof = 4 * tp["colmap"][ci // 4] + ci % 4
        assert sorted(gdof.tolist()) == sorted(g["col_gid"][g["rowptr"][s]:g["rowptr"][s + 1]].tolist())
    owned_total = torch.tensor([len(g["nodes"])])
    dist.all_reduce(owned_total)
    assert int(owned_total) == len(co_s)
    dist.barrier()
    if rank == 0:
        print(f"OK {case} {model} world={world} worst={worst / scale:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
