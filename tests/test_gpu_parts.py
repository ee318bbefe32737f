"""Mesh parts on the GPU: the interface exchange kernels with a host transport (single GPU, several
contexts in one process) and the NCCL path (needs >= 2 GPUs; `gpurun --gpus 2`)."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr
from goal_b200.synthetic import MATERIAL, fields

pytestmark = pytest.mark.gpu


def _serial_truth(parts, model):
    from goal_b200.partition import serial_from_parts
    from oracle.oracle import PRIMAL, Oracle
    co, cn = serial_from_parts(parts)
    fs = fields(co, len(cn), strain=0.004)
    o = Oracle(co, cn, model, [MATERIAL])
    o.set_solution(fs["u"], fs["p"])
    if model == "J2":
        o.state("Fp_old")[:] = fs["Fp_old"]
        o.state("eqps_old")[:] = fs["eqps_old"]
    R, V = o.jacobian(PRIMAL, save=False)
    return o, fs, R, V


def _check_owned(a, part, o, Rs, Vs):
    R, V, g = a.fetch_owned()
    scale = np.abs(Vs).max()
    for s, node in enumerate(g["nodes"]):
        G = int(part["node_gid"][node])
        assert np.abs(R[4 * s:4 * s + 4] - Rs[4 * G:4 * G + 4]).max() < 1e-12 * np.abs(Rs).max()
        for i in range(4):
            lo, hi = g["rowptr"][4 * s + i], g["rowptr"][4 * s + i + 1]
            ref = dict(zip(o.colind[o.rowptr[4 * G + i]:o.rowptr[4 * G + i + 1]].tolist(),
                           Vs[o.rowptr[4 * G + i]:o.rowptr[4 * G + i + 1]].tolist()))
            cols = g["col_gid"][lo:hi]
            assert sorted(cols.tolist()) == sorted(ref)
            assert max(abs(V[lo + k] - ref[int(c)]) for k, c in enumerate(cols)) < 1e-12 * scale


def _check_owned_tpetra(a, part, o, Rs, Vs):
    """gx_fetch_owned_tpetra: the owned rows in Tpetra's local layout (column map = owned dofs in owned order, then the
    remote dofs by owning rank and global id; rows ascending) hold the serial operator's values."""
    R, V, tp = a.fetch_owned_tpetra()
    own = np.nonzero(part["node_owner"] == part["rank"])[0]
    gid = part["node_gid"][own]
    assert tp["n_owned"] == len(own) and np.array_equal(tp["colmap"][:len(own)], gid)
    scale = np.abs(Vs).max()
    for s, G in enumerate(gid):
        assert np.abs(R[4 * s:4 * s + 4] - Rs[4 * G:4 * G + 4]).max() < 1e-12 * np.abs(Rs).max()
        for i in range(4):
            lo, hi = tp["rowptr"][4 * s + i], tp["rowptr"][4 * s + i + 1]
            ci = tp["colind"][lo:hi]
            assert np.all(np.diff(ci) > 0)
            gdof = 4 * tp["colmap"][ci // 4] + ci % 4
            ref = dict(zip(o.colind[o.rowptr[4 * G + i]:o.rowptr[4 * G + i + 1]].tolist(), Vs[o.rowptr[4 * G + i]:o.rowptr[4 * G + i + 1]].tolist()))
            assert sorted(gdof.tolist()) == sorted(ref)
            assert max(abs(V[lo + k] - ref[int(c)]) for k, c in enumerate(gdof)) < 1e-12 * scale


@pytest.mark.parametrize("case,model,kernel", [("fixture", "J2", 0), ("blocks", "neohookean", 0), ("fixture", "J2", 1), ("blocks", "J2", 1)])
def test_interface_exchange_host_transport(cube, case, model, kernel):
    """4 parts = 4 contexts on one GPU; the packed interface rows are handed from context to context
    (what an MPI host would do) and every owned row must equal the serial assembly."""
    import goal_b200
    from goal_b200.partition import block_part, fixture_parts
    parts = fixture_parts(cube["parts"]) if case == "fixture" else [block_part(4, (2, 2, 1), r) for r in range(4)]
    o, fs, Rs, Vs = _serial_truth(parts, model)
    egid = np.concatenate([p["elem_gid"] for p in parts])
    A = []
    for p in parts:
        a = goal_b200.Assembler(p["coords"], p["conn"], model, [MATERIAL], partition=p)
        a.set_option("kernel", kernel)
        A.append(a)
    for r, a in enumerate(A):  # structure exchange, host transport
        for pi in range(a.num_peers):
            q = int(parts[r]["peer_rank"][pi])
            pj = list(parts[q]["peer_rank"]).index(r)
            A[q].struct_unpack(pj, a.struct_pack(pi))
    off = 0
    for r, a in enumerate(A):
        a.struct_finalize()
        p = parts[r]
        a.set_solution(fs["u"][p["node_gid"]], fs["p"][p["node_gid"]])
        if model == "J2":
            sl = slice(off, off + a.ne)
            a.set_state("Fp_old", fs["Fp_old"][sl]); a.set_state("eqps_old", fs["eqps_old"][sl])
        off += a.ne
        a.jacobian(goal_b200.PRIMAL, save=False, out=False)
        # ghost-layout fetch still returns the reference layout of this part
        Rg, Vg = a.fetch()
        assert Vg.shape == (a.nnz,)
    L = A[0].L

    def exchange(what):
        for q, aq in enumerate(A):  # owner side, peers ascending
            for pj in range(aq.num_peers):
                r = int(parts[q]["peer_rank"][pj])
                pi = list(parts[r]["peer_rank"]).index(q)
                buf = C.c_void_p()
                assert L.gx_pack_interface(A[r].h, pi, what, C.byref(buf)) == 0
                sb, rb = C.c_int64(), C.c_int64()
                L.gx_interface_bytes(aq.h, pj, what, C.byref(sb), C.byref(rb))
                if rb.value:
                    assert L.gx_unpack_add_interface(aq.h, pj, what, buf) == 0

    exchange(1)
    exchange(2)
    for r, a in enumerate(A):
        _check_owned(a, parts[r], o, Rs, Vs)
        _check_owned_tpetra(a, parts[r], o, Rs, Vs)
    # functional + gather_dMdu (src/goal_sol_info.cpp:37-39): part values add up, owned dMdu rows equal the serial ones
    Jo, do = o.functional("avg vm", with_dMdu=True)
    Js = [a.functional("avg vm", with_dMdu=True)[0] for a in A]
    exchange(4)
    assert abs(sum(Js) - Jo) < 1e-12 * abs(Jo)
    for r, a in enumerate(A):
        own = parts[r]["node_owner"] == r
        d = a.fetch_dMdu().reshape(-1, 4)
        assert np.abs(d[own] - do.reshape(-1, 4)[parts[r]["node_gid"][own]]).max() < 1e-12 * np.abs(do).max()
    # Disc::add_soln + apf::synchronize (src/goal_disc.cpp:398-422): every part adds an increment that is right on
    # the nodes it owns and garbage on its copies; after the owner -> copies push all parts hold the global field
    ngl = len(fs["u"])
    dug = 1e-3 * np.random.RandomState(9).randn(ngl, 4)
    for r, a in enumerate(A):
        p = parts[r]
        du = dug[p["node_gid"]].copy()
        du[p["node_owner"] != r] = 123.0
        a.add_solution(du)
    for q, aq in enumerate(A):  # owner q packs for each peer, the peer overwrites its copies
        for pj in range(aq.num_peers):
            r = int(parts[q]["peer_rank"][pj])
            pi = list(parts[r]["peer_rank"]).index(q)
            buf, nb = C.c_void_p(), C.c_int64()
            assert L.gx_pack_solution(aq.h, pj, C.byref(buf), C.byref(nb)) == 0
            if nb.value:
                assert L.gx_unpack_solution(A[r].h, pi, buf) == 0
    for r, a in enumerate(A):
        p = parts[r]
        u, pr = a.get_solution()
        assert np.array_equal(u, fs["u"][p["node_gid"]] + dug[p["node_gid"], :3])
        assert np.array_equal(pr, fs["p"][p["node_gid"]] + dug[p["node_gid"], 3])
        a.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_nccl_reduce_interfaces_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 4 if torch.cuda.device_count() >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK nccl" in r.stdout
