"""Worker of tests/test_gpu_parts.py::test_nccl_reduce_interfaces_two_gpus: one part per GPU,
gx_comm_init + gx_reduce_interfaces over NCCL, owned rows checked against a serial oracle assembly."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import goal_b200
    from goal_b200.partition import block_part
    from goal_b200.synthetic import MATERIAL
    from test_gpu_parts import _check_owned, _serial_truth
    grid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    parts = [block_part(5, grid, r) for r in range(world)]
    me = parts[rank]
    for model in ("J2", "neohookean"):
        o, fs, Rs, Vs = _serial_truth(parts, model)
        a = goal_b200.Assembler(me["coords"], me["conn"], model, [MATERIAL], device=local, partition=me)
        a.comm_init_torch(dist)
        a.set_solution(fs["u"][me["node_gid"]], fs["p"][me["node_gid"]])
        if model == "J2":
            off = sum(len(p["conn"]) for p in parts[:rank])
            a.set_state("Fp_old", fs["Fp_old"][off:off + a.ne]); a.set_state("eqps_old", fs["eqps_old"][off:off + a.ne])
        for _ in range(2):  # repeated passes must give the same bits
            a.jacobian(goal_b200.PRIMAL, save=False, out=False)
            a.reduce_interfaces(3)
            R1, V1, _ = a.fetch_owned()
        a.jacobian(goal_b200.PRIMAL, save=False, out=False)
        a.reduce_interfaces(3)
        R2, V2, _ = a.fetch_owned()
        assert np.array_equal(R1, R2) and np.array_equal(V1, V2)
        _check_owned(a, me, o, Rs, Vs)
        # the same pass with the exchange fused in (option "overlap": interface patches first, pack -> NCCL -> unpack on a
        # second stream while the interior patches run): bit-identical owned rows, no separate gx_reduce_interfaces
        a.set_option("overlap", 3)
        a.jacobian(goal_b200.PRIMAL, save=False, out=False)
        R3, V3, _ = a.fetch_owned()
        assert np.array_equal(R1, R3) and np.array_equal(V1, V3)
        assert a.last_timing()["launches"] >= 3  # element records, interface patches, interior patches (+ pack / unpack)
        Rt, Vt, tp = a.fetch_owned_tpetra()
        assert np.array_equal(Rt, R3) and np.isclose(np.abs(Vt).sum(), np.abs(V3).sum(), rtol=1e-13)
        a.set_option("overlap", 0)
        # Disc::add_soln + apf::synchronize over NCCL (gx_add_solution, gx_sync_solution): right on owned nodes,
        # garbage on copies, the owner's value wins everywhere
        dug = 1e-3 * np.random.RandomState(9).randn(len(fs["u"]), 4)
        du = dug[me["node_gid"]].copy()
        du[me["node_owner"] != rank] = 123.0
        a.add_solution(du)
        a.sync_solution()
        u, pr = a.get_solution()
        assert np.array_equal(u, fs["u"][me["node_gid"]] + dug[me["node_gid"], :3])
        assert np.array_equal(pr, fs["p"][me["node_gid"]] + dug[me["node_gid"], 3])
        # functional derivative reduced like gather_dMdu: owned rows equal the serial dMdu
        from oracle.oracle import Oracle
        a.set_solution(fs["u"][me["node_gid"]], fs["p"][me["node_gid"]])
        J, _ = a.functional("avg vm", with_dMdu=True)
        a.reduce_interfaces(4)
        d = a.fetch_dMdu().reshape(-1, 4)
        Jg = a.allreduce_sum(np.array([J]))[0]
        Jo, do = o.functional("avg vm", with_dMdu=True)
        own = me["node_owner"] == rank
        assert abs(Jg - Jo) < 1e-12 * abs(Jo)
        assert np.abs(d[own] - do.reshape(-1, 4)[me["node_gid"][own]]).max() < 1e-12 * np.abs(do).max()
        s = a.allreduce_sum(np.array([float(rank + 1)]))
        assert s[0] == world * (world + 1) / 2
        a.close()
    dist.barrier()
    if rank == 0:
        print(f"OK nccl world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
