#!/usr/bin/env python3
"""bench.py -- mixed P1/P1 tet residual+Jacobian assembly throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--model J2|neohookean]
    python bench.py --impl reference ...      # the reference algorithm on the host cores (oracle)

Workload (BASELINE.json configs[4]): synthetic structured Kuhn tet cube, J2,
E=1000 nu=0.25 K=100 Y=10 c0=1, seeded fields.  Each GPU holds a C^3-cell block (default 128^3 =
12,582,912 tets); with N GPUs the blocks tile a (Px*C, Py*C, Pz*C) box, so 8 GPUs hold the
256^3-cell, 100,663,296-tet cube of BASELINE.md 4 ("weak" scaling: per-GPU work is fixed).

A step = one Jacobian pass = zero R and values + residual & Jacobian of every element + J2 state
save (+ interface reduction when N > 1)  ==  Primal::compute_jacob minus BCs (src/goal_primal.cpp:98-104).
  value : Melem/s, inputs resident in HBM, results left in HBM
  e2e   : same pass through the C-ABI with HOST buffers: u,p copied in, R and CRS values copied out
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mixed P1/P1 tet residual+Jacobian assembly Melem/s"
# algorithmic HBM bytes per element of the Jacobian pass:
#   16 conn + 16 scatter map + (1/6)(24 coords + 32 u,p + 32 R) + state in + state out + 8 * 40 nnz
B_ALG = {"J2": 16 + 16 + (24 + 32 + 32) / 6 + 80 + 152 + 320, "neohookean": 16 + 16 + (24 + 32 + 32) / 6 + 0 + 72 + 320}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=128, help="cells per side of each GPU's block")
    ap.add_argument("--model", default="J2", choices=["J2", "neohookean"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cells", type=int, default=44, help="cells per side of each host thread's sample block")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (gx_set_option), repeatable")
    return ap.parse_args()


def grid_of(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


# ---------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's FAD-16 assembly; the reference itself
# needs Trilinos + SCOREC and cannot be built here).  One mesh part per host thread, like the
# reference's rank-per-core model; each thread assembles its own part into private arrays.
# This is the one place bench.py executes oracle/.
# ---------------------------------------------------------------------------
def cpu_reference_rate(model, cells, threads=None, repeats=1):
    from goal_b200.synthetic import MATERIAL, fields, kuhn_cube
    from oracle.oracle import PRIMAL, Oracle
    threads = threads or os.cpu_count() or 1
    co, cn = kuhn_cube(cells)
    f = fields(co, len(cn))
    parts = []
    for _ in range(threads):
        o = Oracle(co, cn, model, [MATERIAL])
        o.set_solution(f["u"], f["p"])
        if model == "J2":
            o.state("Fp_old")[:] = f["Fp_old"]
            o.state("eqps_old")[:] = f["eqps_old"]
        parts.append((o, np.zeros(4 * o.nn), np.zeros(o.nnz)))
    times = []
    for _ in range(repeats):
        def work(t):
            o, R, V = parts[t]
            R[:] = 0.0
            V[:] = 0.0
            o.jacobian(PRIMAL, save=True, R=R, values=V)
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        times.append(time.perf_counter() - t0)
    ne = len(cn) * threads
    dt = statistics.median(times)
    return dict(value=ne / dt / 1e6, unit="Melem/s", cores=threads, kind="port",
                sample=f"one {cells}^3-cell Kuhn block ({len(cn)} tets) per host thread x {threads} threads, "
                       f"full FAD-16 residual+Jacobian with sorted-row CRS scatter and state save, {dt:.2f} s"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_reference_rate(args.model, args.cpu_cells, repeats=max(1, min(args.steps, 3)))
    grid = grid_of(args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Melem/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, grid),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference algorithm timed as the CPU oracle port (the reference binary needs Trilinos+SCOREC+MPI, "
                "absent here); bounded sample per step, rate is per whole host",
    }
    print(json.dumps(line))


def workload_config(args, grid):
    c = args.cells
    ne = 6 * c ** 3 * args.gpus
    return {"workload": f"synthetic structured Kuhn tet cube, {args.model} mixed u/p Jacobian pass "
                        f"(zero + residual + Jacobian + state save{' + interface reduction' if args.gpus > 1 else ''})",
            "elements": ne, "cells_per_gpu": f"{c}^3", "global_cells": [grid[0] * c, grid[1] * c, grid[2] * c],
            "material": "E=1000 nu=0.25 K=100 Y=10 c0=1", "parallelism": f"element partition {grid[0]}x{grid[1]}x{grid[2]}",
            "l2": "inputs and outputs larger than L2 (CRS values alone exceed 126 MB); no flush needed"}


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        for ln in out.splitlines():
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import goal_b200
    from goal_b200.synthetic import MATERIAL, fields, kuhn_block

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = grid_of(world)
    c = args.cells

    if world == 1:
        co, cn = kuhn_block(c, c, c, (0, 0, 0), c)
        f = fields(co, len(cn))
        a = goal_b200.Assembler(co, cn, args.model, [MATERIAL], device=local)
    else:
        from goal_b200.partition import block_part
        part = block_part(c, grid, rank)
        co, cn = part["coords"], part["conn"]
        f = fields(co, len(cn), node_gid=part["node_gid"], elem_gid=part["elem_gid"])
        a = goal_b200.Assembler(co, cn, args.model, [MATERIAL], device=local, partition=part)
        a.comm_init_torch(dist)
    for kv in args.opt:
        k, v = kv.split("=")
        a.set_option(k, int(v))
    ne_local = a.ne
    # pinned host buffers for the end-to-end arm
    u_h = torch.from_numpy(np.ascontiguousarray(f["u"])).pin_memory()
    p_h = torch.from_numpy(np.ascontiguousarray(f["p"])).pin_memory()
    a.set_solution(u_h, p_h)
    if args.model == "J2":
        a.set_state("Fp_old", f["Fp_old"])
        a.set_state("eqps_old", f["eqps_old"])
    stream = torch.cuda.ExternalStream(a.stream(), device=torch.device("cuda", local))

    def step(R=None, V=None):
        a.jacobian(goal_b200.PRIMAL, save=True, out=False, R_out=R, values_out=V)
        if world > 1:
            a.reduce_interfaces(3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    kern_ms, zero_ms, exch_ms, launches = [], [], [], 0

    def step_dev():
        nonlocal launches
        step()
        t = a.last_timing()
        kern_ms.append(t["assemble_ms"]); zero_ms.append(t["zero_ms"]); exch_ms.append(t["exchange_ms"])
        launches += t["launches"] + (2 * 2 * a.num_peers if world > 1 else 0)  # + pack/unpack kernels (R and rows) per peer

    ms = timed(step_dev, args.steps)
    clk = clocks.stop() if rank == 0 else None
    plastic = a.plastic_count()
    ne_total = ne_local * world
    value = ne_total * args.steps / (ms * 1e-3) / 1e6

    e2e = None
    if not args.no_e2e:
        R_h = torch.empty(4 * a.nn, dtype=torch.float64).pin_memory()
        V_h = torch.empty(a.nnz, dtype=torch.float64).pin_memory()

        def step_e2e():
            a.set_solution(u_h, p_h)
            step(R_h, V_h)

        step_e2e()
        k2 = max(2, min(args.steps, 5))
        ms2 = timed(step_e2e, k2)
        e2e = {"value": ne_total * k2 / (ms2 * 1e-3) / 1e6, "unit": "Melem/s", "ms_per_step": ms2 / k2,
               "h2d_bytes_per_step": int(32 * a.nn), "d2h_bytes_per_step": int(8 * (4 * a.nn + a.nnz)),
               "host_buffers": "pinned"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    k_ms = statistics.mean(kern_ms)
    achieved = B_ALG[args.model] * ne_local / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.model}_bytes_per_element")
        traffic = None if traffic is None else traffic * ne_local
    line = {
        "metric": METRIC, "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, grid),
        "plastic_fraction": plastic / ne_local, "colours": a.num_colors,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": f"gx::elem_record_kernel<{args.model},save> + gx::patch_pair_kernel<primal> "
                               "(the two launches of one Jacobian pass; achieved = B_alg * elements / their summed device time)",
                     "kernel_ms_per_pass": k_ms, "zero_ms_per_pass": statistics.mean(zero_ms),
                     "exchange_ms_per_pass": statistics.mean(exch_ms),
                     "algorithmic_bytes_per_element": B_ALG[args.model]},
        "clocks": clk, "gpu_launches": launches, "e2e": e2e,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_rate(args.model, args.cpu_cells)[0]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
