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
Also in the line: `passes` (residual+save, error localisation + element indicators, adjoint Jacobian -- each with its
own algorithmic bytes and roofline fraction), `sizes` (the Jacobian pass on the 1M and 10M meshes of configs[4]),
`checks` (result checks made after the timed loop, at every N), `setup_s`, the measured FP64 roof.
`--scaling strong --global-cells G` splits one fixed G^3-cell cube over the N ranks instead.
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
# the other passes (per element, each datum once):
#   residual + state save   conn + nodes + R + state in + state out                                  (compute_resid)
#   adjoint Jacobian        as the Jacobian pass without the state save                              (compute_adjoint)
#   error localisation      conn + nodes (x,u,p) + adjoint weights (5 doubles / node) + R + state in,
#                           then compute_error: conn + 4 doubles / node + 1 double / element          (localize + compute_error)
B_PASS = {
    "J2": {"residual_save": 16 + (24 + 32 + 32) / 6 + 80 + 152, "jacobian_adjoint": 16 + 16 + (24 + 32 + 32) / 6 + 80 + 320,
           "error_localisation": 16 + (24 + 32 + 40 + 32) / 6 + 80 + 16 + 32 / 6 + 8},
    "neohookean": {"residual_save": 16 + (24 + 32 + 32) / 6 + 72, "jacobian_adjoint": 16 + 16 + (24 + 32 + 32) / 6 + 320,
                   "error_localisation": 16 + (24 + 32 + 40 + 32) / 6 + 16 + 32 / 6 + 8},
}
# counted flop per element of the Jacobian pass (scripts/count_ops.py, profiles/op_counts.txt): decides which roof binds
F_ALG = {"J2": 2657.0, "neohookean": 2160.0}


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: a --cells^3 block per GPU (default); strong: one --global-cells^3 cube split over the GPUs")
    ap.add_argument("--global-cells", type=int, default=256, help="cells per side of the fixed cube of --scaling strong")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: reduce the interfaces after the pass instead of overlapped with it")
    ap.add_argument("--no-passes", action="store_true", help="skip the residual / localisation / adjoint pass timings")
    ap.add_argument("--no-sizes", action="store_true", help="skip the 1M / 10M mesh timings (N = 1 only)")
    ap.add_argument("--no-checks", action="store_true", help="skip the result checks after the timed loop")
    ap.add_argument("--clock-interval-ms", type=int, default=50, help="nvidia-smi sampling interval during the timed region (0: no sampling)")
    return ap.parse_args()


def grid_of(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


# ---------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's FAD-16 assembly; the reference itself
# needs Trilinos + SCOREC and cannot be built here).  One mesh part per host thread, like the
# reference's rank-per-core model; each thread assembles its own part into private arrays.
# This is the one place bench.py executes oracle/.
# ---------------------------------------------------------------------------
def cpu_reference_rate(model, cells, threads=None, repeats=1, warmup=0):
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
    for it in range(warmup + repeats):
        def work(t):
            o, R, V = parts[t]
            R[:] = 0.0
            V[:] = 0.0
            o.jacobian(PRIMAL, save=True, R=R, values=V)
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        if it >= warmup:
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
    # a step of this arm = one Jacobian pass of every host thread over its own sample block (bounded: a few
    # seconds); warm-up and step counts are what was asked for, capped so that the run stays within minutes
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    cb, dt = cpu_reference_rate(args.model, args.cpu_cells, repeats=steps, warmup=warm)
    ne_sample = 6 * args.cpu_cells ** 3 * cb["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Melem/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic structured Kuhn tet cube, {args.model} mixed u/p Jacobian pass (zero + residual + Jacobian "
                               f"+ state save), CPU oracle port of the reference's FAD-16 assembly",
                   "elements": ne_sample, "cells_per_thread": f"{args.cpu_cells}^3", "threads": cb["cores"],
                   "material": "E=1000 nu=0.25 K=100 Y=10 c0=1", "parallelism": "one replicated block per host thread",
                   "sample_of": f"the GPU arm's {6 * args.cells ** 3 * args.gpus}-tet workload (same mesh family, fields and material; "
                                "throughput per element does not depend on the block size)",
                   "requested": {"steps": args.steps, "warmup": args.warmup}},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference algorithm timed as the CPU oracle port (the reference binary needs Trilinos+SCOREC+MPI, "
                "absent here); steps/warmup/elements are the ones actually run, rate is per whole host",
    }
    print(json.dumps(line))


def block_cells(args, grid):
    """cells per axis of one rank's block"""
    if args.scaling == "strong":
        g = args.global_cells
        if any(g % p for p in grid):
            raise SystemExit(f"--global-cells {g} is not divisible by the rank grid {grid}")
        return (g // grid[0], g // grid[1], g // grid[2])
    return (args.cells, args.cells, args.cells)


def workload_config(args, grid):
    cx, cy, cz = block_cells(args, grid)
    ne = 6 * cx * cy * cz * args.gpus
    return {"workload": f"synthetic structured Kuhn tet cube, {args.model} mixed u/p Jacobian pass "
                        f"(zero + residual + Jacobian + state save{' + interface reduction' if args.gpus > 1 else ''})",
            "elements": ne, "cells_per_gpu": f"{cx}x{cy}x{cz}" if args.scaling == "strong" else f"{cx}^3",
            "global_cells": [grid[0] * cx, grid[1] * cy, grid[2] * cz],
            "material": "E=1000 nu=0.25 K=100 Y=10 c0=1", "parallelism": f"element partition {grid[0]}x{grid[1]}x{grid[2]}",
            "l2": "inputs and outputs larger than L2 (CRS values alone exceed 126 MB); no flush needed"}


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=50):
        self.index, self.proc, self.interval_ms = index, None, interval_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.interval_ms), "-i", str(self.index)], stdout=subprocess.PIPE,
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


def _checksum(x):
    """order-independent bit checksum of a float64 array (xor and wrapping sum of the raw words)"""
    w = np.ascontiguousarray(x).view(np.uint64)
    with np.errstate(over="ignore"):
        return [int(np.bitwise_xor.reduce(w)), int(w.sum(dtype=np.uint64))]


def _translation_defect(rowptr, values):
    """max over dof rows of |sum over displacement columns of K[row, (m,k)] t_k| for the rigid translation t = (1,1,1),
    relative to max|K|.  Columns come in node blocks of 4 (u0 u1 u2 p), so the sum runs over the first three entries
    of every block.  F, hence R, does not change under a translation: K t = 0 on every row that received all of its
    contributions -- on an owned interface row that is only true after every peer's share has arrived."""
    v = values.reshape(-1, 4)
    per_block = v[:, 0] + v[:, 1] + v[:, 2]
    starts = (rowptr[:-1] // 4).astype(np.int64)
    rows = np.add.reduceat(per_block, starts)
    scale = float(np.abs(values).max())
    return float(np.abs(rows).max() / scale) if scale > 0 else float("inf")


def _pass_times(a, fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        fn()
        t = a.last_timing()
        ts.append(t["zero_ms"] + t["assemble_ms"])
    return float(statistics.median(ts))


def run_b200(args):
    # torchrun pins OMP_NUM_THREADS=1 for multi-process launches; the once-per-mesh host setup (graph, patch schedule)
    # is OpenMP code, so give every rank its share of the host cores (must happen before libgomp is loaded)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world_env))
    import torch
    import torch.distributed as dist
    import goal_b200
    from goal_b200.synthetic import MATERIAL, fields, kuhn_block, kuhn_cube

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = grid_of(world)
    cx, cy, cz = block_cells(args, grid)

    t0 = time.perf_counter()
    if world == 1:
        co, cn = kuhn_block(cx, cy, cz, (0, 0, 0), cx)
        f = fields(co, len(cn))
        t1 = time.perf_counter()
        a = goal_b200.Assembler(co, cn, args.model, [MATERIAL], device=local)
    else:
        from goal_b200.partition import block_part
        part = block_part((cx, cy, cz) if args.scaling == "strong" else cx, grid, rank)
        co, cn = part["coords"], part["conn"]
        f = fields(co, len(cn), node_gid=part["node_gid"], elem_gid=part["elem_gid"])
        t1 = time.perf_counter()
        a = goal_b200.Assembler(co, cn, args.model, [MATERIAL], device=local, partition=part)
        a.comm_init_torch(dist)
    create_s = time.perf_counter() - t1
    overlap = world > 1 and not args.no_overlap
    if overlap:  # SolInfo::gather_R / gather_dRdu inside the pass, hidden behind the interior patches
        a.set_option("overlap", 3)
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
        if world > 1 and not overlap:
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

    # the first pass also builds and uploads the patch schedule (host, once per mesh): setup, not a step
    ts = time.perf_counter()
    step()
    first_pass_s = time.perf_counter() - ts
    for _ in range(max(0, args.warmup - 1)):
        step()
    clocks = ClockSampler(local, args.clock_interval_ms)
    if rank == 0 and args.clock_interval_ms > 0:
        clocks.start()
    t_clk0 = time.perf_counter()
    kern_ms, zero_ms, exch_ms, launches = [], [], [], 0

    def step_dev():
        nonlocal launches
        step()
        t = a.last_timing()
        kern_ms.append(t["assemble_ms"]); zero_ms.append(t["zero_ms"]); exch_ms.append(t["exchange_ms"])
        launches += t["launches"] + (2 * 2 * a.num_peers if world > 1 and not overlap else 0)  # + pack/unpack kernels (R and rows) per peer

    ms = timed(step_dev, args.steps)
    # nvidia-smi polls at tens of milliseconds and the timed region may be shorter than that: keep the same steps running
    # (untimed) until the sampler has seen about half a second of this load -- every rank, to keep the load the same
    if args.clock_interval_ms > 0:
        while time.perf_counter() - t_clk0 < 0.5:
            step()
    clk = clocks.stop() if rank == 0 and args.clock_interval_ms > 0 else None
    if clk is not None:
        clk["window"] = "the timed region and the same steps continued, untimed, to 0.5 s of sampling"
    plastic = a.plastic_count()
    ne_total = ne_local * world
    value = ne_total * args.steps / (ms * 1e-3) / 1e6
    setup_s = create_s + max(0.0, first_pass_s - ms * 1e-3 / args.steps)

    # ---- result checks (every rank; rank 0 reports the worst case): the operator the timed loop left behind
    checks = None
    if not args.no_checks:
        if world == 1:
            R1, V1 = a.fetch()
            rowptr = a.rowptr
        else:
            R1, V1, g = a.fetch_owned()
            rowptr = g["rowptr"]
        c1 = (_checksum(R1), _checksum(V1))
        defect = _translation_defect(rowptr, V1)
        finite = bool(np.isfinite(V1).all() and np.isfinite(R1).all())
        del V1
        step()
        R2, V2 = a.fetch() if world == 1 else a.fetch_owned()[:2]
        same = c1 == (_checksum(R2), _checksum(V2))
        del V2
        flags = torch.tensor([1.0 if same else 0.0, 1.0 if finite else 0.0, -defect], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        checks = {"bit_identical_across_passes": bool(flags[0].item() == 1.0), "finite": bool(flags[1].item() == 1.0),
                  "translation_null_vector_defect": float(-flags[2].item()), "translation_tolerance": 1e-9,
                  "rows": "owned rows after the interface reduction" if world > 1 else "all rows",
                  "what": "max_row |sum_m sum_k<3 K[row,(m,k)]| / max|K| (rigid translation is a null vector of the tangent "
                          "only when every interface contribution has arrived); R and CRS values bit-identical over two passes"}
        checks["ok"] = bool(checks["bit_identical_across_passes"] and checks["finite"]
                            and checks["translation_null_vector_defect"] < checks["translation_tolerance"])

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
        del R_h, V_h

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    frac_of = lambda b_alg, ne, ms_: b_alg * ne / (ms_ * 1e-3) / 1e9 / hbm_peak

    # ---- the other passes of the path, device-resident, on this rank's block (rank 0 reports)
    passes = None
    if not args.no_passes:
        zu, zp, zc = [np.ascontiguousarray(f[k]) for k in ("zu_diff", "zp_diff", "zp_coarse")]
        bp = B_PASS[args.model]
        t_res = _pass_times(a, lambda: a.residual(save=True, out=False))
        t_adj = _pass_times(a, lambda: a.jacobian(goal_b200.ADJOINT, save=False, out=False))
        t_loc = _pass_times(a, lambda: a.localize(zu, zp, zc))
        tw = time.perf_counter()
        Re = a.localize(zu, zp, zc).reshape(-1, 4).copy()
        eta, _, bound = a.element_error(Re[:, :3].copy(), Re[:, 3].copy())
        loc_wall = time.perf_counter() - tw
        mk = lambda ms_, key: {"ms": ms_, "Melem_s": ne_local / ms_ / 1e3, "algorithmic_bytes_per_element": bp[key],
                               "frac": frac_of(bp[key], ne_local, ms_)}
        passes = {"residual_save": dict(mk(t_res, "residual_save"), ref="Primal::compute_resid src/goal_primal.cpp:75-90"),
                  "jacobian_adjoint": dict(mk(t_adj, "jacobian_adjoint"), ref="NestedAdjoint::compute_adjoint src/goal_nested_adjoint.cpp:163-187"),
                  "error_localisation": dict(mk(t_loc, "error_localisation"), ref="NestedAdjoint::localize + compute_error src/goal_nested_adjoint.cpp:217-234, goal_error.cpp:7-35",
                                             host_to_host_wall_s=loc_wall, bound=float(bound),
                                             note="ms = device time of the localisation kernels; the wall time adds the H2D of the adjoint "
                                                  "fields, gx_element_error and the D2H of R and the indicators"),
                  "unit": "device ms per pass on one rank's block (CUDA events on the library stream), frac = of the measured HBM roof"}

    fp64 = a.measure_fp64_peak() if rank == 0 else None
    a.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the named sizes of configs[4] that fit one GPU next to the main block: 1M (N=55) and 10M (N=119)
    sizes = None
    if world == 1 and not args.no_sizes:
        sizes = {}
        for n in (55, 119):
            tc = time.perf_counter()
            co2, cn2 = kuhn_cube(n)
            f2 = fields(co2, len(cn2))
            a2 = goal_b200.Assembler(co2, cn2, args.model, [MATERIAL], device=local)
            a2.set_solution(f2["u"], f2["p"])
            if args.model == "J2":
                a2.set_state("Fp_old", f2["Fp_old"]); a2.set_state("eqps_old", f2["eqps_old"])
            fn = lambda: a2.jacobian(goal_b200.PRIMAL, save=True, out=False)
            fn(); fn()
            t = _pass_times(a2, fn, reps=9)
            sizes[f"N={n}"] = {"elements": a2.ne, "ms": t, "Melem_s": a2.ne / t / 1e3, "frac": frac_of(B_ALG[args.model], a2.ne, t),
                               "wall_incl_setup_s": time.perf_counter() - tc}
            a2.close()
        sizes["note"] = ("Jacobian pass (zero + residual + Jacobian + state save), device ms, median of 9; the 1M mesh's working set "
                         "(CRS values 328 MB) is close to the 126 MB L2, the 100M mesh runs with --cells 256 (profiles/)")

    k_ms = statistics.mean(kern_ms)
    achieved = B_ALG[args.model] * ne_local / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.model}_bytes_per_element")
        traffic = None if traffic is None else traffic * ne_local
    fp64_tflops, fp64_mhz = fp64
    exec_flop = json.load(open(tp)).get(f"{args.model}_executed_fp64_flop_per_element") if os.path.exists(tp) else None
    hbm_roof = hbm_peak * 1e9 / B_ALG[args.model] / 1e6            # Melem/s
    fp64_roof = fp64_tflops * 1e12 / F_ALG[args.model] / 1e6       # Melem/s
    line = {
        "metric": METRIC, "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, grid),
        "plastic_fraction": plastic / ne_local,
        "roofline": {"bound": "hbm" if hbm_roof <= fp64_roof else "fp64", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": f"gx::elem_record_kernel<{args.model},save> + gx::patch_pair_kernel<primal> "
                               "(the two launches of one Jacobian pass; achieved = B_alg * elements / their summed device time)",
                     "kernel_ms_per_pass": k_ms, "zero_ms_per_pass": statistics.mean(zero_ms),
                     "exchange_ms_per_pass": statistics.mean(exch_ms),
                     "exchange": ("overlapped with the interior patches on a second stream; the figure is the exposed time after the "
                                  "last assembly kernel" if overlap else "after the pass, on the compute stream") if world > 1 else None,
                     "algorithmic_bytes_per_element": B_ALG[args.model],
                     "fp64_peak_tflops": fp64_tflops, "fp64_peak_source": "measured in this run (gx_measure_fp64_peak: register-only DFMA chains)",
                     "counted_flop_per_element": F_ALG[args.model],
                     "executed_fp64_flop_per_element": exec_flop,
                     "hbm_roof_Melem_s": hbm_roof, "fp64_roof_Melem_s": fp64_roof,
                     "fp64_roof_executed_Melem_s": (fp64_tflops * 1e12 / exec_flop / 1e6) if exec_flop else None,
                     "frac_of_binding_roof": value / world / min(hbm_roof, fp64_roof)},
        "clocks": clk, "gpu_launches": launches, "e2e": e2e, "checks": checks, "passes": passes, "sizes": sizes,
        "setup_s": {"total": setup_s, "gx_create": create_s, "patch_schedule_and_upload": max(0.0, first_pass_s - ms * 1e-3 / args.steps),
                    "mesh_and_fields_numpy": t1 - t0, "note": "host-side, once per mesh (rank 0's block)"},
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
