"""Host logic of the block-reduced residual / error-localisation passes (goal_b200/csrc/gx_setup.cpp:
build_residual_schedule), replayed on the CPU from the words the device reads (tests/hostcheck: hc_residual_schedule
walks them exactly like elem_residual_block_kernel + node_partial_sum_kernel): every (element, local node) incidence
is used exactly once, every entry of R and of the partial-sum buffer is written exactly once, a node is written by a
block only when the block holds all its elements, and the result is the plain per-node sum of the element lines --
SolInfo's ghost R after scatter_primal (src/goal_displacement.cpp:163-175, goal_pressure.cpp:152-164)."""
import ctypes as C

import numpy as np
import pytest

from goal_b200.synthetic import kuhn_block, kuhn_cube


def _replay(hostcheck, co, cn, seed=1):
    co = np.ascontiguousarray(co, dtype=np.float64)
    cn = np.ascontiguousarray(cn, dtype=np.int32)
    nn, ne = len(co), len(cn)
    rng = np.random.default_rng(seed)
    rvec = rng.standard_normal((ne, 4, 4)) * 10.0 ** rng.integers(-6, 3, (ne, 1, 1))  # mixed magnitudes
    R = np.empty(4 * nn)
    stats = np.zeros(4, dtype=np.int64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    hostcheck.hc_residual_schedule.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    rc = hostcheck.hc_residual_schedule(nn, ne, cn.ctypes.data_as(C.POINTER(C.c_int32)), dp(co), dp(rvec), dp(R), stats.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0, f"replay failed with code {rc}"
    want = np.zeros((nn, 4))
    np.add.at(want, cn.reshape(-1), rvec.reshape(-1, 4))
    mag = np.zeros((nn, 4))
    np.add.at(mag, cn.reshape(-1), np.abs(rvec.reshape(-1, 4)))
    assert np.all(np.abs(R.reshape(nn, 4) - want) <= 1e-14 * mag + 1e-300)  # only the order of the additions differs
    return dict(blocks=int(stats[0]), words=int(stats[1]), partials=int(stats[2]), shared_nodes=int(stats[3]))


def test_one_block_writes_every_node_itself(hostcheck):
    co, cn = kuhn_cube(2)  # 48 tets: a single block
    s = _replay(hostcheck, co, cn)
    assert s == dict(blocks=1, words=s["words"], partials=0, shared_nodes=0)


def test_kuhn_cube_in_generator_order(hostcheck):
    co, cn = kuhn_cube(8)  # 3072 tets = 24 blocks; x-fastest numbering: nearly every node is shared between blocks
    s = _replay(hostcheck, co, cn)
    assert s["blocks"] == 24 and 0 < s["shared_nodes"] <= len(co)
    # what the block reduction buys: the partial sums are far fewer than the 4 incidences per element
    assert s["partials"] < 0.45 * 4 * len(cn)
    assert s["words"] * 4 < 24 * len(cn)  # schedule stream: under 24 B per element here (14 B on the 128^3 bench mesh)


def test_ragged_last_block_and_isolated_nodes(hostcheck):
    co, cn = kuhn_block(5, 3, 2)  # 180 tets: one full block and a ragged one
    extra = np.array([[9.0, 9.0, 9.0], [8.0, 9.0, 9.0]])  # two nodes that belong to no element
    co2 = np.concatenate([co[:7], extra[:1], co[7:], extra[1:]])
    remap = np.concatenate([np.arange(7), np.arange(8, len(co) + 1)])
    s = _replay(hostcheck, co2, remap[cn].astype(np.int32))
    assert s["blocks"] == 2 and s["shared_nodes"] >= 2  # the isolated nodes are zeroed by the second kernel


@pytest.mark.parametrize("seed", [3, 4])
def test_shuffled_element_order(hostcheck, seed):
    """No locality at all: the elements of a node are spread over nearly all blocks, so the partial sums multiply --
    the schedule degrades towards the element-line form's traffic, never to a wrong answer."""
    co, cn = kuhn_cube(6)
    ordered = _replay(hostcheck, co, cn, seed)
    rng = np.random.default_rng(seed)
    s = _replay(hostcheck, co, cn[rng.permutation(len(cn))], seed)
    assert s["partials"] > 2 * ordered["partials"] and s["shared_nodes"] >= len(co) - 8  # a corner node has one element


def test_many_elements_on_one_node(hostcheck):
    """a fan of 200 tets around one edge: nodes 0 and 1 have more incidences than a block has elements
    (block rows are limited to 255 blocks, DESIGN.md: 200 is near the largest fan the library accepts)"""
    n = 200
    ang = 2 * np.pi * np.arange(n) / n
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], -1)
    co = np.concatenate([[[0, 0, -1.0], [0, 0, 1.0]], ring])
    cn = np.stack([np.zeros(n, int), np.ones(n, int), 2 + np.arange(n), 2 + (np.arange(n) + 1) % n], -1).astype(np.int32)
    s = _replay(hostcheck, co, cn)
    assert s["blocks"] == 2 and s["partials"] >= 4  # the two hub nodes: one partial sum per block each
