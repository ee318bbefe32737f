"""Host logic of the patch-gather Jacobian pass (goal_b200/csrc/gx_setup.cpp: build_patch_schedule), checked on the
CPU from the schedule words the device reads: every 4x4 block of the operator is written by exactly one work item
(a paired item, where the schedule has them, writes a block and its mirror), every block receives exactly the contributions of the elements that
contain both of its nodes, with the right local node roles, from record slots that hold those elements, and the
partial-sum bookkeeping of split items is consistent."""
import numpy as np
import pytest

from goal_b200.synthetic import MATERIAL, kuhn_cube


def _check(co, cn):
    import goal_b200
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL], device=-1)
    W, RECS, T = a.patch_schedule()
    nrow, ncol = a.node_graph()
    nn = len(co)
    # expected contributions of block (a,b): {(e, n, m)} with conn[e][n] = a, conn[e][m] = b
    want = {}
    for e, t in enumerate(cn):
        for n in range(4):
            for m in range(4):
                want.setdefault((int(t[n]), int(t[m])), set()).add((e, n, m))
    blk_of_off = {}
    for r in range(nn):
        nb = int(nrow[r + 1] - nrow[r])
        for j in range(nb):
            blk_of_off[(16 * int(nrow[r]) + 4 * j, 4 * nb)] = (r, int(ncol[nrow[r] + j]))
    got = {}
    writers = {}
    n_paired = 0
    for pw in W:
        n_recs, n_items = int(pw[0]), int(pw[1])
        recs = pw[4:4 + RECS]
        assert n_recs <= RECS and n_items <= T
        # the bulk copies: runs of consecutive elements into consecutive slots; they fill exactly the slots the items read
        n_runs = int(pw[2])
        runs = pw[4 + RECS + 8 * T:4 + RECS + 8 * T + 2 * n_runs].reshape(n_runs, 2)
        staged = {}
        for e0, sl in runs:
            s0, ln = int(sl) & 0xff, int(sl) >> 8
            assert 1 <= ln <= 8 and s0 + ln <= RECS
            for j in range(ln):
                assert s0 + j not in staged
                staged[s0 + j] = int(e0) + j
        assert len(staged) == n_recs and len(set(staged.values())) == n_recs  # a record is staged once per patch
        assert all(int(recs[k]) == v for k, v in staged.items())
        it = pw[4 + RECS:4 + RECS + 4 * T].reshape(T, 4)
        ot = pw[4 + RECS + 4 * T:4 + RECS + 8 * T].reshape(T, 4)
        o2 = np.zeros((T, 4), dtype=np.uint32)  # no paired items in this layout
        parts = {}
        for t in range(T):
            kind = int(ot[t, 2]) >> 30
            if kind == 0:
                assert t >= n_items or not it[t].any()
                continue
            ents = []
            for k in range(4):
                for h in (0, 16):
                    v = (int(it[t, k]) >> h) & 0xffff
                    if v & 0x8000:
                        ents.append((int(recs[v & 0xff]), (v >> 10) & 3, (v >> 8) & 3))
                        assert (v & 0xff) in staged
            voff = int(ot[t, 0]) | (int(ot[t, 1]) << 32)
            rl = int(ot[t, 2]) & 0xffff
            key = blk_of_off[(voff, rl)]
            diag = bool(int(ot[t, 3]) >> 31)
            assert (int(ot[t, 3]) & 0x7fffffff) == key[0] and diag == (key[0] == key[1])
            paired = bool(int(o2[t, 2]) >> 31)
            if paired:
                voff2 = int(o2[t, 0]) | (int(o2[t, 1]) << 32)
                assert blk_of_off[(voff2, int(o2[t, 2]) & 0xffff)] == (key[1], key[0])
                n_paired += kind == 1
            got.setdefault(key, set())
            for en in ents:
                assert en not in got[key]
                got[key].add(en)
            if paired:
                got.setdefault((key[1], key[0]), set()).update((e, m, n) for e, n, m in ents)
            part, nsec = (int(ot[t, 2]) >> 16) & 0xff, (int(ot[t, 2]) >> 24) & 0x3f
            if kind == 1:
                for k2 in ([key, (key[1], key[0])] if paired else [key]):
                    writers[k2] = writers.get(k2, 0) + 1
                for s in range(nsec):
                    parts[part + s] = ("want", key)
            else:
                assert parts.get(part, ("want", key)) == ("want", key)  # secondaries follow their primary
                parts[part] = ("have", key)
        assert all(v[0] == "have" for v in parts.values())
    assert set(writers) == set(want) and all(v == 1 for v in writers.values())  # write-once
    assert got == want
    return n_paired, len(W)


@pytest.mark.parametrize("n", [2, 5])
def test_patch_schedule_invariants_kuhn(n):
    co, cn = kuhn_cube(n)
    n_paired, n_patches = _check(co, cn)
    edges = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    assert n_paired in (0, edges)  # a pairing schedule makes every mesh edge one paired item


def test_patch_schedule_invariants_reference_fixture(cube):
    # irregular mesh: nodes of valence up to 56, blocks with more than 8 contributions (split items)
    n_paired, _ = _check(cube["coords"], cube["tets"])
    assert n_paired in (0, 230)
