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
    blk_of = {}  # (first block of the row, position, blocks in the row) -> (row node, column node)
    for r in range(nn):
        nb = int(nrow[r + 1] - nrow[r])
        for j in range(nb):
            blk_of[(int(nrow[r]), j, nb)] = (r, int(ncol[nrow[r] + j]))
    got = {}
    writers = {}
    n_paired = 0
    n_mixed = 0
    for pw in W:
        n_recs, n_items, n_runs = int(pw[0]), int(pw[1]), int(pw[2])
        assert n_recs <= RECS and n_items <= T
        it = pw[4:4 + 4 * T].reshape(T, 4)
        ot = pw[4 + 4 * T:4 + 8 * T].reshape(T, 4)
        # the bulk copies: runs of consecutive elements into consecutive slots; they fill exactly the slots the items read
        runs = pw[4 + 8 * T:4 + 8 * T + 2 * n_runs].reshape(n_runs, 2)
        staged = {}
        for e0, sl in runs:
            s0, ln = int(sl) & 0xff, int(sl) >> 8
            assert 1 <= ln <= 32 and s0 + ln <= RECS
            for j in range(ln):
                assert s0 + j not in staged
                staged[s0 + j] = int(e0) + j
        assert len(staged) == n_recs and len(set(staged.values())) == n_recs  # a record is staged once per patch
        parts = {}
        types_seen = []
        part_warps = {}
        for t in range(T):
            w3 = int(ot[t, 3])
            kind, typ, part, nsec = w3 & 3, (w3 >> 2) & 3, (w3 >> 4) & 0xff, (w3 >> 12) & 0x3f
            if kind == 0:  # idle lane (alignment padding between the typed warps, or past the last item)
                assert not it[t].any()
                continue
            assert t < n_items
            types_seen.append(typ)
            ents = []
            for k in range(4):
                for h in (0, 16):
                    v = (int(it[t, k]) >> h) & 0xffff
                    if v & 0x8000:
                        assert (v & 0xff) in staged
                        ents.append((staged[v & 0xff], (v >> 10) & 3, (v >> 8) & 3))
            w2 = int(ot[t, 2])
            key = blk_of[(int(ot[t, 0]), w2 & 0xff, (w2 >> 8) & 0xff)]
            if typ == 1:  # DIAG: block (a,a) and R of node a
                assert key[0] == key[1] == int(ot[t, 1]) and all(n == m for _, n, m in ents)
            elif typ == 2:  # PAIR: the mirror block
                assert blk_of[(int(ot[t, 1]), (w2 >> 16) & 0xff, w2 >> 24)] == (key[1], key[0]) and key[0] != key[1]
                n_paired += kind == 1
            else:
                assert not ents
            got.setdefault(key, set())
            for en in ents:
                assert en not in got[key]
                got[key].add(en)
            if typ == 2:
                got.setdefault((key[1], key[0]), set()).update((e, m, n) for e, n, m in ents)
            if kind == 1:
                for k2 in ([key, (key[1], key[0])] if typ == 2 else [key]):
                    writers[k2] = writers.get(k2, 0) + 1
                for s in range(nsec):
                    parts[part + s] = ("want", key)
                    part_warps.setdefault(part + s, set()).add(t // 32)
            else:
                part_warps.setdefault(part, set()).add(t // 32)
                assert parts.get(part, ("want", key)) == ("want", key)  # secondaries belong to one primary
                parts[part] = ("have", key)
        assert all(v[0] == "have" for v in parts.values())
        # a primary and its secondaries share a warp unless the header says the block barrier is needed
        assert int(pw[3]) in (0, 1) and (int(pw[3]) == 1 or all(len(v) == 1 for v in part_warps.values()))
        # typed warps: lanes [0, 32) hold no PAIR item, the other warps no DIAG item (when the schedule could align them)
        lane_types = [(t, (int(ot[t, 3]) >> 2) & 3) for t in range(T) if int(ot[t, 3]) & 3]
        if int(pw[3]) == 0:
            mixed = {t // 32 for t, ty in lane_types if ty == 1} & {t // 32 for t, ty in lane_types if ty == 2}
            n_mixed += len(mixed)
    assert set(writers) == set(want) and all(v == 1 for v in writers.values())  # write-once
    assert got == want
    assert n_mixed <= max(1, len(W) // 10)  # warps that run both code paths are the exception
    return n_paired, len(W)


@pytest.mark.parametrize("n", [2, 5])
def test_patch_schedule_invariants_kuhn(n):
    co, cn = kuhn_cube(n)
    n_paired, n_patches = _check(co, cn)
    edges = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    assert n_paired == edges  # every mesh edge is one paired item (its primary)


def test_patch_schedule_invariants_reference_fixture(cube):
    # irregular mesh: nodes of valence up to 56, blocks with more than 8 contributions (split items)
    n_paired, _ = _check(cube["coords"], cube["tets"])
    assert n_paired == 230
