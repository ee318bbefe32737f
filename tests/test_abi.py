"""The C-ABI library loads and exports every symbol include/goal_b200.h declares; the product has no CPU path."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_exports_match_header(gxlib):
    hdr = open(os.path.join(ROOT, "include", "goal_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(gx_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(gxlib, s)]
    assert not missing, missing
    from goal_b200.binding import SYMBOLS
    assert sorted(SYMBOLS) == declared


def test_header_is_plain_c():
    import subprocess
    src = '#include "goal_b200.h"\nint main(void){gx_desc d; (void)d; return GX_OK;}\n'
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src.encode(), check=True)


def test_no_cpu_fallback(gxlib):
    """Without a GPU gx_create must fail loudly; bad descriptions are rejected either way."""
    import torch
    import goal_b200
    from goal_b200.synthetic import MATERIAL, kuhn_cube
    co, cn = kuhn_cube(2)
    if not torch.cuda.is_available():
        with pytest.raises(goal_b200.GxError) as e:
            goal_b200.Assembler(co, cn, "J2", [MATERIAL])
        assert "no usable CUDA device" in str(e.value)
    with pytest.raises(goal_b200.GxError):
        goal_b200.Assembler(co, cn, "J2", np.zeros((0, 5)))


def test_product_does_not_touch_oracle():
    """Nothing under goal_b200/ may import, link or execute the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "goal_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "goal_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn


def test_single_part_tpetra_view_is_the_ghost_graph(gxlib):
    """gx_owned_tpetra_graph on a one-part context: every node is owned, the column map is the row map, and the local
    CRS arrays are exactly gx_graph's (the ghost layout the reference's get_lids index into, goal_disc.cpp:209-222)."""
    import numpy as np
    import goal_b200
    from goal_b200.synthetic import MATERIAL, kuhn_cube
    co, cn = kuhn_cube(3)
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL], device=-1)
    tp = a.owned_tpetra_graph()
    assert tp["n_owned"] == len(co) and np.array_equal(tp["colmap"], np.arange(len(co)))
    assert np.array_equal(tp["rowptr"], a.rowptr) and np.array_equal(tp["colind"], a.colind)


def test_option_keys_of_the_header_are_the_ones_the_library_takes(gxlib):
    """include/goal_b200.h lists the gx_set_option keys with their defaults: every listed key is accepted with its
    default, out-of-range values and unknown keys are GX_ERR_ARG with a message (no silent ignore)."""
    import re
    import pytest
    import goal_b200
    from goal_b200.synthetic import MATERIAL, kuhn_cube
    hdr = open(os.path.join(ROOT, "include", "goal_b200.h")).read()
    block = hdr[hdr.index("Tuning / cross-check switches"):hdr.index("int gx_set_option")]
    keys = dict(re.findall(r'"(\w+)"\s+\[(\d+)\]', block))
    assert set(keys) == {"kernel", "residual_kernel", "overlap", "prefetch", "prefetch_elems", "prefetch_elems_nosave", "block_size"}
    co, cn = kuhn_cube(2)
    a = goal_b200.Assembler(co, cn, "J2", [MATERIAL], device=-1)
    for k, v in keys.items():
        a.set_option(k, int(v))
    a.set_option("patch_schedule_dryrun", 1)
    for k, bad in [("kernel", 2), ("residual_kernel", 2), ("overlap", 4), ("prefetch", -1), ("prefetch_elems", -1), ("block_size", 100), ("no_such_key", 0)]:
        with pytest.raises(RuntimeError, match="GX_ERR_ARG"):
            a.set_option(k, bad)
    a.close()


def test_invalid_connectivity_is_rejected_with_the_first_offender(gxlib):
    """gx_create validates the connectivity before anything is built: an out-of-range node id, or an element that
    lists a node twice (the message names the first such element -- the check runs in parallel and must stay deterministic)."""
    import goal_b200
    from goal_b200.synthetic import MATERIAL, kuhn_cube
    co, cn = kuhn_cube(6)
    bad = cn.copy(); bad[700, 2] = len(co)
    with pytest.raises(RuntimeError, match="out of range"):
        goal_b200.Assembler(co, bad, "J2", [MATERIAL], device=-1)
    bad = cn.copy(); bad[5, 0] = -1
    with pytest.raises(RuntimeError, match="out of range"):
        goal_b200.Assembler(co, bad, "J2", [MATERIAL], device=-1)
    bad = cn.copy()
    for e in (1200, 333, 901):
        bad[e, 3] = bad[e, 1]
    with pytest.raises(RuntimeError, match="element 333 repeats a node"):
        goal_b200.Assembler(co, bad, "J2", [MATERIAL], device=-1)
