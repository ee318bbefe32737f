"""What the compiled library (sm_100a) actually contains, read from its SASS without a GPU: the Blackwell features
DESIGN.md 5 claims for the hot kernels are in the binary -- bulk asynchronous copies in both directions (UBLKCP.S.G:
global -> shared staging of records / schedule words; UBLKCP.G.S: the element records' shared -> global store), their
mbarrier waits (SYNCS), L2 bulk prefetches (UBLKPF.L2), 256-bit global accesses -- and no local-memory spills in the
two kernels of the Jacobian pass.  (The profiling recipe's check, /opt/skills/guides/B200_PROFILING.md: cuobjdump -sass.)"""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def sass(gxlib):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", os.path.join(ROOT, "goal_b200", "libgoal_b200.so")], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out or "EF_CUDA_SM100" in out or "arch = sm_100" in out
    funcs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def _of(sass, *parts):
    hit = [v for k, v in sass.items() if all(p in k for p in parts)]
    assert hit, f"no kernel matching {parts}"
    return hit


def test_stage_a_stores_records_with_a_bulk_copy_and_prefetches(sass):
    for body in _of(sass, "elem_record_kernel"):
        assert "UBLKCP.G.S" in body  # shared -> global bulk store of the warp's 32 records
        assert "UBLKPF.L2" in body   # connectivity / old state of the elements ahead
        assert "STL" not in body or body.count("STL") <= 4  # (almost) no spills at 128 registers


def test_stage_b_stages_records_with_bulk_copies_on_an_mbarrier(sass):
    for body in _of(sass, "patch_pair_kernel"):
        assert "UBLKCP.S.G" in body and "SYNCS.ARRIVE.TRANS64" in body and "TRYWAIT" in body
        assert "UBLKPF.L2" in body
        assert "STG.E.ENL2.256" in body  # 4x4 block rows leave as 256-bit stores
        assert "STL" not in body and "LDL" not in body  # no spills at 160 registers
        assert "DFMA" in body


def test_residual_block_kernel_reads_its_schedule_by_bulk_copy(sass):
    for body in _of(sass, "elem_residual_block_kernel"):
        assert "UBLKCP.S.G" in body and "TRYWAIT" in body and "BAR.SYNC" in body
    for body in _of(sass, "node_partial_sum_kernel"):
        assert "DADD" in body


def test_register_budgets_of_the_hot_kernels(gxlib):
    """The occupancies DESIGN.md 5 quotes rest on register counts: stage B at most 160 (4 blocks of 96 threads per SM),
    the element kernels at most 128 (8 x 64 or 4 x 128 threads per SM).  Read from the build's ptxas log."""
    log = os.path.join(ROOT, "goal_b200", "csrc", "ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas.log (library not built by the Makefile)")
    txt = open(log).read()
    regs = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?Used (\d+) registers", txt, re.S):
        regs[m.group(1)] = int(m.group(2))
    def worst(part):
        v = [r for k, r in regs.items() if part in k]
        assert v, part
        return max(v)
    assert worst("patch_pair_kernel") <= 160
    assert worst("elem_record_kernel") <= 128
    assert worst("elem_residual_block_kernel") <= 128
    assert worst("elem_residual_kernel") <= 128
