"""N > 1 path on CPU: one process per mesh part over the gloo backend (the reference tests its
distributed path the same way, `mpirun -np 4` on the pre-split cube, test/CMakeLists.txt:10-14)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case,world,model", [("fixture", 4, "J2"), ("blocks", 2, "J2"), ("blocks", 4, "neohookean")])
def test_owned_rows_after_exchange_match_serial(gxlib, case, world, model):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_parts_worker.py"), case, model]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"OK {case} {model} world={world}" in r.stdout
