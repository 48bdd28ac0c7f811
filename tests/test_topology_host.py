"""Host-only property test of csrc/kai_topology.cuh: incremental per-domain state == state rebuilt from the node tables
(4000 random node deltas, subSetNodesFn compared every few steps).  Compiled with nvcc as host code; no GPU needed."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_incremental_topology_state_matches_rebuild():
    src = os.path.join(ROOT, "tests", "native", "topology_incremental_check.cu")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "check")
        subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-x", "cu", "-o", exe, src], stdout=subprocess.DEVNULL)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")
