"""Sharded drivers (fdb200.dist) on the GPU.  With one GPU the process group has a single rank, which
still exercises the per-rank planning, the shifted column-block addressing and the gather assembly;
the multi-rank run is `torchrun ... tools/check_sharded.py` (2/4/8 GPUs, NCCL)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_sharded_drivers_single_rank():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_sharded.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "match the single-rank results" in out.stdout
