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


def test_c_abi_demo_runs_from_plain_c(tmp_path):
    """examples/c_abi_demo.c: the reference's golden vectors through the C ABI from a C program."""
    libdir = os.path.join(ROOT, "forwarddiff.jl_b200", "lib")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + libdir,
                    "-lfdb200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_abi_demo ok" in r.stdout, r.stdout + r.stderr
