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


def test_sharded_drivers_two_ranks_peer_memory_exchange():
    """Two processes, two GPUs: the sweep kernels store into the peer's arena (fdb_exchange_*), checked bit for bit
    against the single-rank results, and against the NCCL all-gather path.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = str(29700 + os.getpid() % 200)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", port, os.path.join(ROOT, "tools", "check_sharded.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("match the single-rank results") == 2


def test_c_abi_demo_runs_from_plain_c(tmp_path):
    """examples/c_abi_demo.c: the reference's golden vectors through the C ABI from a C program."""
    libdir = os.path.join(ROOT, "forwarddiff.jl_b200", "lib")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + libdir,
                    "-lfdb200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_abi_demo ok" in r.stdout, r.stdout + r.stderr


def test_julia_shim_call_sequences_run_from_plain_c(tmp_path):
    """examples/julia_call_sequences.c: every C-ABI call sequence of julia/ForwardDiffB200.jl (which cannot run here: no Julia),
    with the ccall argument order and types, checked against closed forms."""
    libdir = os.path.join(ROOT, "forwarddiff.jl_b200", "lib")
    exe = str(tmp_path / "julia_call_sequences")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "julia_call_sequences.c"), "-L" + libdir,
                    "-lfdb200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "julia_call_sequences ok" in r.stdout, r.stdout + r.stderr
