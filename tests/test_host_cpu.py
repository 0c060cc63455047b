"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/fdb200.h
declares, the host-side planning code reproduces the reference's bookkeeping, the API mirror
raises the reference's errors before touching the device, the product never imports the oracle,
and the multi-rank sharding + all-gather logic works under gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import fdb200
from fdb200 import functions as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = C.CDLL(fdb200.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table binds exactly the declared set
    assert set(fdb200._lib.SIGNATURES) == declared
    assert fdb200._lib.load().fdb_version().decode().startswith("fdb200")


def test_sweep_kernel_has_no_contracted_multiply_adds():
    """The reference never contracts a*b + c (src/partials.jl:222-224 are separate operations), so the library is built
    with -fmad=false: the SASS of the C2 gradient sweep kernel must hold DMUL and DADD only, no DFMA."""
    import shutil
    if not shutil.which("cuobjdump") or not shutil.which("nm"):
        pytest.skip("cuobjdump / nm not available")
    syms = subprocess.run(["nm", fdb200.LIB_PATH], capture_output=True, text=True).stdout.split()
    names = [t for t in syms if "gradient_sweep_kernelINS_10RosenbrockELi12ELb0ELi256ELb1" in t and t.startswith("_ZN3fdb")]
    assert names, "kernel symbol not found"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], fdb200.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert sass.count("DMUL") > 100 and sass.count("DADD") > 100
    assert "DFMA" not in sass


def test_tensor_core_and_tma_instructions_are_in_the_library():
    """SASS evidence that needs no GPU: the FP64 contraction kernel issues DMMA and is fed by TMA (UTMALDG tensor loads,
    UBLKCP bulk copies, mbarrier SYNCS); the pipelined reduction streams with UBLKCP as well."""
    import shutil
    if not shutil.which("cuobjdump") or not shutil.which("nm"):
        pytest.skip("cuobjdump / nm not available")
    syms = subprocess.run(["nm", fdb200.LIB_PATH], capture_output=True, text=True).stdout.split()

    def sass(fragment):
        name = next(t for t in syms if t.startswith("_ZN3fdb") and fragment in t)
        return subprocess.run(["cuobjdump", "-sass", "-fun", name, fdb200.LIB_PATH], capture_output=True, text=True, timeout=300).stdout

    gemm = sass("dmma_gemm_kernel")
    assert "DMMA" in gemm and "UTMALDG" in gemm and "UBLKCP" in gemm and "SYNCS" in gemm
    red = sass("mapreduce_tma_kernel")
    assert "UBLKCP" in red and "SYNCS" in red


def test_no_cpu_fallback_without_gpu():
    n = C.c_int(-1)
    rc = fdb200._lib.load().fdb_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fdb200.CudaError):
        fdb200.Context(0)
    with pytest.raises(fdb200.CudaError):      # the public API fails loudly, it does not compute on the host
        fdb200.gradient(F.rosenbrock, np.ones(5), fdb200.GradientConfig(F.rosenbrock, np.ones(5), fdb200.Chunk(2)))


def test_product_never_references_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "forwarddiff.jl_b200")):
        if os.sep + "build" in base or os.sep + "lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\b(import orc|from orc|liboracle|oracle/|orc_[a-z])", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_pickchunksize_and_plan_match_reference(oracle):
    for n in list(range(0, 60)) + [100, 1000, 8192, 10_000, 65_536, 1_000_000]:
        assert fdb200.pickchunksize(n) == oracle.pickchunksize(n)
        for thr in (1, 5, 8):
            assert fdb200.pickchunksize(n, thr) == oracle.pickchunksize(n, thr)
    for n, N in ((13, 4), (13, 6), (13, 12), (10_000, 10), (1_000_000, 12), (8192, 12), (65_536, 8), (2048, 8), (25, 5)):
        p = fdb200.chunk_plan(n, N)
        q = oracle.chunk_plan(n, N)
        for k in ("remainder", "lastchunksize", "lastchunkindex", "middle_hi", "nchunks"):
            assert p[k] == q[k], (n, N, k)
        assert p["middle_lo"] == 2
    assert fdb200.chunk_plan(12, 12)["nchunks"] == 1           # N == n: vector mode, one sweep
    with pytest.raises(fdb200.ChunkSizeError):
        fdb200.chunk_plan(3, 4)


@pytest.mark.parametrize("n,N", [(1_000_000, 12), (65_536, 8), (13, 4), (100, 7), (24, 12)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_rank_chunk_ranges_partition_the_chunks(n, N, world):
    plans = [fdb200.chunk_plan(n, N, r, world) for r in range(world)]
    nchunks = plans[0]["nchunks"]
    covered = []
    for p in plans:
        covered += list(range(p["chunk_lo"], p["chunk_hi"])) if nchunks < 10_000 else []
        assert 0 <= p["chunk_lo"] <= p["chunk_hi"] <= nchunks
        assert p["elem_lo"] == min(p["chunk_lo"] * N, n) and p["elem_hi"] == min(p["chunk_hi"] * N, n)
    assert plans[0]["chunk_lo"] == 0 and plans[-1]["chunk_hi"] == nchunks or any(p["chunk_hi"] == nchunks for p in plans)
    for a, b in zip(plans, plans[1:]):
        assert a["chunk_hi"] == b["chunk_lo"] or b["chunk_lo"] == b["chunk_hi"] == nchunks
    if nchunks < 10_000:
        assert covered == list(range(nchunks))
    lay = fdb200.dist.slice_layout(n, N, world)
    assert lay["per"] * world >= n and lay["per"] % N == 0
    assert 0 <= fdb200.dist.owner_of_last_chunk(n, N, world) < world


def test_chunk_plan_properties_random():
    """Property test of fdb_plan_chunks (hypothesis): for any n >= N the ranks' chunk ranges are contiguous, disjoint and
    cover every chunk; every structural position is seeded by exactly one chunk (the last chunk starts at n - lastchunksize,
    src/gradient.jl:121-125); the bookkeeping identities of the reference hold."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 12), st.integers(0, 5_000_000), st.integers(1, 8))
    def check(N, extra, world):
        n = N + extra
        plans = [fdb200.chunk_plan(n, N, r, world) for r in range(world)]
        p = plans[0]
        assert p["remainder"] == n % N and p["lastchunksize"] == (N if n % N == 0 else n % N)
        assert p["lastchunkindex"] == n - p["lastchunksize"] + 1
        assert p["nchunks"] == p["middle_hi"] + 1 == -(-n // N)
        pos = 0
        for q in plans:
            assert q["chunk_lo"] == min(pos, p["nchunks"]) or q["chunk_lo"] == q["chunk_hi"] == p["nchunks"]
            assert q["elem_lo"] == min(q["chunk_lo"] * N, n) and q["elem_hi"] == min(q["chunk_hi"] * N, n)
            pos = q["chunk_hi"]
        assert pos == p["nchunks"] and plans[-1]["elem_hi"] == n or any(q["chunk_hi"] == p["nchunks"] for q in plans)
        assert sum(q["elem_hi"] - q["elem_lo"] for q in plans) == n          # every position belongs to exactly one rank
        lay = fdb200.dist.slice_layout(n, N, world)
        assert lay["per"] % N == 0 and lay["per"] * world >= n
        assert all(q["elem_hi"] - q["elem_lo"] <= lay["per"] for q in plans)

    check()


def test_structural_index_sets(oracle):
    lib = fdb200._lib.load()
    for name, kind in fdb200._lib.STRUCTURE.items():
        for rows in (1, 2, 5, 6, 17):
            sidx = oracle.structural_indices(kind, rows)
            assert lib.fdb_structural_length(kind, rows) == len(sidx)
            assert [lib.fdb_structural_index(kind, rows, p) for p in range(len(sidx))] == sidx.tolist(), (name, rows)


def test_api_errors_raised_on_the_host_before_any_device_work():
    x = np.ones(5)
    with pytest.raises(fdb200.ChunkSizeError):
        fdb200.gradient(F.rosenbrock, x, fdb200.GradientConfig(F.rosenbrock, x, fdb200.Chunk(6)))
    with pytest.raises(fdb200.InvalidTagException):
        fdb200.gradient(F.rosenbrock, x, fdb200.GradientConfig(F.ackley, x, fdb200.Chunk(2)))
    with pytest.raises(fdb200.DimensionMismatch):
        fdb200.gradient_b(np.zeros(4), F.rosenbrock, x, fdb200.GradientConfig(F.rosenbrock, x, fdb200.Chunk(2)))
    with pytest.raises(fdb200.ChunkSizeError):
        fdb200.hessian(F.rosenbrock, x, fdb200.HessianConfig(F.rosenbrock, x, fdb200.Chunk(12)))
    assert fdb200.GradientConfig(F.rosenbrock, np.zeros(13)).N == 7       # Chunk(x) heuristic
    assert fdb200.gradient(F.rosenbrock, np.zeros(0), fdb200.GradientConfig(F.rosenbrock, np.zeros(0), fdb200.Chunk(0))).size == 0
    t1, t2 = fdb200.Tag(F.rosenbrock), fdb200.Tag(F.ackley)
    assert t1.count < t2.count                                            # tag ordering counter, config.jl:8-19
    assert fdb200.checktag(None, F.rosenbrock, x) and fdb200.checktag(fdb200.Tag((F.rosenbrock, 1)), F.ackley, x)


_WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "forwarddiff.jl_b200")); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import numpy as np, torch, torch.distributed as dist
import fdb200, orc
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
o = orc.Oracle()                       # the oracle stands in for the device sweep: this test is about the plumbing
for n, N, fn in ((1003, 12, "rosenbrock"), (50, 4, "ackley"), (24, 12, "sumsq")):
    x = np.random.default_rng(7).uniform(0.5, 1.5, n)
    p = fdb200.chunk_plan(n, N, rank, world)
    lay = fdb200.dist.slice_layout(n, N, world)
    g_local, v_local = o.gradient(fn, x, N, chunks=(p["chunk_lo"], p["chunk_hi"])) if p["chunk_hi"] > p["chunk_lo"] else (np.zeros(n), 0.0)
    mine = torch.zeros(lay["per"], dtype=torch.float64)
    mine[: p["elem_hi"] - p["elem_lo"]] = torch.from_numpy(g_local[p["elem_lo"]:p["elem_hi"]])
    full = fdb200.dist.allgather_slices(mine, world)[:n].numpy()
    v = torch.tensor([v_local], dtype=torch.float64)
    dist.broadcast(v, src=fdb200.dist.owner_of_last_chunk(n, N, world))
    g_ref, v_ref = o.gradient(fn, x, N)
    assert np.array_equal(full, g_ref), (rank, n, N)
    assert v.item() == v_ref
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_sharding_and_allgather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def _build_c_demo(tmp_path):
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "forwarddiff.jl_b200", "lib")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + libdir,
                    "-lfdb200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def test_c_abi_links_from_plain_c_and_fails_loudly_without_gpu(tmp_path):
    """The boundary is a self-contained C library: a C program links against it with gcc alone (no torch, no Python)."""
    exe = _build_c_demo(tmp_path)
    n = C.c_int(-1)
    has_gpu = fdb200._lib.load().fdb_device_count(C.byref(n)) == 0 and n.value > 0
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if has_gpu:
        assert r.returncode == 0 and "c_abi_demo ok" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_julia_shim_is_consistent_with_the_c_abi():
    """Julia is not installed, so ForwardDiffB200.jl cannot be loaded; what CAN be checked statically is: no constant defined
    twice, every function imported from ForwardDiff for extension gets at least one method, and every `ccall` names a symbol of
    include/fdb200.h with the right number of argument types (C types mapped onto the ctypes signatures)."""
    import re
    import fdb200
    src = open(os.path.join(ROOT, "forwarddiff.jl_b200", "julia", "ForwardDiffB200.jl")).read()
    consts = re.findall(r"^const\s+([A-Za-z_0-9, ]+?)\s*=", src, flags=re.M)
    names = [n.strip() for c in consts for n in c.split(",")]
    assert len(names) == len(set(names)), [n for n in names if names.count(n) > 1]
    imported = re.search(r"import ForwardDiff:(.*?)\n\n", src, flags=re.S).group(1).replace("\n", " ")
    for fn in [t.strip() for t in imported.split(",") if t.strip()]:
        pat = r"^(function\s+)?" + re.escape(fn) + r"\("
        assert re.search(pat, src, flags=re.M), f"{fn} is imported for extension but has no method"
    ctype_ok = {"Ptr{Cvoid}": (fdb200._lib.C.c_void_p,), "Ref{Ptr{Cvoid}}": (fdb200._lib.C.POINTER(fdb200._lib.C.c_void_p),),
                "Cint": (fdb200._lib.C.c_int,), "Int64": (fdb200._lib.C.c_int64,), "Cdouble": (fdb200._lib.C.c_double,)}
    calls = re.findall(r"ccall\(\(:(fdb_\w+), libfdb200\),\s*(\w+),\s*\((.*?)\),", src, flags=re.S)
    assert len(calls) > 40
    for name, ret, argtypes in calls:
        assert name in fdb200._lib.SIGNATURES, f"{name} is not declared in include/fdb200.h"
        res, args = fdb200._lib.SIGNATURES[name]
        jl = [t.strip() for t in re.sub(r"\s+", " ", argtypes).split(",") if t.strip()]
        assert len(jl) == len(args), (name, jl, args)
        for t, a in zip(jl, args):
            if t == "Ptr{Cvoid}":           # an opaque handle or a pointer to a plain struct (fdb_chunk_plan)
                import ctypes
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, ctypes._Pointer), (name, t, a)
            elif t in ctype_ok:
                assert a in ctype_ok[t], (name, t, a)


def test_handle_sizes_agree_between_header_and_binding():
    """The byte sizes of the handles that travel through the host transport are part of the C ABI."""
    import re
    hdr = open(os.path.join(ROOT, "include", "fdb200.h")).read()
    sizes = dict(re.findall(r"#define\s+(FDB_\w+_HANDLE_BYTES)\s+(\d+)", hdr))
    assert int(sizes["FDB_IPC_HANDLE_BYTES"]) == fdb200._lib.IPC_HANDLE_BYTES == 64
    assert int(sizes["FDB_SHARE_HANDLE_BYTES"]) == fdb200._lib.SHARE_HANDLE_BYTES == 16
    # the multicast sequence is exported completely (a partial set would leave a caller stuck between two collective steps)
    for name in ("fdb_multicast_supported", "fdb_exchange_create_multicast", "fdb_exchange_share_handle", "fdb_exchange_multicast_create",
                 "fdb_exchange_connect_multicast", "fdb_exchange_multicast_bind", "fdb_exchange_set_multicast", "fdb_set_chunk_grouping"):
        assert name in fdb200._lib.SIGNATURES and re.search(r"\b" + name + r"\s*\(", hdr), name


def test_launch_list_tool_summarises_an_ncu_csv(tmp_path):
    """tools/launch_list.py: per-kernel launches, average and share from the CSV that `ncu --csv --log-file` writes."""
    import subprocess, sys
    rows = ['==PROF== Connected to process 1', '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size",'
            '"Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']
    def row(i, kernel, ns):
        return f'"{i}","1","python","box","{kernel}","1","7","(256, 1, 1)","(100, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","{ns}"'
    rows += [row(0, "void fdb::gradient_sweep_kernel<Rosenbrock, 12, 0, 256, 1>(const double *, long)", "94,000,000"),
             row(1, "void fdb::gradient_sweep_kernel<Rosenbrock, 12, 0, 256, 1>(const double *, long)", "96,000,000"),
             row(2, "fdb::seed_kernel(double *)", "10,000,000")]
    p = tmp_path / "launches.csv"
    p.write_text("\n".join(rows) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_list.py"), str(p), "--md"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("| `")]
    assert lines[0].startswith("| `gradient_sweep_kernel<Rosenbrock, 12, 0, 256, 1>` | 2 | 95.0000 | 190.000 | 95.00 |"), lines[0]
    assert lines[1].startswith("| `seed_kernel` | 1 | 10.0000 | 10.000 | 5.00 |"), lines[1]
    assert "total listed GPU time: 200.000 ms over 3 launches" in out.stdout
