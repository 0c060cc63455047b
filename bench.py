#!/usr/bin/env python
"""bench.py -- ForwardDiff chunk-mode gradient throughput on B200 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU oracle port on the host cores)

A step = one full `ForwardDiff.gradient!(out, rosenbrock, x, GradientConfig(rosenbrock, x, Chunk{12}()))`
at n = 1 000 000 Float64: all 83 334 chunk sweeps (sharded over the ranks by contiguous chunk
ranges) plus, for N > 1, the all-gather of the gradient slices.  Total work is fixed as N grows
("strong" scaling).  `value` is timed with inputs resident in HBM; `e2e` goes through the host
buffer entry point (pinned host x -> H2D -> sweeps -> gather -> D2H of the gradient) every step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "forwarddiff.jl_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import ctypes
import numpy as np

C_DP = ctypes.POINTER(ctypes.c_double)
N_ELEMS = 1_000_000
CHUNK = 12
WORKLOAD = "C2: ForwardDiff.gradient of Rosenbrock, n=1_000_000 Float64, Chunk{12} (83334 chunk sweeps)"
METRIC = "rosenbrock_gradient_evals_per_sec"
# FP64 instructions the sweep kernel issues per term for terms outside the seeded window
# (SASS of gradient_sweep_kernel<Rosenbrock,12,false,256,true>: 9 DMUL + 11 DADD)
DP_OPS_PER_TERM_SHARED = 20
DP_OPS_PER_TERM_FULL = 133          # all 12 lanes: 120 in the term + 13 accumulate (lane sharing off)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------
# CPU baseline (oracle port) -- the only place bench.py touches oracle/
# ---------------------------------------------------------------------------
def cpu_oracle_rate(sample_chunks, threads):
    """Time `sample_chunks` consecutive chunk sweeps of C2 per thread on the host; returns (evals/s, seconds, nchunks)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    o = orc.Oracle()
    x = np.random.default_rng(1).random(N_ELEMS)
    nchunks = o.chunk_plan(N_ELEMS, CHUNK)["nchunks"]
    o.gradient("rosenbrock", x, CHUNK, chunks=(0, 2))            # warm the pages

    def work(t):
        lo = 1000 + t * sample_chunks
        o.gradient("rosenbrock", x, CHUNK, chunks=(lo, lo + sample_chunks))

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = time.perf_counter() - t0
    done = sample_chunks * threads
    return 1.0 / (dt * nchunks / done), dt, nchunks


def reference_cpp_rate(ctx):
    """The reference's OWN native benchmark (benchmarks/cpp: gradient<rosenbrock<Dual5>>, n = 12 000, the shape it ships with),
    built from /root/reference into oracle/_ref, against the same gradient through the library: kind "reference"."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    if not orc.RefBench.available():
        return None
    import fdb200
    rb = orc.RefBench()
    n, K = 12_000, 5
    x = np.random.default_rng(11).random(n)
    rb.gradient("rosenbrock", x, K)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        g_ref = rb.gradient("rosenbrock", x, K)
    cpu_ms = (time.perf_counter() - t0) / reps * 1e3
    xd, gd, vd = ctx.array(x), ctx.zeros(n), ctx.zeros(1)
    run = lambda: ctx.check(ctx.lib.fdb_gradient_chunked(ctx.h, fdb200._lib.FN["rosenbrock"], xd.h, K, 0, -1, gd.h, vd.h))
    run(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    ctx.sync()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    g = gd.numpy()
    return {"kind": "reference", "workload": "benchmarks/cpp gradient<rosenbrock<Dual5>>, n = 12000 (2400 chunk sweeps)", "cores": 1,
            "reference_cpp_ms": cpu_ms, "fdb200_ms": gpu_ms, "speedup": cpu_ms / gpu_ms,
            "max_rel_diff": float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-300))),
            "build": "g++ -std=c++11 -O2 on the reference's own sources (oracle/Makefile target ref)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_thread = 24
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, nchunks = cpu_oracle_rate(per_thread, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals])); dt = float(np.mean([b for _, b in vals]))
    sample = (f"{per_thread} consecutive chunk sweeps per thread x {threads} threads of {nchunks} (n={N_ELEMS}, Chunk{{12}}), "
              f"scaled by nchunks/sample; oracle.cpp (C++ restatement, AoS Dual, g++ -O3 -march=x86-64-v3, no FMA contraction); "
              "Julia is not installed in this image (JULIA_NUM_THREADS n/a); the reference itself is single-threaded by design")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "timed": "bounded sample, extrapolated"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample_seconds": dt}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fdb200", choices=["fdb200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the C3/C4/C5/Ackley and generic-kernel measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to stdout) are
    # pointed at stderr for the duration of the run; the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist
    import fdb200
    from fdb200 import functions as F

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3); K = args.steps
    # one explicit stream for everything: library kernels, copies, NCCL hand-off and the timing events
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    ctx = fdb200.Context(local)                                   # raises without the CUDA library / a GPU
    cfg = fdb200.GradientConfig(F.rosenbrock, np.empty(N_ELEMS), fdb200.Chunk(CHUNK))
    sg = fdb200.dist.ShardedGradient(ctx, F.rosenbrock, N_ELEMS, cfg)      # sets ctx stream = torch current stream
    ctx.set_async(True)
    x_host = torch.from_numpy(np.random.default_rng(1).random(N_ELEMS)).pin_memory()
    out_host = torch.empty(N_ELEMS, dtype=torch.float64).pin_memory()
    sg.x_t.copy_(x_host); torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, kernel_only=None):
        """Per-step CUDA-event timing on the launching stream, L2 flushed between steps; returns (ms list, kernel ms list)."""
        ms, kms = [], []
        for _ in range(steps):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            if kernel_only is not None:
                kernel_only(); e2.record(); fn(skip_sweep=True)
            else:
                fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            if kernel_only is not None:
                kms.append(e0.elapsed_time(e2))
        return ms, kms

    def step_dev(skip_sweep=False):
        if not skip_sweep:
            sg.sweep()
        sg.gather()

    x_np, out_np = x_host.numpy(), out_host.numpy()           # views of the pinned host buffers

    def step_e2e():
        if world == 1:
            # the public API call with host buffers: ForwardDiff.gradient!(out, f, x, cfg) -> fdb_gradient_host
            fdb200.gradient_b(out_np, F.rosenbrock, x_np, cfg, ctx=ctx)
        else:
            sg(x_host, out_host)

    # ---- device-resident steps -------------------------------------------------------------
    for _ in range(W):
        step_dev()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launches()
    ms, kms = timed(step_dev, K, kernel_only=sg.sweep)
    launches = ctx.launches() - l0
    barrier()
    # ---- the same steps with the result exchange done by an NCCL all-gather after the sweeps (comparison only) ----
    nccl_ms = None
    if world > 1 and sg.exchange == "p2p":
        sg_n = fdb200.dist.ShardedGradient(ctx, F.rosenbrock, N_ELEMS, cfg, exchange="nccl")
        sg_n.x_t.copy_(x_host)

        def step_nccl():
            sg_n.sweep(); sg_n.gather()
        for _ in range(W):
            step_nccl()
        barrier()
        nms, _ = timed(step_nccl, K)
        barrier()
        nccl_ms = sum(nms)
    # ---- end-to-end steps (host buffers, copies inside the timed region) --------------------
    for _ in range(2):
        step_e2e()
    barrier()
    ems, _ = timed(step_e2e, K)
    barrier()
    clocks = sampler.stop() if sampler else None

    tot = torch.tensor([sum(ms), sum(ems), sum(kms), nccl_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_ms, tot_ems, tot_kms, tot_nccl = (float(t) for t in tot.tolist())
    ms_per_step = tot_ms / K
    value = 1e3 / ms_per_step
    e2e_value = 1e3 / (tot_ems / K)

    # correctness of what was timed: closed-form Rosenbrock gradient (rank-local full copy after the gather)
    g = out_host.numpy(); xh = x_host.numpy()
    a = np.zeros(N_ELEMS)
    a[:-1] += -2.0 * (1.0 - xh[:-1]) - 400.0 * xh[:-1] * (xh[1:] - xh[:-1] ** 2)
    a[1:] += 200.0 * (xh[1:] - xh[:-1] ** 2)
    ok = bool(np.allclose(g, a, rtol=1e-12, atol=1e-11))

    exchange = sg.exchange
    xch_timeouts = sg.xch.timeouts() if sg.xch is not None else 0
    if rank != 0:
        sg.close()
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = load_peaks()
    dp_peak, dfma_peak = ctx.fp64_peak()                           # FP64 instruction issue rate (non-fused / fused), ops/s
    my_chunks = sg.plan["chunk_hi"] - sg.plan["chunk_lo"]
    kernel_s = (tot_kms / K) * 1e-3
    terms = (N_ELEMS - 1) * my_chunks
    flops = DP_OPS_PER_TERM_SHARED * terms
    achieved = flops / kernel_s / 1e12
    roofline = {"bound": "fp64", "kernel": "gradient_sweep_kernel<Rosenbrock,12,false,256,true>",
                "achieved": achieved, "peak": dp_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (dp_peak / 1e12),
                # dram__bytes_read.sum + dram__bytes_write.sum of one full C2 launch, ncu --set full
                # (profiles/r01_sweep_rosenbrock_pipelined_raw.csv): x read once, the 8 MB gradient still in L2
                "traffic": 8.165e6 if my_chunks == 83334 else None,
                "peak_source": "fdb_measure_fp64_peak: non-fused DMUL/DADD issue rate measured live on this GPU "
                               "(MEASURED_PEAKS.json has no FP64 figure); FMA peak %.1f TFLOP/s" % (2 * dfma_peak / 1e12),
                "algorithmic": f"{DP_OPS_PER_TERM_SHARED} FP64 ops per term per chunk sweep x {N_ELEMS - 1} terms x {my_chunks} chunks "
                               "(SURVEY.md 8d: implicit seeds => bound is FP64 issue, not HBM)",
                "hbm_equivalent": {"achieved": 8.0 * (CHUNK + 1) * N_ELEMS * my_chunks / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                   "note": "materialised-dual bytes 8(N+1)n per chunk sweep that the reference would stream; "
                                           "this kernel builds seeds in registers and reads x (8 MB) from L2, so it is NOT DRAM traffic",
                                   "peak_source": peak_src}}
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "fdb200",
            "config": {"workload": WORKLOAD, "n": N_ELEMS, "chunk": CHUNK, "sharding": f"{world} ranks x contiguous chunk ranges" + ("" if world == 1 else
                                    " + sweep kernel stores its gradient slice into every peer's arena over NVLink (fdb_exchange_*), device-side flag barrier"
                                    if exchange == "p2p" else " + NCCL all-gather"),
                       "exchange": None if world == 1 else exchange,
                       "l2": "256 MB buffer written between timed steps (x is 8 MB and meant to be L2-resident within a step)",
                       "lane_sharing": True, "verified_against_closed_form": ok},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": 8 * N_ELEMS, "d2h_bytes_per_step": 8 * N_ELEMS,
                    "ms_per_step": tot_ems / K,
                    "api": ("fdb200.gradient_b(out, rosenbrock, x, GradientConfig(rosenbrock, x, Chunk(12))) on pinned host arrays -> C ABI fdb_gradient_host "
                            "(H2D x, all chunk sweeps, D2H gradient + value)") if world == 1 else
                           f"fdb200.dist.ShardedGradient.__call__ (pinned host x -> H2D -> this rank's sweeps, result exchange: {exchange} -> D2H)"},
            "gpu_launches": int(launches), "clocks": clocks}

    if world == 1 and not args.no_cpu_baseline:
        v, dt, nchunks = cpu_oracle_rate(1200, 1)
        line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": 1, "kind": "port",
                                "sample": f"1200 consecutive chunk sweeps of {nchunks} (n={N_ELEMS}, Chunk{{12}}) in {dt:.1f} s on 1 core, "
                                          "scaled by nchunks/1200; oracle.cpp (C++ restatement of src/gradient.jl:114-159 on AoS Duals, "
                                          "g++ -O3 -march=x86-64-v3, no FMA contraction); Julia not installed (JULIA_NUM_THREADS n/a); "
                                          "the reference is single-threaded by design (docs/src/user/upgrade.md:39-50)"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_cpp_rate(ctx)
            if r is not None:
                line["cpu_baseline_reference_cpp"] = r
        except Exception as e:
            line["cpu_baseline_reference_cpp"] = {"error": repr(e)}
    if world == 1 and not args.no_extras:
        try:
            line["extra"] = extras(ctx, torch, dev, flush, hbm_peak, dp_peak)
        except Exception as e:                                    # extras never invalidate the headline line
            line["extra"] = {"error": repr(e)}
    if world > 1 and exchange == "p2p":
        line["exchange_compare"] = {"p2p_fused_ms_per_step": ms_per_step, "nccl_allgather_ms_per_step": tot_nccl / K,
                                    "barrier_timeouts": int(xch_timeouts)}
    emit(line)
    sg.close()
    if world > 1:
        dist.destroy_process_group()


def extras(ctx, torch, dev, flush, hbm_peak, dp_peak):
    """Other BASELINE.json configs and the generic HBM-bound kernels, each timed with CUDA events (3 reps, L2 flushed)."""
    import fdb200
    from fdb200 import functions as F
    lib = ctx.lib
    out = {}

    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            best.append(e0.elapsed_time(e1))
        return float(np.median(best))

    rng = np.random.default_rng(1)
    # lane sharing off: all 12 lanes evaluated for every term of every chunk
    x = ctx.array(rng.random(N_ELEMS)); g = ctx.zeros(N_ELEMS); v = ctx.zeros(1)
    fid = fdb200._lib.FN
    ctx.set_lane_sharing(False)
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["rosenbrock"], x.h, CHUNK, 0, 8192, g.h, v.h)), reps=2)
    ctx.set_lane_sharing(True)
    full_s = ms * 1e-3 * 83334 / 8192
    out["rosenbrock_gradient_all_lanes"] = {"evals_per_s": 1.0 / full_s, "timed": "8192 of 83334 chunks, scaled",
                                            "fp64_frac": DP_OPS_PER_TERM_FULL * (N_ELEMS - 1) * 8192 / (ms * 1e-3) / dp_peak}
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["ackley"], x.h, CHUNK, 0, -1, g.h, v.h)), reps=2)
    out["ackley_gradient_c2"] = {"evals_per_s": 1e3 / ms, "ms": ms}
    # the same Rosenbrock written as an ordinary callable on the device Dual array (broadcast + sum), flattened into
    # an op-program: what an arbitrary user f costs on the batched sweep -- through the kernel specialised for it at
    # run time (NVRTC instantiation of the sweep kernel template for the generated functor; the first call compiles,
    # the timed calls hit the cache) and through the per-thread interpreter
    from fdb200 import trace
    from fdb200.api import _i32p
    prog = trace.trace_scalar(F.rosenbrock_generic, N_ELEMS)
    cp = prog.consts.ctypes.data_as(C_DP)
    run_prog = lambda hi: ctx.check(lib.fdb_program_gradient_chunked(ctx.h, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi), cp,
                                                                     prog.consts.size, prog.n_acc, prog.halo, prog.result_reg, x.h, CHUNK, 0, hi,
                                                                     g.h, v.h))
    t0 = time.perf_counter(); run_prog(1); ctx.sync(); compile_s = time.perf_counter() - t0
    ms = t(lambda: run_prog(-1), reps=2)
    st = ctx.jit_stats()
    out["rosenbrock_gradient_generic_callable_specialised"] = {"evals_per_s": 1e3 / ms, "ms": ms, "first_call_s": compile_s,
                                                               "instructions": int(len(prog.term)), "jit": {k: st[k] for k in ("compiled", "cache_hits", "failures")}}
    ctx.set_jit(False)
    ms = t(lambda: run_prog(8192), reps=2)
    ctx.set_jit(True)
    out["rosenbrock_gradient_generic_callable_interpreted"] = {"evals_per_s": 1.0 / (ms * 1e-3 * 83334 / 8192), "timed": "8192 of 83334 chunks, scaled"}
    # C1: n = 10 000, Chunk{10} (the reference's published 0.2825 s case, docs/src/user/advanced.md:70-72)
    x1 = ctx.array(rng.random(10_000)); g1 = ctx.zeros(10_000)
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["rosenbrock"], x1.h, 10, 0, -1, g1.h, v.h)), reps=5)
    out["rosenbrock_gradient_c1"] = {"evals_per_s": 1e3 / ms, "ms": ms, "reference_published_s": 0.282529}
    # C4: Brusselator Jacobian n = 65 536, Chunk{8}: 34.4 GB dense column-major result
    M = 32768; n = 2 * M
    xb = ctx.array(np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)]))
    J = fdb200.B200Array.alloc(ctx, n * n); y = ctx.zeros(n)
    al = (M + 1) ** 2 / 50.0
    ms = t(lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["brusselator"], xb.h, n, 8, None, n, None, al, 0, -1, J.h, n, y.h)))
    out["brusselator_jacobian_c4"] = {"evals_per_s": 1e3 / ms, "ms": ms,
                                      "roofline": {"bound": "hbm", "achieved": 8.0 * n * n / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                   "frac": 8.0 * n * n / (ms * 1e-3) / 1e9 / hbm_peak,
                                                   "algorithmic": "8*m*n bytes of J written once (implicit seeds: no dual reads)"}}
    del J
    # C5: Rosenbrock Hessian n = 2048, Chunk{8}
    xh = ctx.array(rng.random(2048)); H = ctx.zeros(2048 * 2048); gh = ctx.zeros(2048)
    ms = t(lambda: ctx.check(lib.fdb_hessian_chunked(ctx.h, fid["rosenbrock"], xh.h, 8, 0, -1, H.h, 2048, gh.h, v.h)))
    out["rosenbrock_hessian_c5"] = {"evals_per_s": 1e3 / ms, "ms": ms}
    # C3: tanh.(A*x .+ b) Jacobian n = m = 8192, Chunk{12} (CUDA-core kernel)
    m3 = 8192
    A = ctx.array(rng.uniform(-1, 1, m3 * m3) / np.sqrt(m3)); b = ctx.array(rng.uniform(-1, 1, m3)); x3 = ctx.array(rng.uniform(-1, 1, m3))
    J3 = ctx.zeros(m3 * m3); y3 = ctx.zeros(m3)
    run_c3 = lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["tanh_affine"], x3.h, m3, 12, A.h, m3, b.h, 0.0, 0, -1, J3.h, m3, y3.h))
    # FP64 tensor-core denominator: cuBLAS DGEMM 8192^3 on this GPU (MEASURED_PEAKS.json has no FP64 figure)
    ta = torch.randn(m3, m3, dtype=torch.float64, device=dev); tb = torch.randn(m3, m3, dtype=torch.float64, device=dev)
    dgemm_ms = t(lambda: torch.matmul(ta, tb), reps=3)
    dgemm_tf = 2.0 * m3 ** 3 / (dgemm_ms * 1e-3) / 1e12
    del ta, tb
    ms = t(run_c3, reps=3)
    cols = 2 * 683                                   # lane sharing: value column + representative zero column per chunk
    out["tanh_affine_jacobian_c3"] = {"evals_per_s": 1e3 / ms, "ms": ms, "kernels": "pack_b + dmma_gemm_kernel + tanh_affine_epilogue",
                                      "roofline": {"bound": "tensor", "achieved": 2.0 * m3 * m3 * cols / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                                                   "unit": "TFLOP/s", "frac": 2.0 * m3 * m3 * cols / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                   "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run",
                                                   "algorithmic": "2*m*n flops per operand column, 2 columns per chunk x 683 chunks (whole call incl. pack and epilogue)"}}
    ctx.set_lane_sharing(False)
    ms = t(run_c3, reps=2)
    ctx.set_lane_sharing(True)
    cols = 13 * 682 + 9
    out["tanh_affine_jacobian_c3_all_lanes"] = {"evals_per_s": 1e3 / ms, "ms": ms,
                                                "roofline": {"bound": "tensor", "achieved": 2.0 * m3 * m3 * 13 * 683 / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                                                             "unit": "TFLOP/s", "frac": 2.0 * m3 * m3 * 13 * 683 / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                             "algorithmic": "2*m*n*13*683 = 1.19 TFLOP: A against [value | 12 one-hot partials] of every chunk"}}
    out["cublas_dgemm_8192_tflops"] = dgemm_tf
    # two layers tanh.(A2*tanh.(A1*x .+ b1) .+ b2), n = h = m = 8192: the second contraction has DENSE partials
    A2 = ctx.array(rng.uniform(-1, 1, m3 * m3) / np.sqrt(m3)); b2 = ctx.array(rng.uniform(-1, 1, m3))
    ms = t(lambda: ctx.check(lib.fdb_jacobian_mlp2_chunked(ctx.h, x3.h, m3, m3, 12, A.h, m3, b.h, A2.h, m3, b2.h, 0, -1, J3.h, m3, y3.h)), reps=2)
    fl = 2.0 * m3 * m3 * (2 * 683 + 13 * 683)
    out["mlp2_jacobian_8192_dense_partials"] = {"evals_per_s": 1e3 / ms, "ms": ms,
                                                "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": dgemm_tf, "unit": "TFLOP/s",
                                                             "frac": fl / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                             "algorithmic": "layer 1: 2 shared columns per chunk; layer 2: 13 dense columns per chunk; 683 chunks"}}
    del A, J3, A2
    # generic HBM-bound kernels on materialised Dual{12} arrays, n = 2^22 (436 MB per array > L2)
    nn = 1 << 22
    a = fdb200.B200DualArray.alloc(ctx, nn, 12); bb = fdb200.B200DualArray.alloc(ctx, nn, 12); o = fdb200.B200DualArray.alloc(ctx, nn, 12)
    o1 = fdb200.B200DualArray.alloc(ctx, 1, 12)
    xs = ctx.array(rng.uniform(0.5, 1.5, nn))
    ctx.check(lib.fdb_seed_zero(ctx.h, a.h, xs.h, 0, 0)); ctx.check(lib.fdb_seed_zero(ctx.h, bb.h, xs.h, 0, 0))
    dual_b = 8.0 * 13 * nn
    gk = {}
    for name, fn, nbytes in (
            ("map1_exp", lambda: ctx.check(lib.fdb_map1(ctx.h, 2, o.h, a.h)), 2 * dual_b),
            ("map1_pow2", lambda: ctx.check(lib.fdb_map1(ctx.h, 13, o.h, a.h)), 2 * dual_b),
            ("map2_mul", lambda: ctx.check(lib.fdb_map2(ctx.h, 3, o.h, a.h, bb.h)), 3 * dual_b),
            ("map2_scalar_add", lambda: ctx.check(lib.fdb_map2_scalar(ctx.h, 1, o.h, a.h, 1.5, 0)), 2 * dual_b),
            ("reduce_sum", lambda: ctx.check(lib.fdb_reduce(ctx.h, 1, o1.h, a.h)), dual_b),
            ("eval_rosenbrock_materialised", lambda: ctx.check(lib.fdb_eval_scalar_fn(ctx.h, 1, o1.h, a.h)), dual_b),
            ("seed_zero_full", lambda: ctx.check(lib.fdb_seed_zero(ctx.h, a.h, xs.h, 0, 0)), dual_b + 8.0 * nn)):
        ms = t(fn, reps=5)
        gk[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak}
    out["generic_kernels_dual12_n4M"] = gk
    return out


if __name__ == "__main__":
    main()
