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
# FP64-pipe instructions (thread level) per element per chunk sweep / per term per chunk pair, measured with ncu on the C2 Ackley launch
# and the C5 Hessian launch (profiles/r02_fp64_counts.md): sum of smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on / units
ACKLEY_FP64_ISSUE_PER_TERM = 28.06   # (938.6 DADD + 1117.4 DMUL + 3161.3 DFMA) inst/cycle x 448.13e6 cycles / (1e6 x 83334)   [profiles/r02_ackley_sweep_raw.csv]
HESSIAN_FP64_ISSUE_PER_TERM = 7.19   # (2074.1 DADD + 988.2 DMUL) inst/cycle x 3.14919e5 cycles / (2047 x 65536)                  [profiles/r02_hessian_c5_grouped_raw.csv]
#   (the first round-2 kernel, r02_hessian_c5_raw.csv, executed 52.88: every pair re-evaluated all far terms, every near term on all lanes)


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
    ap.add_argument("--under-profiler", action="store_true",
                    help="skip the extras that replay thousands of tiny kernels (the graph-replayed chunk loop: ~11 000 launches per call, "
                         "each serialised and replayed under ncu) so that an ncu launch list of the command finishes in about a minute")
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
    sharded = None
    if world > 1 and not args.no_extras:
        try:
            sharded = extras_sharded(ctx, torch, dist, dev, rank, world, flush)       # collective: every rank takes part
        except Exception as e:
            sharded = {"error": repr(e)}
    if rank != 0:
        sg.close()
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = load_peaks()
    ctx_sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
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
                # independent of the probe: SMs x 64 FP64 lanes x the SM clock sampled by nvidia-smi during the timed region
                "peak_spec_crosscheck": (ctx_sm_count * 64 * (clocks or {}).get("sm_mhz", 0) * 1e6 / 1e12) if clocks and clocks.get("sm_mhz") else None,
                "algorithmic": f"{DP_OPS_PER_TERM_SHARED} FP64 ops per term per chunk sweep x {N_ELEMS - 1} terms x {my_chunks} chunks "
                               "(SURVEY.md 8d: implicit seeds => bound is FP64 issue, not HBM).  20 is the count of the CHOSEN algorithm: with lane "
                               "sharing a term far from the seeded window is evaluated on the value lane and ONE representative zero lane, a 6.6x "
                               "saving over the 133 ops of all 12 lanes (extra.rosenbrock_gradient_all_lanes times that form: 0.86 of the FP64 issue rate)",
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
            line["extra"] = extras(ctx, torch, dev, flush, hbm_peak, dp_peak, under_profiler=args.under_profiler)
        except Exception as e:                                    # extras never invalidate the headline line
            line["extra"] = {"error": repr(e)}
    if sharded is not None:
        line["extra_sharded"] = sharded
    if world > 1 and exchange == "p2p":
        line["exchange_compare"] = {"p2p_fused_ms_per_step": ms_per_step, "nccl_allgather_ms_per_step": tot_nccl / K,
                                    "barrier_timeouts": int(xch_timeouts)}
    emit(line)
    sg.close()
    if world > 1:
        dist.destroy_process_group()


def extras_sharded(ctx, torch, dist, dev, rank, world, flush):
    """C3, C4 (the full 34.4 GB Jacobian) and C5 with their chunk sweeps sharded over the ranks (collective: every rank calls it).
    Per config, ms per evaluation (median of the reps, max over ranks, CUDA events, L2 flushed between reps) for
      sharded_resident : every rank sweeps its chunk range into its own column block, no exchange (the result stays sharded);
      gathered_p2p     : the sweep kernels also store every entry into all peers' arenas over NVLink + device-side flag barrier
                         (the complete matrix is resident on every rank);
      gathered_multicast: the same with ONE store per entry through the NVSwitch multicast mapping of all arenas (when supported);
      gathered_nccl    : the sharded-resident sweep followed by one NCCL all-gather of the column blocks;
    and a bit-for-bit check of sampled chunk column blocks (first / middle / last chunk, wherever they were computed) of both
    gathered matrices against the same chunks recomputed on THIS rank alone."""
    import fdb200
    from fdb200 import functions as F
    lib = ctx.lib
    rng = np.random.default_rng(1)

    def sync_all():
        dist.barrier(); torch.cuda.synchronize()

    def timed(fn, reps):
        fn(); sync_all()
        ms = []
        for _ in range(reps):
            flush.zero_(); sync_all()
            torch.cuda._sleep(400_000)                      # the host enqueues the whole step while the GPU is parked (see extras.t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        tt = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def all_ok(flag):
        tt = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        return bool(tt.item() > 0.5)

    def check_columns(sh, recompute, nchunks, N, n, m):
        """sampled chunks of the gathered matrix vs a single-rank recomputation, bit for bit"""
        full = sh.gather()
        torch.cuda.synchronize()
        ok = True
        for c in sorted({0, nchunks // 2, nchunks - 1}):
            c0, c1 = c * N, min((c + 1) * N, n)
            blk = torch.zeros(m * (c1 - c0), dtype=torch.float64, device=dev)
            shifted = ctx.wrap(blk.data_ptr() - 8 * c0 * m, m * n, keepalive=blk)
            recompute(c, shifted)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(full[m * c0: m * c1], blk))
        return all_ok(ok)

    out = {}
    configs = []
    # C3: tanh.(A*x .+ b), n = m = 8192, Chunk{12}
    m3 = 8192
    A_h = (rng.uniform(-1, 1, m3 * m3) / np.sqrt(m3)).reshape((m3, m3), order="F"); b_h = rng.uniform(-1, 1, m3); x3 = rng.uniform(-1, 1, m3)
    f3 = F.tanh_affine.with_params(A=A_h, b=b_h)
    configs.append(("tanh_affine_jacobian_c3", "jac", f3, x3, m3, 12, 5,
                    "sharded-resident: DMMA contraction of 683/W chunk sweeps per rank (128 x 16 tiles, three blocks per SM: 704 tiles on 444 slots at W = 8); "
                    "gathered: every GPU must RECEIVE (W-1)/W of the 512 MB result over its NVLink ingress (W/W with multicast stores, which also "
                    "come back to the sender) -- 0.65-0.73 ms at ~700 GB/s against ~0.85 ms of contraction at W = 8, overlapped only in part because "
                    "all ranks produce at the same time; per-peer stores additionally send (W-1) x the slice through each GPU's egress"))
    # C4: Brusselator n = 65 536, Chunk{8}: 34.4 GB, every rank ends up holding all of it when gathered
    M = 32768
    xb = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    f4 = F.brusselator.with_params(alpha=(M + 1) ** 2 / 50.0)
    configs.append(("brusselator_jacobian_c4", "jac", f4, xb, 2 * M, 8, 3,
                    "gathered: each GPU's NVLink INGRESS of (W-1)/W of the 34.4 GB matrix; sharded-resident: HBM writes of 34.4/W GB"))
    # C5: Rosenbrock Hessian n = 2048, Chunk{8}
    xh = rng.random(2048)
    configs.append(("rosenbrock_hessian_c5", "hess", F.rosenbrock, xh, 2048, 8, 10,
                    "latency: 256/W outer chunks x 256 inner chunks of 2047 short terms per rank, then a 33.5 MB exchange"))

    modes = ["nccl", "p2p"]
    sup = [None] * world
    dist.all_gather_object(sup, bool(fdb200.dist.multicast_supported(ctx)))
    if all(sup):
        modes.append("multicast")      # NVSwitch multicast: each entry stored once, replicated by the switch (csrc/fdb_multicast.cu)
    for name, kind, f, xhost, m, N, reps, limiting in configs:
        n = xhost.size
        nchunks = fdb200.chunk_plan(n, N)["nchunks"]
        entry = {"ranks": world, "n": int(n), "chunk": N, "limiting_element": limiting}
        res = {}
        for mode in modes:
            try:
                if kind == "jac":
                    sh = fdb200.dist.ShardedJacobian(ctx, f, n, m, fdb200.JacobianConfig(f, xhost, fdb200.Chunk(N)), exchange=mode)
                else:
                    sh = fdb200.dist.ShardedHessian(ctx, f, n, fdb200.HessianConfig(f, xhost, fdb200.Chunk(N)), exchange=mode)
            except Exception as e:
                entry[f"gathered_{mode}_error"] = repr(e)
                continue
            sh.x_t.copy_(torch.from_numpy(xhost)); torch.cuda.synchronize()
            if mode == "nccl":
                res["sharded_resident_ms"] = timed(sh.sweep, reps)

            def step():
                sh.sweep(); sh.gather()
            res[f"gathered_{mode}_ms"] = timed(step, reps)
            if kind == "jac":
                fid = sh.fid
                rec = lambda c, arr: ctx.check(lib.fdb_jacobian_chunked(ctx.h, fid, sh.x_a.h, m, N, sh.A_a.h if sh.A_a else None, m,
                                                                         sh.b_a.h if sh.b_a else None, sh.alpha, c, c + 1, arr.h, m, None))
            else:
                rec = lambda c, arr: ctx.check(lib.fdb_hessian_chunked(ctx.h, sh.fid, sh.x_a.h, N, c, c + 1, arr.h, n, None, None))
            res[f"gathered_{mode}_bitwise_vs_single_rank"] = check_columns(sh, rec, nchunks, N, n, m)
            if name == "tanh_affine_jacobian_c3" and mode in ("p2p", "multicast"):
                # how the fused DMMA epilogue gets its result blocks to the peers (fdb_set_dmma_tma_store): every thread its own
                # accumulator entries / staged + TMA tensor stores / staged + 256-byte runs stored by the block's threads
                for label, sm in (("own_entries", 0), ("staged_tma", 2), ("staged_thread_runs", 3)):
                    ctx.check(lib.fdb_set_dmma_tma_store(ctx.h, sm))
                    res[f"gathered_{mode}_{label}_ms"] = timed(step, reps)
                    res[f"gathered_{mode}_{label}_bitwise_vs_single_rank"] = check_columns(sh, rec, nchunks, N, n, m)
                ctx.check(lib.fdb_set_dmma_tma_store(ctx.h, 1))
            if kind == "hess" and mode in ("p2p", "multicast"):
                # the two ways of getting the column block to the peers (fdb200.dist.ShardedHessian.peer_stores); the default picks by rank count
                for ps in ("mirror", "push"):
                    sh.peer_stores = ps
                    res[f"gathered_{mode}_{ps}_ms"] = timed(step, reps)
                    res[f"gathered_{mode}_{ps}_bitwise_vs_single_rank"] = check_columns(sh, rec, nchunks, N, n, m)
                sh.peer_stores = "auto"
            if sh.xch is not None:
                res["barrier_timeouts"] = res.get("barrier_timeouts", 0) + int(sh.xch.timeouts())
            sh.close()
            del sh
            torch.cuda.empty_cache()
        entry.update(res)
        entry["result_bytes"] = 8 * int(m) * int(n)
        if "gathered_p2p_ms" in res:
            entry["gathered_p2p_ingress_GBps_per_gpu"] = 8.0 * m * n * (world - 1) / world / (res["gathered_p2p_ms"] * 1e-3) / 1e9
        out[name] = entry

    # ONE host process driving all GPUs (fdb_multi_*: what a single Julia session binds to).  Rank 0 alone builds a DeviceGroup
    # over the `world` devices while the other ranks wait; C2 through the host-buffer call: x (pinned host) -> every device,
    # every device's chunk range, every slice -> the host gradient.  Host wall clock of the blocking call (it returns when the
    # last slice has landed), median of 5.
    sync_all()
    # the other ranks must wait on the HOST: an NCCL barrier is a kernel spinning on their GPUs, which would time-slice with the
    # group's contexts on the same devices
    store = dist.distributed_c10d._get_default_store()
    if rank != 0:
        store.wait(["fdb_group_done"])
    if rank == 0:
        try:
            import ctypes as C
            grp = fdb200.DeviceGroup(list(range(world)))
            xs = np.random.default_rng(1).random(N_ELEMS); gout = np.zeros(N_ELEMS)
            for a in (xs, gout):
                ctx.lib.fdb_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
            grp.gradient(F.rosenbrock, xs, CHUNK, gout)
            wall = []
            for _ in range(5):
                t0 = time.perf_counter(); _, val = grp.gradient(F.rosenbrock, xs, CHUNK, gout); wall.append(time.perf_counter() - t0)
            a = np.zeros(N_ELEMS)
            a[:-1] += -2.0 * (1.0 - xs[:-1]) - 400.0 * xs[:-1] * (xs[1:] - xs[:-1] ** 2)
            a[1:] += 200.0 * (xs[1:] - xs[:-1] ** 2)
            out["single_process_device_group_c2"] = {"devices": world, "ms_per_eval_host_wall": float(np.median(wall)) * 1e3,
                                                     "evals_per_s": 1.0 / float(np.median(wall)),
                                                     "verified": {"ok": bool(np.allclose(gout, a, rtol=1e-12, atol=1e-11)), "against": "closed-form gradient, all entries"},
                                                     "api": "fdb_multi_gradient_host: one process, one context per device, no inter-GPU traffic"}
            for a_ in (xs, gout):
                ctx.lib.fdb_host_unregister(C.c_void_p(a_.ctypes.data))
            grp.close()
        except Exception as e:
            out["single_process_device_group_c2"] = {"error": repr(e)}
        store.set("fdb_group_done", "1")
    sync_all()
    return out


def extras(ctx, torch, dev, flush, hbm_peak, dp_peak, under_profiler=False):
    """Other BASELINE.json configs and the generic HBM-bound kernels, each timed with CUDA events (L2 flushed between reps) and
    each CHECKED, outside the timed region, against the CPU oracle on a sample of chunk sweeps (oracle/checks.py; the same
    checks as tests/test_gpu_fullsize.py) or against a closed form: every entry carries `verified`."""
    import fdb200
    from fdb200 import functions as F
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import checks
    import orc
    o = orc.Oracle()
    lib = ctx.lib
    out = {}

    def t(fn, reps=3):
        """median device time of fn() in ms.  L2 is flushed first; then the GPU is parked in a short spin kernel while the host
        enqueues [event, fn's launches, event], so the interval between the events is device work only -- without it the
        15-20 us the host needs to get a launch through Python / ctypes sit inside the interval (visible on the sub-100 us kernels)."""
        fn(); torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda._sleep(400_000)                      # ~0.2 ms at 1.965 GHz
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            best.append(e0.elapsed_time(e1))
        return float(np.median(best))

    rng = np.random.default_rng(1)
    xs_host = rng.random(N_ELEMS)
    x = ctx.array(xs_host); g = ctx.zeros(N_ELEMS); v = ctx.zeros(1)
    fid = fdb200._lib.FN
    closed = checks.rosenbrock_closed_form(xs_host)
    # lane sharing off: all 12 lanes evaluated for every term of every chunk
    ctx.set_lane_sharing(False)
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["rosenbrock"], x.h, CHUNK, 0, 8192, g.h, v.h)), reps=2)
    ctx.set_lane_sharing(True)
    full_s = ms * 1e-3 * 83334 / 8192
    gh = g.numpy()[:8192 * CHUNK]
    out["rosenbrock_gradient_all_lanes"] = {"evals_per_s": 1.0 / full_s, "timed": "8192 of 83334 chunks, scaled",
                                            "fp64_frac": DP_OPS_PER_TERM_FULL * (N_ELEMS - 1) * 8192 / (ms * 1e-3) / dp_peak,
                                            "verified": {"ok": bool(np.allclose(gh, closed[:8192 * CHUNK], rtol=1e-12, atol=1e-11)),
                                                         "against": "closed-form gradient, all 98304 entries of the timed chunk range"}}
    # opt-in chunk grouping (fdb_set_chunk_grouping): a block sweeps G chunks and evaluates the terms far from all of them once.  NOT the
    # headline: there the ceil(n/N) sweeps stay independent evaluations of f, as in the reference's loop.  Same gradient (checked).
    grouped = {}
    for G in (8, 32):
        ctx.set_chunk_grouping(G)
        ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["rosenbrock"], x.h, CHUNK, 0, -1, g.h, v.h)), reps=3)
        ops = DP_OPS_PER_TERM_SHARED * (N_ELEMS - 1) * -(-83334 // G)
        grouped[f"group_{G}"] = {"evals_per_s": 1e3 / ms, "ms": ms, "far_fp64_frac": ops / (ms * 1e-3) / dp_peak,
                                 "verified": {"ok": bool(np.array_equal(g.numpy(), closed) or np.allclose(g.numpy(), closed, rtol=1e-12, atol=1e-11)),
                                              "against": "closed-form gradient, all entries"}}
    ctx.set_chunk_grouping(1)
    grouped["note"] = ("fdb_set_chunk_grouping(G): the far terms' Dual{1} evaluation is chunk-invariant and is shared by the G chunks of a block; "
                       "far_fp64_frac counts only the shared far evaluations (20 ops x terms x ceil(chunks/G)) against the FP64 issue rate")
    out["rosenbrock_gradient_c2_grouped_chunks"] = grouped
    # Ackley at C2's size.  Roofline: FP64 thread-level instructions of one launch, from the ncu capture of this kernel
    # (profiles/r02_ackley_sweep_raw.csv: smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on.sum; libdevice sincos is
    # FP64 polynomial code, so they are all on the FP64 pipe), against the FP64 issue rate
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["ackley"], x.h, CHUNK, 0, -1, g.h, v.h)), reps=2)
    ga = g.numpy(); va = float(v.numpy()[0])
    ver = checks.gradient_sample(o, "ackley", xs_host, CHUNK, ga, va)
    ca = checks.ackley_closed_form(xs_host)
    ver["closed_form_max_rel_err"] = float(np.max(np.abs(ga - ca)) / np.max(np.abs(ca)))
    ver["ok"] = bool(ver["ok"] and ver["closed_form_max_rel_err"] <= 1e-12)
    ack_issue = ACKLEY_FP64_ISSUE_PER_TERM * (N_ELEMS) * 83334 / (ms * 1e-3)
    out["ackley_gradient_c2"] = {"evals_per_s": 1e3 / ms, "ms": ms, "verified": ver,
                                 "roofline": {"bound": "fp64", "achieved": ack_issue / 1e12, "peak": dp_peak / 1e12, "unit": "T FP64 instr/s",
                                              "frac": ack_issue / dp_peak,
                                              "algorithmic": f"{ACKLEY_FP64_ISSUE_PER_TERM} FP64 pipe instructions (DADD+DMUL+DFMA, thread level) per element per chunk sweep "
                                                             "(one sincos + x^2 on Dual{1}; ncu count of the launch / (n x chunks)) x 10^6 x 83334",
                                              "peak_source": "fdb_measure_fp64_peak (FP64 instruction issue rate; a DFMA issues like a DMUL)"}}
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
                                                               "instructions": int(len(prog.term)), "jit": {k: st[k] for k in ("compiled", "cache_hits", "failures")},
                                                               "verified": {"ok": bool(np.allclose(g.numpy(), closed, rtol=1e-11, atol=1e-11)),
                                                                            "against": "closed-form gradient, all entries"}}
    ctx.set_jit(False)
    ms = t(lambda: run_prog(8192), reps=2)
    ctx.set_jit(True)
    out["rosenbrock_gradient_generic_callable_interpreted"] = {"evals_per_s": 1.0 / (ms * 1e-3 * 83334 / 8192), "timed": "8192 of 83334 chunks, scaled",
                                                               "verified": {"ok": bool(np.allclose(g.numpy()[:8192 * CHUNK], closed[:8192 * CHUNK], rtol=1e-11, atol=1e-11)),
                                                                            "against": "closed-form gradient over the timed chunk range"}}
    # C1: n = 10 000, Chunk{10} (the reference's published 0.2825 s case, docs/src/user/advanced.md:70-72)
    x1h = rng.random(10_000)
    x1 = ctx.array(x1h); g1 = ctx.zeros(10_000)
    ms = t(lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid["rosenbrock"], x1.h, 10, 0, -1, g1.h, v.h)), reps=5)
    g1_ref, v1_ref = o.gradient("rosenbrock", x1h, 10)
    out["rosenbrock_gradient_c1"] = {"evals_per_s": 1e3 / ms, "ms": ms, "reference_published_s": 0.282529,
                                     "verified": {"ok": bool(np.array_equal(g1.numpy(), g1_ref) and abs(float(v.numpy()[0]) - v1_ref) <= 1e-12 * abs(v1_ref)),
                                                  "against": "oracle, all 1000 chunk sweeps, gradient bit-exact"}}
    # the same C1 gradient for a callable the tracer cannot flatten (fma on the Dual array): the reference's chunk loop on materialised
    # duals -- seed! -> f(xdual) -> extract_gradient_chunk! per chunk (src/gradient.jl:114-159) -- replayed as one CUDA graph
    import warnings
    def f_loop(xd):
        d = xd[1:] - xd[:-1] * xd[:-1]
        return (fdb200.fma(d, d * 100.0, (1.0 - xd[:-1]) ** 2)).sum()
    cfg_loop = fdb200.GradientConfig(f_loop, x1h, fdb200.Chunk(10))
    def run_loop():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fdb200.gradient_b(g1, f_loop, x1, cfg_loop, ctx=ctx)
    if under_profiler:
        out["rosenbrock_gradient_c1_loop_path"] = {"skipped": "--under-profiler: ~11 000 kernel launches per call"}
    else:
        run_loop(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            run_loop()
        ctx.sync()
        wall_loop = (time.perf_counter() - t0) / 5 * 1e3
        out["rosenbrock_gradient_c1_loop_path"] = {"ms_host_wall_per_call": wall_loop, "evals_per_s": 1e3 / wall_loop, "path": cfg_loop.last_path,
                                                   "verified": {"ok": bool(np.allclose(g1.numpy(), g1_ref, rtol=1e-12, atol=1e-12)),
                                                                "against": "oracle gradient (catalogue Rosenbrock), all entries"},
                                                   "api": "fdb200.gradient_b on a callable using fma (not traceable): 1000 chunk iterations of seed/f/extract replayed as a CUDA graph"}
    # C4: Brusselator Jacobian n = 65 536, Chunk{8}: 34.4 GB dense column-major result
    M = 32768; n = 2 * M
    xbh = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    xb = ctx.array(xbh)
    J = fdb200.B200Array.alloc(ctx, n * n); y = ctx.zeros(n)
    al = (M + 1) ** 2 / 50.0
    ms = t(lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["brusselator"], xb.h, n, 8, None, n, None, al, 0, -1, J.h, n, y.h)))
    out["brusselator_jacobian_c4"] = {"evals_per_s": 1e3 / ms, "ms": ms, "verified": checks.brusselator_sample(ctx, o, J, y, xbh, al, 8),
                                      "roofline": {"bound": "hbm", "achieved": 8.0 * n * n / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                   "frac": 8.0 * n * n / (ms * 1e-3) / 1e9 / hbm_peak,
                                                   "algorithmic": "8*m*n bytes of J written once (implicit seeds: no dual reads); the peak is the measured COPY "
                                                                  "rate (read+write), a write-only stream may exceed it"}}
    del J
    # C5: Rosenbrock Hessian n = 2048, Chunk{8}.  Roofline as for Ackley: ncu FP64 instruction count of the launch
    # (profiles/r02_hessian_c5_raw.csv) / (terms x chunk pairs)
    nh = 2048
    xhh = rng.random(nh)
    xh = ctx.array(xhh); H = ctx.zeros(nh * nh); ghd = ctx.zeros(nh)
    ms = t(lambda: ctx.check(lib.fdb_hessian_chunked(ctx.h, fid["rosenbrock"], xh.h, 8, 0, -1, H.h, nh, ghd.h, v.h)))
    Hh = H.numpy().reshape((nh, nh), order="F")
    hess_issue = HESSIAN_FP64_ISSUE_PER_TERM * (nh - 1) * 256 * 256 / (ms * 1e-3)
    out["rosenbrock_hessian_c5"] = {"evals_per_s": 1e3 / ms, "ms": ms,
                                    "verified": checks.hessian_sample(o, "rosenbrock", xhh, 8, Hh, ghd.numpy(), float(v.numpy()[0])),
                                    "roofline": {"bound": "fp64", "achieved": hess_issue / 1e12, "peak": dp_peak / 1e12, "unit": "T FP64 instr/s",
                                                 "frac": hess_issue / dp_peak,
                                                 "algorithmic": f"{HESSIAN_FP64_ISSUE_PER_TERM} FP64 pipe instructions per term per (inner, outer) chunk pair "
                                                                "(far terms on 4 shared lanes, evaluated once per block of 8 inner chunks; near terms that read one window only on shared lanes "
                                                                "of the other level; ncu count of the launch / (2047 x 65536)) x 2047 terms x 256^2 pairs.  The kernel is latency-bound "
                                                                "now (FP64 pipe 37 % active): 7.4x fewer instructions than the one-pair-per-block kernel bought 3.3x in time",
                                                 "peak_source": "fdb_measure_fp64_peak"}}
    # C3: tanh.(A*x .+ b) Jacobian n = m = 8192, Chunk{12} on the FP64 tensor cores
    m3 = 8192
    A_h = (rng.uniform(-1, 1, m3 * m3) / np.sqrt(m3)).reshape((m3, m3), order="F"); b_h = rng.uniform(-1, 1, m3); x3_h = rng.uniform(-1, 1, m3)
    A = ctx.array(A_h.ravel(order="F")); b = ctx.array(b_h); x3 = ctx.array(x3_h)
    J3 = ctx.zeros(m3 * m3); y3 = ctx.zeros(m3)
    run_c3 = lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["tanh_affine"], x3.h, m3, 12, A.h, m3, b.h, 0.0, 0, -1, J3.h, m3, y3.h))
    # FP64 tensor-core denominator: cuBLAS DGEMM 8192^3 on this GPU (MEASURED_PEAKS.json has no FP64 figure)
    ta = torch.randn(m3, m3, dtype=torch.float64, device=dev); tb = torch.randn(m3, m3, dtype=torch.float64, device=dev)
    dgemm_ms = t(lambda: torch.matmul(ta, tb), reps=3)
    dgemm_tf = 2.0 * m3 ** 3 / (dgemm_ms * 1e-3) / 1e12
    del ta, tb
    ms = t(run_c3, reps=3)
    fl_shared = 2.0 * m3 * m3 * 2 * 683             # what the lane-shared sweeps contract: value + representative zero lane per chunk
    out["tanh_affine_jacobian_c3"] = {"evals_per_s": 1e3 / ms, "ms": ms, "kernels": "pack_b + dmma_gemm_kernel with the fused tanh / extraction epilogue",
                                      "verified": checks.tanh_affine_sample(ctx, o, J3, y3, A_h, b_h, x3_h, 12),
                                      "roofline": {"bound": "tensor", "achieved": fl_shared / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                                                   "unit": "TFLOP/s", "frac": fl_shared / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                   "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run",
                                                   "algorithmic": "lane sharing: per chunk sweep only the value lane and ONE representative zero-partial lane go through the "
                                                                  "contraction (2*m*n flops each), 683 sweeps = 0.183 TFLOP -- a 7.4x algorithmic saving over the reference's "
                                                                  "13 lanes per sweep; whole call incl. pack and the fused epilogue (reads A once, writes J once)"}}
    # the same target written by the user as an ordinary function of the Dual array: `A @ x` is a MATVEC node of the trace, the rest
    # an elementwise program specialised by NVRTC; contraction on the same DMMA kernel (fdb_program_matvec_jacobian_chunked).
    # Timed through the public API call on device-resident x and J (tracing + parameter binding on the host are inside the wall time)
    f_user = lambda xd: fdb200.tanh(A_h @ xd + b_h)
    cfg_user = fdb200.JacobianConfig(f_user, x3_h, fdb200.Chunk(12))
    run_user = lambda: fdb200.jacobian_b(J3, f_user, x3, cfg_user, ctx=ctx)
    run_user(); ctx.sync()
    ms_user = t(run_user, reps=3)
    t0 = time.perf_counter()
    for _ in range(5):
        run_user()
    ctx.sync()
    wall_user = (time.perf_counter() - t0) / 5 * 1e3
    # checked on a result array cleared first (the catalogue call above left the same numbers in J3); y is returned only to DiffResult targets
    ctx.check(lib.fdb_fill(ctx.h, J3.h, 0.0)); run_user(); ctx.sync()
    ver_user = checks.tanh_affine_sample(ctx, o, J3, None, A_h, b_h, x3_h, 12)
    out["tanh_affine_jacobian_c3_user_function"] = {"evals_per_s": 1e3 / ms_user, "ms": ms_user, "ms_host_wall_per_call": wall_user, "path": cfg_user.last_path,
                                                    "vs_catalogue": ms_user / ms, "verified": ver_user,
                                                    "api": "fdb200.jacobian_b(J, lambda x: tanh(A @ x + b), x, JacobianConfig(f, x, Chunk(12))) on device-resident x, J"}
    ctx.set_lane_sharing(False)
    ms = t(run_c3, reps=2)
    ver_all = checks.tanh_affine_sample(ctx, o, J3, y3, A_h, b_h, x3_h, 12)
    ctx.set_lane_sharing(True)
    fl_survey = 2.0 * m3 * m3 * (m3 + 1)            # every one-hot lane once + one value column = 1.0996 TFLOP
    out["tanh_affine_jacobian_c3_all_lanes"] = {"evals_per_s": 1e3 / ms, "ms": ms, "verified": ver_all,
                                                "roofline": {"bound": "tensor", "achieved": fl_survey / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                                                             "unit": "TFLOP/s", "frac": fl_survey / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                             "algorithmic": "SURVEY.md 8d count 2*m*n*(n+1) = 1.0996 TFLOP: A against [x | every one-hot seed]; whole call "
                                                                            "(operand build, pack, contraction, epilogue)"}}
    out["cublas_dgemm_8192_tflops"] = dgemm_tf
    # two layers tanh.(A2*tanh.(A1*x .+ b1) .+ b2), n = h = m = 8192: the second contraction has DENSE partials
    A2_h = (rng.uniform(-1, 1, m3 * m3) / np.sqrt(m3)).reshape((m3, m3), order="F"); b2_h = rng.uniform(-1, 1, m3)
    A2 = ctx.array(A2_h.ravel(order="F")); b2 = ctx.array(b2_h)
    ms = t(lambda: ctx.check(lib.fdb_jacobian_mlp2_chunked(ctx.h, x3.h, m3, m3, 12, A.h, m3, b.h, A2.h, m3, b2.h, 0, -1, J3.h, m3, y3.h)), reps=2)
    fl = 2.0 * m3 * m3 * (2 * 683 + 13 * 683)
    # closed form on a sample of rows: J[r, :] = (1 - y_r^2) * A2[r, :] @ (diag(1 - h^2) A1)
    h1 = np.tanh(A_h @ x3_h + b_h); yy = np.tanh(A2_h @ h1 + b2_h)
    rows = np.array([0, 1, 4095, 8190, 8191])
    Jrows = np.stack([checks._dev_block(ctx, J3, c * m3, m3)[rows] for c in (0, 1, 11, 12, 4096, 8184, 8191)], axis=1)
    refrows = ((1 - yy[rows] ** 2)[:, None] * A2_h[rows, :]) @ ((1 - h1 ** 2)[:, None] * A_h[:, [0, 1, 11, 12, 4096, 8184, 8191]])
    err = float(np.max(np.abs(Jrows - refrows)) / np.max(np.abs(refrows)))
    out["mlp2_jacobian_8192_dense_partials"] = {"evals_per_s": 1e3 / ms, "ms": ms,
                                                "verified": {"ok": bool(err <= 1e-11 and np.allclose(y3.numpy(), yy, rtol=1e-11, atol=1e-13)), "max_rel_err": err,
                                                             "against": "closed form J = diag(1-y^2) A2 diag(1-h^2) A1 on 5 rows x 7 columns (first/last chunk), and y"},
                                                "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": dgemm_tf, "unit": "TFLOP/s",
                                                             "frac": fl / (ms * 1e-3) / 1e12 / dgemm_tf,
                                                             "algorithmic": "layer 1: 2 shared columns per chunk; layer 2: 13 dense columns per chunk; 683 chunks"}}
    del A, J3, A2
    # generic HBM-bound kernels on materialised Dual{N} arrays: n x N sweep, arrays > L2 or L2 flushed between reps
    gk = {}
    for nn, NN in ((1 << 22, 12), (1 << 20, 12), (1 << 24, 12), (1 << 24, 1), (1 << 24, 4), (1 << 24, 8), (1 << 22, 4)):
        a = fdb200.B200DualArray.alloc(ctx, nn, NN); bb = fdb200.B200DualArray.alloc(ctx, nn, NN); oo = fdb200.B200DualArray.alloc(ctx, nn, NN)
        o1 = fdb200.B200DualArray.alloc(ctx, 1, NN)
        xsh = rng.uniform(0.5, 1.5, nn)
        xs = ctx.array(xsh)
        ctx.check(lib.fdb_seed_zero(ctx.h, a.h, xs.h, 0, 0)); ctx.check(lib.fdb_seed_zero(ctx.h, bb.h, xs.h, 0, 0))
        ctx.check(lib.fdb_seed(ctx.h, a.h, xs.h, 1, NN))
        dual_b = 8.0 * (NN + 1) * nn
        row = {}
        for name, fn, nbytes in (
                ("map1_exp", lambda: ctx.check(lib.fdb_map1(ctx.h, 2, oo.h, a.h)), 2 * dual_b),
                ("map1_pow2", lambda: ctx.check(lib.fdb_map1(ctx.h, 13, oo.h, a.h)), 2 * dual_b),
                ("map2_mul", lambda: ctx.check(lib.fdb_map2(ctx.h, 3, oo.h, a.h, bb.h)), 3 * dual_b),
                ("map2_scalar_add", lambda: ctx.check(lib.fdb_map2_scalar(ctx.h, 1, oo.h, a.h, 1.5, 0)), 2 * dual_b),
                ("reduce_sum", lambda: ctx.check(lib.fdb_reduce(ctx.h, 1, o1.h, a.h)), dual_b),
                ("eval_rosenbrock_materialised", lambda: ctx.check(lib.fdb_eval_scalar_fn(ctx.h, 1, o1.h, a.h)), dual_b),
                ("seed_zero_full", lambda: ctx.check(lib.fdb_seed_zero(ctx.h, bb.h, xs.h, 0, 0)), dual_b + 8.0 * nn)):
            ms = t(fn, reps=5)
            row[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak}
        # checks (outside the timed loops): sum of the seeded array, and the materialised Rosenbrock value
        ctx.check(lib.fdb_reduce(ctx.h, 1, o1.h, a.h))
        s1 = o1.numpy()[0]
        ok = abs(s1[0] - float(np.sum(xsh))) <= 1e-12 * abs(float(np.sum(xsh))) and np.array_equal(s1[1:], np.ones(NN))
        ctx.check(lib.fdb_eval_scalar_fn(ctx.h, 1, o1.h, a.h))
        r1 = o1.numpy()[0]
        ok = ok and abs(r1[0] - o.value("rosenbrock", xsh)) <= 1e-12 * abs(r1[0]) and \
            np.allclose(r1[1:], checks.rosenbrock_closed_form(xsh)[:NN], rtol=1e-12, atol=1e-11)
        ctx.check(lib.fdb_map2(ctx.h, 3, oo.h, a.h, bb.h))
        head = oo.view(0, 64).numpy()
        ok = ok and np.array_equal(head[:, 0], xsh[:64] * xsh[:64])
        row["verified"] = {"ok": bool(ok), "against": "sum and materialised Rosenbrock of the seeded array vs numpy / oracle value / closed-form partials; Dual*Dual values bit-exact"}
        gk[f"n=2^{int(np.log2(nn))},N={NN}"] = row
        del a, bb, oo, o1, xs
    out["generic_kernels"] = gk
    out["all_verified"] = bool(all(vv.get("verified", {}).get("ok", True) for vv in out.values() if isinstance(vv, dict)) and
                               all(r["verified"]["ok"] for r in gk.values()) and
                               all(r["verified"]["ok"] for r in out["rosenbrock_gradient_c2_grouped_chunks"].values() if isinstance(r, dict)))
    return out


if __name__ == "__main__":
    main()
