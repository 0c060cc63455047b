#!/usr/bin/env python
"""Run one BASELINE.json configuration a few times (for ncu captures and quick timing).

    python tools/run_config.py c3 [--reps 3] [--no-share] [--no-dmma]
configs: c1 c2 c2ackley c2traced c3 c4 c5 c5ackley generic
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "forwarddiff.jl_b200"))
import numpy as np
import fdb200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-share", action="store_true")
    ap.add_argument("--no-dmma", action="store_true")
    ap.add_argument("--bn", type=int, default=0, help="DMMA column-tile width (0 = auto, 16, 32, 64)")
    ap.add_argument("--no-tma-store", action="store_true", help="fused DMMA epilogue: per-thread stores instead of TMA tensor stores")
    ap.add_argument("--group", type=int, default=1, help="fdb_set_chunk_grouping: chunks per block of the catalogue gradient kernel")
    ap.add_argument("--store-mode", type=int, default=-1, help="fdb_set_dmma_tma_store mode 0..3")
    a = ap.parse_args()
    ctx = fdb200.Context(0)
    lib = ctx.lib
    ctx.set_lane_sharing(not a.no_share)
    ctx.check(lib.fdb_set_dmma(ctx.h, 0 if a.no_dmma else 1))
    ctx.check(lib.fdb_set_dmma_tile(ctx.h, a.bn))
    ctx.set_chunk_grouping(a.group)
    ctx.check(lib.fdb_set_dmma_tma_store(ctx.h, a.store_mode if a.store_mode >= 0 else (0 if a.no_tma_store else 2)))
    rng = np.random.default_rng(1)
    FN, JFN = fdb200._lib.FN, fdb200._lib.JFN
    v = ctx.zeros(1)
    if a.config in ("c1", "c2", "c2ackley"):
        n, N = (10_000, 10) if a.config == "c1" else (1_000_000, 12)
        x = ctx.array(rng.random(n)); g = ctx.zeros(n)
        fid = FN["ackley"] if a.config == "c2ackley" else FN["rosenbrock"]
        fn = lambda: ctx.check(lib.fdb_gradient_chunked(ctx.h, fid, x.h, N, 0, -1, g.h, v.h))
    elif a.config == "c2traced":
        # C2 written as a plain callable, traced into an op-program and run by the kernel specialised for it (fdb_jit.cu)
        from fdb200 import trace, functions as F
        from fdb200.api import _i32p
        import ctypes as C
        n, N = 1_000_000, 12
        x = ctx.array(rng.random(n)); g = ctx.zeros(n)
        prog = trace.trace_scalar(F.rosenbrock_generic, n)
        cp = prog.consts.ctypes.data_as(C.POINTER(C.c_double))
        fn = lambda: ctx.check(lib.fdb_program_gradient_chunked(ctx.h, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi), cp,
                                                                prog.consts.size, prog.n_acc, prog.halo, prog.result_reg, x.h, N, 0, -1, g.h, v.h))
    elif a.config == "c5ackley":
        n = 2048
        x = ctx.array(rng.random(n)); H = ctx.zeros(n * n); g = ctx.zeros(n)
        fn = lambda: ctx.check(lib.fdb_hessian_chunked(ctx.h, FN["ackley"], x.h, 8, 0, -1, H.h, n, g.h, v.h))
    elif a.config == "c3":
        m = 8192
        A = ctx.array(rng.uniform(-1, 1, m * m) / np.sqrt(m)); b = ctx.array(rng.uniform(-1, 1, m)); x = ctx.array(rng.uniform(-1, 1, m))
        J = ctx.zeros(m * m); y = ctx.zeros(m)
        fn = lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, JFN["tanh_affine"], x.h, m, 12, A.h, m, b.h, 0.0, 0, -1, J.h, m, y.h))
    elif a.config == "c4":
        M = 32768; n = 2 * M
        x = ctx.array(np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)]))
        J = fdb200.B200Array.alloc(ctx, n * n); y = ctx.zeros(n)
        fn = lambda: ctx.check(lib.fdb_jacobian_chunked(ctx.h, JFN["brusselator"], x.h, n, 8, None, n, None, (M + 1) ** 2 / 50.0, 0, -1, J.h, n, y.h))
    elif a.config == "c5":
        n = 2048
        x = ctx.array(rng.random(n)); H = ctx.zeros(n * n); g = ctx.zeros(n)
        fn = lambda: ctx.check(lib.fdb_hessian_chunked(ctx.h, FN["rosenbrock"], x.h, 8, 0, -1, H.h, n, g.h, v.h))
    elif a.config == "generic":
        nn = 1 << 22
        xs = ctx.array(rng.uniform(0.5, 1.5, nn))
        d = fdb200.B200DualArray.alloc(ctx, nn, 12); o = fdb200.B200DualArray.alloc(ctx, nn, 12); o1 = fdb200.B200DualArray.alloc(ctx, 1, 12)
        ctx.check(lib.fdb_seed_zero(ctx.h, d.h, xs.h, 0, 0))

        def fn():
            ctx.check(lib.fdb_map1(ctx.h, 2, o.h, d.h))
            ctx.check(lib.fdb_map2(ctx.h, 3, o.h, d.h, d.h))
            ctx.check(lib.fdb_reduce(ctx.h, 1, o1.h, d.h))
            ctx.check(lib.fdb_eval_scalar_fn(ctx.h, 1, o1.h, d.h))
            ctx.check(lib.fdb_seed_zero(ctx.h, d.h, xs.h, 0, 0))
    else:
        raise SystemExit("unknown config")
    for i in range(a.reps):
        t0 = time.perf_counter(); fn(); ctx.sync(); dt = time.perf_counter() - t0
        print(f"{a.config} rep {i}: {dt * 1e3:.3f} ms (host wall, synchronous call)", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
