#!/usr/bin/env python
"""Reference chunk loop on materialised device duals (path A): eager launches vs the replayed CUDA graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "forwarddiff.jl_b200"))
import numpy as np
import fdb200
from fdb200 import functions as F

ctx = fdb200.Context(0)
for n, N in ((10_000, 10), (100_000, 12)):
    x = np.random.default_rng(1).uniform(0.5, 1.5, n)
    for name, f in (("rosenbrock_generic", F.rosenbrock_generic), ("ackley_generic", F.ackley_generic)):
        cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(N))
        out = {}
        for mode, kw in (("eager", dict(fuse=False, graph=False)), ("graph", dict(fuse=False, graph=True)), ("op-program interpreted", dict(fuse=True, jit=False)), ("op-program specialised", dict(fuse=True, jit=True))):
            if mode == "eager" and n > 10_000:
                continue
            if "jit" in kw:
                ctx.set_jit(kw.pop("jit"))
            fdb200.gradient(f, x, cfg, ctx=ctx, **kw)
            t0 = time.perf_counter(); g = fdb200.gradient(f, x, cfg, ctx=ctx, **kw); out[mode] = (time.perf_counter() - t0, g)
        base = out.get("graph")[1]
        print(f"n={n} N={N} {name}: " + ", ".join(f"{m} {t*1e3:.1f} ms" for m, (t, _) in out.items()),
              "| max rel diff vs graph:", max(float(np.max(np.abs(g - base)) / np.max(np.abs(base))) for _, g in out.values()))
ctx.close()
