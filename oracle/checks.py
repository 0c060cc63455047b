"""Sampled full-size parity checks of BASELINE.json's configurations against the CPU oracle.

TEST INFRASTRUCTURE ONLY (see orc.py): used by tests/test_gpu_fullsize.py and, outside every timed
region, by bench.py to attach a `verified` record to each timed entry.  A full-size result cannot be
recomputed on the host in seconds (C2 is 83 334 sweeps over 10^6 elements), so each check recomputes
a *sample* of chunk sweeps with the oracle -- always the first chunk, a middle chunk and the last
(short) chunk, the three branches of src/gradient.jl:133-152 / src/jacobian.jl:191-210 -- and adds
the size-independent properties the target offers (closed forms, symmetry, band structure).

Tolerances are BASELINE.json's: bit-exact where every partial lane has at most two non-zero addends
(Rosenbrock, Brusselator), 1e-12 relative for anything through a reordered reduction or `A*x`.
Every function returns a plain dict {"ok": bool, "max_rel_err": float, "sampled_chunks": [...], ...}.
"""
import numpy as np

RED_RTOL = 1e-12


def sample_chunks(nchunks, extra=()):
    """first, a middle and the last chunk (plus `extra`), as (lo, hi) 0-based half-open ranges"""
    picks = sorted({0, nchunks // 2, nchunks - 1, *[c for c in extra if 0 <= c < nchunks]})
    return [(c, c + 1) for c in picks]


def _rel_err(got, ref):
    scale = max(float(np.max(np.abs(ref))) if ref.size else 0.0, 1e-300)
    return float(np.max(np.abs(got - ref)) / scale) if ref.size else 0.0


def _dev_block(ctx, arr, offset, count):
    """host copy of `count` doubles at `offset` of a plain device array"""
    return ctx.wrap(arr.ptr + 8 * int(offset), int(count), keepalive=arr).numpy()


def gradient_sample(o, fn, x, N, g, value=None, exact=False, extra=()):
    """g: the full host gradient of catalogue target `fn` (oracle name) at x, Chunk{N}."""
    n = x.size
    nchunks = o.chunk_plan(n, N)["nchunks"]
    worst, ok, sampled = 0.0, True, []
    scale = float(np.max(np.abs(g)))
    for lo, hi in sample_chunks(nchunks, extra):
        g_ref, v_ref = o.gradient(fn, x, N, chunks=(lo, hi))
        e0, e1 = lo * N, min(hi * N, n)
        err = float(np.max(np.abs(g[e0:e1] - g_ref[e0:e1]))) / max(scale, 1e-300)
        worst = max(worst, err)
        ok &= bool(np.array_equal(g[e0:e1], g_ref[e0:e1])) if exact else err <= RED_RTOL
        if hi == nchunks and value is not None:
            ok &= abs(value - v_ref) <= RED_RTOL * abs(v_ref)
        sampled.append(lo)
    return {"ok": bool(ok), "max_rel_err": worst, "sampled_chunks": sampled, "of_chunks": nchunks,
            "tolerance": "bit-exact" if exact else RED_RTOL}


def rosenbrock_closed_form(x):
    a = np.zeros(x.size)
    a[:-1] += -2.0 * (1.0 - x[:-1]) - 400.0 * x[:-1] * (x[1:] - x[:-1] ** 2)
    a[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
    return a


def ackley_closed_form(x):
    """gradient of benchmarks/cpp/benchmarks.cpp:10-21 (a=20, b=-0.2, c=2pi)"""
    a, b, c = 20.0, -0.2, 2.0 * np.pi
    n = x.size
    s = np.sqrt(np.sum(x * x) / n)
    t = np.sum(np.cos(c * x)) / n
    return (-a * b * np.exp(b * s)) * x / (n * s) + np.exp(t) * c * np.sin(c * x) / n


def tanh_affine_sample(ctx, o, J_dev, y_dev, A, b, x, N, extra=()):
    """C3: J_dev is the device m x n column-major Jacobian (leading dimension m), A the host matrix."""
    m, n = A.shape
    nchunks = o.chunk_plan(n, N)["nchunks"]
    worst, ok, sampled = 0.0, True, []
    for lo, hi in sample_chunks(nchunks, extra):
        c0, c1 = lo * N, min(hi * N, n)
        J = _dev_block(ctx, J_dev, c0 * m, m * (c1 - c0)).reshape((m, c1 - c0), order="F")
        J_ref, y_ref = o.jacobian("tanh_affine", x, m, N, A=A, b=b, chunks=(lo, hi), cols=(c0, c1))
        ok &= bool(np.allclose(J, J_ref, rtol=RED_RTOL, atol=RED_RTOL * float(np.max(np.abs(J_ref)))))
        worst = max(worst, _rel_err(J, J_ref))
        if hi == nchunks and y_dev is not None:
            ok &= bool(np.allclose(y_dev.numpy(), y_ref, rtol=RED_RTOL, atol=1e-13))
        sampled.append(lo)
    return {"ok": bool(ok), "max_rel_err": worst, "sampled_chunks": sampled, "of_chunks": nchunks, "tolerance": RED_RTOL}


def brusselator_sample(ctx, o, J_dev, y_dev, x, alpha, N, extra=()):
    """C4: sampled column blocks of the device n x n Jacobian, bit-exact, plus the band structure."""
    n = x.size
    M = n // 2
    nchunks = o.chunk_plan(n, N)["nchunks"]
    ok, sampled = True, []
    for lo, hi in sample_chunks(nchunks, extra):
        c0, c1 = lo * N, min(hi * N, n)
        J = _dev_block(ctx, J_dev, c0 * n, n * (c1 - c0)).reshape((n, c1 - c0), order="F")
        J_ref, y_ref = o.jacobian("brusselator", x, n, N, alpha=alpha, inplace=True, chunks=(lo, hi), cols=(c0, c1))
        ok &= bool(np.array_equal(J, J_ref))
        nz = np.argwhere(J != 0)
        ok &= bool((np.abs((nz[:, 0] % M) - ((nz[:, 1] + c0) % M)) <= 1).all())
        if hi == nchunks and y_dev is not None:
            ok &= bool(np.array_equal(y_dev.numpy(), y_ref))
        sampled.append(lo)
    return {"ok": bool(ok), "max_rel_err": 0.0 if ok else None, "sampled_chunks": sampled, "of_chunks": nchunks, "tolerance": "bit-exact"}


def hessian_sample(o, fn, x, N, H, grad=None, value=None, extra=()):
    """C5: H is the full host n x n Hessian; sampled OUTER chunks (column blocks) against the oracle's nested-dual
    sweeps (src/hessian.jl:14-19), symmetry of the whole matrix, and value/gradient from the last outer chunk."""
    n = x.size
    nchunks = o.chunk_plan(n, N)["nchunks"]
    worst, ok, sampled = 0.0, True, []
    scale = float(np.max(np.abs(H)))
    for lo, hi in sample_chunks(nchunks, extra):
        H_ref, g_ref, v_ref = o.hessian(fn, x, N, chunks=(lo, hi))
        c0, c1 = lo * N, min(hi * N, n)
        err = float(np.max(np.abs(H[:, c0:c1] - H_ref[:, c0:c1]))) / max(scale, 1e-300)
        worst = max(worst, err)
        ok &= err <= RED_RTOL
        if hi == nchunks:
            if grad is not None:
                ok &= bool(np.allclose(grad, g_ref, rtol=RED_RTOL, atol=RED_RTOL * float(np.max(np.abs(g_ref)))))
            if value is not None:
                ok &= abs(value - v_ref) <= RED_RTOL * abs(v_ref)
        sampled.append(lo)
    sym = float(np.max(np.abs(H - H.T))) / max(scale, 1e-300)
    ok &= sym <= RED_RTOL
    return {"ok": bool(ok), "max_rel_err": worst, "symmetry_rel_err": sym, "sampled_chunks": sampled, "of_chunks": nchunks,
            "tolerance": RED_RTOL}


def rosenbrock_hessian_closed_form(x):
    """dense closed-form Hessian of the Rosenbrock chain (tridiagonal)"""
    n = x.size
    H = np.zeros((n, n))
    i = np.arange(n - 1)
    H[i, i] += 2.0 - 400.0 * (x[1:] - x[:-1] ** 2) + 800.0 * x[:-1] ** 2
    H[i + 1, i + 1] += 200.0
    H[i, i + 1] += -400.0 * x[:-1]
    H[i + 1, i] += -400.0 * x[:-1]
    return H
