"""BASELINE.json's configurations at their STATED sizes, through the C ABI, against the CPU oracle on a sample of
chunk sweeps (first / middle / last short chunk: the three branches of src/gradient.jl:133-152,
src/jacobian.jl:191-210, src/hessian.jl:14-19) plus the size-independent properties of each target.
The sampled comparison lives in oracle/checks.py so that bench.py attaches the very same check to every timed entry.

  C2  Rosenbrock + Ackley gradient, n = 1 000 000, Chunk{12}
  C3  jacobian of tanh.(A*x .+ b), n = m = 8192, Chunk{12}: FP64 tensor-core (DMMA) path, both lane-sharing modes
      (64 x 139 thread blocks, 256 k-slabs, the 4-stage ring wraps 64 times, short last chunk of 8 columns)
  C4  Brusselator Jacobian, n = 65 536, Chunk{8} (34.4 GB result on the device)
  C5  Rosenbrock Hessian, n = 2048, Chunk{8} (256 x 256 chunk pairs)
"""
import numpy as np
import pytest

import checks
import fdb200
from fdb200 import functions as F

pytestmark = pytest.mark.gpu


def test_c2_ackley_gradient_full_size(ctx, oracle):
    n, N = 1_000_000, 12
    x = np.random.default_rng(1).random(n)
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, F.ackley, x, fdb200.GradientConfig(F.ackley, x, fdb200.Chunk(N)), ctx=ctx)
    g = res.derivs[0]
    r = checks.gradient_sample(oracle, "ackley", x, N, g, res.value, extra=(1, 41000, 83332))
    assert r["ok"], r
    assert r["of_chunks"] == 83334 and 83333 in r["sampled_chunks"]
    # closed form: every entry, at the reduction tolerance relative to the largest entry
    a = checks.ackley_closed_form(x)
    assert np.max(np.abs(g - a)) <= 1e-12 * np.max(np.abs(a))


def test_c2_rosenbrock_gradient_full_size_lane_sharing_off_sample(ctx, oracle):
    """all 12 lanes evaluated for every term (fdb_set_lane_sharing 0): a chunk range of the full-size problem, bit-exact"""
    n, N = 1_000_000, 12
    x = np.random.default_rng(1).random(n)
    xd = ctx.array(x); gd = ctx.zeros(n); vd = ctx.zeros(1)
    ctx.set_lane_sharing(False)
    try:
        for lo, hi in ((0, 64), (41000, 41064), (83334 - 64, 83334)):
            ctx.check(ctx.lib.fdb_gradient_chunked(ctx.h, fdb200._lib.FN["rosenbrock"], xd.h, N, lo, hi, gd.h, vd.h))
    finally:
        ctx.set_lane_sharing(True)
    g = gd.numpy()
    a = checks.rosenbrock_closed_form(x)
    for lo, hi in ((0, 64), (41000, 41064), (83334 - 64, 83334)):
        e0, e1 = lo * N, min(hi * N, n)
        assert np.allclose(g[e0:e1], a[e0:e1], rtol=1e-12, atol=1e-11)
    g_ref, v_ref = oracle.gradient("rosenbrock", x, N, chunks=(83333, 83334))
    assert np.array_equal(g[83333 * N:], g_ref[83333 * N:])
    assert vd.numpy()[0] == pytest.approx(v_ref, rel=1e-12)


@pytest.mark.parametrize("share", [True, False])
def test_c3_tanh_affine_jacobian_full_size_on_dmma(ctx, oracle, share):
    m = n = 8192; N = 12
    rng = np.random.default_rng(1)
    A = (rng.uniform(-1, 1, m * n) / np.sqrt(n)).reshape((m, n), order="F")
    b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    Ad = ctx.array(A.ravel(order="F")); bd = ctx.array(b); xd = ctx.array(x)
    Jd = ctx.zeros(m * n); yd = ctx.zeros(m)
    l0 = ctx.launches()
    ctx.set_lane_sharing(share)
    try:
        ctx.check(ctx.lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["tanh_affine"], xd.h, m, N, Ad.h, m, bd.h, 0.0, 0, -1, Jd.h, m, yd.h))
    finally:
        ctx.set_lane_sharing(True)
    assert ctx.launches() - l0 >= 2
    r = checks.tanh_affine_sample(ctx, oracle, Jd, yd, A, b, x, N, extra=(1, 300, 681))
    assert r["ok"], r
    assert r["of_chunks"] == 683 and r["sampled_chunks"][-1] == 682            # the short last chunk (8 columns)
    # every column, closed form: J = diag(1 - tanh(z)^2) * A with z = A*x + b from float64 numpy (tolerance 1e-12 of the largest entry)
    y = np.tanh(A @ x + b)
    J = Jd.numpy().reshape((m, n), order="F")
    ref = (1.0 - y * y)[:, None] * A
    assert np.max(np.abs(J - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.allclose(yd.numpy(), y, rtol=1e-12, atol=1e-13)           # |z| error ~1e-15 absolute; y passes through 0


def test_c3_chunk_range_at_full_size_matches_the_full_sweep(ctx):
    """a rank's chunk range [lo, hi) of the full-size problem (what ShardedJacobian launches) gives the same columns bit for bit"""
    m = n = 8192; N = 12
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, m * n) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    Ad = ctx.array(A); bd = ctx.array(b); xd = ctx.array(x)
    Jd = ctx.zeros(m * n); J2 = ctx.zeros(m * n)
    fid = fdb200._lib.JFN["tanh_affine"]
    ctx.check(ctx.lib.fdb_jacobian_chunked(ctx.h, fid, xd.h, m, N, Ad.h, m, bd.h, 0.0, 0, -1, Jd.h, m, None))
    p = fdb200.chunk_plan(n, N, 7, 8)                                       # the last of eight ranks: includes the short chunk
    ctx.check(ctx.lib.fdb_jacobian_chunked(ctx.h, fid, xd.h, m, N, Ad.h, m, bd.h, 0.0, p["chunk_lo"], p["chunk_hi"], J2.h, m, None))
    full = Jd.numpy().reshape((m, n), order="F"); part = J2.numpy().reshape((m, n), order="F")
    assert np.array_equal(part[:, p["elem_lo"]:p["elem_hi"]], full[:, p["elem_lo"]:p["elem_hi"]])
    assert not part[:, :p["elem_lo"]].any()


def test_c5_rosenbrock_hessian_full_size(ctx, oracle):
    n, N = 2048, 8
    x = np.random.default_rng(5).random(n)
    res = fdb200.HessianResult(x)
    fdb200.hessian_b(res, F.rosenbrock, x, fdb200.HessianConfig(F.rosenbrock, x, fdb200.Chunk(N)), ctx=ctx)
    H = res.derivs[1]
    r = checks.hessian_sample(oracle, "rosenbrock", x, N, H, res.derivs[0], res.value, extra=(1, 100, 254))
    assert r["ok"], r
    assert r["of_chunks"] == 256
    ref = checks.rosenbrock_hessian_closed_form(x)
    assert np.max(np.abs(H - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.allclose(res.derivs[0], checks.rosenbrock_closed_form(x), rtol=1e-12, atol=1e-11)


def test_c5_ackley_hessian_full_size_sample(ctx, oracle):
    """the general nested-dual finalize (two sums + exp/sqrt after the reductions) at C5's size"""
    n, N = 2048, 8
    x = np.random.default_rng(6).uniform(0.5, 1.5, n)
    res = fdb200.HessianResult(x)
    fdb200.hessian_b(res, F.ackley, x, fdb200.HessianConfig(F.ackley, x, fdb200.Chunk(N)), ctx=ctx)
    r = checks.hessian_sample(oracle, "ackley", x, N, res.derivs[1], res.derivs[0], res.value)
    assert r["ok"], r


def test_c4_brusselator_jacobian_full_size_resident(ctx, oracle):
    """the whole 34.4 GB Jacobian on the device in one call; sampled column blocks bit-exact + band structure"""
    M = 32768; n = 2 * M; N = 8
    x = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    al = (M + 1) ** 2 / 50.0
    xd = ctx.array(x); yd = ctx.zeros(n)
    Jd = fdb200.B200Array.alloc(ctx, n * n)
    ctx.check(ctx.lib.fdb_jacobian_chunked(ctx.h, fdb200._lib.JFN["brusselator"], xd.h, n, N, None, n, None, al, 0, -1, Jd.h, n, yd.h))
    r = checks.brusselator_sample(ctx, oracle, Jd, yd, x, al, N, extra=(1, 4095, 4097, 8190))
    assert r["ok"], r
    assert r["of_chunks"] == 8192
