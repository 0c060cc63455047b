"""GPU parity of the FP64 tensor-core (DMMA) contraction: fdb_matmul, fdb_dual_matvec and the
batched tanh.(A*x .+ b) Jacobian that runs on it, against numpy / the CPU oracle.
A*x is a reordered (and, on DMMA, fused) summation: tolerance 1e-12 relative (north_star)."""
import numpy as np
import pytest

import fdb200
from fdb200 import functions as F

pytestmark = pytest.mark.gpu


def _fortran(ctx, M):
    return ctx.array(np.asfortranarray(M).ravel(order="F"))


@pytest.mark.parametrize("m,k,ncols", [(128, 32, 64), (256, 96, 13), (130, 64, 70), (512, 256, 200), (64, 32, 5), (129, 33, 7)])
def test_matmul_vs_numpy(ctx, m, k, ncols):
    rng = np.random.default_rng(m + k + ncols)
    A = rng.uniform(-1, 1, (m, k)); B = rng.uniform(-1, 1, (k, ncols))
    Ad, Bd = _fortran(ctx, A), _fortran(ctx, B)
    Cd = ctx.zeros(m * ncols)
    ctx.check(ctx.lib.fdb_matmul(ctx.h, Ad.h, m, m, k, Bd.h, k, ncols, Cd.h, m))
    C = Cd.numpy().reshape((m, ncols), order="F")
    ref = A @ B
    assert np.allclose(C, ref, rtol=1e-12, atol=1e-12 * np.abs(A).sum(axis=1).max())


@pytest.mark.parametrize("N", [1, 7, 12])
def test_dual_matvec_vs_oracle(ctx, oracle, N):
    rng = np.random.default_rng(N)
    for m, n in ((128, 64), (256, 160), (50, 33)):
        A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n)
        xdual = rng.uniform(-1, 1, (n, N + 1))
        Ad = _fortran(ctx, A); xd = ctx.duals(xdual)
        yd = fdb200.B200DualArray.alloc(ctx, m, N)
        ctx.check(ctx.lib.fdb_dual_matvec(ctx.h, Ad.h, m, m, n, xd.h, yd.h))
        got = yd.numpy()
        ref = A @ xdual                                  # planes are columns: Y = A * [value | partials]
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("N", [1, 8, 12])
@pytest.mark.parametrize("share", [True, False])
def test_tanh_affine_jacobian_on_dmma(ctx, oracle, N, share):
    rng = np.random.default_rng(3)
    for m, n in ((128, 96), (256, 128), (384, 160)):       # DMMA-eligible: m >= 128 even, n % 32 == 0
        A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
        f = F.tanh_affine.with_params(A=A, b=b)
        res = fdb200.JacobianResult(np.zeros(m), x)
        ctx.set_lane_sharing(share)
        try:
            fdb200.jacobian_b(res, f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
        finally:
            ctx.set_lane_sharing(True)
        J_ref, y_ref = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b)
        assert np.allclose(res.value, y_ref, rtol=1e-13, atol=1e-15)
        assert np.allclose(res.derivs[0], J_ref, rtol=1e-12, atol=1e-15), (m, n, N, share)
        # same result from the CUDA-core kernel
        ctx.check(ctx.lib.fdb_set_dmma(ctx.h, 0))
        try:
            res2 = fdb200.JacobianResult(np.zeros(m), x)
            fdb200.jacobian_b(res2, f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
        finally:
            ctx.check(ctx.lib.fdb_set_dmma(ctx.h, 1))
        assert np.allclose(res.derivs[0], res2.derivs[0], rtol=1e-12, atol=1e-15)


def test_tanh_affine_chunk_ranges_tile_on_dmma(ctx):
    rng = np.random.default_rng(9)
    m, n, N = 256, 128, 12
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    f = F.tanh_affine.with_params(A=A, b=b)
    cfg = fdb200.JacobianConfig(f, x, fdb200.Chunk(N))
    full = fdb200.jacobian(f, x, cfg, ctx=ctx, m=m)
    out = np.zeros((m, n), order="F")
    for r in range(3):
        p = fdb200.chunk_plan(n, N, r, 3)
        part = np.zeros((m, n), order="F")
        fdb200.jacobian_b(part, f, x, cfg, ctx=ctx, chunks=(p["chunk_lo"], p["chunk_hi"]))
        out[:, p["elem_lo"]:p["elem_hi"]] = part[:, p["elem_lo"]:p["elem_hi"]]
    assert np.array_equal(out, full)


@pytest.mark.parametrize("N", [1, 7, 12])
@pytest.mark.parametrize("share", [True, False])
def test_two_layer_dense_partials_on_dmma(ctx, oracle, N, share):
    """tanh.(A2*tanh.(A1*x .+ b1) .+ b2): the second contraction sees dense partials."""
    rng = np.random.default_rng(21)
    for n, h, m in ((96, 128, 160), (130, 256, 128), (37, 50, 20)):       # last shape: not TMA-eligible -> fallback contraction
        A1 = rng.uniform(-1, 1, (h, n)) / np.sqrt(n); b1 = rng.uniform(-1, 1, h)
        A2 = rng.uniform(-1, 1, (m, h)) / np.sqrt(h); b2 = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
        f = F.mlp2.with_params(A1=A1, b1=b1, A2=A2, b2=b2)
        res = fdb200.JacobianResult(np.zeros(m), x)
        ctx.set_lane_sharing(share)
        try:
            fdb200.jacobian_b(res, f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
        finally:
            ctx.set_lane_sharing(True)
        J_ref, y_ref = oracle.jacobian_mlp2(x, N, A1, b1, A2, b2)
        assert np.allclose(res.value, y_ref, rtol=1e-13, atol=1e-15)
        assert np.allclose(res.derivs[0], J_ref, rtol=1e-11, atol=1e-14), (n, h, m, N, share)
        h1 = np.tanh(A1 @ x + b1); yy = np.tanh(A2 @ h1 + b2)
        assert np.allclose(res.derivs[0], (1 - yy ** 2)[:, None] * A2 @ ((1 - h1 ** 2)[:, None] * A1), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("N", [1, 5, 12])
def test_dual_times_dual_matrix_product(ctx, oracle, N):
    """C = A * B with Duals on both sides (the Dual*Dual body of src/dual.jl:260-300 summed over the inner index)."""
    rng = np.random.default_rng(N)
    for m, k, n in ((128, 64, 40), (256, 96, 130), (37, 19, 11), (130, 33, 7)):      # the last two take the CUDA-core contraction
        A = rng.uniform(-1, 1, (m * k, N + 1)); B = rng.uniform(-1, 1, (k * n, N + 1))
        Cd = fdb200.dual_matmul(ctx.duals(A), (m, k), ctx.duals(B), (k, n))
        got = Cd.numpy()
        ref = oracle.dual_matmul(A, B, m, k, n)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), (m, k, n, N)
        # the value plane alone is the plain product
        Av = A[:, 0].reshape((m, k), order="F"); Bv = B[:, 0].reshape((k, n), order="F")
        assert np.allclose(got[:, 0].reshape((m, n), order="F"), Av @ Bv, rtol=1e-12, atol=1e-13)
    with pytest.raises(fdb200.DimensionMismatch):
        fdb200.dual_matmul(ctx.duals(np.zeros((6, N + 1))), (2, 3), ctx.duals(np.zeros((8, N + 1))), (4, 2))


@pytest.mark.parametrize("N", [1, 5, 12])
def test_fused_epilogue_tma_stores_match_thread_stores(ctx, oracle, N):
    """the fused tile epilogue writes J either with the threads' own stores or as staged 128 x 4N blocks through TMA tensor stores
    (the default when peers are mirrored): same bits, including the short last chunk, ragged m and partial column groups"""
    rng = np.random.default_rng(N)
    for m, n in ((256, 160), (384, 203), (130, 64)):
        if n < N:
            continue
        A = rng.uniform(-1, 1, (m, n + (n % 2))) [:, :n] / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
        f = F.tanh_affine.with_params(A=A, b=b)
        out = {}
        for mode in (0, 2, 3):
            ctx.check(ctx.lib.fdb_set_dmma_tma_store(ctx.h, mode))
            try:
                out[mode] = fdb200.jacobian(f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx, m=m)
            finally:
                ctx.check(ctx.lib.fdb_set_dmma_tma_store(ctx.h, 1))
        assert np.array_equal(out[0], out[2]) and np.array_equal(out[0], out[3]), (m, n, N)
        J_ref, _ = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b)
        assert np.allclose(out[2], J_ref, rtol=1e-12, atol=1e-15)
