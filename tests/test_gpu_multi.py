"""One host process driving several GPUs (fdb_multi_*, fdb_exchange_connect_local) and the host-buffer Jacobian.

On a one-GPU box the group is built from several contexts on device 0: every code path is the same (per-context
streams, scratch and staging, chunk ranges, in-process arena linking, flag barrier, peer stores through the byte
deltas of fdb_mirror.cuh) except the cudaDeviceEnablePeerAccess call itself, which needs two devices and runs when the
box has them.  Results must match the single-context drivers bit for bit: the same kernels on the same chunk ranges.
"""
import ctypes as C
import warnings

import numpy as np
import pytest

import fdb200
from fdb200 import functions as F

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    nd = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]                 # several contexts of this process on one GPU
    if nd >= 2:
        sets.append(list(range(min(nd, 8))))    # the real thing: one context per GPU, peer access both ways
    return sets


@pytest.fixture(params=_device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def group(request):
    g = fdb200.DeviceGroup(request.param)
    yield g
    g.close()


def test_group_gradient_host_matches_single_context(ctx, group, oracle):
    rng = np.random.default_rng(3)
    for n, N in ((1003, 12), (100, 7), (12, 12), (25, 12), (200_000, 12)):
        x = rng.uniform(0.5, 1.5, n)
        for f in (F.rosenbrock, F.ackley):
            g, v = group.gradient(f, x, N)
            ref = fdb200.GradientResult(x)
            fdb200.gradient_b(ref, f, x, fdb200.GradientConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
            assert np.array_equal(g, ref.derivs[0]), (n, N, f)
            assert v == ref.value
    x = rng.uniform(0.5, 1.5, 1003)
    g, v = group.gradient(F.rosenbrock, x, 12)
    g_ref, v_ref = oracle.gradient("rosenbrock", x, 12)
    assert np.array_equal(g, g_ref) and v == pytest.approx(v_ref, rel=1e-12)
    with pytest.raises(fdb200.ChunkSizeError):
        group.gradient(F.rosenbrock, x[:5], 12)


def test_group_jacobian_host_matches_oracle_and_single_context(ctx, group, oracle):
    rng = np.random.default_rng(4)
    m, n, N = 256, 160, 12
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    f = F.tanh_affine.with_params(A=A, b=b)
    J, y = group.jacobian(f, x, N)
    J_ref, y_ref = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b)
    assert np.allclose(J, J_ref, rtol=1e-12, atol=1e-15) and np.allclose(y, y_ref, rtol=1e-13)
    single = fdb200.jacobian(f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx, m=m)
    assert np.array_equal(J, single)
    # parameters staged by the previous call are reused: same result without uploading A again
    J2, _ = group.jacobian(f, 0.5 * x, N, reuse_params=True)
    J2_ref, _ = oracle.jacobian("tanh_affine", 0.5 * x, m, N, A=A, b=b)
    assert np.allclose(J2, J2_ref, rtol=1e-12, atol=1e-15)
    # f!(dx, x) form without parameters: Brusselator, bit-exact
    M = 96; nb = 2 * M
    xb = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    al = (M + 1) ** 2 / 50.0
    fb = F.brusselator.with_params(alpha=al)
    Jb, yb = group.jacobian(fb, xb, 8)
    Jb_ref, yb_ref = oracle.jacobian("brusselator", xb, nb, 8, alpha=al, inplace=True)
    assert np.array_equal(Jb, Jb_ref) and np.array_equal(yb, yb_ref)


def test_single_context_jacobian_host_entry_point(ctx, oracle):
    """fdb_jacobian_host: chunk ranges tile the result; reuse_params survives a change of range"""
    rng = np.random.default_rng(5)
    m, n, N = 130, 70, 8
    A = np.asfortranarray(rng.uniform(-1, 1, (m, n)) / np.sqrt(n)); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    dp = C.POINTER(C.c_double)
    J = np.zeros((m, n), order="F"); y = np.zeros(m)
    nchunks = fdb200.chunk_plan(n, N)["nchunks"]
    fid = fdb200._lib.JFN["tanh_affine"]
    for k, (lo, hi) in enumerate(((0, 3), (3, nchunks))):
        ctx.check(ctx.lib.fdb_jacobian_host(ctx.h, fid, x.ctypes.data_as(dp), n, m, N, A.ctypes.data_as(dp), m, b.ctypes.data_as(dp), 0.0, int(k > 0),
                                            lo, hi, J.ctypes.data_as(dp), m, y.ctypes.data_as(dp)))
    J_ref, y_ref = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b)
    assert np.allclose(J, J_ref, rtol=1e-12, atol=1e-15) and np.allclose(y, y_ref, rtol=1e-13)
    with pytest.raises(fdb200.ChunkSizeError):
        ctx.check(ctx.lib.fdb_jacobian_host(ctx.h, fid, x.ctypes.data_as(dp), 3, m, N, A.ctypes.data_as(dp), m, b.ctypes.data_as(dp), 0.0, 0, 0, -1,
                                            J.ctypes.data_as(dp), m, None))


def test_group_hessian_host_matches_single_context(ctx, group, oracle):
    rng = np.random.default_rng(6)
    for n, N in ((100, 8), (37, 5), (8, 8)):
        x = rng.uniform(0.5, 1.5, n)
        for f, name in ((F.rosenbrock, "rosenbrock"), (F.ackley, "ackley")):
            H, g, v = group.hessian(f, x, N)
            ref = fdb200.HessianResult(x)
            fdb200.hessian_b(ref, f, x, fdb200.HessianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
            assert np.array_equal(H, ref.derivs[1]) and np.array_equal(g, ref.derivs[0]) and v == ref.value
            H_ref, g_ref, v_ref = oracle.hessian(name, x, N)
            assert np.allclose(H, H_ref, rtol=1e-12, atol=1e-12)


def test_resident_results_are_complete_on_every_device(ctx, group, oracle):
    """the sweep kernels store into every peer's arena (linked in-process) + flag barrier: gradient, Jacobian, Hessian"""
    rng = np.random.default_rng(7)
    n, N = 4099, 12
    x = rng.uniform(0.5, 1.5, n)
    npad = -(-n // 32) * 32
    M = 64; nb = 2 * M
    ar = group.arena(3 * npad + 64 + nb * nb + nb + 32 + 200 * 200 + 256)
    try:
        # gradient: x at 0, gradient at npad, value at 2*npad
        ar.upload(0, x)
        for rep in range(2):                                 # twice: the barrier epochs advance
            ar.gradient(F.rosenbrock, 0, n, N, npad, 2 * npad)
            g_ref, v_ref = oracle.gradient("rosenbrock", x, N)
            for d in range(group.size):
                assert np.array_equal(ar.download(d, npad, n), g_ref), d
                assert ar.download(d, 2 * npad, 1)[0] == pytest.approx(v_ref, rel=1e-12)
        # Brusselator Jacobian (f! form, y from the owner of the last chunk), bit-exact on every device
        xb = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
        al = (M + 1) ** 2 / 50.0
        base = 3 * npad + 64
        ar.upload(0, xb)
        ar.jacobian(F.brusselator.with_params(alpha=al), 0, nb, nb, 8, base, base + nb * nb)
        J_ref, y_ref = oracle.jacobian("brusselator", xb, nb, 8, alpha=al, inplace=True)
        for d in range(group.size):
            assert np.array_equal(ar.download(d, base, nb * nb).reshape((nb, nb), order="F"), J_ref), d
            assert np.array_equal(ar.download(d, base + nb * nb, nb), y_ref)
        # Hessian n = 200, Chunk{8}
        nh = 200
        xh = rng.uniform(0.5, 1.5, nh)
        hb = base + nb * nb + nb + 32
        ar.upload(0, xh)
        ar.hessian(F.rosenbrock, 0, nh, 8, hb, grad_off=npad, value_off=2 * npad)
        H_ref, gh_ref, vh_ref = oracle.hessian("rosenbrock", xh, 8)
        for d in range(group.size):
            assert np.allclose(ar.download(d, hb, nh * nh).reshape((nh, nh), order="F"), H_ref, rtol=1e-12, atol=1e-12), d
            assert np.allclose(ar.download(d, npad, nh), gh_ref, rtol=1e-12, atol=1e-12)
    finally:
        ar.close()


def test_resident_tanh_affine_jacobian_with_per_device_parameters(group, oracle):
    rng = np.random.default_rng(8)
    m, n, N = 256, 128, 12
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    Ad = [c.array(np.asfortranarray(A).ravel(order="F")) for c in group.contexts]
    bd = [c.array(b) for c in group.contexts]
    ar = group.arena(256 + m * n + m + 64)
    try:
        ar.upload(0, x)
        ar.jacobian(F.tanh_affine, 0, n, m, N, 256, 256 + m * n, A=Ad, b=bd)
        J_ref, y_ref = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b)
        for d in range(group.size):
            assert np.allclose(ar.download(d, 256, m * n).reshape((m, n), order="F"), J_ref, rtol=1e-12, atol=1e-15), d
            assert np.allclose(ar.download(d, 256 + m * n, m), y_ref, rtol=1e-13)
    finally:
        ar.close()
        del Ad, bd


def test_a_barrier_that_times_out_poisons_the_context():
    """a dead or slow peer must not yield a partly stale result returned as valid.  Rank 1 never arrives:
    rank 0's barrier gives up after the (shortened) timeout, and the context then fails every synchronising call."""
    g = fdb200.DeviceGroup([0, 0])
    try:
        ar = g.arena(1024)
        ar.set_timeout(0.05)
        c0 = g.contexts[0]
        with pytest.raises(fdb200.CudaError, match="never reached an exchange barrier"):
            c0.check(c0.lib.fdb_exchange_barrier(ar.xs[0]))      # synchronous context: waits for the kernel, which gives up
        n = C.c_int64()
        assert c0.lib.fdb_exchange_timeouts(ar.xs[0], C.byref(n)) == 0 and n.value == 1
        with pytest.raises(fdb200.CudaError, match="never reached an exchange barrier"):
            c0.sync()
        with pytest.raises(fdb200.CudaError):
            c0.array(np.ones(4)).numpy()
        g.contexts[1].sync()                                     # the other context is unaffected
        ar.close()
    finally:
        g.close()


def test_sharded_gradient_on_torch_default_stream(ctx):
    """torch's default stream has handle 0; it must be adopted as the legacy default stream, not read
    as "keep the context's own non-blocking stream" -- otherwise the sweep races the H2D copy of x."""
    import torch
    assert torch.cuda.current_stream().cuda_stream == 0
    n = 200_003
    c = fdb200.Context(0)
    try:
        sg = fdb200.dist.ShardedGradient(c, F.rosenbrock, n, fdb200.GradientConfig(F.rosenbrock, np.empty(n), fdb200.Chunk(12)))
        assert c.stream() == 1                                   # cudaStreamLegacy
        c.set_async(True)
        out = torch.empty(n, dtype=torch.float64).pin_memory()
        for seed in range(3):                                    # a stale x from the previous call would give the previous gradient
            x = np.random.default_rng(seed).uniform(0.5, 1.5, n)
            sg(torch.from_numpy(x).pin_memory(), out)
            a = np.zeros(n)
            a[:-1] += -2.0 * (1.0 - x[:-1]) - 400.0 * x[:-1] * (x[1:] - x[:-1] ** 2)
            a[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
            assert np.allclose(out.numpy(), a, rtol=1e-12, atol=1e-11)
        sg.close()
        c.use_own_stream()
        assert c.stream() not in (0, 1)
    finally:
        c.close()


def test_extract_gradient_chunk_rejects_bad_windows(ctx):
    """index < 1 wrote before res->data; a nested result with another N read ydual out of bounds."""
    y = fdb200.B200DualArray.alloc(ctx, 1, 4)
    res = ctx.zeros(16)
    for bad in (0, -3):
        with pytest.raises(fdb200.FdbError):
            ctx.check(ctx.lib.fdb_extract_gradient_chunk(ctx.h, res.h, y.h, bad, 4))
    y2 = fdb200.B200DualArray.alloc(ctx, 1, 4, level=2)
    res_bad = fdb200.B200DualArray.alloc(ctx, 16, 3)
    with pytest.raises(fdb200.FdbError):
        ctx.check(ctx.lib.fdb_extract_gradient_chunk(ctx.h, res_bad.h, y2.h, 1, 4))


def test_graph_capture_falls_back_to_the_eager_loop(ctx):
    """an f that reads device data back invalidates the stream capture of the replayed chunk loop; the loop
    must then run eagerly instead of surfacing a cudaErrorStreamCapture* exception.  A NaN value is a value."""
    x = np.random.default_rng(0).uniform(0.5, 1.5, 60)
    peeks = []

    def f(xd):
        y = (xd ** 2).sum()
        peeks.append(float(y.numpy()[0, 0]))                     # host read-back: illegal while capturing
        return y

    cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(6))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", fdb200.SlowPathWarning)
        res = fdb200.GradientResult(x)
        fdb200.gradient_b(res, f, x, cfg, ctx=ctx, fuse=False)
    assert np.array_equal(res.derivs[0], 2.0 * x) and res.value == pytest.approx(float(np.sum(x * x)), rel=1e-14)
    assert cfg.last_path == "loop:eager" and cfg.capture_error is not None and len(peeks) >= 10
    # same f without the read-back is captured and replayed
    g2 = lambda xd: (xd ** 2).sum()
    cfg2 = fdb200.GradientConfig(g2, x, fdb200.Chunk(6))
    fdb200.gradient(g2, x, cfg2, ctx=ctx, fuse=False)
    assert cfg2.last_path == "loop:graph"
    # a NaN function value comes back as NaN, not as "value not written"
    xn = x.copy(); xn[7] = np.nan
    r = fdb200.GradientResult(xn)
    fdb200.gradient_b(r, F.rosenbrock, xn, fdb200.GradientConfig(F.rosenbrock, xn, fdb200.Chunk(6)), ctx=ctx, chunks=(0, 10))
    assert np.isnan(r.value)


def test_paths_are_reported(ctx):
    x = np.random.default_rng(1).uniform(0.5, 1.5, 300)
    cfg = fdb200.GradientConfig(F.rosenbrock, x, fdb200.Chunk(12))
    fdb200.gradient(F.rosenbrock, x, cfg, ctx=ctx)
    assert cfg.last_path == "catalogue"
    cfg = fdb200.GradientConfig(F.rosenbrock_generic, x, fdb200.Chunk(12))
    fdb200.gradient(F.rosenbrock_generic, x, cfg, ctx=ctx)
    assert cfg.last_path == "program:specialised"
    ctx.set_jit(False)
    try:
        fdb200.gradient(F.rosenbrock_generic, x, cfg, ctx=ctx)
        assert cfg.last_path == "program:interpreted"
    finally:
        ctx.set_jit(True)
    x2 = np.random.default_rng(1).uniform(0.5, 1.5, 1200)
    f = lambda xd: fdb200.fma(xd, xd, xd).sum()              # Base.fma on arrays has no op-program instruction: the reference loop runs
    cfg = fdb200.GradientConfig(f, x2, fdb200.Chunk(12))
    with pytest.warns(fdb200.SlowPathWarning):
        g = fdb200.gradient(f, x2, cfg, ctx=ctx)
    assert cfg.last_path in ("loop:graph", "loop:eager") and cfg.trace_error
    assert np.allclose(g, 2.0 * x2 + 1.0, rtol=1e-14)


def test_persistent_dmma_kernel_with_tma_peer_stores_at_a_sharded_c3_shape(oracle):
    """the mirrored form of the fused DMMA kernel (one persistent block per SM, result blocks pushed to every peer arena by TMA
    tensor stores from two staging buffers while the next tile is contracted): several tiles per block, ragged last column group,
    short last chunk -- complete and equal on every device, and equal to the closed form"""
    g = fdb200.DeviceGroup([0, 0, 0])
    try:
        rng = np.random.default_rng(11)
        m, n, N = 2048 + 128, 4100, 12                          # 342 chunks over 3 contexts; last chunk has 8 columns
        A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
        Ad = [c.array(np.asfortranarray(A).ravel(order="F")) for c in g.contexts]
        bd = [c.array(b) for c in g.contexts]
        xoff, joff = 0, 4128
        ar = g.arena(joff + m * n + m + 64)
        try:
            ar.upload(xoff, x)
            ar.jacobian(F.tanh_affine, xoff, n, m, N, joff, joff + m * n, A=Ad, b=bd)
            y = np.tanh(A @ x + b)
            ref = (1.0 - y * y)[:, None] * A
            for d in range(g.size):
                J = ar.download(d, joff, m * n).reshape((m, n), order="F")
                assert np.max(np.abs(J - ref)) <= 1e-12 * np.max(np.abs(ref)), d
                assert np.allclose(ar.download(d, joff + m * n, m), y, rtol=1e-12, atol=1e-13)
            J0 = ar.download(0, joff, m * n)
            assert np.array_equal(J0, ar.download(1, joff, m * n)) and np.array_equal(J0, ar.download(2, joff, m * n))
            J_ref, _ = oracle.jacobian("tanh_affine", x, m, N, A=A, b=b, chunks=(341, 342), cols=(4092, 4100))
            assert np.allclose(J0.reshape((m, n), order="F")[:, 4092:], J_ref, rtol=1e-12, atol=1e-15)
        finally:
            ar.close()
            del Ad, bd
    finally:
        g.close()


def _multicast_ok(c):
    v = C.c_int(0)
    c.check(c.lib.fdb_multicast_supported(c.h, C.byref(v)))
    return bool(v.value)


def test_multicast_arena_single_rank(ctx):
    """csrc/fdb_multicast.cu on one GPU: a virtual-memory arena and its {pid, fd} handle; CUDA IPC calls refuse it, a multicast
    object over fewer than two devices is refused cleanly, and the arena behaves like any exchange arena for the drivers."""
    if not _multicast_ok(ctx):
        pytest.skip("no NVSwitch multicast support on this box")
    lib = ctx.lib
    x = C.c_void_p()
    ctx.check(lib.fdb_exchange_create_multicast(ctx.h, 4096, 0, 1, C.byref(x)))
    try:
        h = C.create_string_buffer(fdb200._lib.SHARE_HANDLE_BYTES); mc = C.create_string_buffer(fdb200._lib.SHARE_HANDLE_BYTES)
        ctx.check(lib.fdb_exchange_share_handle(x, h))
        pid, fd = np.frombuffer(h.raw, dtype=np.int32)[:2]
        import os
        assert pid == os.getpid() and fd > 2
        ipc = C.create_string_buffer(fdb200._lib.IPC_HANDLE_BYTES)
        assert lib.fdb_exchange_handle(x, ipc) != 0 and b"fdb_exchange_share_handle" in lib.fdb_last_error(ctx.h)
        assert lib.fdb_exchange_multicast_create(x, mc) != 0 and b"at least two" in lib.fdb_last_error(ctx.h)
        assert lib.fdb_exchange_set_multicast(x, 1) != 0                      # no mapping
        # the arena is ordinary device memory for the drivers
        xs = np.random.default_rng(5).uniform(0.5, 1.5, 1000)
        xd = ctx.array(xs)
        gv = C.c_void_p(); vv = C.c_void_p()
        ctx.check(lib.fdb_exchange_view(x, 0, 1000, 1000, C.byref(gv))); ctx.check(lib.fdb_exchange_view(x, 1024, 1, 1, C.byref(vv)))
        ctx.check(lib.fdb_exchange_enable(ctx.h, x))
        ctx.check(lib.fdb_gradient_chunked(ctx.h, fdb200._lib.FN["rosenbrock"], xd.h, 12, 0, -1, gv, vv))
        ctx.check(lib.fdb_exchange_enable(ctx.h, None))
        ctx.check(lib.fdb_exchange_barrier(x))
        out = np.zeros(1000)
        ctx.check(lib.fdb_download_values(gv, out.ctypes.data_as(C.POINTER(C.c_double))))
        ref = fdb200.gradient(F.rosenbrock, xs, fdb200.GradientConfig(F.rosenbrock, xs, fdb200.Chunk(12)), ctx=ctx)
        assert np.array_equal(out, ref)
        lib.fdb_array_free(gv); lib.fdb_array_free(vv)
    finally:
        ctx.check(lib.fdb_exchange_destroy(x))


def test_multicast_exchange_between_two_devices_of_one_process(oracle):
    """Two contexts on two GPUs of this process, arenas bound into one multicast object (the descriptors are dup()ed inside one
    process): each context sweeps half of the chunks with the exchange enabled, every entry is stored ONCE through the multicast
    mapping, and both arenas end up holding the complete gradient and the value -- bit-identical to a single-context gradient."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cs = [fdb200.Context(0), fdb200.Context(1)]
    try:
        if not all(_multicast_ok(c) for c in cs):
            pytest.skip("no NVSwitch multicast support on this box")
        lib = cs[0].lib
        n, N = 10_007, 12
        xs = np.random.default_rng(6).uniform(0.5, 1.5, n)
        xch = []
        for r, c in enumerate(cs):
            x = C.c_void_p()
            c.check(lib.fdb_exchange_create_multicast(c.h, n + 64, r, 2, C.byref(x)))
            xch.append(x)
        hs = []
        for r, c in enumerate(cs):
            h = C.create_string_buffer(fdb200._lib.SHARE_HANDLE_BYTES)
            c.check(lib.fdb_exchange_share_handle(xch[r], h)); hs.append(h.raw)
        mc = C.create_string_buffer(fdb200._lib.SHARE_HANDLE_BYTES)
        cs[0].check(lib.fdb_exchange_multicast_create(xch[0], mc))
        for r, c in enumerate(cs):
            c.check(lib.fdb_exchange_connect_multicast(xch[r], b"".join(hs), mc.raw))
        for r, c in enumerate(cs):                                   # every device added before anyone binds
            c.check(lib.fdb_exchange_multicast_bind(xch[r]))
        nchunks = fdb200.chunk_plan(n, N)["nchunks"]
        half = nchunks // 2
        views = []
        for r, c in enumerate(cs):
            c.set_async(True)
            xd = c.array(xs)
            gv = C.c_void_p(); vv = C.c_void_p()
            c.check(lib.fdb_exchange_view(xch[r], 0, n, n, C.byref(gv))); c.check(lib.fdb_exchange_view(xch[r], n + 32, 1, 1, C.byref(vv)))
            views.append((xd, gv, vv))
        for rep in range(2):                                          # the second sweep overwrites the first behind the pre-sweep barrier
            for r, c in enumerate(cs):
                c.check(lib.fdb_exchange_barrier(xch[r]))
            for r, c in enumerate(cs):
                xd, gv, vv = views[r]
                c.check(lib.fdb_exchange_enable(c.h, xch[r]))
                lo, hi = (0, half) if r == 0 else (half, -1)
                c.check(lib.fdb_gradient_chunked(c.h, fdb200._lib.FN["ackley"], xd.h, N, lo, hi, gv, vv))
                c.check(lib.fdb_exchange_enable(c.h, None))
                c.check(lib.fdb_exchange_barrier(xch[r]))
            for c in cs:
                c.sync()
        ref = fdb200.GradientResult(xs)
        fdb200.gradient_b(ref, F.ackley, xs, fdb200.GradientConfig(F.ackley, xs, fdb200.Chunk(N)), ctx=cs[0])
        for r, c in enumerate(cs):
            xd, gv, vv = views[r]
            out = np.zeros(n); val = np.zeros(1)
            c.check(lib.fdb_download_values(gv, out.ctypes.data_as(C.POINTER(C.c_double))))
            c.check(lib.fdb_download_values(vv, val.ctypes.data_as(C.POINTER(C.c_double))))
            assert np.array_equal(out, ref.derivs[0]), f"arena of device {r} is not the complete gradient"
            assert val[0] == ref.value
            t = C.c_int64(); c.check(lib.fdb_exchange_timeouts(xch[r], C.byref(t))); assert t.value == 0
            lib.fdb_array_free(gv); lib.fdb_array_free(vv)
        for r, c in enumerate(cs):
            c.check(lib.fdb_exchange_destroy(xch[r]))
    finally:
        for c in cs:
            c.close()


def test_readme_examples_run_and_agree_with_finite_differences(ctx):
    """The `What f may be` block of README.md, at small sizes: every example runs on the path the README names and its derivative
    agrees with central differences of the same function evaluated in NumPy."""
    rng = np.random.default_rng(11)
    n = 600
    x = rng.uniform(0.2, 1.0, n); w = rng.uniform(0.5, 1.5, n - 1); d = rng.uniform(0.0, 1.0, n)
    A = rng.uniform(-1, 1, (128, n)) / 30; b = rng.uniform(-1, 1, 128)
    cases = [
        (lambda xd: ((1.0 - xd[:-1]) ** 2 + 100.0 * (xd[1:] - xd[:-1] ** 2) ** 2).sum(),
         lambda v: ((1.0 - v[:-1]) ** 2 + 100.0 * (v[1:] - v[:-1] ** 2) ** 2).sum(), "program"),
        (lambda xd: (w * (xd[1:] - xd[:-1]) ** 2).sum() + 0.5 * ((xd - d) ** 2).sum(),
         lambda v: (w * (v[1:] - v[:-1]) ** 2).sum() + 0.5 * ((v - d) ** 2).sum(), "program"),
        (lambda xd: ((xd - xd.mean()) ** 2).sum() / (len(xd) - 1), lambda v: ((v - v.mean()) ** 2).sum() / (len(v) - 1), "program"),
        (lambda xd: fdb200.exp(xd * 1e-3).prod(), lambda v: np.exp(v * 1e-3).prod(), "program"),
        (lambda xd: fdb200.fma(xd[1:], xd[:-1], xd[:-1]).sum(), lambda v: (v[1:] * v[:-1] + v[:-1]).sum(), "loop"),
    ]
    for f, fnp, path in cases:
        cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(12))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", fdb200.SlowPathWarning)
            g = fdb200.gradient(f, x, cfg, ctx=ctx)
        assert cfg.last_path.startswith(path), (cfg.last_path, cfg.trace_error)
        for i in (0, 1, n // 2, n - 1):
            e = np.zeros(n); e[i] = 1e-6
            fd = (fnp(x + e) - fnp(x - e)) / 2e-6
            assert abs(g[i] - fd) <= 1e-6 * max(1.0, abs(fd)), (path, i, g[i], fd)
    f5 = lambda xd: fdb200.tanh(A @ xd + b)
    c5 = fdb200.JacobianConfig(f5, x, fdb200.Chunk(12))
    J = fdb200.jacobian(f5, x, c5, ctx=ctx)
    assert c5.last_path.startswith("program")
    assert np.allclose(J, (1.0 - np.tanh(A @ x + b) ** 2)[:, None] * A, rtol=1e-12, atol=1e-15)
    xs, ws = x[:96], w[:95]
    f2s = lambda xd: (ws * (xd[1:] - xd[:-1]) ** 2).sum() + fdb200.exp(xd * 0.1).sum()
    H = fdb200.hessian(f2s, xs, fdb200.HessianConfig(f2s, xs, fdb200.Chunk(8)), ctx=ctx)
    Href = np.diag(0.01 * np.exp(0.1 * xs))
    for i in range(95):
        Href[i, i] += 2 * ws[i]; Href[i + 1, i + 1] += 2 * ws[i]; Href[i, i + 1] -= 2 * ws[i]; Href[i + 1, i] -= 2 * ws[i]
    assert np.allclose(H, Href, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("group", [2, 8, 32])
def test_grouped_gradient_sweep_matches_independent_chunk_sweeps(ctx, group):
    """fdb_set_chunk_grouping: a block sweeps `group` chunks and evaluates the far terms once.  Same gradient as the independent
    sweeps -- bit for bit where every lane has at most two non-zero addends (Rosenbrock, sum of squares), to reduction tolerance
    otherwise -- for whole gradients, chunk sub-ranges, short last chunks, N = 1 and non-finite inputs (0 * Inf must still poison)."""
    lib = ctx.lib
    rng = np.random.default_rng(21)
    FN = fdb200._lib.FN
    try:
        for n, N in ((1003, 12), (25, 12), (100_003, 12), (4099, 7), (300, 1), (12, 12), (13, 12)):
            xs = rng.uniform(0.5, 1.5, n)
            x = ctx.array(xs)
            nchunks = 1 if n == N else fdb200.chunk_plan(n, N)["nchunks"]
            for name in ("rosenbrock", "ackley", "sumsq", "prod"):
                if name == "prod" and n > 2000:
                    continue
                for lo, hi in ((0, -1), (min(3, nchunks - 1), max(min(3, nchunks - 1) + 1, min(41, nchunks)))):
                    out = []
                    for g in (1, group):
                        ctx.set_chunk_grouping(g)
                        gr = ctx.array(np.full(n, -7.0)); v = ctx.array(np.array([-7.0]))
                        ctx.check(lib.fdb_gradient_chunked(ctx.h, FN[name], x.h, N, lo, hi, gr.h, v.h))
                        out.append((gr.numpy(), v.numpy()[0]))
                    (g1, v1), (g2, v2) = out
                    if name in ("rosenbrock", "sumsq"):
                        assert np.array_equal(g1, g2), (name, n, N, lo, hi)
                    else:
                        assert np.allclose(g1, g2, rtol=1e-12, atol=1e-12 * np.max(np.abs(g1))), (name, n, N, lo, hi)
                    assert v1 == v2 or abs(v1 - v2) <= 1e-12 * abs(v1), (name, n, N, v1, v2)
        # a non-finite far term poisons every lane of every chunk in both forms (the representative zero lane is really computed)
        xs = rng.uniform(0.5, 1.5, 500); xs[400] = np.inf
        x = ctx.array(xs)
        res = []
        for g in (1, group):
            ctx.set_chunk_grouping(g)
            gr = ctx.zeros(500)
            ctx.check(lib.fdb_gradient_chunked(ctx.h, FN["rosenbrock"], x.h, 12, 0, -1, gr.h, None))
            res.append(gr.numpy())
        assert np.array_equal(np.isnan(res[0]), np.isnan(res[1])) and np.isnan(res[0]).all()
    finally:
        ctx.set_chunk_grouping(1)
