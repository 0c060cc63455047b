"""Op-programs: an arbitrary user f flattened into the interpreter's instruction set and run through
the batched chunk sweep (fdb_program_gradient_chunked / fdb_program_jacobian_chunked).
CPU part: the tracer.  GPU part: parity with the oracle and with the reference loop on materialised duals."""
import numpy as np
import pytest

import fdb200
from fdb200 import functions as F
from fdb200 import trace


# ---- tracer (no GPU) -------------------------------------------------------------------------------
def test_tracer_flattens_rosenbrock_and_ackley():
    p = trace.trace_scalar(F.rosenbrock_generic, 100)
    assert p.halo == 1 and p.n_acc == 1 and len(p.epi) == 0 and p.consts.tolist() == [1.0, 100.0]
    kinds = [int(op) >> 6 for op in p.term[:, 0]]
    assert kinds.count(trace.K_LOADX) == 2 and kinds[-1] == trace.K_ACC
    assert sorted(p.term[[k == trace.K_LOADX for k in kinds], 3].tolist()) == [0, 1]
    q = trace.trace_scalar(F.ackley_generic, 50)
    assert q.halo == 0 and q.n_acc == 2 and len(q.epi) > 0 and q.consts[1] == 1.0 / 50


def test_tracer_rejects_what_the_interpreter_cannot_run():
    pr = trace.trace_scalar(F.prod_generic, 10)                      # prod: a product accumulator, specialised kernels only
    assert pr.has_prod and pr.needs_jit and pr.term[-1, 3] == 1
    with pytest.raises(trace.TraceError):
        trace.trace_scalar(lambda x: (x * x.prod()).sum(), 10)        # a product re-entering a broadcast
    assert trace.trace_scalar(lambda x: (x * x.sum()).sum(), 10).two_pass    # reduced value inside a broadcast: two passes
    with pytest.raises(trace.TraceError):
        trace.trace_scalar(lambda x: x * 2.0, 10)                    # not a scalar
    with pytest.raises(fdb200.DimensionMismatch):
        trace.trace_scalar(lambda x: (x + x[1:]).sum(), 10)
    with pytest.raises(trace.TraceError):
        trace.trace_scalar(lambda x: (x[::2]).sum(), 10)
    a = trace.trace_array(lambda x: x[2:] - 2.0 * x[1:-1] + x[:-2], 20)
    assert a.halo == 2 and a.out_len == 18


# ---- GPU parity ---------------------------------------------------------------------------------------
@pytest.fixture(params=["specialised", "interpreted"])
def mode(request, ctx):
    """Every op-program test runs twice: through the kernel specialised at run time for the traced f (fdb_jit.cu)
    and through the per-thread interpreter.  The specialised run must really have used a specialised kernel."""
    jit = request.param == "specialised"
    ctx.set_jit(jit)
    before = ctx.jit_stats()
    yield request.param
    after = ctx.jit_stats()
    ctx.set_jit(True)
    assert after["failures"] == before["failures"], after["last_log"]
    used = (after["compiled"] + after["cache_hits"]) - (before["compiled"] + before["cache_hits"])
    assert (used > 0) == jit


def test_tracer_data_arrays_and_reductions_of_different_lengths():
    n = 20
    w, d = np.linspace(1.0, 2.0, n - 1), np.linspace(0.0, 1.0, n)
    K_DP, K_PD, K_GUARD = trace.K_DP, trace.K_PD, trace.K_GUARD
    p = trace.trace_scalar(lambda x: (w * (x[1:] - x[:-1]) ** 2).sum(), n)
    kinds = [int(i[0]) >> 6 for i in p.term]
    assert K_PD in kinds and K_GUARD not in kinds and len(p.params) == 1 and np.array_equal(p.params[0], w)
    assert p.halo == 1 and p.needs_jit
    # the same array used twice is one parameter; an expression of data only is evaluated on the host into another
    p = trace.trace_scalar(lambda x: (w * x[1:] + x[:-1] / w - (w * w) * x[1:]).sum(), n)
    assert len(p.params) == 2
    # sums over views of different lengths: one guarded group per reduction, group k runs while i + halo_k < n
    p = trace.trace_scalar(lambda x: ((x[1:] - x[:-1]) ** 2).sum() + ((x - d) ** 2).sum() + (x[2:] * 3.0).sum(), n)
    guards = [(int(i[2]), int(i[3])) for i in p.term if (int(i[0]) >> 6) == K_GUARD]
    assert [g[0] for g in guards] == [1, 0, 2] and p.halo == 2 and p.n_acc == 3 and p.needs_jit
    assert sum(g[1] + 1 for g in guards) == len(p.term)                    # every instruction belongs to a group
    with pytest.raises(fdb200.DimensionMismatch):
        trace.trace_scalar(lambda x: (w * x).sum(), n)                     # w has n - 1 entries
    with pytest.raises(trace.TraceError):
        trace.trace_scalar(lambda x: fdb200.maximum(x, d).sum(), n)        # max with a data array has no Real rule here
    assert not trace.trace_scalar(F.rosenbrock_generic, n).needs_jit
    # a reduced value re-entering a broadcast: term1 | STAGE 1 | epilogue1 | STAGE 2 | term2, pass-1 sums numbered first
    p = trace.trace_scalar(lambda x: ((x - x.mean()) ** 2).sum() + (x * 2.0).sum(), n)
    kinds = [int(i[0]) >> 6 for i in p.term]
    assert p.two_pass and kinds.count(trace.K_STAGE) == 2 and trace.K_LOADS in kinds and p.n_acc == 3
    st1, st2 = [i for i, k in enumerate(kinds) if k == trace.K_STAGE]
    assert sorted(int(i[1]) for i in p.term[:st1] if (int(i[0]) >> 6) == trace.K_ACC) == [0, 1]      # pass-1 sums: 0, 1
    assert [int(i[1]) for i in p.term[st2:] if (int(i[0]) >> 6) == trace.K_ACC] == [2]                 # pass-2 sum: 2
    with pytest.raises(trace.TraceError):                                    # a second re-entry would need a third pass
        trace.trace_scalar(lambda x: ((x - ((x - x.mean()) ** 2).sum()) ** 2).sum(), n)
    # mean and dot are sums: dot with a data vector is one Real (x) Dual multiply per term
    p = trace.trace_scalar(lambda x: fdb200.dot(d, x) + (x ** 2).mean(), n)
    assert p.n_acc == 2 and len(p.params) == 1 and 1.0 * n in p.consts.tolist()


def test_programs_mean_what_the_reference_evaluator_says(oracle):
    """tests/program_ref.py evaluates a traced program with full dual numbers in NumPy: the tracer and the instruction
    encoding (guards, stage markers, data arrays, accumulator numbering) checked on the CPU against closed forms and
    against the oracle's chunk-mode gradients."""
    import program_ref
    n = 23
    rng = np.random.default_rng(21)
    x, w, d = rng.uniform(0.5, 1.5, n), rng.uniform(1.0, 2.0, n - 1), rng.uniform(0.0, 1.0, n)
    for gen, name in ((F.rosenbrock_generic, "rosenbrock"), (F.ackley_generic, "ackley"), (F.sumsq_generic, "sumsq")):
        v, g = program_ref.evaluate(trace.trace_scalar(gen, n), x)
        g_ref, v_ref = oracle.gradient(name, x, 4)
        assert np.allclose(g, g_ref, rtol=1e-12, atol=1e-13) and v == pytest.approx(v_ref, rel=1e-13), name
    # data arrays + reductions of different lengths (guarded groups)
    f = lambda xd: (w * (xd[1:] - xd[:-1]) ** 2).sum() + 0.5 * ((xd - d) ** 2).sum() + fdb200.exp(xd[2:] * 0.1).sum()
    v, g = program_ref.evaluate(trace.trace_scalar(f, n), x)
    g_ref = (x - d).copy()
    g_ref[1:] += 2.0 * w * (x[1:] - x[:-1]); g_ref[:-1] -= 2.0 * w * (x[1:] - x[:-1]); g_ref[2:] += 0.1 * np.exp(0.1 * x[2:])
    assert np.allclose(g, g_ref, rtol=1e-13, atol=1e-14)
    assert v == pytest.approx((w * (x[1:] - x[:-1]) ** 2).sum() + 0.5 * ((x - d) ** 2).sum() + np.exp(0.1 * x[2:]).sum(), rel=1e-14)
    # two passes: variance, and logsumexp with a reduced shift (gradient = softmax)
    v, g = program_ref.evaluate(trace.trace_scalar(lambda xd: ((xd - xd.mean()) ** 2).sum() / (len(xd) - 1), n), x)
    assert v == pytest.approx(x.var(ddof=1), rel=1e-13) and np.allclose(g, 2.0 * (x - x.mean()) / (n - 1), rtol=1e-12, atol=1e-15)
    lse = lambda xd: fdb200.log(fdb200.exp(xd - xd.mean()).sum()) + xd.mean()
    v, g = program_ref.evaluate(trace.trace_scalar(lse, n), x)
    e = np.exp(x - x.max())
    assert np.allclose(g, e / e.sum(), rtol=1e-12, atol=1e-16) and v == pytest.approx(np.log(e.sum()) + x.max(), rel=1e-13)
    # array-valued (Jacobian) form with a stencil and a data array
    y, J = program_ref.evaluate(trace.trace_array(lambda xd: d[:-2] * xd[2:] - 2.0 * xd[1:-1] + xd[:-2] ** 2, n), x)
    J_ref = np.zeros((n - 2, n))
    for i in range(n - 2):
        J_ref[i, i + 2] = d[i]; J_ref[i, i + 1] = -2.0; J_ref[i, i] = 2.0 * x[i]
    assert np.allclose(J, J_ref, rtol=1e-15, atol=0) and np.allclose(y, d[:-2] * x[2:] - 2.0 * x[1:-1] + x[:-2] ** 2, rtol=1e-15)
    # every op family through one function, against central differences (the reference's own 3e-5 tolerance)
    def fall(xd):
        u = fdb200.tanh(xd[:-1]) * fdb200.log1p(xd[1:]) + 2.0 / xd[:-1] - (xd[1:] ** 2.5) / 3.0 + fdb200.atan(xd[:-1]) * fdb200.cbrt(xd[1:])
        return (fdb200.maximum(u, xd[1:]) ** 3).sum() + fdb200.sqrt((xd ** 2).sum())
    def fnum(z):
        u = np.tanh(z[:-1]) * np.log1p(z[1:]) + 2.0 / z[:-1] - z[1:] ** 2.5 / 3.0 + np.arctan(z[:-1]) * np.cbrt(z[1:])
        return (np.maximum(u, z[1:]) ** 3).sum() + np.sqrt((z ** 2).sum())
    v, g = program_ref.evaluate(trace.trace_scalar(fall, n), x)
    h = 1e-6
    fd = np.array([(fnum(x + h * np.eye(n)[i]) - fnum(x - h * np.eye(n)[i])) / (2 * h) for i in range(n)])
    assert v == pytest.approx(fnum(x), rel=1e-13) and np.allclose(g, fd, atol=3e-5)


def test_random_expression_trees_trace_to_equivalent_programs():
    """Property test (hypothesis): random broadcast expression DAGs over x[1:], x[:-1], constants and a data array are traced,
    and the program -- whatever instruction order and register reuse the tracer chose -- must evaluate (program_ref) to the
    same value and gradient as the expression applied directly to NumPy dual numbers."""
    import program_ref
    from hypothesis import assume, given, settings, strategies as st

    class ND:                                   # NumPy dual array with the operators a user function uses
        __array_ufunc__ = None

        def __init__(self, d):
            self.d = d

        def __getitem__(self, sl):
            return ND(program_ref.D(self.d.v[sl], self.d.g[sl]))

        def _lift(self, o):
            m, k = self.d.v.size, self.d.g.shape[1]
            if isinstance(o, ND):
                return o.d
            if isinstance(o, np.ndarray):
                return program_ref.D(o, np.zeros((m, k)))
            return program_ref.D.const(o, m, k)

        def _b(self, op, o, rev=False):
            a, b = (self._lift(o), self.d) if rev else (self.d, self._lift(o))
            return ND(program_ref.binary(op, a, b))

        def __add__(self, o): return self._b("add", o)
        def __radd__(self, o): return self._b("add", o, True)
        def __sub__(self, o): return self._b("sub", o)
        def __rsub__(self, o): return self._b("sub", o, True)
        def __mul__(self, o): return self._b("mul", o)
        def __rmul__(self, o): return self._b("mul", o, True)
        def __truediv__(self, o): return self._b("div", o)
        def __neg__(self): return ND(program_ref.unary("neg", self.d))
        def __pow__(self, k): return ND(program_ref.unary(f"pow{k}", self.d))
        def _trace_unary(self, op): return ND(program_ref.unary(op, self.d))      # the hook fdb200.sin etc. dispatch on
        def sum(self): return ND(program_ref.D(self.d.v.sum(keepdims=True), self.d.g.sum(axis=0, keepdims=True)))

    n = 9
    w = np.linspace(0.5, 1.5, n - 1)
    unary = {"sin": fdb200.sin, "cos": fdb200.cos, "tanh": fdb200.tanh, "atan": fdb200.atan, "abs2": fdb200.abs2, "neg": lambda t: -t,
             "pow2": lambda t: t ** 2, "pow3": lambda t: t ** 3}
    leaf = st.sampled_from(["a", "b"])
    expr = st.recursive(leaf, lambda kids: st.one_of(
        st.tuples(st.just("un"), st.sampled_from(sorted(unary)), kids),
        st.tuples(st.just("bin"), st.sampled_from(["add", "sub", "mul"]), kids, kids),
        st.tuples(st.just("con"), st.sampled_from(["add", "sub", "mul", "div", "rsub"]), kids, st.sampled_from([0.5, 2.0, -1.5, 3.0])),
        st.tuples(st.just("dat"), st.sampled_from(["mul", "add", "rsub"]), kids)), max_leaves=8)

    def build(e, X):
        if e == "a": return X[:-1]
        if e == "b": return X[1:]
        if e[0] == "un": return unary[e[1]](build(e[2], X))
        if e[0] == "bin":
            l, r = build(e[2], X), build(e[3], X)
            return l + r if e[1] == "add" else (l - r if e[1] == "sub" else l * r)
        if e[0] == "con":
            t, c = build(e[2], X), e[3]
            return {"add": lambda: t + c, "sub": lambda: t - c, "mul": lambda: c * t, "div": lambda: t / c, "rsub": lambda: c - t}[e[1]]()
        t = build(e[2], X)
        return {"mul": lambda: w * t, "add": lambda: t + w, "rsub": lambda: w - t}[e[1]]()

    def shape_of(kind, e1, e2):
        if kind == "plain":                     # two reductions + scalar arithmetic after them
            return lambda X: build(e1, X).sum() * 0.5 + fdb200.tanh(build(e2, X).sum())
        if kind == "lengths":                   # reductions over views of different lengths: guarded groups
            return lambda X: build(e1, X).sum() + (X * 2.0 - 1.0).sum() * 0.25 + fdb200.sin(X[2:]).sum()
        # a reduced value re-enters a broadcast: two passes
        return lambda X: ((build(e1, X) - build(e2, X).mean()) ** 2).sum() / 3.0 + build(e2, X).sum()

    ND.mean = lambda self: self.sum() / self.d.v.size

    @settings(max_examples=150, deadline=None)
    @given(expr, expr, st.integers(0, 10_000), st.sampled_from(["plain", "lengths", "twopass"]))
    def check(e1, e2, seed, kind):
        x = np.random.default_rng(seed).uniform(0.5, 1.5, n)
        f = shape_of(kind, e1, e2)
        try:
            prog = trace.trace_scalar(f, n)
        except trace.TraceError:                    # e.g. more than 12 live registers: such an f takes the reference loop
            assume(False)
        assert prog.two_pass == (kind == "twopass") and (kind != "lengths" or prog.needs_jit)
        v, g = program_ref.evaluate(prog, x)
        ref = f(ND(program_ref.D(x, np.eye(n)))).d
        assert np.isfinite(v)
        assert v == pytest.approx(float(ref.v[0]), rel=1e-12, abs=1e-12)
        assert np.allclose(g, ref.g[0], rtol=1e-11, atol=1e-11 * max(1.0, np.abs(ref.g).max()))

    check()

    # and the functor text generated for such programs always compiles (NVRTC, host only; a handful of examples)
    import ctypes as C
    from fdb200.api import _i32p
    lib = fdb200._lib.load()

    @settings(max_examples=9, deadline=None)
    @given(expr, expr, st.sampled_from(["plain", "lengths", "twopass"]))
    def compiles(e1, e2, kind):
        f = shape_of(kind, e1, e2)
        try:
            prog = trace.trace_scalar(f, n)
        except trace.TraceError:
            assume(False)
        log = C.create_string_buffer(1 << 15); nbytes = C.c_int64()
        rc = lib.fdb_jit_compile_check(0, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                       prog.consts.ctypes.data_as(C.POINTER(C.c_double)) if prog.consts.size else None, prog.consts.size,
                                       prog.n_acc, prog.halo, prog.result_reg, len(prog.params), 5, 0, 1, log, len(log), C.byref(nbytes))
        assert rc == 0 and nbytes.value > 1000, log.value.decode()

    compiles()


def test_specialised_kernels_compile_on_the_host():
    """NVRTC needs no GPU: the functor generated for a traced f and the sweep kernel templates compile for sm_100a."""
    import ctypes as C
    from fdb200.api import _i32p
    lib = fdb200._lib.load()
    dp = C.POINTER(C.c_double)
    jac = trace.trace_array(lambda xd: fdb200.tanh(xd[1:] * xd[:-1]) + 2.0 ** xd[:-1], 50)
    w = np.linspace(1.0, 2.0, 99)
    data = trace.trace_scalar(lambda xd: (w * (xd[1:] - xd[:-1]) ** 2 / (1.0 + w)).sum(), 100)     # closes over a data array
    assert len(data.params) == 2            # w and the host-evaluated 1.0 + w
    for kind, prog, N in ((0, trace.trace_scalar(F.rosenbrock_generic, 100), 12), (0, trace.trace_scalar(F.ackley_generic, 100), 7),
                          (2, trace.trace_scalar(F.rosenbrock_generic, 100), 8), (1, jac, 5), (0, data, 6), (2, data, 4),
                          (2, trace.trace_scalar(F.ackley_generic, 100), 8),
                          (0, trace.trace_scalar(lambda xd: ((xd - xd.mean()) ** 2).sum() / 99.0, 100), 12),                 # two passes
                          (0, trace.trace_scalar(lambda xd: ((xd[1:] - xd[:-1]) ** 2).sum() + (xd * xd).sum(), 100), 3)):  # guards
        log = C.create_string_buffer(1 << 16); nbytes = C.c_int64()
        rc = lib.fdb_jit_compile_check(kind, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                       prog.consts.ctypes.data_as(dp) if prog.consts.size else None, prog.consts.size, prog.n_acc,
                                       prog.halo, prog.result_reg, len(prog.params), N, 0, 1, log, len(log), C.byref(nbytes))
        assert rc == 0 and nbytes.value > 1000, log.value.decode()
    bad = np.array([[7 << 6, 0, 0, 0]], dtype=np.int32)                 # unknown instruction kind
    assert lib.fdb_jit_compile_check(0, _i32p(bad), 1, None, 0, None, 0, 1, 0, 0, 0, 4, 0, 1, None, 0, None) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 4, 7, 12])
def test_program_gradient_vs_oracle_and_loop(ctx, oracle, N, mode):
    for n in (13, 257, 1003):
        x = np.random.default_rng(n).uniform(0.5, 1.5, n)
        for gen, fn in ((F.rosenbrock_generic, "rosenbrock"), (F.ackley_generic, "ackley"), (F.sumsq_generic, "sumsq")):
            res = fdb200.GradientResult(x)
            fdb200.gradient_b(res, gen, x, fdb200.GradientConfig(gen, x, fdb200.Chunk(N)), ctx=ctx)      # fused op-program
            g_ref, v_ref = oracle.gradient(fn, x, N)
            scale = np.abs(g_ref).max()
            assert np.allclose(res.derivs[0], g_ref, rtol=1e-11, atol=1e-12 * scale), (fn, n, N)
            assert res.value == pytest.approx(v_ref, rel=1e-12)
            if n <= 257:
                res2 = fdb200.GradientResult(x)
                fdb200.gradient_b(res2, gen, x, fdb200.GradientConfig(gen, x, fdb200.Chunk(N)), ctx=ctx, fuse=False)   # reference loop
                assert np.allclose(res.derivs[0], res2.derivs[0], rtol=1e-11, atol=1e-12 * scale)


@pytest.mark.gpu
def test_program_gradient_lane_sharing_chunk_ranges_and_fallback(ctx, oracle, mode):
    n, N = 1000, 12
    x = np.random.default_rng(4).uniform(0.5, 1.5, n)
    cfg = fdb200.GradientConfig(F.rosenbrock_generic, x, fdb200.Chunk(N))
    a = fdb200.gradient(F.rosenbrock_generic, x, cfg, ctx=ctx)
    ctx.set_lane_sharing(False)
    try:
        b = fdb200.gradient(F.rosenbrock_generic, x, cfg, ctx=ctx)
    finally:
        ctx.set_lane_sharing(True)
    assert np.array_equal(a, b)
    out = np.zeros(n)
    for r in range(3):
        p = fdb200.chunk_plan(n, N, r, 3)
        fdb200.gradient_b(out, F.rosenbrock_generic, x, cfg, ctx=ctx, chunks=(p["chunk_lo"], p["chunk_hi"]))
    assert np.array_equal(out, a)
    # prod cannot be flattened: the call silently takes the reference loop on materialised duals
    xs = x[:40]
    g = fdb200.gradient(F.prod_generic, xs, fdb200.GradientConfig(F.prod_generic, xs, fdb200.Chunk(4)), ctx=ctx)
    assert np.allclose(g, oracle.gradient("prod", xs, 4)[0], rtol=1e-11)
    # a user function with every op family: unary rules, Real (x) Dual both ways, literal and real powers, max
    def f(xd):
        u = fdb200.tanh(xd[:-1]) * fdb200.log1p(xd[1:]) + 2.0 / xd[:-1] - (xd[1:] ** 2.5) / 3.0
        return (fdb200.maximum(u, xd[1:]) ** 3).sum() + fdb200.sqrt((xd ** 2).sum())
    gf = fdb200.gradient(f, xs, fdb200.GradientConfig(f, xs, fdb200.Chunk(5)), ctx=ctx)
    gl = fdb200.gradient(f, xs, fdb200.GradientConfig(f, xs, fdb200.Chunk(5)), ctx=ctx, fuse=False)
    assert np.allclose(gf, gl, rtol=1e-11, atol=1e-12 * np.abs(gl).max())
    h = 1e-6
    def fv(z):
        u = np.tanh(z[:-1]) * np.log1p(z[1:]) + 2.0 / z[:-1] - z[1:] ** 2.5 / 3.0
        return (np.maximum(u, z[1:]) ** 3).sum() + np.sqrt((z ** 2).sum())
    fd = np.array([(fv(xs + h * np.eye(40)[i]) - fv(xs - h * np.eye(40)[i])) / (2 * h) for i in range(40)])
    assert np.allclose(gf, fd, atol=3e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 6, 12])
def test_program_jacobian_elementwise_and_stencil(ctx, N, mode):
    n = 61
    x = np.random.default_rng(6).uniform(0.5, 1.5, n)

    def f(xd):
        return fdb200.sin(xd) * xd + xd ** 3 / 2.0

    J = fdb200.jacobian(f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(J, np.diag(np.cos(x) * x + np.sin(x) + 1.5 * x * x), rtol=1e-14, atol=1e-15)
    Jl = fdb200.jacobian(f, x, fdb200.JacobianConfig(f, x, fdb200.Chunk(N)), ctx=ctx, m=n, fuse=False)
    assert np.array_equal(J, Jl)                                   # same device functions per op: bit identical

    def lap(xd):
        return xd[2:] - 2.0 * xd[1:-1] + xd[:-2] ** 2

    res = fdb200.JacobianResult(np.zeros(n - 2), x)
    fdb200.jacobian_b(res, lap, x, fdb200.JacobianConfig(lap, x, fdb200.Chunk(N)), ctx=ctx)
    expect = np.zeros((n - 2, n))
    for i in range(n - 2):
        expect[i, i + 2] = 1.0; expect[i, i + 1] = -2.0; expect[i, i] = 2.0 * x[i]
    assert np.array_equal(res.derivs[0], expect)
    assert np.allclose(res.value, x[2:] - 2.0 * x[1:-1] + x[:-2] ** 2, rtol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 5, 12])
def test_functions_closing_over_data_arrays(ctx, N, mode):
    """f(x) = sum(w .* (x[2:end] .- x[1:end-1]).^2 ./ (1 .+ w)) + sum((x[1:end-1] .- d).^2): NumPy arrays inside f become array
    parameters of the program (Real (x) Dual rules).  Specialised kernels only; with the interpreter f takes the reference
    loop on materialised duals, which handles Vector{Float64} .(op) Vector{Dual} through fdb_map2."""
    n = 101
    rng = np.random.default_rng(8)
    x, w, d = rng.uniform(0.5, 1.5, n), rng.uniform(1.0, 2.0, n - 1), rng.uniform(0.0, 1.0, n - 1)

    def f(xd):
        return (w * (xd[1:] - xd[:-1]) ** 2 / (1.0 + w)).sum() + ((xd[:-1] - d) ** 2).sum()

    from fdb200 import trace
    assert len(trace.trace_scalar(f, n).params) == 3                 # w, 1.0 + w (evaluated on the host), d
    c = w / (1.0 + w)
    g_ref = np.zeros(n)
    g_ref[:-1] += 2.0 * (x[:-1] - d)
    g_ref[1:] += 2.0 * c * (x[1:] - x[:-1]); g_ref[:-1] -= 2.0 * c * (x[1:] - x[:-1])
    H_ref = 2.0 * np.eye(n); H_ref[-1, -1] = 0.0
    for i in range(n - 1):
        H_ref[i, i] += 2 * c[i]; H_ref[i + 1, i + 1] += 2 * c[i]; H_ref[i, i + 1] -= 2 * c[i]; H_ref[i + 1, i] -= 2 * c[i]
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, f, x, fdb200.GradientConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(res.derivs[0], g_ref, rtol=1e-12, atol=1e-13)
    assert res.value == pytest.approx((c * (x[1:] - x[:-1]) ** 2).sum() + ((x[:-1] - d) ** 2).sum(), rel=1e-13)
    gl = fdb200.gradient(f, x, fdb200.GradientConfig(f, x, fdb200.Chunk(N)), ctx=ctx, fuse=False)          # reference loop
    assert np.allclose(res.derivs[0], gl, rtol=1e-12, atol=1e-13)

    # reductions over views of different lengths: one guarded group per reduction
    d2 = rng.uniform(0.0, 1.0, n)

    def f2(xd):
        return (w * (xd[1:] - xd[:-1]) ** 2).sum() + ((xd - d2) ** 2).sum() * 0.5 + fdb200.exp(xd[2:] * 0.1).sum()

    g2_ref = (x - d2).copy()
    g2_ref[1:] += 2.0 * w * (x[1:] - x[:-1]); g2_ref[:-1] -= 2.0 * w * (x[1:] - x[:-1])
    g2_ref[2:] += 0.1 * np.exp(0.1 * x[2:])
    res2 = fdb200.GradientResult(x)
    fdb200.gradient_b(res2, f2, x, fdb200.GradientConfig(f2, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(res2.derivs[0], g2_ref, rtol=1e-12, atol=1e-13)
    assert res2.value == pytest.approx((w * (x[1:] - x[:-1]) ** 2).sum() + 0.5 * ((x - d2) ** 2).sum() + np.exp(0.1 * x[2:]).sum(), rel=1e-13)
    if mode == "specialised":
        assert trace.trace_scalar(f2, n).needs_jit
        H2 = fdb200.hessian(f2, x, fdb200.HessianConfig(f2, x, fdb200.Chunk(N)), ctx=ctx)
        H2_ref = np.eye(n)
        for i in range(n - 1):
            H2_ref[i, i] += 2 * w[i]; H2_ref[i + 1, i + 1] += 2 * w[i]; H2_ref[i, i + 1] -= 2 * w[i]; H2_ref[i + 1, i] -= 2 * w[i]
        H2_ref[np.arange(2, n), np.arange(2, n)] += 0.01 * np.exp(0.1 * x[2:])
        assert np.allclose(H2, H2_ref, rtol=1e-12, atol=1e-13)

    def f3(xd):
        return fdb200.dot(d2, xd) + (xd ** 2).mean()

    g3 = fdb200.gradient(f3, x, fdb200.GradientConfig(f3, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(g3, d2 + 2.0 * x / n, rtol=1e-13, atol=0)

    def fa(xd):
        return d * xd[1:] - xd[:-1] / w + w ** 2.0

    J = fdb200.jacobian(fa, x, fdb200.JacobianConfig(fa, x, fdb200.Chunk(N)), ctx=ctx)
    J_ref = np.zeros((n - 1, n))
    for i in range(n - 1):
        J_ref[i, i + 1] = d[i]; J_ref[i, i] = -1.0 / w[i]
    assert np.allclose(J, J_ref, rtol=1e-15, atol=0)
    if mode == "specialised":
        H = fdb200.hessian(f, x, fdb200.HessianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
        assert np.allclose(H, H_ref, rtol=1e-12, atol=1e-13)
    else:
        with pytest.raises(fdb200.UnsupportedOp):
            fdb200.hessian(f, x, fdb200.HessianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 4, 12])
def test_reduced_values_reentering_a_broadcast_two_pass(ctx, N, mode):
    """var(x) = sum((x .- mean(x)).^2) / (n - 1) and logsumexp with a reduced shift: the reduced Dual re-enters a second
    broadcast, so the program has two passes (specialised kernel: gradient_sweep2_body); with the interpreter the same f
    takes the reference loop (Vector{Dual} .- Dual through fdb_map2's broadcast of a 1-element array)."""
    n = 203
    x = np.random.default_rng(12).uniform(-1.0, 2.0, n)

    def var(xd):
        m = xd.mean()
        return ((xd - m) ** 2).sum() / (len(xd) - 1)

    def lse(xd):
        m = xd.mean()
        return fdb200.log(fdb200.exp(xd - m).sum()) + m

    from fdb200 import trace
    assert trace.trace_scalar(var, n).two_pass and trace.trace_scalar(lse, n).two_pass
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, var, x, fdb200.GradientConfig(var, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(res.derivs[0], 2.0 * (x - x.mean()) / (n - 1), rtol=1e-11, atol=1e-14)
    assert res.value == pytest.approx(x.var(ddof=1), rel=1e-13)
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, lse, x, fdb200.GradientConfig(lse, x, fdb200.Chunk(N)), ctx=ctx)
    e = np.exp(x - x.max())
    assert np.allclose(res.derivs[0], e / e.sum(), rtol=1e-11, atol=1e-15)        # softmax
    assert res.value == pytest.approx(np.log(np.exp(x - x.max()).sum()) + x.max(), rel=1e-13)
    gl = fdb200.gradient(lse, x, fdb200.GradientConfig(lse, x, fdb200.Chunk(N)), ctx=ctx, fuse=False)
    assert np.allclose(res.derivs[0], gl, rtol=1e-11, atol=1e-15)
    with pytest.raises(fdb200.UnsupportedOp):
        fdb200.hessian(var, x, fdb200.HessianConfig(var, x, fdb200.Chunk(N)), ctx=ctx)


@pytest.mark.gpu
def test_random_programs_on_the_device_match_the_reference_evaluator(ctx):
    """Random broadcast expressions (plain, guarded and two-pass shapes, with a data array) through the specialised kernels
    and, where it can run them, the interpreter -- against tests/program_ref.py on the same traced program."""
    import program_ref
    rng = np.random.default_rng(77)
    n = 47
    w = rng.uniform(0.5, 1.5, n - 1)
    un = [fdb200.sin, fdb200.cos, fdb200.tanh, fdb200.atan, fdb200.abs2, lambda t: -t, lambda t: t ** 2, lambda t: t ** 3]

    def rand_expr(depth):
        r = rng.integers(0, 6 if depth > 0 else 2)
        if r == 0: return lambda X: X[:-1]
        if r == 1: return lambda X: X[1:]
        if r == 2:
            u, e = un[rng.integers(len(un))], rand_expr(depth - 1)
            return lambda X: u(e(X))
        if r == 3:
            a, b, k = rand_expr(depth - 1), rand_expr(depth - 1), rng.integers(3)
            return lambda X: (a(X) + b(X)) if k == 0 else ((a(X) - b(X)) if k == 1 else a(X) * b(X))
        if r == 4:
            e, c, k = rand_expr(depth - 1), float(rng.choice([0.5, 2.0, -1.5])), rng.integers(3)
            return lambda X: (c * e(X)) if k == 0 else ((e(X) / c) if k == 1 else c - e(X))
        e, k = rand_expr(depth - 1), rng.integers(2)
        return lambda X: (w * e(X)) if k == 0 else (e(X) + w)

    checked = 0
    for trial in range(24):
        e1, e2, kind, N = rand_expr(3), rand_expr(2), ("plain", "lengths", "twopass")[trial % 3], int(rng.choice([1, 3, 5, 12]))
        if kind == "plain":
            f = lambda X: e1(X).sum() * 0.5 + fdb200.tanh(e2(X).sum())
        elif kind == "lengths":
            f = lambda X: e1(X).sum() + (X * 2.0 - 1.0).sum() * 0.25 + fdb200.sin(X[2:]).sum()
        else:
            f = lambda X: ((e1(X) - e2(X).mean()) ** 2).sum() / 3.0 + e2(X).sum()
        x = rng.uniform(0.5, 1.5, n)
        try:
            prog = trace.trace_scalar(f, n)
        except trace.TraceError:
            continue
        v_ref, g_ref = program_ref.evaluate(prog, x)
        scale = max(1.0, np.abs(g_ref).max())
        for jit in (True, False):
            if not jit and prog.needs_jit:
                continue
            ctx.set_jit(jit)
            try:
                res = fdb200.GradientResult(x)
                fdb200.gradient_b(res, f, x, fdb200.GradientConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
            finally:
                ctx.set_jit(True)
            assert np.allclose(res.derivs[0], g_ref, rtol=1e-10, atol=1e-11 * scale), (trial, kind, N, jit)
            assert res.value == pytest.approx(v_ref, rel=1e-11, abs=1e-11), (trial, kind, N, jit)
            checked += 1
    assert checked >= 20 and ctx.jit_stats()["failures"] == 0


# ---- the reference loop as a replayed CUDA graph (fdb_graph.cu) ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 5, 12])
def test_graph_replayed_chunk_loop_matches_eager_loop(ctx, oracle, N):
    for n in (53, 257):
        x = np.random.default_rng(n + N).uniform(0.5, 1.5, n)
        for gen, fn in ((F.prod_generic, "prod"), (F.rosenbrock_generic, "rosenbrock"), (F.ackley_generic, "ackley")):
            if fn == "prod" and n > 100:
                continue
            cfg = fdb200.GradientConfig(gen, x, fdb200.Chunk(N))
            r_graph = fdb200.GradientResult(x); r_eager = fdb200.GradientResult(x)
            fdb200.gradient_b(r_graph, gen, x, cfg, ctx=ctx, fuse=False, graph=True)
            fdb200.gradient_b(r_eager, gen, x, cfg, ctx=ctx, fuse=False, graph=False)
            assert np.array_equal(r_graph.derivs[0], r_eager.derivs[0]), (fn, n, N)        # same kernels, same order: bit identical
            assert r_graph.value == r_eager.value
            g_ref, v_ref = oracle.gradient(fn, x, N)
            assert np.allclose(r_graph.derivs[0], g_ref, rtol=1e-11, atol=1e-12 * np.abs(g_ref).max())


@pytest.mark.gpu
def test_chunk_cursor_follows_the_reference_bookkeeping(ctx):
    import ctypes as C
    lib = ctx.lib
    for n, N in ((13, 4), (24, 12), (100, 7)):
        p = fdb200.chunk_plan(n, N)
        ctx.check(lib.fdb_cursor_set(ctx.h, 1, N))
        seen = []
        for _ in range(p["nchunks"]):
            i, cs = C.c_int64(), C.c_int64()
            ctx.check(lib.fdb_cursor_get(ctx.h, C.byref(i), C.byref(cs)))
            seen.append((i.value, cs.value))
            ctx.check(lib.fdb_cursor_advance(ctx.h, N, n))
        expect = [((c - 1) * N + 1, N) for c in range(1, p["nchunks"])] + [(p["lastchunkindex"], p["lastchunksize"])]
        assert seen == expect


# ---- Hessians of flattened user functions (fdb_program_hess.cu) ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 3, 8, 12])
def test_program_hessian_vs_oracle_and_finite_differences(ctx, oracle, N, mode):
    for n in (13, 60):
        x = np.random.default_rng(n).uniform(0.5, 1.5, n)
        for gen, fn in ((F.rosenbrock_generic, "rosenbrock"), (F.sumsq_generic, "sumsq")):
            res = fdb200.HessianResult(x)
            fdb200.hessian_b(res, gen, x, fdb200.HessianConfig(gen, x, fdb200.Chunk(N)), ctx=ctx)
            H_ref, g_ref, v_ref = oracle.hessian(fn, x, N if N in (1, 2, 3, 4, 5, 6, 7, 8, 12) else 8)
            assert np.allclose(res.derivs[1], H_ref, rtol=1e-11, atol=1e-11), (fn, n, N)
            assert np.allclose(res.derivs[0], g_ref, rtol=1e-11, atol=1e-11)
            assert res.value == pytest.approx(v_ref, rel=1e-12)

    def f(xd):      # every op family that has a nested rule
        a, b = xd[:-1], xd[1:]
        return (fdb200.exp(a * b) + fdb200.sin(a) * fdb200.cos(b) + fdb200.tanh(a - b) ** 2 + fdb200.sqrt(a + b) / (1.0 + a ** 2)
                + fdb200.log(b) * 2.0 - 1.0 / a + fdb200.abs_(a - 1.0) ** 3).sum()

    def fv(z):
        a, b = z[:-1], z[1:]
        return (np.exp(a * b) + np.sin(a) * np.cos(b) + np.tanh(a - b) ** 2 + np.sqrt(a + b) / (1.0 + a ** 2)
                + np.log(b) * 2.0 - 1.0 / a + np.abs(a - 1.0) ** 3).sum()

    n = 17
    x = np.random.default_rng(3).uniform(0.6, 1.4, n)
    res = fdb200.HessianResult(x)
    fdb200.hessian_b(res, f, x, fdb200.HessianConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
    H = res.derivs[1]
    assert np.allclose(H, H.T, rtol=1e-12, atol=1e-12)
    assert res.value == pytest.approx(fv(x), rel=1e-13)
    g = fdb200.gradient(f, x, fdb200.GradientConfig(f, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.allclose(res.derivs[0], g, rtol=1e-12, atol=1e-13)
    h = 1e-5
    E = np.eye(n)
    fd = np.array([[(fv(x + h * E[i] + h * E[j]) - fv(x + h * E[i] - h * E[j]) - fv(x - h * E[i] + h * E[j]) + fv(x - h * E[i] - h * E[j])) / (4 * h * h)
                    for j in range(n)] for i in range(n)])
    assert np.allclose(H, fd, atol=2e-2 * max(1.0, np.abs(H).max()) * 1e-2 + 1e-4)      # Calculus.hessian-style tolerance (HessianTest.jl:65-88)
    if N == 1:      # test/MiscTest.jl:183: hessian(t -> abs(t[1])^2, [0.0]) == 2 pins abs'(0) = +1 through the nested rule
        fa0 = lambda xd: (fdb200.abs_(xd) ** 2).sum()
        x0 = np.array([0.0])
        assert fdb200.hessian(fa0, x0, fdb200.HessianConfig(fa0, x0, fdb200.Chunk(1)), ctx=ctx)[0, 0] == 2.0
    # arithmetic after the reductions (Ackley: two sums, then exp / sqrt on Dual{Dual{N},N}): the specialised kernel
    # finalizes on full nested duals; the interpreter has no such path and says so
    if mode == "specialised":
        for n in (13, 60):
            xa = np.random.default_rng(n).uniform(0.5, 1.5, n)
            res = fdb200.HessianResult(xa)
            fdb200.hessian_b(res, F.ackley_generic, xa, fdb200.HessianConfig(F.ackley_generic, xa, fdb200.Chunk(N)), ctx=ctx)
            H_ref, g_ref, v_ref = oracle.hessian("ackley", xa, N if N in (1, 2, 3, 4, 5, 6, 7, 8, 12) else 8)
            assert np.allclose(res.derivs[1], H_ref, rtol=1e-10, atol=1e-12), (n, N)
            assert np.allclose(res.derivs[0], g_ref, rtol=1e-10, atol=1e-12)
            assert res.value == pytest.approx(v_ref, rel=1e-12)
    else:
        with pytest.raises(fdb200.UnsupportedOp):
            fdb200.hessian(F.ackley_generic, x, fdb200.HessianConfig(F.ackley_generic, x, fdb200.Chunk(N)), ctx=ctx)


# ---- A @ x inside a user function: contraction on the FP64 tensor cores + elementwise program -----------------------------
def test_tracer_matvec_node_and_host_compile():
    """`lambda x: tanh(A @ x + b)` flattens to an elementwise program over z = A*x (LOADX 0 = z_i, b an array parameter read at
    the row index); anything that uses x besides A @ x is rejected.  The specialised kernel (kind 3) compiles on the host."""
    import ctypes as C
    from fdb200.api import _i32p
    rng = np.random.default_rng(0)
    A = rng.random((40, 24)); b = rng.random(40)
    prog = trace.trace_array(lambda x: fdb200.tanh(A @ x + b), 24)
    assert prog.matvec is A and prog.out_len == 40 and prog.halo == 0 and len(prog.params) == 1 and prog.needs_jit
    kinds = [int(k) >> 6 for k in prog.term[:, 0]]
    assert kinds == [trace.K_LOADX, trace.K_DP, trace.K_UNARY]
    plain = trace.trace_array(lambda x: fdb200.exp(0.5 * (A @ x)) * 2.0, 24)
    assert plain.matvec is A and not plain.needs_jit
    sq = rng.random((24, 24))
    for bad in (lambda x: fdb200.tanh(sq @ x) + x, lambda x: A @ (2.0 * x), lambda x: (A @ x)[1:], lambda x: A @ x[:24] + (A @ x)):
        with pytest.raises((trace.TraceError, fdb200.DimensionMismatch)):
            trace.trace_array(bad, 24)
    with pytest.raises(trace.TraceError, match="reduction over A @ x"):
        trace.trace_scalar(lambda x: fdb200.tanh(A @ x).sum(), 24)
    with pytest.raises(fdb200.DimensionMismatch):
        trace.trace_array(lambda x: A @ x, 23)
    lib = fdb200._lib.load()
    for p, N in ((prog, 12), (plain, 5)):
        log = C.create_string_buffer(1 << 16); nbytes = C.c_int64()
        rc = lib.fdb_jit_compile_check(3, _i32p(p.term), len(p.term), None, 0, p.consts.ctypes.data_as(C.POINTER(C.c_double)) if p.consts.size else None,
                                       p.consts.size, 0, 0, p.result_reg, len(p.params), N, 0, 1, log, len(log), C.byref(nbytes))
        assert rc == 0 and nbytes.value > 1000, log.value.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 8, 12])
@pytest.mark.parametrize("share", [True, False])
def test_user_function_through_a_linear_map_runs_on_the_tensor_cores(ctx, oracle, N, share, mode):
    """jacobian(x -> tanh.(A*x .+ b), x, cfg) written by the USER (src/jacobian.jl:220 `ydual = f(xdual)` with Julia's generic
    matvec over Dual): same numbers as the catalogue target and the oracle, contraction on the DMMA kernel."""
    rng = np.random.default_rng(5)
    for m, n in ((256, 128), (384, 160), (50, 33)):          # last shape: not TMA-eligible -> CUDA-core contraction
        A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
        f = lambda xd: fdb200.tanh(A @ xd + b)
        if mode == "interpreted":                             # array parameters exist in the specialised kernels only
            f = lambda xd: fdb200.tanh(A @ xd + 0.25)
        res = fdb200.JacobianResult(np.zeros(m), x)
        cfg = fdb200.JacobianConfig(f, x, fdb200.Chunk(N))
        ctx.set_lane_sharing(share)
        try:
            l0 = ctx.launches()
            fdb200.jacobian_b(res, f, x, cfg, ctx=ctx)
            nl = ctx.launches() - l0
        finally:
            ctx.set_lane_sharing(True)
        assert cfg.last_path == ("program:specialised" if mode == "specialised" else "program:interpreted")
        assert nl == 4                                        # operand build, pack, contraction, elementwise program
        bb = b if mode == "specialised" else np.full(m, 0.25)
        J_ref, y_ref = oracle.jacobian("tanh_affine", x, m, N, A=A, b=bb)
        assert np.allclose(res.value, y_ref, rtol=1e-12, atol=1e-13)
        assert np.allclose(res.derivs[0], J_ref, rtol=1e-12, atol=1e-15), (m, n, N, share)
        if mode == "specialised":
            cat = F.tanh_affine.with_params(A=A, b=b)
            Jc = fdb200.jacobian(cat, x, fdb200.JacobianConfig(cat, x, fdb200.Chunk(N)), ctx=ctx, m=m)
            assert np.allclose(res.derivs[0], Jc, rtol=1e-13, atol=1e-15)
    # chunk ranges tile
    m, n = 256, 128
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); x = rng.uniform(-1, 1, n)
    f = lambda xd: fdb200.sin(A @ xd) * 3.0
    cfg = fdb200.JacobianConfig(f, x, fdb200.Chunk(N))
    full = fdb200.jacobian(f, x, cfg, ctx=ctx)
    assert np.allclose(full, 3.0 * np.cos(A @ x)[:, None] * A, rtol=1e-12, atol=1e-14)
    out = np.zeros((m, n), order="F")
    for r in range(3):
        p = fdb200.chunk_plan(n, N, r, 3)
        part = np.zeros((m, n), order="F")
        fdb200.jacobian_b(part, f, x, cfg, ctx=ctx, chunks=(p["chunk_lo"], p["chunk_hi"]))
        out[:, p["elem_lo"]:p["elem_hi"]] = part[:, p["elem_lo"]:p["elem_hi"]]
    assert np.array_equal(out, full)


@pytest.mark.gpu
def test_matvec_on_materialised_duals_in_the_reference_loop(ctx):
    """gradient(x -> sum(tanh.(A*x)), x, cfg): no batched kernel, so the reference loop runs -- with `A @ xdual` as one DMMA
    contraction per chunk (B200DualArray.__rmatmul__ -> fdb_dual_matvec)."""
    rng = np.random.default_rng(6)
    m, n, N = 160, 96, 12
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); x = rng.uniform(-1, 1, n)
    f = lambda xd: fdb200.tanh(A @ xd).sum()
    cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(N))
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, f, x, cfg, ctx=ctx)
    assert cfg.last_path.startswith("loop") and "A @ x" in cfg.trace_error
    y = np.tanh(A @ x)
    assert np.allclose(res.derivs[0], A.T @ (1.0 - y * y), rtol=1e-11, atol=1e-13)
    assert res.value == pytest.approx(float(y.sum()), rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 5, 12])
def test_prod_reductions_run_on_the_batched_sweep(ctx, oracle, N):
    """prod(x): an accumulator that starts at one(Dual) and combines by Dual*Dual; lane sharing still applies (far factors have
    identical partial lanes).  Specialised kernels only; without them f takes the reference loop."""
    rng = np.random.default_rng(N)
    for n in (13, 300, 2000):
        x = rng.uniform(0.9, 1.1, n)
        f = lambda xd: xd.prod()
        cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(N))
        res = fdb200.GradientResult(x)
        fdb200.gradient_b(res, f, x, cfg, ctx=ctx)
        assert cfg.last_path == "program:specialised"
        g_ref, v_ref = oracle.gradient("prod", x, N)
        assert np.allclose(res.derivs[0], g_ref, rtol=1e-11) and res.value == pytest.approx(v_ref, rel=1e-12)
        ctx.set_lane_sharing(False)
        try:
            g2 = fdb200.gradient(f, x, cfg, ctx=ctx)
        finally:
            ctx.set_lane_sharing(True)
        assert np.allclose(g2, res.derivs[0], rtol=1e-12)
    # README.md:26-45 shape: a sum over one view plus a product over another, and arithmetic after the reductions
    x = np.array([np.pi / 4, 2.0, 3.0, 4.0, 0.5, 1.5, 2.5])
    f = lambda xd: fdb200.sin(xd[:1]).sum() + xd[1:].prod() * 2.0
    cfg = fdb200.GradientConfig(f, x, fdb200.Chunk(min(N, 7)))
    g = fdb200.gradient(f, x, cfg, ctx=ctx)
    assert cfg.last_path == "program:specialised"
    p = np.prod(x[1:])
    assert np.allclose(g, np.concatenate([[np.cos(x[0])], 2.0 * p / x[1:]]), rtol=1e-13)
    # interpreter: products are not available -> the reference loop, same numbers
    ctx.set_jit(False)
    try:
        g3 = fdb200.gradient(f, x, cfg, ctx=ctx)
        assert cfg.last_path.startswith("loop") and np.allclose(g3, g, rtol=1e-13)
    finally:
        ctx.set_jit(True)


def test_two_pass_array_programs_trace_and_compile_on_the_host():
    """reductions inside an array-valued f (softmax, centring): term1 | STAGE 1 | epilogue1 | STAGE 2 | term2 with the result register of
    term2 as y[i]; specialised kernel kind 4 (jacobian_sweep2_body) compiles for sm_100a without a GPU."""
    import ctypes as C
    from fdb200.api import _i32p
    lib = fdb200._lib.load()
    softmax = lambda x: fdb200.exp(x) / fdb200.exp(x).sum()
    centre = lambda x: (x - x.mean()) * 2.0
    for f in (softmax, centre):
        p = trace.trace_array(f, 60)
        assert p.two_pass and p.needs_jit and p.out_len == 60 and p.n_acc == 1
        log = C.create_string_buffer(1 << 16); nbytes = C.c_int64()
        rc = lib.fdb_jit_compile_check(4, _i32p(p.term), len(p.term), None, 0, p.consts.ctypes.data_as(C.POINTER(C.c_double)) if p.consts.size else None,
                                       p.consts.size, p.n_acc, p.halo, p.result_reg, len(p.params), 12, 0, 1, log, len(log), C.byref(nbytes))
        assert rc == 0 and nbytes.value > 1000, log.value.decode()
    with pytest.raises(trace.TraceError):                        # three stages
        trace.trace_array(lambda x: (x - x.mean()) / fdb200.sqrt(((x - x.mean()) ** 2).sum()), 60)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 7, 12])
def test_softmax_like_jacobians_run_as_two_pass_programs(ctx, N):
    """jacobian(x -> exp.(x) ./ sum(exp.(x)), x): dense, every output depends on every input through the reduced Dual."""
    rng = np.random.default_rng(N)
    for n in (13, 200, 1031):
        x = rng.uniform(-1, 1, n)
        softmax = lambda xd: fdb200.exp(xd) / fdb200.exp(xd).sum()
        cfg = fdb200.JacobianConfig(softmax, x, fdb200.Chunk(N))
        res = fdb200.JacobianResult(np.zeros(n), x)
        fdb200.jacobian_b(res, softmax, x, cfg, ctx=ctx)
        assert cfg.last_path == "program:specialised"
        s = np.exp(x) / np.exp(x).sum()
        assert np.allclose(res.value, s, rtol=1e-13)
        assert np.allclose(res.derivs[0], np.diag(s) - np.outer(s, s), rtol=1e-11, atol=1e-15)
        centre = lambda xd: (xd - xd.mean()) * 2.0
        J = fdb200.jacobian(centre, x, fdb200.JacobianConfig(centre, x, fdb200.Chunk(N)), ctx=ctx)
        assert np.allclose(J, 2.0 * (np.eye(n) - 1.0 / n), rtol=1e-12, atol=1e-15)
        # lane sharing off and chunk ranges give the same matrix
        ctx.set_lane_sharing(False)
        try:
            J2 = fdb200.jacobian(softmax, x, cfg, ctx=ctx)
        finally:
            ctx.set_lane_sharing(True)
        assert np.allclose(J2, res.derivs[0], rtol=1e-12, atol=1e-16)
        out = np.zeros((n, n), order="F")
        for r in range(2):
            p = fdb200.chunk_plan(n, N, r, 2)
            part = np.zeros((n, n), order="F")
            fdb200.jacobian_b(part, softmax, x, cfg, ctx=ctx, chunks=(p["chunk_lo"], p["chunk_hi"]))
            out[:, p["elem_lo"]:p["elem_hi"]] = part[:, p["elem_lo"]:p["elem_hi"]]
        assert np.array_equal(out, res.derivs[0])


def _many_live_values(xd):
    """fifteen broadcast results alive at once and > 96 instructions: beyond the interpreter (12 registers), fine for the specialised kernel"""
    terms = [fdb200.sin(float(k) * xd[1:]) * fdb200.cos(0.5 * k * xd[:-1]) for k in range(1, 16)]
    acc = terms[0]
    for t in terms[1:]:
        acc = acc * 0.5 + t
    for t in terms:
        acc = acc + t * t
    return acc.sum()


def test_programs_beyond_the_interpreter_limits_and_new_forms_mean_what_the_evaluator_says():
    """programs that only the specialised kernels accept (more registers / instructions, product accumulators, array-valued two-pass
    form) against the NumPy reference evaluator and closed forms, plus host compilation"""
    import ctypes as C
    from fdb200.api import _i32p
    import program_ref
    lib = fdb200._lib.load()
    rng = np.random.default_rng(3)
    x = rng.uniform(0.5, 1.5, 40)
    big = trace.trace_scalar(_many_live_values, 40)
    assert big.beyond_interpreter and big.needs_jit and len(big.term) > 96
    v, g = program_ref.evaluate(big, x)
    h = 1e-6
    fd = np.array([(_np_many(x + h * e) - _np_many(x - h * e)) / (2 * h) for e in np.eye(40)])
    assert v == pytest.approx(_np_many(x), rel=1e-13) and np.allclose(g, fd, rtol=1e-6, atol=1e-6)
    pr = trace.trace_scalar(lambda xd: xd[1:].prod() * 2.0 + (xd * xd).sum(), 40)
    v, g = program_ref.evaluate(pr, x)
    p = np.prod(x[1:])
    gref = 2.0 * x; gref[1:] += 2.0 * p / x[1:]
    assert v == pytest.approx(2.0 * p + np.sum(x * x), rel=1e-13) and np.allclose(g, gref, rtol=1e-12)
    sm = trace.trace_array(lambda xd: fdb200.exp(xd) / fdb200.exp(xd).sum(), 40)
    y, J = program_ref.evaluate(sm, x)
    s = np.exp(x) / np.exp(x).sum()
    assert np.allclose(y, s, rtol=1e-14) and np.allclose(J, np.diag(s) - np.outer(s, s), rtol=1e-12, atol=1e-16)
    for kind, prog in ((0, big), (0, pr), (4, sm)):
        log = C.create_string_buffer(1 << 16); nbytes = C.c_int64()
        rc = lib.fdb_jit_compile_check(kind, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                       prog.consts.ctypes.data_as(C.POINTER(C.c_double)) if prog.consts.size else None, prog.consts.size, prog.n_acc,
                                       prog.halo, prog.result_reg, len(prog.params), 12, 0, 1, log, len(log), C.byref(nbytes))
        assert rc == 0 and nbytes.value > 1000, log.value.decode()


def _np_many(x):
    a, b = x[1:], x[:-1]
    terms = [np.sin(k * a) * np.cos(0.5 * k * b) for k in range(1, 16)]
    acc = terms[0]
    for t in terms[1:]:
        acc = acc * 0.5 + t
    for t in terms:
        acc = acc + t * t
    return float(acc.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("N", [3, 12])
def test_programs_beyond_the_interpreter_limits_run_specialised(ctx, N):
    rng = np.random.default_rng(N)
    x = rng.uniform(0.5, 1.5, 500)
    cfg = fdb200.GradientConfig(_many_live_values, x, fdb200.Chunk(N))
    res = fdb200.GradientResult(x)
    fdb200.gradient_b(res, _many_live_values, x, cfg, ctx=ctx)
    assert cfg.last_path == "program:specialised"
    import program_ref
    v, g = program_ref.evaluate(trace.trace_scalar(_many_live_values, 500), x)
    assert res.value == pytest.approx(v, rel=1e-12) and np.allclose(res.derivs[0], g, rtol=1e-10, atol=1e-12)
    # without the specialiser such a program takes the reference loop, not the interpreter
    ctx.set_jit(False)
    try:
        g2 = fdb200.gradient(_many_live_values, x[:60], fdb200.GradientConfig(_many_live_values, x[:60], fdb200.Chunk(N)), ctx=ctx)
    finally:
        ctx.set_jit(True)
    v60, g60 = program_ref.evaluate(trace.trace_scalar(_many_live_values, 60), x[:60])
    assert np.allclose(g2, g60, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("halo,N", [(4, 12), (3, 12), (4, 11), (2, 1)])
def test_specialised_hessian_kernel_fits_static_shared_memory_at_the_largest_shapes(halo, N):
    """hessian_sweep_body keeps its staging in static shared memory (48 KB): the buffer grows with N and the stencil reach, so the
    largest supported shapes (N = 12, halo = PH_MAXHALO = 4) are compiled here -- NVRTC rejects a kernel that does not fit."""
    import ctypes as C
    from fdb200.api import _i32p
    lib = fdb200._lib.load()
    prog = trace.trace_scalar(lambda xd, h=halo: ((xd[h:] - xd[:-h]) ** 2 * xd[:-h]).sum(), 200)
    assert prog.halo == halo
    log = C.create_string_buffer(1 << 16); nbytes = C.c_int64()
    rc = lib.fdb_jit_compile_check(2, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                   prog.consts.ctypes.data_as(C.POINTER(C.c_double)) if prog.consts.size else None, prog.consts.size, prog.n_acc,
                                   prog.halo, prog.result_reg, len(prog.params), N, 0, 1, log, len(log), C.byref(nbytes))
    assert rc == 0 and nbytes.value > 1000, log.value.decode()
