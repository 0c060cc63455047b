"""The wider DiffRules op set: tan exp2 log2 log10 log1p expm1 sinh cosh asin acos
atan cbrt, binary atan(y,x) and hypot.  CPU: the oracle's formulas vs central finite differences (the
reference's own tolerance, test/utils.jl: 3e-5).  GPU: device kernels vs the oracle -- value <= 4 ulp
(two different libms), partials bit-exact against the same formula applied to the device primal."""
import numpy as np
import pytest

UNARY = {
    "tan": np.tan, "exp2": np.exp2, "log2": np.log2, "log10": np.log10, "log1p": np.log1p, "expm1": np.expm1,
    "sinh": np.sinh, "cosh": np.cosh, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "cbrt": np.cbrt,
}


def _inputs(n=64):
    rng = np.random.default_rng(11)
    v = rng.uniform(0.05, 0.9, n)
    d = np.stack([v, np.ones(n), -2.0 * np.ones(n), rng.uniform(-1, 1, n)], axis=1)
    return v, d


@pytest.mark.parametrize("op", sorted(UNARY))
def test_oracle_unary_rules_vs_finite_differences(oracle, op):
    v, d = _inputs()
    r = oracle.map1(op, d)
    fn = UNARY[op]
    assert np.allclose(r[:, 0], fn(v), rtol=1e-15, atol=1e-16)
    h = 1e-6
    fd = (fn(v + h) - fn(v - h)) / (2 * h)
    assert np.allclose(r[:, 1], fd, atol=3e-5)
    assert np.allclose(r[:, 2], -2.0 * r[:, 1], rtol=1e-15)


def test_oracle_binary_rules_vs_finite_differences(oracle):
    rng = np.random.default_rng(12)
    n = 40
    a = np.stack([rng.uniform(0.2, 1.0, n), np.ones(n), np.zeros(n)], axis=1)
    b = np.stack([rng.uniform(0.2, 1.0, n), np.zeros(n), np.ones(n)], axis=1)
    h = 1e-6
    for op, fn in (("atan2", np.arctan2), ("hypot", np.hypot)):
        r = oracle.map2(op, a, b)
        assert np.allclose(r[:, 0], fn(a[:, 0], b[:, 0]), rtol=1e-15)
        assert np.allclose(r[:, 1], (fn(a[:, 0] + h, b[:, 0]) - fn(a[:, 0] - h, b[:, 0])) / (2 * h), atol=3e-5)
        assert np.allclose(r[:, 2], (fn(a[:, 0], b[:, 0] + h) - fn(a[:, 0], b[:, 0] - h)) / (2 * h), atol=3e-5)


def _ulp(a, b):
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.spacing(np.maximum(np.abs(b), np.finfo(float).tiny))
    return float(np.max(np.where(same, 0.0, d)))


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 5, 12])
def test_device_unary_rules(ctx, oracle, N):
    rng = np.random.default_rng(N)
    n = 777
    a = rng.uniform(-1, 1, (n, N + 1)); a[:, 0] = rng.uniform(0.05, 0.9, n)
    ad = ctx.duals(a)
    log2c, log10c = 0.6931471805599453, 2.302585092994046
    for op in sorted(UNARY):
        got = ad._map1(op).numpy(); ref = oracle.map1(op, a)
        assert _ulp(got[:, 0], ref[:, 0]) <= 4, op
        v, val = a[:, 0], got[:, 0]
        with np.errstate(all="ignore"):
            deriv = {"tan": 1.0 + val * val, "exp2": val * log2c, "log2": (1.0 / v) / log2c, "log10": (1.0 / v) / log10c,
                     "log1p": 1.0 / (v + 1.0), "asin": 1.0 / np.sqrt(1.0 - v * v), "acos": -(1.0 / np.sqrt(1.0 - v * v)),
                     "atan": 1.0 / (1.0 + v * v), "cbrt": 1.0 / (3.0 * (val * val))}.get(op)
        if deriv is not None:
            assert np.array_equal(got[:, 1:], a[:, 1:] * deriv[:, None]), op     # same formula on the device primal: bit exact
        else:                                                                   # expm1/sinh/cosh: derivative is another libm call
            assert _ulp(got[:, 1:], ref[:, 1:]) <= 8, op


@pytest.mark.gpu
def test_device_binary_rules(ctx, oracle):
    rng = np.random.default_rng(5)
    n, N = 500, 6
    a = rng.uniform(-1, 1, (n, N + 1)); b = rng.uniform(-1, 1, (n, N + 1))
    a[:, 0] = rng.uniform(0.2, 1.0, n); b[:, 0] = rng.uniform(0.2, 1.0, n)
    ad, bd = ctx.duals(a), ctx.duals(b)
    import fdb200
    for name, fn in (("atan2", fdb200.atan2), ("hypot", fdb200.hypot)):
        got = fn(ad, bd).numpy(); ref = oracle.map2(name, a, b)
        assert _ulp(got[:, 0], ref[:, 0]) <= 4
        assert np.allclose(got[:, 1:], ref[:, 1:], rtol=1e-14, atol=1e-15)


# ---- NaNMath variants (src/dual.jl:477,549) --------------------------------------------------------------------------
def _nanmath_pow_cases():
    """test/DualTest.jl:514 and test/DerivativeTest.jl:98-102 as (base dual, exponent dual | float, kind, expected partial / predicate)"""
    nan = float("nan")
    return [
        # partials(NaNMath.pow(Dual(-2.0, 1.0), Dual(2.0, 0.0)), 1) == -4.0: a constant exponent never takes log of the negative base
        (np.array([[-2.0, 1.0]]), np.array([[2.0, 0.0]]), "dd", lambda p: p == -4.0),
        # derivative(x -> NaNMath.pow(NaN, x), 1.0) is NaN
        (nan, np.array([[1.0, 1.0]]), "rd", np.isnan),
        # derivative(x -> NaNMath.pow(x, NaN), 1.0) is NaN
        (np.array([[1.0, 1.0]]), nan, "dr", np.isnan),
        # derivative(x -> NaNMath.pow(1.0, x), 1.0) is not NaN
        (1.0, np.array([[1.0, 1.0]]), "rd", lambda p: not np.isnan(p)),
        # derivative(x -> NaNMath.pow(x, 0.5), -1.0) is NaN (Base.^ throws a DomainError here; the device cannot throw)
        (np.array([[-1.0, 1.0]]), 0.5, "dr", np.isnan),
    ]


def test_oracle_nanmath_pow_and_max_min(oracle):
    for a, b, kind, pred in _nanmath_pow_cases():
        r = oracle.map2("pow", a, b)
        assert pred(r[0, 1]), (a, b, kind, r)
    nan = float("nan")
    x = np.array([[1.0, 1.0, 0.0], [nan, 1.0, 0.0], [3.0, 1.0, 0.0], [nan, 1.0, 0.0]])
    y = np.array([[2.0, 0.0, 1.0], [2.0, 0.0, 1.0], [nan, 0.0, 1.0], [nan, 0.0, 1.0]])
    mx = oracle.map2("nanmax", x, y); mn = oracle.map2("nanmin", x, y)
    assert np.array_equal(mx[:3], [[2.0, 0.0, 1.0], [2.0, 0.0, 1.0], [3.0, 1.0, 0.0]]) and np.isnan(mx[3, 0])
    assert np.array_equal(mn[:3], [[1.0, 1.0, 0.0], [2.0, 0.0, 1.0], [3.0, 1.0, 0.0]]) and np.isnan(mn[3, 0])
    # Base.max propagates the NaN instead (value), which is what distinguishes the two
    assert oracle.map2("max", x, y)[0, 0] == 2.0


@pytest.mark.gpu
def test_device_nanmath_pow_and_max_min(ctx, oracle):
    import fdb200
    for a, b, kind, pred in _nanmath_pow_cases():
        if kind == "dd":
            r = fdb200.nanmath.pow(ctx.duals(a), ctx.duals(b)).numpy()
        elif kind == "dr":
            r = fdb200.nanmath.pow(ctx.duals(a), b).numpy()
        else:
            r = fdb200.nanmath.pow(a, ctx.duals(b)).numpy()
        assert pred(r[0, 1]), (a, b, kind, r)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (200, 4)); y = rng.uniform(-1, 1, (200, 4))
    x[::7, 0] = np.nan; y[::5, 0] = np.nan
    for op, fn in (("nanmax", fdb200.nanmath.max), ("nanmin", fdb200.nanmath.min)):
        got = fn(ctx.duals(x), ctx.duals(y)).numpy(); ref = oracle.map2(op, x, y)
        assert np.array_equal(got, ref, equal_nan=True), op
    # in a traced user function (op-program, both executors share the device functions)
    xs = rng.uniform(0.5, 1.5, 300)
    f = lambda xd: fdb200.nanmath.max(xd[1:], xd[:-1] ** 2).sum()
    g = fdb200.gradient(f, xs, fdb200.GradientConfig(f, xs, fdb200.Chunk(7)), ctx=ctx)
    take = xs[1:] > xs[:-1] ** 2
    ref = np.zeros(300); ref[1:] += take; ref[:-1] += (~take) * 2.0 * xs[:-1]
    assert np.allclose(g, ref, rtol=1e-14)
