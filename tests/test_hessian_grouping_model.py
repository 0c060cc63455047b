"""CPU model of the decomposition used by `hessian_sweep_body` (csrc/fdb_hessian_kernel.cuh) and `gradient_sweep_group_body`
(csrc/fdb_sweep_kernel.cuh): is evaluating

  * the terms far from every window of a block ONCE on shared lanes,
  * a term that reads only the inner window on `Dual{Dual{1},1}` per inner lane,
  * a term that reads only the outer window once with a zero inner partial,
  * and only the terms that read both windows on the full nested dual

really the same as what the reference does -- `f` on `Dual{Dual{N},N}` for every (inner chunk, outer chunk) pair
(src/hessian.jl:14-19 over src/gradient.jl:141-152 / src/jacobian.jl:199-210)?  The model below states both in plain Python
with IEEE doubles (NumPy scalars, so 0 * inf = nan as on the device) and compares every Hessian entry, including the
placement of NaNs when x holds an Inf.  It does not run the kernel (tests/test_gpu_*.py do); it pins the reasoning.
"""
import itertools

import numpy as np
import pytest

F = np.float64


class Dual:
    """value + a list of partials; both may be floats or Duals (nesting).  No NaN-safe shortcuts: p * x is always formed."""
    __slots__ = ("v", "p")

    def __init__(self, v, p):
        self.v, self.p = v, list(p)

    @staticmethod
    def _lift(o, like):
        if isinstance(o, Dual):
            return o
        return Dual(_const(o, like.v), [_zero(q) for q in like.p])

    def __add__(self, o):
        o = Dual._lift(o, self)
        return Dual(self.v + o.v, [a + b for a, b in zip(self.p, o.p)])
    __radd__ = __add__

    def __sub__(self, o):
        o = Dual._lift(o, self)
        return Dual(self.v - o.v, [a - b for a, b in zip(self.p, o.p)])

    def __rsub__(self, o):
        return Dual._lift(o, self) - self

    def __mul__(self, o):
        if not isinstance(o, Dual):                      # Real * Dual: p * x (src/dual.jl:471-472)
            return Dual(self.v * o, [a * o for a in self.p])
        # Dual * Dual: val = vx*vy; p = px*vy + py*vx (test/DualTest.jl:470), separate ops
        return Dual(self.v * o.v, [a * o.v + b * self.v for a, b in zip(self.p, o.p)])
    __rmul__ = __mul__


def _zero(x):
    return Dual(_zero(x.v), [_zero(q) for q in x.p]) if isinstance(x, Dual) else F(0.0)


def _const(c, like):
    return Dual(_const(c, like.v), [_zero(q) for q in like.p]) if isinstance(like, Dual) else F(c)


def term(load, i):
    """Rosenbrock term i: (1 - x_i)^2 + 100 (x_{i+1} - x_i^2)^2 written with + - * only."""
    a, b = load(i), load(i + 1)
    u = 1.0 - a
    w = b - a * a
    return u * u + (w * w) * 100.0


HALO = 1


def reference_pair(x, N, istart, ics, ostart, ocs):
    """The reference: f on Dual{Dual{N},N}.  Returns (grad rows [k], H block [k][j]) of the pair."""
    n = len(x)

    def load(pos):
        inner_v = Dual(F(x[pos]), [F(1.0) if (pos - ostart == j and j < ocs) else F(0.0) for j in range(N)])
        parts = []
        for k in range(N):
            hot = (pos - istart == k) and k < ics
            parts.append(Dual(F(1.0) if hot else F(0.0), [F(0.0)] * N))
        return Dual(inner_v, parts)

    acc = None
    for i in range(n - 1):
        t = term(load, i)
        acc = t if acc is None else acc + t
    g = [acc.p[k].v for k in range(N)]
    H = [[acc.p[k].p[j] for j in range(N)] for k in range(N)]
    return g, H


def grouped_block(x, N, G, cg0, nchunks, lastchunksize, co):
    """The decomposition of hessian_sweep_body for the block (inner chunks cg0.., outer chunk co).  Returns {ci: (g, H)}."""
    n = len(x); nt = n - 1
    start = lambda c: n - lastchunksize if c == nchunks - 1 else c * N
    size = lambda c: lastchunksize if c == nchunks - 1 else N
    gcount = min(G, nchunks - cg0)
    ostart, ocs = start(co), size(co)
    clamp = lambda r0, r1: (max(r0, 0), max(min(r1, nt), max(r0, 0)))
    br0, br1 = clamp(ostart - HALO, ostart + ocs)
    u0, u1 = clamp(start(cg0) - HALO, start(cg0 + gcount - 1) + size(cg0 + gcount - 1))
    arange = lambda g: clamp(start(cg0 + g) - HALO, start(cg0 + g) + size(cg0 + g))

    def far_load(pos):                                  # Dual{Dual{1},1}, all partials zero
        return Dual(Dual(F(x[pos]), [F(0.0)]), [Dual(F(0.0), [F(0.0)])])

    def lanes4(d):
        return np.array([d.v.v, d.v.p[0], d.p[0].v, d.p[0].p[0]])

    common = np.zeros(4)
    for i in range(nt):
        if not (u0 <= i < u1) and not (br0 <= i < br1):
            common = common + lanes4(term(far_load, i))
    c4 = {i: (lanes4(term(far_load, i)) if not (br0 <= i < br1) else np.zeros(4)) for i in range(u0, u1)}

    def b_only(i):                                      # outer window only: Dual{Dual{N},1}, inner partial zero(Dual{N})
        def load(pos):
            return Dual(Dual(F(x[pos]), [F(1.0) if (pos - ostart == j and j < ocs) else F(0.0) for j in range(N)]),
                        [Dual(F(0.0), [F(0.0)] * N)])
        y = term(load, i)
        return [y.p[0].v] + list(y.p[0].p)
    sb = {i: b_only(i) for i in range(br0, br1)}

    out = {}
    for g in range(gcount):
        ci = cg0 + g
        istart, ics = start(ci), size(ci)
        ar0, ar1 = arange(g)
        ff = common.copy()
        for i in range(u0, u1):
            if not (ar0 <= i < ar1):
                ff = ff + c4[i]
        grad = [F(0.0)] * N
        H = [[F(0.0)] * N for _ in range(N)]
        near = sorted(set(range(ar0, ar1)) | set(range(br0, br1)))
        for i in near:
            inA, inB = ar0 <= i < ar1, br0 <= i < br1
            for k in range(N):
                if inA and inB:                          # both windows: the full nested evaluation for inner lane k
                    def load(pos, k=k):
                        hot = (pos - istart == k) and k < ics
                        return Dual(Dual(F(x[pos]), [F(1.0) if (pos - ostart == j and j < ocs) else F(0.0) for j in range(N)]),
                                    [Dual(F(1.0) if hot else F(0.0), [F(0.0)] * N)])
                    y = term(load, i)
                    row = [y.p[0].v] + list(y.p[0].p)
                elif inA:                                # inner window only: one representative outer lane
                    def load(pos, k=k):
                        hot = (pos - istart == k) and k < ics
                        return Dual(Dual(F(x[pos]), [F(0.0)]), [Dual(F(1.0) if hot else F(0.0), [F(0.0)])])
                    y = term(load, i)
                    row = [y.p[0].v] + [y.p[0].p[0]] * N
                else:                                    # outer window only: the same for every inner lane k
                    row = sb[i]
                grad[k] = grad[k] + row[0]
                for j in range(N):
                    H[k][j] = H[k][j] + row[1 + j]
        for k in range(N):
            grad[k] = grad[k] + ff[2]
            for j in range(N):
                H[k][j] = H[k][j] + ff[3]
        out[ci] = (grad, H)
    return out


def _same(a, b):
    a, b = float(a), float(b)
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    return a == b or abs(a - b) <= 1e-12 * max(abs(a), abs(b))


@pytest.mark.parametrize("n,N,G", [(23, 4, 3), (17, 4, 8), (26, 5, 2)])
@pytest.mark.parametrize("poison", [None, "inf"])
def test_grouped_decomposition_equals_the_full_nested_evaluation(n, N, G, poison):
    rng = np.random.default_rng(n * 100 + N)
    x = rng.uniform(0.5, 1.5, n).astype(F)
    if poison:
        x[n // 2] = np.inf                                # derivatives of the terms around it are Inf / NaN: 0 * Inf must poison alike
    nchunks = -(-n // N)
    lastchunksize = n - (nchunks - 1) * N
    start = lambda c: n - lastchunksize if c == nchunks - 1 else c * N
    size = lambda c: lastchunksize if c == nchunks - 1 else N
    with np.errstate(invalid="ignore", over="ignore"):
        for co, cg0 in itertools.product(range(nchunks), range(0, nchunks, G)):
            blk = grouped_block(x, N, G, cg0, nchunks, lastchunksize, co)
            for ci, (g, H) in blk.items():
                g_ref, H_ref = reference_pair(x, N, start(ci), size(ci), start(co), size(co))
                for k in range(size(ci)):
                    assert _same(g[k], g_ref[k]), (ci, co, k, g[k], g_ref[k])
                    for j in range(size(co)):
                        assert _same(H[k][j], H_ref[k][j]), (ci, co, k, j, H[k][j], H_ref[k][j])


def test_far_terms_are_chunk_invariant():
    """gradient_sweep_group_body's premise: a term that reads no seeded position gives the same Dual{1} (value, zero lane) whichever
    chunk is being swept, so a group of chunks may share it -- including when that lane is NaN."""
    x = np.array([0.7, 1.2, np.inf, 0.9, 1.1, 0.6, 1.3], dtype=F)

    def lanes(i, start, N):
        def load(pos):
            return Dual(F(x[pos]), [F(1.0) if pos - start == j else F(0.0) for j in range(N)])
        with np.errstate(invalid="ignore", over="ignore"):
            return term(load, i)
    N = 2
    for i in range(len(x) - 1):
        seen = []
        for start in range(0, len(x), N):
            if start - HALO <= i < start + N:
                continue                                   # near this chunk
            y = lanes(i, start, N)
            assert all(_same(y.p[0], q) for q in y.p)      # all N lanes of a far term are one computation
            seen.append((y.v, y.p[0]))
        assert all(_same(a[0], seen[0][0]) and _same(a[1], seen[0][1]) for a in seen)
