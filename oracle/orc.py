"""ctypes loader for the CPU oracle (oracle/oracle.cpp) and, when built, the
reference's own benchmarks/cpp program (oracle/_ref/libfdref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(forwarddiff.jl_b200/) never imports this module.

All dual arrays are AoS float64 arrays of shape (n, N+1): [:, 0] is the value,
[:, 1:] the partials (wire order of src/dual.jl:357-360).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_long)

# op codes shared with oracle.cpp (enum Op1 / Op2)
OP1 = dict(neg=1, exp=2, log=3, sin=4, cos=5, tanh=6, sqrt=7, abs=8, abs2=9, inv=10,
           pow0=11, pow1=12, pow2=13, pow3=14, tan=15, exp2=16, log2=17, log10=18, log1p=19, expm1=20, sinh=21, cosh=22,
           asin=23, acos=24, atan=25, cbrt=26)
OP2 = dict(add=1, sub=2, mul=3, div=4, pow=5, max=6, min=7, atan2=8, hypot=9, nanmax=10, nanmin=11)
FN = dict(rosenbrock=1, ackley=2, sumsq=3, prod=4)
JFN = dict(tanh_affine=1, brusselator=2, jactest=3)


def build(native=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    targets = ["all", "ref"] + (["native"] if native else [])
    subprocess.run(["make", "-C", _HERE] + targets, check=True, capture_output=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, nansafe=False, native=False):
        name = "liboracle_nansafe.so" if nansafe else ("liboracle_native.so" if native else "liboracle.so")
        path = os.path.join(_HERE, "lib", name)
        if not os.path.exists(path):
            build(native=native)
        self.lib = L = C.CDLL(path)
        L.orc_pickchunksize.restype = C.c_long
        L.orc_pickchunksize.argtypes = [C.c_long, C.c_long]
        L.orc_chunk_plan.argtypes = [C.c_long, C.c_long, _lp]
        L.orc_structural_indices.restype = C.c_long
        L.orc_structural_indices.argtypes = [C.c_int, C.c_long, _lp]
        L.orc_structural_length.restype = C.c_long
        L.orc_structural_length.argtypes = [C.c_int, C.c_long]
        L.orc_seed.argtypes = [_dp, _dp, C.c_long, _lp, C.c_int, C.c_long, C.c_long]
        L.orc_seed_zero.argtypes = [_dp, _dp, C.c_long, _lp, C.c_int, C.c_long, C.c_long]
        L.orc_extract_gradient_chunk.argtypes = [_dp, C.c_long, _lp, _dp, C.c_int, C.c_long, C.c_long]
        L.orc_extract_jacobian_chunk.argtypes = [_dp, C.c_long, _dp, C.c_long, C.c_int, C.c_long, C.c_long]
        L.orc_extract_jacobian.argtypes = [_dp, C.c_long, _dp, C.c_long, C.c_int, C.c_long]
        L.orc_extract_value.argtypes = [_dp, _dp, C.c_long, C.c_int]
        L.orc_map1.argtypes = [C.c_int, _dp, _dp, C.c_long, C.c_int]
        L.orc_map2.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_long, C.c_int]
        L.orc_fma3.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_long, C.c_int]
        L.orc_reduce.argtypes = [C.c_int, _dp, _dp, C.c_long, C.c_int]
        L.orc_dual_matmul.argtypes = [_dp, _dp, _dp, C.c_long, C.c_long, C.c_long, C.c_int]
        L.orc_eval_scalar_fn.argtypes = [C.c_int, _dp, _dp, C.c_long, C.c_int]
        L.orc_eval_scalar_fn_value.restype = C.c_double
        L.orc_eval_scalar_fn_value.argtypes = [C.c_int, _dp, C.c_long]
        L.orc_eval_array_fn.argtypes = [C.c_int, _dp, C.c_long, _dp, C.c_long, C.c_int, _dp, C.c_long, _dp, C.c_double]
        L.orc_gradient.argtypes = [C.c_int, _dp, C.c_long, C.c_int, _dp, _dp, C.c_long, C.c_long]
        L.orc_jacobian.argtypes = [C.c_int, C.c_int, _dp, C.c_long, C.c_long, C.c_int, _dp, C.c_long, _dp,
                                   C.c_double, _dp, C.c_long, _dp, C.c_long, C.c_long]
        L.orc_jacobian_mlp2.argtypes = [_dp, C.c_long, C.c_long, C.c_long, C.c_int, _dp, C.c_long, _dp, _dp, C.c_long, _dp,
                                        _dp, C.c_long, _dp, C.c_long, C.c_long]
        L.orc_hessian.argtypes = [C.c_int, _dp, C.c_long, C.c_int, _dp, C.c_long, _dp, _dp, C.c_long, C.c_long]
        L.orc_third_order.argtypes = [C.c_int, _dp, C.c_int, _dp]

    # -- bookkeeping -------------------------------------------------------
    def pickchunksize(self, n, threshold=12):
        return int(self.lib.orc_pickchunksize(n, threshold))

    def chunk_plan(self, n, N):
        out = (C.c_long * 5)()
        rc = self.lib.orc_chunk_plan(n, N, out)
        if rc != 0:
            raise ValueError("chunk size cannot be greater than structural_length(x)")
        return dict(remainder=out[0], lastchunksize=out[1], lastchunkindex=out[2], middle_hi=out[3], nchunks=out[4])

    def structural_indices(self, kind, rows):
        n = int(self.lib.orc_structural_length(kind, rows))
        out = np.zeros(max(n, 1), dtype=np.int64)
        k = self.lib.orc_structural_indices(kind, rows, out.ctypes.data_as(_lp))
        return out[:k].copy()

    # -- seeding / extraction (operate in place on AoS arrays) ---------------
    def seed(self, duals, x, index=0, chunksize=0, sidx=None):
        N = duals.shape[-1] - 1
        ns = len(sidx) if sidx is not None else x.size
        sp = sidx.ctypes.data_as(_lp) if sidx is not None else None
        assert self.lib.orc_seed(_p(duals), _p(x), ns, sp, N, index, chunksize) == 0

    def seed_zero(self, duals, x, index=0, count=0, sidx=None):
        N = duals.shape[-1] - 1
        ns = len(sidx) if sidx is not None else x.size
        sp = sidx.ctypes.data_as(_lp) if sidx is not None else None
        assert self.lib.orc_seed_zero(_p(duals), _p(x), ns, sp, N, index, count) == 0

    def extract_gradient_chunk(self, result, ydual, index, chunksize, sidx=None):
        N = ydual.size - 1
        ns = len(sidx) if sidx is not None else result.size
        sp = sidx.ctypes.data_as(_lp) if sidx is not None else None
        assert self.lib.orc_extract_gradient_chunk(_p(result), ns, sp, _p(ydual), N, index, chunksize) == 0

    def extract_jacobian_chunk(self, result_colmajor, ld, ydual, index, chunksize):
        m, N = ydual.shape[0], ydual.shape[1] - 1
        assert self.lib.orc_extract_jacobian_chunk(_p(result_colmajor), ld, _p(ydual), m, N, index, chunksize) == 0

    # -- elementwise ---------------------------------------------------------
    def map1(self, op, a):
        a = _f64(a); out = np.empty_like(a)
        assert self.lib.orc_map1(OP1[op], _p(out), _p(a), a.shape[0], a.shape[1] - 1) == 0
        return out

    def map2(self, op, a, b):
        """a, b: AoS dual arrays or python floats (exactly one may be a float)."""
        if np.isscalar(b):
            a = _f64(a); out = np.empty_like(a)
            rc = self.lib.orc_map2(OP2[op], 1, _p(out), _p(a), None, float(b), a.shape[0], a.shape[1] - 1)
        elif np.isscalar(a):
            b = _f64(b); out = np.empty_like(b)
            rc = self.lib.orc_map2(OP2[op], 2, _p(out), _p(b), None, float(a), b.shape[0], b.shape[1] - 1)
        else:
            a = _f64(a); b = _f64(b); out = np.empty_like(a)
            rc = self.lib.orc_map2(OP2[op], 0, _p(out), _p(a), _p(b), 0.0, a.shape[0], a.shape[1] - 1)
        assert rc == 0
        return out

    def fma3(self, kind, a, b, c, s=0.0):
        a = _f64(a); out = np.empty_like(a)
        b = None if b is None else _f64(b); c = None if c is None else _f64(c)
        assert self.lib.orc_fma3(kind, _p(out), _p(a), _p(b), _p(c), float(s), a.shape[0], a.shape[1] - 1) == 0
        return out

    def reduce(self, op, a):
        a = _f64(a); out = np.empty(a.shape[1])
        assert self.lib.orc_reduce(dict(sum=1, prod=2)[op], _p(out), _p(a), a.shape[0], a.shape[1] - 1) == 0
        return out

    def dual_matmul(self, A, B, m, k, n):
        """A: (m*k, N+1) AoS duals of a column-major m x k matrix, B: (k*n, N+1) -> (m*n, N+1)"""
        A = _f64(A); B = _f64(B); out = np.empty((m * n, A.shape[1]))
        assert self.lib.orc_dual_matmul(_p(out), _p(A), _p(B), m, k, n, A.shape[1] - 1) == 0
        return out

    def eval_scalar_fn(self, fn, xdual):
        xdual = _f64(xdual); out = np.empty(xdual.shape[1])
        assert self.lib.orc_eval_scalar_fn(FN[fn], _p(out), _p(xdual), xdual.shape[0], xdual.shape[1] - 1) == 0
        return out

    def value(self, fn, x):
        x = _f64(x)
        return float(self.lib.orc_eval_scalar_fn_value(FN[fn], _p(x), x.size))

    def eval_array_fn(self, fn, xdual, m, A=None, b=None, alpha=0.0):
        xdual = _f64(xdual); N = xdual.shape[1] - 1
        y = np.zeros((m, N + 1))
        lda = 0 if A is None else A.shape[0]
        Af = None if A is None else np.asfortranarray(A, dtype=np.float64)
        assert self.lib.orc_eval_array_fn(JFN[fn], _p(y), m, _p(xdual), xdual.shape[0], N, _p(Af), lda,
                                          _p(None if b is None else _f64(b)), float(alpha)) == 0
        return y

    # -- drivers ---------------------------------------------------------------
    def gradient(self, fn, x, N, chunks=None):
        x = _f64(x); g = np.zeros_like(x); val = C.c_double(0.0)
        clo, chi = (0, -1) if chunks is None else chunks
        rc = self.lib.orc_gradient(FN[fn], _p(x), x.size, N, _p(g), C.byref(val), clo, chi)
        if rc == -3:
            raise ValueError("chunk size cannot be greater than structural_length(x)")
        assert rc == 0
        return g, val.value

    def jacobian(self, fn, x, m, N, A=None, b=None, alpha=0.0, inplace=False, y0=None, chunks=None, want_J=True,
                 cols=None):
        """cols=(c0, c1): return only Jacobian columns [c0, c1) (must cover the requested chunks) so that a
        chunk range of a Jacobian too large for host memory can be checked."""
        x = _f64(x); n = x.size
        y = np.zeros(m) if y0 is None else _f64(y0).copy()
        Af = None if A is None else np.asfortranarray(A, dtype=np.float64)
        lda = 0 if A is None else A.shape[0]
        clo, chi = (0, -1) if chunks is None else chunks
        if cols is not None:
            J = np.zeros((m, cols[1] - cols[0]), order="F")
            Jp = C.cast(C.c_void_p(J.ctypes.data - cols[0] * m * 8), _dp)   # column c lands at J[:, c - c0]
        else:
            J = np.zeros((m, n), order="F") if want_J else None
            Jp = _p(J)
        rc = self.lib.orc_jacobian(JFN[fn], int(inplace), _p(x), n, m, N, _p(Af), lda,
                                   _p(None if b is None else _f64(b)), float(alpha),
                                   Jp, m, _p(y), clo, chi)
        if rc == -3:
            raise ValueError("chunk size cannot be greater than structural_length(x)")
        assert rc == 0
        return J, y

    def jacobian_mlp2(self, x, N, A1, b1, A2, b2, chunks=None):
        x = _f64(x); n = x.size; h, m = A1.shape[0], A2.shape[0]
        A1f = np.asfortranarray(A1, dtype=np.float64); A2f = np.asfortranarray(A2, dtype=np.float64)
        J = np.zeros((m, n), order="F"); y = np.zeros(m)
        clo, chi = (0, -1) if chunks is None else chunks
        rc = self.lib.orc_jacobian_mlp2(_p(x), n, h, m, N, _p(A1f), h, _p(_f64(b1)), _p(A2f), m, _p(_f64(b2)), _p(J), m, _p(y), clo, chi)
        assert rc == 0, rc
        return J, y

    def hessian(self, fn, x, N, chunks=None, want_H=True):
        x = _f64(x); n = x.size
        H = np.zeros((n, n), order="F") if want_H else None
        g = np.zeros(n); val = C.c_double(0.0)
        olo, ohi = (0, -1) if chunks is None else chunks
        rc = self.lib.orc_hessian(FN[fn], _p(x), n, N, _p(H), n, _p(g), C.byref(val), olo, ohi)
        if rc == -3:
            raise ValueError("chunk size cannot be greater than structural_length(x)")
        assert rc == 0, rc
        return H, g, val.value

    def third_order(self, fn, x):
        x = _f64(x); n = x.size
        T3 = np.zeros((n * n, n), order="F")
        assert self.lib.orc_third_order(FN[fn], _p(x), n, _p(T3)) == 0
        return T3


class RefBench:
    """The reference's own benchmarks/cpp program (oracle/_ref/libfdref.so)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libfdref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.fdref_gradient.argtypes = [C.c_int, C.c_int, _dp, C.c_long, _dp]
        L.fdref_value.restype = C.c_double
        L.fdref_value.argtypes = [C.c_int, _dp, C.c_long]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libfdref.so"))

    def testgrad(self):
        return self.lib.fdref_testgrad()

    def gradient(self, fn, x, K):
        x = _f64(x); out = np.zeros_like(x)
        rc = self.lib.fdref_gradient(FN[fn], K, _p(x), x.size, _p(out))
        assert rc == 0
        return out

    def value(self, fn, x):
        x = _f64(x)
        return float(self.lib.fdref_value(FN[fn], _p(x), x.size))
