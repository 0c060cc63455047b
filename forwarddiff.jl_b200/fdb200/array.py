"""Device arrays of the engine: the Python twin of the Julia shim's
`B200Array{Float64}` / `B200DualArray{T,Float64,N}` (julia/ForwardDiffB200.jl).

`similar(x, Dual{T,V,N})` in the reference's config constructors
(src/config.jl:122,159,185-186,225-226) decides the work-buffer type; here that is
`B200Array.similar_dual(N)`.  Arithmetic on a B200DualArray dispatches to the
hand-written kernels through the C ABI exactly as Julia's broadcast / sum / prod on
`Vector{Dual}` dispatch to src/dual.jl.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import OP1, OP2, REDOP, DimensionMismatch, FdbError, check

_dp = C.POINTER(C.c_double)


class Context:
    """One per process and GPU (fdb_init).  No GPU -> raises: there is no CPU path."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.fdb_init(device, C.byref(h))
        if rc != 0:
            msg = self.lib.fdb_last_error(None)
            raise _lib.CudaError((msg.decode() if msg else "fdb_init failed") + f" (status {rc})")
        self.h = h
        self.device = device
        self._capturing = False
        self._graph_handles = []

    def capture(self):
        """Context manager: capture the enclosed C-ABI calls into a CUDA graph (fdb_graph_begin / fdb_graph_end).
        Device arrays created or dropped inside stay alive until the returned Graph is destroyed."""
        return _Capture(self)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fdb_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        check(rc, self.h)

    def sync(self):
        self.check(self.lib.fdb_sync(self.h))

    def set_async(self, flag):
        self.check(self.lib.fdb_set_async(self.h, int(flag)))

    def set_nansafe(self, flag):
        self.check(self.lib.fdb_set_nansafe(self.h, int(flag)))

    def set_lane_sharing(self, flag):
        self.check(self.lib.fdb_set_lane_sharing(self.h, int(flag)))

    def set_chunk_grouping(self, group):
        """Opt-in: the catalogue gradient kernel sweeps `group` chunks per block and shares the far terms between them (1 = off)."""
        self.check(self.lib.fdb_set_chunk_grouping(self.h, int(group)))

    def set_jit(self, flag):
        """Run-time specialisation of the op-program kernels (fdb_jit.cu); False = per-thread interpreter."""
        self._jit = bool(flag)
        self.check(self.lib.fdb_set_jit(self.h, int(flag)))

    def jit_enabled(self):
        return getattr(self, "_jit", True)

    def data_array(self, host):
        """Device copy of a NumPy data array used inside f on the materialised path.  Cached by content, so the chunk
        loop uploads it once and the captured iteration of the graph-replayed loop (where f runs again and temporaries
        like `1.0 + w` are new objects) finds the copy made by the eager first chunk instead of uploading mid-capture."""
        cache = self.__dict__.setdefault("_data_cache", {})
        flat = np.ascontiguousarray(host, dtype=np.float64).ravel()
        key = (flat.size, hash(flat.tobytes()))
        hit = cache.get(key)
        if hit is None:
            if self._capturing:
                raise FdbError("a data array was first seen while the chunk loop was being captured")
            if len(cache) > 64:
                cache.clear()
            hit = cache[key] = self.array(flat)
        return hit

    def matrix_array(self, A):
        """Column-major device copy of a closed-over Float64 matrix (the `A` of `A @ x`).  Cached per matrix OBJECT together
        with a cheap fingerprint (shape, data pointer, a strided sample of 4096 entries), so that repeated jacobian calls do
        not upload A again; a matrix modified in place beyond what the sample sees needs `forget_matrices()`."""
        cache = self.__dict__.setdefault("_matrix_cache", {})
        A = np.asarray(A, dtype=np.float64)
        flat = A.ravel(order="K")[:: max(1, A.size // 4096)]          # memory order: a view for C- and F-contiguous matrices (no 512 MB copy per call)
        key = id(A)
        fp = (A.shape, A.__array_interface__["data"][0], float(np.sum(flat)), float(flat[-1]) if flat.size else 0.0)
        hit = cache.get(key)
        if hit is None or hit[0] != fp:
            if self._capturing:
                raise FdbError("a matrix was first seen while the chunk loop was being captured")
            if len(cache) > 8:
                cache.clear()
            hit = cache[key] = (fp, self.array(np.asfortranarray(A).ravel(order="F")), A)      # keep A alive: id() stays unique
        return hit[1]

    def forget_matrices(self):
        self.__dict__.pop("_matrix_cache", None)

    def jit_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self.lib.fdb_jit_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        log = self.lib.fdb_jit_last_log(self.h)
        return dict(compiled=a.value, cache_hits=b.value, failures=c.value, last_log=log.decode() if log else "")

    def set_stream(self, cuda_stream_ptr):
        """Run this context's kernels and copies on a caller-owned CUDA stream.  A handle of 0 -- what torch reports for
        its default stream -- is the LEGACY DEFAULT STREAM and is passed as cudaStreamLegacy (0x1): the C ABI reserves
        NULL for "go back to the context's own non-blocking stream" (use_own_stream()), which does not synchronise with
        the default stream at all."""
        ptr = int(cuda_stream_ptr or 0)
        self.check(self.lib.fdb_set_stream(self.h, C.c_void_p(ptr if ptr != 0 else 1)))

    def use_own_stream(self):
        self.check(self.lib.fdb_set_stream(self.h, None))

    def stream(self):
        s = C.c_void_p()
        self.check(self.lib.fdb_get_stream(self.h, C.byref(s)))
        return s.value or 0

    def launches(self):
        n = C.c_int64()
        self.check(self.lib.fdb_kernel_launches(self.h, C.byref(n)))
        return n.value

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self.check(self.lib.fdb_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- array factories ------------------------------------------------------
    def array(self, host):
        """Upload a Float64 vector (Vector{Float64} -> B200Array)."""
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        a = B200Array.alloc(self, host.size)
        if host.size:
            self.check(self.lib.fdb_upload_values(a.h, host.ctypes.data_as(_dp)))
        return a

    def zeros(self, n):
        return self.array(np.zeros(n))

    def duals(self, host_aos, level=1):
        """Upload an AoS dual array of shape (n, N+1) (level 1) or (n, (N+1)**2) (level 2)."""
        host_aos = np.ascontiguousarray(host_aos, dtype=np.float64)
        n, P = host_aos.shape
        N = P - 1 if level == 1 else int(round(P ** 0.5)) - 1
        d = B200DualArray.alloc(self, n, N, level)
        if n:
            self.check(self.lib.fdb_upload_duals_aos(d.h, host_aos.ctypes.data_as(_dp)))
        return d

    def wrap(self, device_ptr, n, ld=None, N=0, level=0, keepalive=None):
        """Borrow caller-owned device memory (e.g. a torch tensor's data_ptr())."""
        h = C.c_void_p()
        self.check(self.lib.fdb_array_wrap(self.h, C.c_void_p(device_ptr), n, ld if ld is not None else n, N, level, C.byref(h)))
        cls = B200Array if level == 0 else B200DualArray
        a = cls.__new__(cls)
        a._init(self, h, n, N, level, ld if ld is not None else n, device_ptr, keepalive)
        return a


class Graph:
    """An instantiated CUDA graph of one chunk-loop iteration plus the device arrays it references."""

    def __init__(self, ctx, exec_handle, handles):
        self.ctx, self.exec, self.handles = ctx, exec_handle, handles

    def launch(self, times=1):
        self.ctx.check(self.ctx.lib.fdb_graph_launch(self.ctx.h, self.exec, int(times)))

    def destroy(self):
        if self.exec:
            self.ctx.sync()
            self.ctx.lib.fdb_graph_destroy(self.exec)
            self.exec = None
            for h in self.handles:
                self.ctx.lib.fdb_array_free(h)
            self.handles = []

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class _Capture:
    def __init__(self, ctx):
        self.ctx, self.graph = ctx, None

    def __enter__(self):
        c = self.ctx
        c.check(c.lib.fdb_graph_begin(c.h))
        c._capturing, c._graph_handles = True, []
        return self

    def __exit__(self, et, ev, tb):
        c = self.ctx
        ex = C.c_void_p()
        rc = c.lib.fdb_graph_end(c.h, C.byref(ex))
        c._capturing = False
        handles, c._graph_handles = c._graph_handles, []
        if et is None:
            c.check(rc)
            self.graph = Graph(c, ex, handles)
        else:
            for h in handles:
                c.lib.fdb_array_free(h)
        return False


class _DeviceArray:
    def _init(self, ctx, h, n, N, level, ld, ptr, keepalive=None):
        self.ctx, self.h, self.n, self.N, self.level, self.ld, self.ptr = ctx, h, n, N, level, ld, ptr
        self._keepalive = keepalive

    @classmethod
    def _alloc(cls, ctx, n, N, level):
        h = C.c_void_p()
        ctx.check(ctx.lib.fdb_array_alloc(ctx.h, n, N, level, C.byref(h)))
        ld = C.c_int64(); ptr = C.c_void_p()
        ctx.check(ctx.lib.fdb_array_info(h, None, None, None, C.byref(ld), C.byref(ptr)))
        a = cls.__new__(cls)
        a._init(ctx, h, n, N, level, ld.value, ptr.value or 0)
        return a

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                if self.ctx._capturing:          # freeing synchronises the stream: defer to Graph.destroy()
                    self.ctx._graph_handles.append(self.h)
                else:
                    self.ctx.lib.fdb_array_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.n

    @property
    def planes(self):
        return 1 if self.level == 0 else (self.N + 1 if self.level == 1 else (self.N + 1) ** 2)

    def view(self, start, stop):
        """Contiguous element range [start, stop) sharing storage (x[a:b] as a view)."""
        if not (0 <= start <= stop <= self.n):
            raise IndexError("view out of range")
        return self.ctx.wrap(self.ptr + 8 * start, stop - start, self.ld, self.N, self.level, keepalive=self)

    def __getitem__(self, sl):
        if isinstance(sl, slice):
            start, stop, step = sl.indices(self.n)
            if step != 1:
                raise FdbError("only unit-stride views are supported (no scalar indexing on device arrays)")
            return self.view(start, max(stop, start))
        raise FdbError("scalar indexing of a device array is not supported")


class B200Array(_DeviceArray):
    """Vector{Float64} on the device (level 0)."""

    @classmethod
    def alloc(cls, ctx, n):
        return cls._alloc(ctx, n, 0, 0)

    def numpy(self):
        out = np.empty(self.n)
        if self.n:
            self.ctx.check(self.ctx.lib.fdb_download_values(self.h, out.ctypes.data_as(_dp)))
        return out

    def similar_dual(self, N, level=1):
        """similar(x, Dual{T,Float64,N})"""
        return B200DualArray.alloc(self.ctx, self.n, N, level)

    def similar(self, n=None):
        return B200Array.alloc(self.ctx, self.n if n is None else n)


class B200DualArray(_DeviceArray):
    """Vector{Dual{T,Float64,N}} (level 1) or Vector{Dual{T,Dual{T,Float64,N},N}} (level 2), SoA planes."""
    __array_ufunc__ = None          # `ndarray * xdual` defers to xdual.__rmul__ (Real (x) Dual broadcast)

    @classmethod
    def alloc(cls, ctx, n, N, level=1):
        return cls._alloc(ctx, n, N, level)

    def numpy(self):
        """AoS host copy, shape (n, planes): [:,0] value, [:,1:] partials (src/dual.jl:357-360)."""
        out = np.empty((self.n, self.planes))
        if self.n:
            self.ctx.check(self.ctx.lib.fdb_download_duals_aos(self.h, out.ctypes.data_as(_dp)))
        return out

    def similar(self, n=None):
        return B200DualArray.alloc(self.ctx, self.n if n is None else n, self.N, self.level)

    def similar_dual(self, N=None):
        """similar(xdual, Dual{T,Dual{T,V,N},N}) -- the HessianConfig inner buffer (src/config.jl:225-226)"""
        if self.level != 1:
            raise FdbError("only two dual levels are supported")
        return B200DualArray.alloc(self.ctx, self.n, self.N if N is None else N, 2)

    def values(self):
        """value.(xdual) as a device Vector{Float64}"""
        out = B200Array.alloc(self.ctx, self.n) if self.level == 1 else B200DualArray.alloc(self.ctx, self.n, self.N, 1)
        self.ctx.check(self.ctx.lib.fdb_extract_value(self.ctx.h, out.h, self.h))
        return out

    # -- elementwise arithmetic (src/dual.jl ops applied by broadcast) ---------
    def _out(self, other=None):
        n = self.n
        if isinstance(other, _DeviceArray) and other.n != n:
            if n == 1:
                n = other.n
            elif other.n != 1:
                raise DimensionMismatch(f"arrays could not be broadcast to a common size ({self.n} vs {other.n})")
        return B200DualArray.alloc(self.ctx, n, self.N, 1)

    def _map1(self, op):
        out = self._out()
        self.ctx.check(self.ctx.lib.fdb_map1(self.ctx.h, OP1[op], out.h, self.h))
        return out

    def _map2(self, op, other, reverse=False):
        lib, ctx = self.ctx.lib, self.ctx
        if isinstance(other, _DeviceArray):
            out = self._out(other)
            a, b = (other, self) if reverse else (self, other)
            ctx.check(lib.fdb_map2(ctx.h, OP2[op], out.h, a.h, b.h))
            return out
        if isinstance(other, (int, float, np.floating, np.integer)):
            out = self._out()
            ctx.check(lib.fdb_map2_scalar(ctx.h, OP2[op], out.h, self.h, float(other), int(reverse)))
            return out
        if isinstance(other, np.ndarray) and other.ndim == 1:        # closed-over data: Vector{Float64} .(op) Vector{Dual}
            return self._map2(op, ctx.data_array(other), reverse)
        return NotImplemented

    def __add__(self, o): return self._map2("add", o)
    def __radd__(self, o): return self._map2("add", o, True)
    def __sub__(self, o): return self._map2("sub", o)
    def __rsub__(self, o): return self._map2("sub", o, True)
    def __mul__(self, o): return self._map2("mul", o)
    def __rmul__(self, o): return self._map2("mul", o, True)
    def __truediv__(self, o): return self._map2("div", o)
    def __rtruediv__(self, o): return self._map2("div", o, True)
    def __neg__(self): return self._map1("neg")

    def __pow__(self, o):
        # Base.literal_pow for the literal integer exponents 0..3 (src/dual.jl:587-597)
        if isinstance(o, int) and not isinstance(o, bool) and 0 <= o <= 3:
            return self._map1(f"pow{o}")
        return self._map2("pow", o)

    def __rpow__(self, o): return self._map2("pow", o, True)

    def __rmatmul__(self, A):
        """A @ xdual for a Float64 matrix A (m x n): Julia's generic matvec over Dual (Real*Dual, Dual+Dual; src/dual.jl:476-490)
        as ONE contraction of A against the planes [value | partials] on the FP64 tensor cores (fdb_dual_matvec)."""
        A = np.asarray(A, dtype=np.float64)
        if A.ndim != 2 or A.shape[1] != self.n:
            raise DimensionMismatch(f"A has dimensions {A.shape}, vector has length {self.n}")
        if self.level != 1:
            raise FdbError("A @ x is implemented for Vector{Dual{T,Float64,N}}")
        m = A.shape[0]
        Ad = self.ctx.matrix_array(A)
        out = B200DualArray.alloc(self.ctx, m, self.N, 1)
        self.ctx.check(self.ctx.lib.fdb_dual_matvec(self.ctx.h, Ad.h, m, m, self.n, self.h, out.h))
        return out

    def sum(self):
        return self._reduce("sum")

    def prod(self):
        return self._reduce("prod")

    def mean(self):
        """Statistics.mean: sum(x) / length(x)"""
        return self.sum() / self.n

    def dot(self, other):
        """LinearAlgebra.dot of real vectors: sum(x .* y)"""
        return (self * other).sum()

    def _reduce(self, op):
        out = B200DualArray.alloc(self.ctx, 1, self.N, 1)
        self.ctx.check(self.ctx.lib.fdb_reduce(self.ctx.h, REDOP[op], out.h, self.h))
        return out


def _unary(op):
    def f(x):
        if isinstance(x, B200DualArray):
            return x._map1(op)
        if hasattr(x, "_trace_unary"):          # symbolic array while a user f is flattened into an op-program
            return x._trace_unary(op)
        return getattr(np, {"abs2": "square", "inv": "reciprocal"}.get(op, op))(x)
    f.__name__ = op
    return f


exp, log, sin, cos, tanh, sqrt, abs_, abs2, inv = (_unary(o) for o in
                                                  ("exp", "log", "sin", "cos", "tanh", "sqrt", "abs", "abs2", "inv"))
tan, exp2, log2, log10, log1p, expm1, sinh, cosh, asin, acos, atan, cbrt = (
    _unary(o) for o in ("tan", "exp2", "log2", "log10", "log1p", "expm1", "sinh", "cosh", "asin", "acos", "atan", "cbrt"))


def maximum(a, b):
    return a._map2("max", b)


def minimum(a, b):
    return a._map2("min", b)


class nanmath:
    """NaNMath.jl variants (src/dual.jl:477,549).  NaNMath.f returns NaN where Base.f throws a DomainError; device math never
    throws and already returns NaN there, so pow / log / sqrt / sin / cos / tan / asin / acos / log2 / log10 / log1p are the
    plain device ops.  max / min differ in value: a NaN operand is ignored."""
    pow = staticmethod(lambda x, y: x ** y if not isinstance(x, (int, float)) else y.__rpow__(x))
    log, sqrt, sin, cos, tan, asin, acos, log2, log10, log1p = log, sqrt, sin, cos, tan, asin, acos, log2, log10, log1p
    max = staticmethod(lambda a, b: a._map2("nanmax", b))
    min = staticmethod(lambda a, b: a._map2("nanmin", b))


def atan2(y, x):
    """atan(y, x) on two Dual arrays (DiffRules binary rule)."""
    return y._map2("atan2", x)


def hypot(x, y):
    return x._map2("hypot", y)


def dot(a, b):
    """dot(a, b) where at least one operand is a Dual array (or a symbolic array while f is being flattened)."""
    return a.dot(b) if hasattr(a, "dot") and not isinstance(a, np.ndarray) else b.dot(a)


def dual_matmul(A, shape_a, B, shape_b):
    """C = A * B for column-major matrices of Duals stored as flat B200DualArrays: shape_a = (m, k), shape_b = (k, n).
    Julia: the generic matrix product over Dual*Dual (src/dual.jl:260-300); here 2N+1 FP64 tensor-core contractions."""
    (m, k), (k2, n) = shape_a, shape_b
    if k != k2 or A.n != m * k or B.n != k * n:
        raise DimensionMismatch(f"A has dimensions {shape_a}, B has dimensions {shape_b}")
    out = B200DualArray.alloc(A.ctx, m * n, A.N, 1)
    A.ctx.check(A.ctx.lib.fdb_dual_matmul(A.ctx.h, A.h, m, m, k, B.h, k, n, out.h, m))
    return out


def fma(x, y, z):
    """Base.fma on Dual arrays (src/dual.jl:625-665): the all-dual, (x,y,Real) and (x,Real,z) bodies."""
    ctx = x.ctx
    out = B200DualArray.alloc(ctx, x.n, x.N, 1)
    if isinstance(y, B200DualArray) and isinstance(z, B200DualArray):
        ctx.check(ctx.lib.fdb_fma(ctx.h, 0, out.h, x.h, y.h, z.h, 0.0))
    elif isinstance(y, B200DualArray):
        ctx.check(ctx.lib.fdb_fma(ctx.h, 1, out.h, x.h, y.h, None, float(z)))
    else:
        ctx.check(ctx.lib.fdb_fma(ctx.h, 2, out.h, x.h, None, z.h, float(y)))
    return out
