"""Host-side mirror of ForwardDiff's API for the chunk-mode path.

Julia is not available in this image, so this module plays the role of the Julia
shim (julia/ForwardDiffB200.jl): same names, argument meaning and error behaviour
as the reference, every device action a call into the C ABI.

    reference                                   here
    ---------------------------------------     ------------------------------------------
    Chunk{N}(), Chunk(x)                        Chunk(N), Chunk.pick(x)     src/prelude.jl:8-38
    Tag(f, V), checktag                         Tag(f), checktag            src/config.jl:5-43
    GradientConfig(f, x, Chunk{N}(), tag)       GradientConfig(f, x, chunk, tag)   config.jl:117-124
    JacobianConfig(f[, y], x, chunk, tag)       JacobianConfig(f, x, chunk, tag, y=None)   :154-189
    HessianConfig(f[, result], x, chunk, tag)   HessianConfig(f, x, chunk, tag, result=None)  :221-254
    gradient / gradient!                        gradient / gradient_b       src/gradient.jl:16-44
    jacobian / jacobian!  (f and f! forms)      jacobian / jacobian_b       src/jacobian.jl:18-87
    hessian / hessian!                          hessian / hessian_b         src/hessian.jl:14-71
    DiffResults.GradientResult etc.             GradientResult, JacobianResult, HessianResult

`x` may be a host numpy vector (the call uploads, sweeps, downloads -- the
`*_host` entry points) or a device B200Array (results stay on the device).

Two execution paths, as in the Julia shim:
  * catalogue targets (fdb200.functions.*) -> one fused, batched launch for all chunks
    (fdb_gradient_chunked / fdb_jacobian_chunked / fdb_hessian_chunked);
  * any other callable written against B200DualArray arithmetic -> the reference's
    chunk loop statement by statement (src/gradient.jl:114-159, src/jacobian.jl:169-216)
    on materialised device duals, every step a C-ABI kernel (seed!, f, extract, unseed).
"""
import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._lib import ChunkPlan, ChunkSizeError, DimensionMismatch, FdbError, InvalidTagException
from .array import B200Array, B200DualArray, Context, _DeviceArray

_dp = C.POINTER(C.c_double)
DEFAULT_CHUNK_THRESHOLD = 12        # src/prelude.jl:2


# ---------------------------------------------------------------------------
# Chunk, Tag
# ---------------------------------------------------------------------------
def pickchunksize(input_length, threshold=DEFAULT_CHUNK_THRESHOLD):
    """src/prelude.jl:29-36 (evaluated by the library's host code)."""
    return int(_lib.load().fdb_pickchunksize(int(input_length), int(threshold)))


class Chunk:
    """Chunk{N} (src/prelude.jl:8-13)."""

    def __init__(self, N):
        self.N = int(N)

    @classmethod
    def pick(cls, x_or_len, threshold=DEFAULT_CHUNK_THRESHOLD):
        n = x_or_len if isinstance(x_or_len, int) else _length(x_or_len)
        return cls(pickchunksize(n, threshold))

    def __repr__(self):
        return f"Chunk{{{self.N}}}()"


_TAGCOUNT = itertools.count()


class Tag:
    """Tag{F,V} with the global ordering counter (src/config.jl:5-26)."""

    def __init__(self, f, V=np.float64):
        self.f, self.V = f, V
        self.count = next(_TAGCOUNT)

    def __repr__(self):
        return f"Tag({getattr(self.f, '__name__', self.f)!r}, {np.dtype(self.V).name})"


_DEFAULT = object()


def checktag(tag, f, x):
    """src/config.jl:34-43: a Tag built for another function is an error; custom tags pass."""
    if isinstance(tag, Tag):
        if isinstance(tag.f, tuple):           # Tag{<:Tuple}: Hessian tags are not checked
            return True
        if tag.f is not f:
            raise InvalidTagException(f"Invalid Tag object:\n  Expected Tag({f!r}),\n  Observed {tag!r}.")
    return True


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _length(x):
    if isinstance(x, _Structured):
        return x.structural_length()
    return x.n if isinstance(x, _DeviceArray) else int(np.asarray(x).size)


def _vec(x):
    """vec(x): an N-d input is differentiated in linear-index (column-major) order, src/apiutils.jl:44-47."""
    return x.ravel(order="F") if isinstance(x, np.ndarray) and x.ndim > 1 else x


def _ctx_of(x, ctx=None):
    if ctx is not None:
        return ctx
    return x.ctx if isinstance(x, _DeviceArray) else default_context()


def _catalogue_id(f, table):
    name = getattr(f, "fdb_catalogue", None)
    return table.get(name) if name else None


def chunk_plan(xlen, N, rank=0, world=1):
    """Reference bookkeeping (src/gradient.jl:121-125) + this rank's chunk range (fdb_plan_chunks)."""
    p = ChunkPlan()
    rc = _lib.load().fdb_plan_chunks(int(xlen), int(N), int(rank), int(world), C.byref(p))
    if rc == 6:
        raise ChunkSizeError(f"chunk size cannot be greater than ForwardDiff.structural_length(x) ({N} > {xlen})")
    if rc != 0:
        raise FdbError(f"fdb_plan_chunks failed ({rc})")
    return p.asdict()


# ---------------------------------------------------------------------------
# structured inputs (LinearAlgebra wrappers): only structurally non-zero entries are seeded
# (src/prelude.jl:15-20, src/apiutils.jl:44-71)
# ---------------------------------------------------------------------------
class _Structured:
    kind = "dense"

    def __init__(self, M):
        self.M = np.asfortranarray(np.array(M, dtype=np.float64))
        if self.M.ndim != 2 or self.M.shape[0] != self.M.shape[1]:
            raise DimensionMismatch("structured wrappers need a square matrix")
        self.rows = self.M.shape[0]
        self.M = self.M * self.mask(self.rows)            # off-structure entries read as zero, whatever the parent holds

    @property
    def structure(self):
        return _lib.STRUCTURE[self.kind]

    def structural_length(self):
        return int(_lib.load().fdb_structural_length(self.structure, self.rows))

    def storage(self):
        return self.M.ravel(order="F")

    def __eq__(self, other):
        return type(other) is type(self) and np.array_equal(self.M, other.M)


class UpperTriangular(_Structured):
    kind = "upper"
    mask = staticmethod(lambda n: np.triu(np.ones((n, n))))


class LowerTriangular(_Structured):
    kind = "lower"
    mask = staticmethod(lambda n: np.tril(np.ones((n, n))))


class Diagonal(_Structured):
    kind = "diagonal"
    mask = staticmethod(lambda n: np.eye(n))


# ---------------------------------------------------------------------------
# DiffResults
# ---------------------------------------------------------------------------
class DiffResult:
    """MutableDiffResult: value plus derivative arrays (DiffResults.jl)."""

    def __init__(self, value, *derivs):
        self.value = value
        self.derivs = list(derivs)


def GradientResult(x):
    """DiffResults.GradientResult(x): the gradient has the shape of x (a Matrix input gives a Matrix gradient)."""
    shape = np.shape(x) if isinstance(x, np.ndarray) and x.ndim > 1 else (_length(x),)
    return DiffResult(0.0, np.zeros(shape, order="F"))


def JacobianResult(y, x):
    return DiffResult(np.zeros(_length(y)), np.zeros((_length(y), _length(x)), order="F"))


def HessianResult(x):
    n = _length(x)
    return DiffResult(0.0, np.zeros(n), np.zeros((n, n), order="F"))


# ---------------------------------------------------------------------------
# Configs  (src/config.jl)
# ---------------------------------------------------------------------------
class _Config:
    def __init__(self, f, x, chunk, tag):
        self.chunk = chunk if chunk is not None else Chunk.pick(x)
        self.N = self.chunk.N
        self.tag = Tag(f) if tag is _DEFAULT else tag
        self.xlen = _length(x)
        self._duals = None           # allocated lazily: only the materialised path needs them
        # which execution path the last call through this config took: "catalogue" (fused batched kernel of a catalogue
        # target), "program:specialised" / "program:interpreted" (flattened f on the batched sweep), "loop:graph" /
        # "loop:eager" (the reference chunk loop on materialised duals); trace_error: why f was not flattened
        self.last_path = None
        self.trace_error = None
        self.capture_error = None

    @property
    def chunksize(self):
        return self.N


class GradientConfig(_Config):
    """GradientConfig(f, x, Chunk{N}(), tag): owns `duals = similar(x, Dual{T,V,N})` (config.jl:117-124)."""

    def __init__(self, f, x, chunk=None, tag=_DEFAULT, level=1):
        super().__init__(f, x, chunk, tag)
        self.level = level

    def duals(self, ctx):
        if self._duals is None or self._duals.ctx is not ctx:
            self._duals = B200DualArray.alloc(ctx, self.xlen, self.N, self.level)
        return self._duals


class JacobianConfig(_Config):
    """JacobianConfig(f, x, chunk, tag) / JacobianConfig(f!, y, x, chunk, tag) (config.jl:154-189)."""

    def __init__(self, f, x, chunk=None, tag=_DEFAULT, y=None):
        super().__init__(f, x, chunk, tag)
        self.ylen = None if y is None else _length(y)
        self._yduals = None

    def duals(self, ctx):
        if self._duals is None or self._duals.ctx is not ctx:
            self._duals = B200DualArray.alloc(ctx, self.xlen, self.N, 1)
        if self.ylen is not None and (self._yduals is None or self._yduals.ctx is not ctx):
            self._yduals = B200DualArray.alloc(ctx, self.ylen, self.N, 1)
        return (self._yduals, self._duals) if self.ylen is not None else self._duals


class HessianConfig(_Config):
    """HessianConfig(f, x, chunk, tag): a JacobianConfig plus a GradientConfig over its duals
    (Dual{T,Dual{T,V,N},N}, same tag on both levels; config.jl:221-254)."""

    def __init__(self, f, x, chunk=None, tag=_DEFAULT, result=None):
        super().__init__(f, x, chunk, tag)
        self.jacobian_config = JacobianConfig(f, x, self.chunk, self.tag)
        self.gradient_config = GradientConfig(f, x, self.chunk, self.tag, level=2)


# ---------------------------------------------------------------------------
# gradient
# ---------------------------------------------------------------------------
class SlowPathWarning(UserWarning):
    """f could not be flattened into an op-program: the reference chunk loop runs on materialised device duals, one small
    kernel per array statement per chunk (launch-bound; orders of magnitude slower than the batched sweep)."""


_trace_failure = {}


def _try_trace(kind, f, n):
    """Flatten a non-catalogue callable into an op-program; None when its shape is not supported (the reason is kept
    for `cfg.trace_error` / the SlowPathWarning)."""
    from . import trace
    try:
        _trace_failure.pop(id(f), None)
        return (trace.trace_scalar if kind == "scalar" else trace.trace_array)(f, n)
    except (trace.TraceError, TypeError, AttributeError, ValueError) as e:
        _trace_failure[id(f)] = f"{type(e).__name__}: {e}"
        return None


def _warn_loop(what, f, n, N, cfg):
    """The drop to the materialised loop is never silent: cfg.last_path / cfg.trace_error say what ran and why, and a
    SlowPathWarning is issued when the loop is long enough to matter."""
    import warnings
    cfg.trace_error = _trace_failure.get(id(f))
    nchunks = -(-n // N)
    if nchunks >= 64 and cfg.trace_error is not None:      # fuse=False is the caller's own choice: no warning
        warnings.warn(f"{what}: {getattr(f, '__name__', f)!r} runs on the materialised chunk loop ({nchunks} chunks x one kernel per "
                      f"array statement) because it could not be flattened ({cfg.trace_error or 'fuse=False'}); see cfg.last_path",
                      SlowPathWarning, stacklevel=4)


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a.size else None


class _bound_params:
    """Upload the data arrays a traced f closes over and bind them for the program call (fdb_program_bind_params)."""

    def __init__(self, ctx, prog):
        self.ctx, self.arrays = ctx, [ctx.array(a) for a in prog.params]

    def __enter__(self):
        if self.arrays:
            hs = (C.c_void_p * len(self.arrays))(*[a.h for a in self.arrays])
            self.ctx.check(self.ctx.lib.fdb_program_bind_params(self.ctx.h, len(self.arrays), hs))
        return self

    def __exit__(self, *exc):
        if self.arrays:
            self.ctx.lib.fdb_program_bind_params(self.ctx.h, 0, None)
        return False


def _usable(ctx, prog):
    """A program with data arrays or reductions of different lengths runs on the specialised kernels only; without them
    f takes the reference loop."""
    return prog is not None and (not prog.needs_jit or ctx.jit_enabled())


def gradient(f, x, cfg=None, check=True, ctx=None, chunks=None, fuse=True, graph=True):
    """ForwardDiff.gradient(f, x, cfg, Val{check}()) (src/gradient.jl:16-24)."""
    if isinstance(x, _Structured):
        return gradient_b(None, f, x, cfg, check, ctx, None, fuse)
    if isinstance(x, np.ndarray) and x.ndim > 1:
        result = np.zeros(x.shape, order="F")
    else:
        result = np.zeros(_length(x)) if not isinstance(x, _DeviceArray) else x.similar()
    return gradient_b(result, f, x, cfg, check, ctx, chunks, fuse, graph)


def gradient_b(result, f, x, cfg=None, check=True, ctx=None, chunks=None, fuse=True, graph=True):
    """ForwardDiff.gradient!(result, f, x, cfg, Val{check}()) (src/gradient.jl:35-44).
    `chunks=(lo, hi)` restricts the call to a 0-based chunk range (the sharding hook).
    A non-catalogue callable is flattened into an op-program and run as one batched launch when its
    shape allows (`fuse=True`); otherwise -- or with `fuse=False` -- the reference chunk loop runs on
    materialised device duals."""
    cfg = cfg if cfg is not None else GradientConfig(f, x)
    if check:
        checktag(cfg.tag, f, x)
    if isinstance(x, _Structured):
        return _structured_gradient(_ctx_of(x, ctx), result, f, x, cfg, fuse)
    if isinstance(x, np.ndarray) and x.ndim > 1:
        # an N-d input is seeded in linear-index (column-major) order (structural_eachindex, src/apiutils.jl:44-47);
        # f sees vec(x) and the gradient comes back in the shape of x
        out = result.derivs[0] if isinstance(result, DiffResult) else result
        if np.shape(out) != x.shape:
            raise DimensionMismatch(f"gradient result has shape {np.shape(out)}, x has {x.shape}")
        flat = DiffResult(0.0, np.zeros(x.size))
        gradient_b(flat, f, x.ravel(order="F"), cfg, check, ctx, chunks, fuse, graph)
        if chunks is None:
            out[...] = flat.derivs[0].reshape(x.shape, order="F")
        else:                                                       # only the entries of the requested chunks are written
            p = chunk_plan(x.size, cfg.N)
            e0, e1 = chunks[0] * cfg.N, (x.size if chunks[1] < 0 or chunks[1] >= p["nchunks"] else chunks[1] * cfg.N)
            view = out.reshape(-1, order="F") if out.flags.f_contiguous else None
            if view is None:
                raise FdbError("gradient!: an N-d result must be column-major when a chunk range is given")
            view[e0:e1] = flat.derivs[0][e0:e1]
        if isinstance(result, DiffResult):
            result.value = flat.value
        return result
    n = _length(x)
    out = result.derivs[0] if isinstance(result, DiffResult) else result
    if _length(out) != n:
        raise DimensionMismatch(f"gradient result has {_length(out)} entries, x has {n}")
    fid = _catalogue_id(f, _lib.FN)
    if n == 0:      # Chunk{0}: vector mode, f returns a plain Real, fill!(result, 0) (gradient.jl:65)
        if isinstance(result, DiffResult):
            result.value = float(f(np.zeros(0))) if callable(f) and fid is None else 0.0
        return result
    N = cfg.N
    if N > n:
        raise ChunkSizeError(f"chunk size cannot be greater than ForwardDiff.structural_length(x) ({N} > {n})")
    ctx = _ctx_of(x, ctx)
    lo, hi = (0, -1) if chunks is None else chunks
    if fid is not None:
        value = _fused_gradient(ctx, fid, x, N, lo, hi, out)
        cfg.last_path = "catalogue"
    else:
        prog = _try_trace("scalar", f, n) if fuse else None
        value, done = None, False
        if _usable(ctx, prog):
            try:
                c0 = ctx.jit_stats()
                value, done = _program_gradient(ctx, prog, x, N, lo, hi, out), True
                c1 = ctx.jit_stats()
                ran_jit = (c1["compiled"] + c1["cache_hits"]) > (c0["compiled"] + c0["cache_hits"])
                cfg.last_path = "program:specialised" if ran_jit else "program:interpreted"
            except _lib.UnsupportedOp:
                if not prog.needs_jit:                  # only "libnvrtc is not installed" is recoverable: take the loop
                    raise
        if not done:
            cfg.last_path = "loop:eager"
            _warn_loop("gradient", f, n, N, cfg)
            value = _loop_gradient(ctx, f, x, cfg, lo, hi, out, graph)
    if isinstance(result, DiffResult) and value is not None:
        result.value = value
    return result


def _structured_gradient(ctx, result, f, X, cfg, fuse=True):
    """gradient(f, x::Union{UpperTriangular,LowerTriangular,Diagonal}): the reference loop with structural seeding.
    `f` receives the rows*rows Dual storage (column-major) whose off-structure entries are zero(Dual), exactly what
    the Julia wrapper types return for them.  Sets cfg.fevals = number of evaluations of f (GradientTest.jl:257-274)."""
    lib = ctx.lib
    S, rows, n, N = X.structure, X.rows, X.structural_length(), cfg.N
    if N > n:
        raise ChunkSizeError(f"chunk size cannot be greater than ForwardDiff.structural_length(x) ({N} > {n})")
    xd = ctx.array(X.storage())
    res = ctx.zeros(rows * rows)
    # batched sweep: f flattened into an op-program, chunks over the structural positions, only those are seeded and extracted
    # (fdb_program_gradient_chunked_structured; specialised kernels)
    prog = _try_trace("scalar", f, rows * rows) if (fuse and ctx.jit_enabled()) else None
    if prog is not None and not prog.two_pass:
        val = B200Array.alloc(ctx, 1)
        try:
            with _bound_params(ctx, prog):
                ctx.check(lib.fdb_program_gradient_chunked_structured(
                    ctx.h, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                    prog.consts.ctypes.data_as(_dp) if prog.consts.size else None, prog.consts.size, prog.n_acc, prog.halo, prog.result_reg,
                    S, rows, xd.h, N, 0, -1, res.h, val.h))
            cfg.last_path = "program:specialised"
            cfg.fevals = 1 if n == N else chunk_plan(n, N)["nchunks"]          # one evaluation of f per chunk sweep
            G = type(X)(res.numpy().reshape((rows, rows), order="F"))
            if isinstance(result, DiffResult):
                result.value = float(val.numpy()[0]); result.derivs[0] = G
                return result
            return G
        except _lib.UnsupportedOp:
            pass
    cfg.last_path = "loop:eager"
    xdual = B200DualArray.alloc(ctx, rows * rows, N, 1)
    ctx.check(lib.fdb_fill(ctx.h, xdual.h, 0.0))
    cfg.fevals = 0

    def call_f():
        cfg.fevals += 1
        ydual = f(xdual)
        if not isinstance(ydual, B200DualArray) or ydual.n != 1:
            raise DimensionMismatch("gradient(f, x) expects that f(x) is a real number. Perhaps you meant jacobian(f, x)?")
        return ydual

    seed = lambda i, c: ctx.check(lib.fdb_seed_structured(ctx.h, xdual.h, xd.h, S, rows, i, c))
    unseed = lambda i, c: ctx.check(lib.fdb_seed_zero_structured(ctx.h, xdual.h, xd.h, S, rows, i, c))
    extract = lambda yd, i, c: ctx.check(lib.fdb_extract_gradient_chunk_structured(ctx.h, res.h, yd.h, S, rows, i, c))
    if n == N:                                                      # vector mode
        seed(1, N); ydual = call_f(); extract(ydual, 1, N)
    else:
        p = chunk_plan(n, N)
        seed(1, N); unseed(N + 1, n - N); ydual = call_f(); extract(ydual, 1, N); unseed(1, N)
        for c in range(2, p["middle_hi"] + 1):
            i = (c - 1) * N + 1
            seed(i, N); ydual = call_f(); extract(ydual, i, N); unseed(i, N)
        i, cs = p["lastchunkindex"], p["lastchunksize"]
        seed(i, cs); ydual = call_f(); extract(ydual, i, cs)
    G = type(X)(res.numpy().reshape((rows, rows), order="F"))
    if isinstance(result, DiffResult):
        result.value = float(ydual.numpy()[0, 0]); result.derivs[0] = G
        return result
    return G


def _program_hessian(ctx, result, H, f, x, N, lo, hi):
    """hessian(f, x, cfg) of a user f flattened into an op-program (fdb_program_hessian_chunked)."""
    n = _length(x)
    prog = _try_trace("scalar", f, n)
    if prog is None:
        raise _lib.UnsupportedOp("hessian: f must be a catalogue target or flatten to sums of elementwise terms (no prod, no strided views)")
    if prog.needs_jit and not ctx.jit_enabled():
        raise _lib.UnsupportedOp("hessian: a function with data arrays or reductions of different lengths needs the run-time specialised kernel")
    if getattr(prog, "has_prod", False):
        raise _lib.UnsupportedOp("hessian: prod has no nested-dual batched kernel")
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    Hd = B200Array.alloc(ctx, n * n); gd = B200Array.alloc(ctx, n); vd = B200Array.alloc(ctx, 1)
    with _bound_params(ctx, prog):
        ctx.check(ctx.lib.fdb_program_hessian_chunked(ctx.h, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                                      prog.consts.ctypes.data_as(_dp) if prog.consts.size else None, prog.consts.size,
                                                      prog.n_acc, prog.halo, prog.result_reg, xd.h, N, lo, hi, Hd.h, n, gd.h, vd.h))
    nchunks = 1 if n == N else chunk_plan(n, N)["nchunks"]
    last = hi < 0 or hi == nchunks
    c0, c1 = lo * N, (n if last else hi * N)
    H[:, c0:c1] = Hd.numpy().reshape((n, n), order="F")[:, c0:c1]
    if last and isinstance(result, DiffResult):
        result.value = float(vd.numpy()[0])
        result.derivs[0][...] = gd.numpy()
    return result


def _program_gradient(ctx, prog, x, N, lo, hi, out):
    """A flattened user f through the batched sweep (fdb_program_gradient_chunked)."""
    lib = ctx.lib
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    res = out if isinstance(out, _DeviceArray) else B200Array.alloc(ctx, xd.n)
    val = B200Array.alloc(ctx, 1)
    with _bound_params(ctx, prog):
        ctx.check(lib.fdb_program_gradient_chunked(ctx.h, _i32p(prog.term), len(prog.term), _i32p(prog.epi), len(prog.epi),
                                                   prog.consts.ctypes.data_as(_dp) if prog.consts.size else None, prog.consts.size,
                                                   prog.n_acc, prog.halo, prog.result_reg, xd.h, N, lo, hi, res.h, val.h))
    nchunks = 1 if xd.n == N else chunk_plan(xd.n, N)["nchunks"]
    lo_e, hi_e = lo * N, (xd.n if (hi < 0 or hi == nchunks) else hi * N)
    if res is not out:
        out[lo_e:hi_e] = res.numpy()[lo_e:hi_e]
    return float(val.numpy()[0]) if (hi < 0 or hi == nchunks) else None


def _fused_gradient(ctx, fid, x, N, lo, hi, out):
    lib = ctx.lib
    if isinstance(x, _DeviceArray):
        val = B200Array.alloc(ctx, 1)
        ctx.check(lib.fdb_gradient_chunked(ctx.h, fid, x.h, N, lo, hi, out.h, val.h))
        plan = chunk_plan(x.n, N) if x.n != N else dict(nchunks=1)
        return float(val.numpy()[0]) if hi < 0 or hi == plan["nchunks"] else None
    xh = np.ascontiguousarray(x, dtype=np.float64).ravel()
    if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous):
        raise FdbError("gradient!: result must be a contiguous float64 array")
    val = C.c_double(0.0)
    ctx.check(lib.fdb_gradient_host(ctx.h, fid, xh.ctypes.data_as(_dp), xh.size, N, lo, hi,
                                    out.ctypes.data_as(_dp), C.byref(val)))
    # the value is written exactly when the range includes the last chunk (gradient.jl:155) -- decided from the plan,
    # not from the number that came back (f(x) may legitimately be NaN)
    nchunks = 1 if xh.size == N else chunk_plan(xh.size, N)["nchunks"]
    return val.value if (hi < 0 or hi == nchunks) else None


def _loop_gradient(ctx, f, x, cfg, lo, hi, out, graph=True):
    """The reference chunk loop on materialised device duals (src/gradient.jl:114-159 / :97-108)."""
    lib = ctx.lib
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    n, N = xd.n, cfg.N
    res = out if isinstance(out, _DeviceArray) else B200Array.alloc(ctx, n)
    xdual = cfg.duals(ctx)

    def call_f():
        ydual = f(xdual)
        if not isinstance(ydual, B200DualArray) or ydual.n != 1:   # GRAD_ERROR, gradient.jl:88-91
            raise DimensionMismatch("gradient(f, x) expects that f(x) is a real number. Perhaps you meant jacobian(f, x)?")
        return ydual

    if n == N:                                                     # vector mode
        ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, 0, 0))
        ydual = call_f()
        ctx.check(lib.fdb_extract_gradient(ctx.h, res.h, ydual.h))
    else:
        p = chunk_plan(n, N)
        nchunks = p["nchunks"]
        hi = nchunks if hi < 0 else hi
        full = (lo == 0 and hi == nchunks)
        if not full:
            ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, 0, 0))
        ydual = None
        if graph and full and nchunks >= 4:
            # first chunk eagerly (gradient.jl:133-138; it also sizes the library's scratch buffers), then ONE
            # iteration  seed! -> f -> extract -> unseed -> advance  captured at the device-resident chunk cursor
            # and replayed for chunks 2..nchunks (the cursor's chunksize shrinks for the short last chunk).
            # Capture restrictions: f must only enqueue device work.  An f that reads device data back (`.numpy()`),
            # synchronises, or takes host-side decisions from device values invalidates the capture; the capture is
            # then abandoned and chunks 2.. run through the eager loop below, which handles any f.
            ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, 1, N))
            ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, N + 1, n - N))
            ydual = call_f()
            ctx.check(lib.fdb_extract_gradient_chunk(ctx.h, res.h, ydual.h, 1, N))
            ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, 1, N))
            ctx.check(lib.fdb_cursor_set(ctx.h, N + 1, min(N, n - N)))
            cap = None
            try:
                with ctx.capture() as cap:
                    ctx.check(lib.fdb_seed_at_cursor(ctx.h, xdual.h, xd.h))
                    ydual = call_f()
                    ctx.check(lib.fdb_extract_gradient_chunk_at_cursor(ctx.h, res.h, ydual.h))
                    ctx.check(lib.fdb_seed_zero_at_cursor(ctx.h, xdual.h, xd.h))
                    ctx.check(lib.fdb_cursor_advance(ctx.h, N, n))
            except DimensionMismatch:
                raise
            except (FdbError, RuntimeError, ValueError, TypeError) as e:
                cap = None
                cfg.capture_error = repr(e)
            if cap is not None and cap.graph is not None:
                cap.graph.launch(nchunks - 1)
                ctx.sync()
                value = float(ydual.numpy()[0, 0])
                if res is not out:
                    out[...] = res.numpy().reshape(np.shape(out))
                cap.graph.destroy()
                cfg.last_path = "loop:graph"
                return value
            lo, full = 1, False                                     # nothing of the abandoned capture ran: continue eagerly with chunk 2
        for c in range(lo, hi):
            c1 = c + 1
            if c1 == nchunks:                                       # final chunk, gradient.jl:150-152
                i, cs = p["lastchunkindex"], p["lastchunksize"]
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, i, cs))
                ydual = call_f()
                ctx.check(lib.fdb_extract_gradient_chunk(ctx.h, res.h, ydual.h, i, cs))
            elif c1 == 1 and full:                                  # first chunk, gradient.jl:133-138
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, 1, N))
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, N + 1, n - N))
                ydual = call_f()
                ctx.check(lib.fdb_extract_gradient_chunk(ctx.h, res.h, ydual.h, 1, N))
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, 1, N))
            else:                                                   # middle chunks, gradient.jl:141-147
                i = (c1 - 1) * N + 1
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, i, N))
                ydual = call_f()
                ctx.check(lib.fdb_extract_gradient_chunk(ctx.h, res.h, ydual.h, i, N))
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, i, N))
        if ydual is None:
            return None
    if res is not out:
        out[...] = res.numpy().reshape(np.shape(out))
    return float(ydual.numpy()[0, 0])                               # extract_value!, gradient.jl:155


# ---------------------------------------------------------------------------
# jacobian
# ---------------------------------------------------------------------------
def jacobian(f, x, cfg=None, check=True, ctx=None, y=None, params=None, chunks=None, m=None, fuse=True):
    """ForwardDiff.jacobian(f, x, cfg) and, with `y`, jacobian(f!, y, x, cfg) (src/jacobian.jl:18-44).
    Returns the dense column-major length(y) x length(x) matrix (jacobian.jl:221)."""
    x = _vec(x)
    n = _length(x)
    ylen = _length(y) if y is not None else (m if m is not None else getattr(f, "output_length", lambda n: None)(n))
    fp = getattr(f, "params", None) or {}
    if ylen is None and fp.get("A2") is not None:
        ylen = np.shape(fp["A2"])[0]
    elif ylen is None and fp.get("A") is not None:
        ylen = np.shape(fp["A"])[0]
    if ylen is None and fuse and _catalogue_id(f, _lib.JFN) is None and not fp:
        prog = _try_trace("array", f, n)            # the flattened f knows its output length
        ylen = prog.out_len if prog is not None else None
    if ylen is None:
        raise FdbError("jacobian: output length unknown; pass m= or y=")
    if isinstance(x, _DeviceArray):
        result = B200Array.alloc(x.ctx, ylen * n)
    else:
        result = np.zeros((ylen, n), order="F")
    return jacobian_b(result, f, x, cfg, check, ctx, y, params, chunks, fuse)


def jacobian_b(result, f, x, cfg=None, check=True, ctx=None, y=None, params=None, chunks=None, fuse=True):
    """ForwardDiff.jacobian!(result, f, x, cfg) / jacobian!(result, f!, y, x, cfg) (src/jacobian.jl:57-87)."""
    x = _vec(x)
    cfg = cfg if cfg is not None else JacobianConfig(f, x, y=y)
    if check:
        checktag(cfg.tag, f, x)
    n, N = _length(x), cfg.N
    if N > n:
        raise ChunkSizeError(f"chunk size cannot be greater than ForwardDiff.structural_length(x) ({N} > {n})")
    ctx = _ctx_of(x, ctx)
    lib = ctx.lib
    out = result.derivs[0] if isinstance(result, DiffResult) else result
    fid = _catalogue_id(f, _lib.JFN)
    lo, hi = (0, -1) if chunks is None else chunks
    if getattr(f, "fdb_catalogue", None) == "mlp2":
        cfg.last_path = "catalogue"
        return _mlp2_jacobian(ctx, result, out, f, x, N, lo, hi)
    if fid is None:
        prog = _try_trace("array", f, n) if (fuse and y is None) else None
        if _usable(ctx, prog):
            try:
                c0 = ctx.jit_stats()
                r = _program_jacobian(ctx, prog, result, out, x, N, lo, hi)
                c1 = ctx.jit_stats()
                cfg.last_path = "program:specialised" if (c1["compiled"] + c1["cache_hits"]) > (c0["compiled"] + c0["cache_hits"]) else "program:interpreted"
                return r
            except _lib.UnsupportedOp:
                if not prog.needs_jit:
                    raise
        cfg.last_path = "loop:eager"
        _warn_loop("jacobian", f, n, N, cfg)
        return _loop_jacobian(ctx, result, out, f, x, cfg, y, lo, hi)
    cfg.last_path = "catalogue"
    params = params or getattr(f, "params", {}) or {}
    host = not isinstance(x, _DeviceArray)
    xd = ctx.array(x) if host else x
    if isinstance(out, _DeviceArray):
        m = out.n // n
        Jd = out
    else:
        m = out.shape[0]
        Jd = B200Array.alloc(ctx, m * n)
    yd = B200Array.alloc(ctx, m)
    A, b, alpha = params.get("A"), params.get("b"), float(params.get("alpha", 0.0))
    Ad = None if A is None else (A if isinstance(A, _DeviceArray) else ctx.array(np.asfortranarray(A).ravel(order="F")))
    bd = None if b is None else (b if isinstance(b, _DeviceArray) else ctx.array(b))
    lda = m
    ctx.check(lib.fdb_jacobian_chunked(ctx.h, fid, xd.h, m, N, Ad.h if Ad else None, lda, bd.h if bd else None, alpha,
                                       lo, hi, Jd.h, m, yd.h))
    if Jd is not out:
        out[...] = Jd.numpy().reshape((m, n), order="F")
    last = hi < 0 or hi == (1 if n == N else chunk_plan(n, N)["nchunks"])
    if last:
        yv = yd.numpy()
        if y is not None and not isinstance(y, _DeviceArray):
            y[...] = yv                                              # map!(d -> value(T,d), y, ydual), jacobian.jl:229
        if isinstance(result, DiffResult):
            result.value = yv
    return result


def _mlp2_jacobian(ctx, result, out, f, x, N, lo, hi):
    """jacobian(x -> tanh.(A2*tanh.(A1*x .+ b1) .+ b2), x, cfg): fdb_jacobian_mlp2_chunked (dense partials on DMMA)."""
    lib, p = ctx.lib, f.params
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    n = xd.n
    h, m = np.shape(p["A1"])[0], np.shape(p["A2"])[0]
    dev = lambda M: M if isinstance(M, _DeviceArray) else ctx.array(np.asfortranarray(M).ravel(order="F"))
    A1, A2, b1, b2 = dev(p["A1"]), dev(p["A2"]), dev(p["b1"]), dev(p["b2"])
    Jd = out if isinstance(out, _DeviceArray) else B200Array.alloc(ctx, m * n)
    yd = B200Array.alloc(ctx, m)
    ctx.check(lib.fdb_jacobian_mlp2_chunked(ctx.h, xd.h, h, m, N, A1.h, h, b1.h, A2.h, m, b2.h, lo, hi, Jd.h, m, yd.h))
    nchunks = 1 if n == N else chunk_plan(n, N)["nchunks"]
    last = hi < 0 or hi == nchunks
    if Jd is not out:
        c0, c1 = lo * N, (n if last else hi * N)
        out[:, c0:c1] = Jd.numpy().reshape((m, n), order="F")[:, c0:c1]
    if last and isinstance(result, DiffResult):
        result.value = yd.numpy()
    return result


def _program_jacobian(ctx, prog, result, out, x, N, lo, hi):
    """A flattened elementwise / stencil f through the batched sweep (fdb_program_jacobian_chunked); with `A @ x` inside f,
    the contraction on the FP64 tensor cores followed by the elementwise program (fdb_program_matvec_jacobian_chunked)."""
    lib = ctx.lib
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    n, m = xd.n, prog.out_len
    if isinstance(out, _DeviceArray):
        if out.n != m * n:
            raise DimensionMismatch(f"Jacobian result has {out.n} entries, expected {m}x{n}")
        Jd = out
    else:
        if out.shape != (m, n):
            raise DimensionMismatch(f"Jacobian result has shape {out.shape}, expected {(m, n)}")
        Jd = B200Array.alloc(ctx, m * n)
    yd = B200Array.alloc(ctx, m)
    with _bound_params(ctx, prog):
        if prog.matvec is not None:
            Ad = ctx.matrix_array(prog.matvec)               # column-major device copy, cached per matrix object
            ctx.check(lib.fdb_program_matvec_jacobian_chunked(ctx.h, _i32p(prog.term), len(prog.term),
                                                              prog.consts.ctypes.data_as(_dp) if prog.consts.size else None, prog.consts.size,
                                                              prog.result_reg, Ad.h, m, m, xd.h, N, lo, hi, Jd.h, m, yd.h))
        else:
            ctx.check(lib.fdb_program_jacobian_chunked(ctx.h, _i32p(prog.term), len(prog.term),
                                                       prog.consts.ctypes.data_as(_dp) if prog.consts.size else None, prog.consts.size,
                                                       prog.halo, prog.result_reg, xd.h, N, lo, hi, Jd.h, m, yd.h))
    nchunks = 1 if n == N else chunk_plan(n, N)["nchunks"]
    last = hi < 0 or hi == nchunks
    if Jd is not out:
        c0, c1 = lo * N, (n if last else hi * N)
        out[:, c0:c1] = Jd.numpy().reshape((m, n), order="F")[:, c0:c1]
    if last and isinstance(result, DiffResult):
        result.value = yd.numpy()
    return result


def _loop_jacobian(ctx, result, out, f, x, cfg, y, lo, hi):
    """Reference chunk loop on materialised duals (src/jacobian.jl:169-216); f(xdual) -> ydual or f!(ydual, xdual)."""
    lib = ctx.lib
    xd = x if isinstance(x, _DeviceArray) else ctx.array(x)
    n, N = xd.n, cfg.N
    inplace = y is not None
    duals = cfg.duals(ctx)
    ydual_buf, xdual = duals if inplace else (None, duals)
    yd0 = (y if isinstance(y, _DeviceArray) else ctx.array(y)) if inplace else None

    def compute():
        if inplace:
            ctx.check(lib.fdb_seed_zero(ctx.h, ydual_buf.h, yd0.h, 0, 0))    # jacobian.jl:227
            f(ydual_buf, xdual)
            return ydual_buf
        yd = f(xdual)
        if not isinstance(yd, B200DualArray):                                   # JACOBIAN_ERROR
            raise DimensionMismatch("jacobian(f, x) expects that f(x) is an array. Perhaps you meant gradient(f, x)?")
        return yd

    state = {}

    def extract(ydual, index, cs):
        if "J" not in state:
            m = ydual.n
            state["m"] = m
            state["J"] = out if isinstance(out, _DeviceArray) else B200Array.alloc(ctx, m * n)
            if state["J"].n != m * n:
                raise DimensionMismatch(f"Jacobian result has {state['J'].n} entries, expected {m}x{n}")
        if index == 1 and cs == n:        # vector mode: extract_jacobian!, src/jacobian.jl:95-107 (all N partials of every row)
            ctx.check(lib.fdb_extract_jacobian(ctx.h, state["J"].h, state["m"], ydual.h, n))
        else:
            ctx.check(lib.fdb_extract_jacobian_chunk(ctx.h, state["J"].h, state["m"], ydual.h, index, cs))

    if n == N:
        ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, 0, 0))
        ydual = compute()
        extract(ydual, 1, N)
    else:
        p = chunk_plan(n, N)
        nchunks = p["nchunks"]
        hi = nchunks if hi < 0 else hi
        full = (lo == 0 and hi == nchunks)
        if not full:
            ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, 0, 0))
        ydual = None
        for c in range(lo, hi):
            c1 = c + 1
            if c1 == nchunks:
                i, cs = p["lastchunkindex"], p["lastchunksize"]
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, i, cs))
                ydual = compute(); extract(ydual, i, cs)
            elif c1 == 1 and full:
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, 1, N))
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, N + 1, n - N))
                ydual = compute(); extract(ydual, 1, N)
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, 1, N))
            else:
                i = (c1 - 1) * N + 1
                ctx.check(lib.fdb_seed(ctx.h, xdual.h, xd.h, i, N))
                ydual = compute(); extract(ydual, i, N)
                ctx.check(lib.fdb_seed_zero(ctx.h, xdual.h, xd.h, i, N))
    J, m = state["J"], state["m"]
    if J is not out:
        if out.shape != (m, n):
            raise DimensionMismatch(f"Jacobian result has shape {out.shape}, expected {(m, n)}")
        out[...] = J.numpy().reshape((m, n), order="F")
    yv = ydual.values().numpy()
    if inplace and not isinstance(y, _DeviceArray):
        y[...] = yv
    if isinstance(result, DiffResult):
        result.value = yv
    return result


# ---------------------------------------------------------------------------
# hessian
# ---------------------------------------------------------------------------
def hessian(f, x, cfg=None, check=True, ctx=None, chunks=None):
    """ForwardDiff.hessian(f, x, cfg) (src/hessian.jl:14-19); for an N-d x the Hessian is over vec(x)."""
    x = _vec(x)
    n = _length(x)
    result = np.zeros((n, n), order="F")
    return hessian_b(result, f, x, cfg, check, ctx, chunks)


def hessian_b(result, f, x, cfg=None, check=True, ctx=None, chunks=None):
    """ForwardDiff.hessian!(result, f, x, cfg) for an array or a HessianResult (src/hessian.jl:31-37,66-71)."""
    x = _vec(x)
    cfg = cfg if cfg is not None else HessianConfig(f, x)
    if check:
        checktag(cfg.tag, f, x)
    n, N = _length(x), cfg.N
    if N > n:
        raise ChunkSizeError(f"chunk size cannot be greater than ForwardDiff.structural_length(x) ({N} > {n})")
    fid = _catalogue_id(f, _lib.FN)
    ctx = _ctx_of(x, ctx)
    lo, hi = (0, -1) if chunks is None else chunks
    H = result.derivs[1] if isinstance(result, DiffResult) else result
    if H.shape != (n, n) or not H.flags.f_contiguous:
        raise DimensionMismatch("hessian!: result must be a column-major n x n float64 matrix")
    if fid is None:
        cfg.last_path = "program:specialised" if ctx.jit_enabled() else "program:interpreted"
        return _program_hessian(ctx, result, H, f, x, N, lo, hi)
    cfg.last_path = "catalogue"
    xh = np.ascontiguousarray(x.numpy() if isinstance(x, _DeviceArray) else x, dtype=np.float64).ravel()
    g = np.zeros(n)
    val = C.c_double(0.0)
    ctx.check(ctx.lib.fdb_hessian_host(ctx.h, fid, xh.ctypes.data_as(_dp), n, N, lo, hi, H.ctypes.data_as(_dp), n,
                                       g.ctypes.data_as(_dp), C.byref(val)))
    if isinstance(result, DiffResult):
        result.value = val.value
        result.derivs[0][...] = g
    return result
