"""fdb200 -- Python host for the B200-native engine behind ForwardDiff.jl's
chunked dual-number sweep (ctypes over include/fdb200.h; no CPU fallback).

The directory `forwarddiff.jl_b200/` is not an importable name; add it to
sys.path (tests/conftest.py, bench.py and __graft_entry__.py do) and
`import fdb200`.
"""
from . import _lib
from ._lib import (ChunkSizeError, CudaError, DimensionMismatch, FdbError, InvalidTagException, UnsupportedOp, LIB_PATH)
from .array import (B200Array, B200DualArray, Context, abs2, abs_, acos, asin, atan, atan2, cbrt, cos, cosh, dot, dual_matmul, exp, exp2,
                    expm1, fma, hypot, inv, log, log10, log1p, log2, maximum, minimum, nanmath, sin, sinh, sqrt, tan, tanh)
from .api import (Chunk, Diagonal, DiffResult, LowerTriangular, UpperTriangular, GradientConfig, GradientResult, HessianConfig, HessianResult, JacobianConfig,
                  JacobianResult, Tag, checktag, chunk_plan, default_context, gradient, gradient_b, hessian, hessian_b,
                  jacobian, jacobian_b, pickchunksize, SlowPathWarning)
from . import functions
from . import dist
from . import multi
from .multi import DeviceGroup

__all__ = [n for n in dir() if not n.startswith("_")]
