"""ctypes binding of libfdb200.so (include/fdb200.h).

The library is the product: there is NO CPU fallback.  Importing this module
only loads the shared object (so host-side planning and the export check work
without a GPU); creating a Context without a CUDA device raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfdb200.so")

c_ctx = C.c_void_p
c_arr = C.c_void_p
_dp = C.POINTER(C.c_double)
i64 = C.c_int64


# ---- exceptions mirroring the reference's -----------------------------------
class FdbError(RuntimeError):
    pass


class DimensionMismatch(FdbError):       # src/gradient.jl:46,88-91; src/jacobian.jl:89,164
    pass


class ChunkSizeError(FdbError, ValueError):   # ArgumentError, src/gradient.jl:116-118
    pass


class InvalidTagException(FdbError):     # src/config.jl:28-35
    pass


class UnsupportedOp(FdbError):
    pass


class CudaError(FdbError):
    pass


class ChunkPlan(C.Structure):
    _fields_ = [(k, i64) for k in ("xlen", "remainder", "lastchunksize", "lastchunkindex", "middle_lo", "middle_hi",
                                   "nchunks", "chunk_lo", "chunk_hi", "elem_lo", "elem_hi")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# every symbol include/fdb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fdb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fdb_init": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "fdb_shutdown": (C.c_int, [c_ctx]),
    "fdb_last_error": (C.c_char_p, [c_ctx]),
    "fdb_sync": (C.c_int, [c_ctx]),
    "fdb_set_async": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_stream": (C.c_int, [c_ctx, C.c_void_p]),
    "fdb_get_stream": (C.c_int, [c_ctx, C.POINTER(C.c_void_p)]),
    "fdb_set_nansafe": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_lane_sharing": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_chunk_grouping": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_dmma": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_dmma_tile": (C.c_int, [c_ctx, C.c_int]),
    "fdb_set_dmma_tma_store": (C.c_int, [c_ctx, C.c_int]),
    "fdb_kernel_launches": (C.c_int, [c_ctx, C.POINTER(i64)]),
    "fdb_version": (C.c_char_p, []),
    "fdb_array_alloc": (C.c_int, [c_ctx, i64, C.c_int, C.c_int, C.POINTER(c_arr)]),
    "fdb_array_wrap": (C.c_int, [c_ctx, C.c_void_p, i64, i64, C.c_int, C.c_int, C.POINTER(c_arr)]),
    "fdb_array_free": (C.c_int, [c_arr]),
    "fdb_array_info": (C.c_int, [c_arr, C.POINTER(i64), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(i64),
                                 C.POINTER(C.c_void_p)]),
    "fdb_upload_values": (C.c_int, [c_arr, _dp]),
    "fdb_download_values": (C.c_int, [c_arr, _dp]),
    "fdb_upload_duals_aos": (C.c_int, [c_arr, _dp]),
    "fdb_download_duals_aos": (C.c_int, [c_arr, _dp]),
    "fdb_copy": (C.c_int, [c_arr, c_arr]),
    "fdb_pickchunksize": (i64, [i64, i64]),
    "fdb_structural_length": (i64, [C.c_int, i64]),
    "fdb_structural_index": (i64, [C.c_int, i64, i64]),
    "fdb_plan_chunks": (C.c_int, [i64, C.c_int, C.c_int, C.c_int, C.POINTER(ChunkPlan)]),
    "fdb_seed": (C.c_int, [c_ctx, c_arr, c_arr, i64, i64]),
    "fdb_seed_zero": (C.c_int, [c_ctx, c_arr, c_arr, i64, i64]),
    "fdb_seed_structured": (C.c_int, [c_ctx, c_arr, c_arr, C.c_int, i64, i64, i64]),
    "fdb_seed_zero_structured": (C.c_int, [c_ctx, c_arr, c_arr, C.c_int, i64, i64, i64]),
    "fdb_extract_gradient_chunk": (C.c_int, [c_ctx, c_arr, c_arr, i64, i64]),
    "fdb_extract_gradient": (C.c_int, [c_ctx, c_arr, c_arr]),
    "fdb_extract_gradient_chunk_structured": (C.c_int, [c_ctx, c_arr, c_arr, C.c_int, i64, i64, i64]),
    "fdb_fill": (C.c_int, [c_ctx, c_arr, C.c_double]),
    "fdb_extract_jacobian_chunk": (C.c_int, [c_ctx, c_arr, i64, c_arr, i64, i64]),
    "fdb_extract_jacobian": (C.c_int, [c_ctx, c_arr, i64, c_arr, i64]),
    "fdb_extract_value": (C.c_int, [c_ctx, c_arr, c_arr]),
    "fdb_cursor_set": (C.c_int, [c_ctx, i64, i64]),
    "fdb_cursor_advance": (C.c_int, [c_ctx, i64, i64]),
    "fdb_cursor_get": (C.c_int, [c_ctx, C.POINTER(i64), C.POINTER(i64)]),
    "fdb_seed_at_cursor": (C.c_int, [c_ctx, c_arr, c_arr]),
    "fdb_seed_zero_at_cursor": (C.c_int, [c_ctx, c_arr, c_arr]),
    "fdb_extract_gradient_chunk_at_cursor": (C.c_int, [c_ctx, c_arr, c_arr]),
    "fdb_extract_jacobian_chunk_at_cursor": (C.c_int, [c_ctx, c_arr, i64, c_arr]),
    "fdb_graph_begin": (C.c_int, [c_ctx]),
    "fdb_graph_end": (C.c_int, [c_ctx, C.POINTER(C.c_void_p)]),
    "fdb_graph_launch": (C.c_int, [c_ctx, C.c_void_p, i64]),
    "fdb_graph_destroy": (C.c_int, [C.c_void_p]),
    "fdb_map1": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr]),
    "fdb_map2": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr, c_arr]),
    "fdb_map2_scalar": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr, C.c_double, C.c_int]),
    "fdb_fma": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr, c_arr, c_arr, C.c_double]),
    "fdb_dual_matvec": (C.c_int, [c_ctx, c_arr, i64, i64, i64, c_arr, c_arr]),
    "fdb_dual_matmul": (C.c_int, [c_ctx, c_arr, i64, i64, i64, c_arr, i64, i64, c_arr, i64]),
    "fdb_matmul": (C.c_int, [c_ctx, c_arr, i64, i64, i64, c_arr, i64, i64, c_arr, i64]),
    "fdb_reduce": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr]),
    "fdb_eval_scalar_fn": (C.c_int, [c_ctx, C.c_int, c_arr, c_arr]),
    "fdb_gradient_chunked": (C.c_int, [c_ctx, C.c_int, c_arr, C.c_int, i64, i64, c_arr, c_arr]),
    "fdb_jacobian_chunked": (C.c_int, [c_ctx, C.c_int, c_arr, i64, C.c_int, c_arr, i64, c_arr, C.c_double, i64, i64,
                                       c_arr, i64, c_arr]),
    "fdb_jacobian_mlp2_chunked": (C.c_int, [c_ctx, c_arr, i64, i64, C.c_int, c_arr, i64, c_arr, c_arr, i64, c_arr, i64, i64, c_arr, i64,
                                            c_arr]),
    "fdb_hessian_chunked": (C.c_int, [c_ctx, C.c_int, c_arr, C.c_int, i64, i64, c_arr, i64, c_arr, c_arr]),
    "fdb_program_gradient_chunked": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int,
                                               C.c_int, C.c_int, C.c_int, c_arr, C.c_int, i64, i64, c_arr, c_arr]),
    "fdb_program_gradient_chunked_structured": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int,
                                                          C.c_int, C.c_int, C.c_int, C.c_int, i64, c_arr, C.c_int, i64, i64, c_arr, c_arr]),
    "fdb_program_jacobian_chunked": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int, C.c_int, C.c_int, c_arr,
                                               C.c_int, i64, i64, c_arr, i64, c_arr]),
    "fdb_program_matvec_jacobian_chunked": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int, C.c_int, c_arr, i64, i64, c_arr,
                                                      C.c_int, i64, i64, c_arr, i64, c_arr]),
    "fdb_program_hessian_chunked": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int,
                                              C.c_int, C.c_int, C.c_int, c_arr, C.c_int, i64, i64, c_arr, i64, c_arr, c_arr]),
    "fdb_set_jit": (C.c_int, [c_ctx, C.c_int]),
    "fdb_jit_stats": (C.c_int, [c_ctx, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "fdb_jit_last_log": (C.c_char_p, [c_ctx]),
    "fdb_jit_compile_check": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, _dp, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, i64, C.POINTER(i64)]),
    "fdb_program_bind_params": (C.c_int, [c_ctx, C.c_int, C.POINTER(C.c_void_p)]),
    "fdb_gradient_host": (C.c_int, [c_ctx, C.c_int, _dp, i64, C.c_int, i64, i64, _dp, _dp]),
    "fdb_hessian_host": (C.c_int, [c_ctx, C.c_int, _dp, i64, C.c_int, i64, i64, _dp, i64, _dp, _dp]),
    "fdb_jacobian_host": (C.c_int, [c_ctx, C.c_int, _dp, i64, i64, C.c_int, _dp, i64, _dp, C.c_double, C.c_int, i64, i64, _dp, i64, _dp]),
    "fdb_host_register": (C.c_int, [C.c_void_p, i64]),
    "fdb_host_unregister": (C.c_int, [C.c_void_p]),
    "fdb_multi_init": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "fdb_multi_shutdown": (C.c_int, [C.c_void_p]),
    "fdb_multi_size": (C.c_int, [C.c_void_p]),
    "fdb_multi_ctx": (c_ctx, [C.c_void_p, C.c_int]),
    "fdb_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "fdb_multi_sync": (C.c_int, [C.c_void_p]),
    "fdb_multi_gradient_host": (C.c_int, [C.c_void_p, C.c_int, _dp, i64, C.c_int, _dp, _dp]),
    "fdb_multi_jacobian_host": (C.c_int, [C.c_void_p, C.c_int, _dp, i64, i64, C.c_int, _dp, i64, _dp, C.c_double, C.c_int, _dp, i64, _dp]),
    "fdb_multi_hessian_host": (C.c_int, [C.c_void_p, C.c_int, _dp, i64, C.c_int, _dp, i64, _dp, _dp]),
    "fdb_multi_exchange_create": (C.c_int, [C.c_void_p, i64, C.POINTER(C.c_void_p)]),
    "fdb_multi_exchange_destroy": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "fdb_multi_barrier": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "fdb_multi_upload": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), i64, _dp, i64]),
    "fdb_multi_download": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, i64, _dp, i64]),
    "fdb_multi_gradient_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, i64, i64, C.c_int, i64, i64]),
    "fdb_multi_jacobian_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, i64, i64, i64, C.c_int, C.POINTER(C.c_void_p), i64,
                                              C.POINTER(C.c_void_p), C.c_double, i64, i64, i64]),
    "fdb_multi_hessian_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, i64, i64, C.c_int, i64, i64, i64, i64]),
    "fdb_exchange_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "fdb_exchange_set_timeout": (C.c_int, [C.c_void_p, C.c_double]),
    "fdb_exchange_create": (C.c_int, [c_ctx, i64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fdb_exchange_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fdb_exchange_connect": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fdb_exchange_view": (C.c_int, [C.c_void_p, i64, i64, i64, C.POINTER(c_arr)]),
    "fdb_exchange_enable": (C.c_int, [c_ctx, C.c_void_p]),
    "fdb_exchange_barrier": (C.c_int, [C.c_void_p]),
    "fdb_exchange_push": (C.c_int, [C.c_void_p, i64, i64]),
    "fdb_exchange_timeouts": (C.c_int, [C.c_void_p, C.POINTER(i64)]),
    "fdb_exchange_destroy": (C.c_int, [C.c_void_p]),
    "fdb_exchange_measure_store_bw": (C.c_int, [C.c_void_p, i64, C.c_int, C.c_int, _dp]),
    "fdb_multicast_supported": (C.c_int, [c_ctx, C.POINTER(C.c_int)]),
    "fdb_exchange_create_multicast": (C.c_int, [c_ctx, i64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fdb_exchange_share_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fdb_exchange_multicast_create": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fdb_exchange_connect_multicast": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "fdb_exchange_multicast_bind": (C.c_int, [C.c_void_p]),
    "fdb_exchange_set_multicast": (C.c_int, [C.c_void_p, C.c_int]),
    "fdb_measure_fp64_peak": (C.c_int, [c_ctx, _dp, _dp]),
}
IPC_HANDLE_BYTES = 64
SHARE_HANDLE_BYTES = 16

# enums of fdb200.h
OP1 = dict(neg=1, exp=2, log=3, sin=4, cos=5, tanh=6, sqrt=7, abs=8, abs2=9, inv=10, pow0=11, pow1=12, pow2=13, pow3=14,
           tan=15, exp2=16, log2=17, log10=18, log1p=19, expm1=20, sinh=21, cosh=22, asin=23, acos=24, atan=25, cbrt=26)
OP2 = dict(add=1, sub=2, mul=3, div=4, pow=5, max=6, min=7, atan2=8, hypot=9, nanmax=10, nanmin=11)
REDOP = dict(sum=1, prod=2)
FN = dict(rosenbrock=1, ackley=2, sumsq=3, prod=4)
JFN = dict(tanh_affine=1, brusselator=2, jactest=3)
STRUCTURE = dict(dense=0, upper=1, lower=2, diagonal=3)

_ERRORS = {1: FdbError, 2: DimensionMismatch, 3: CudaError, 4: MemoryError, 5: UnsupportedOp, 6: ChunkSizeError}

_lib = None


def load():
    """Load libfdb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, ctx_handle=None):
    if rc == 0:
        return
    lib = load()
    msg = lib.fdb_last_error(ctx_handle)
    msg = msg.decode() if msg else f"fdb200 error {rc}"
    raise _ERRORS.get(rc, FdbError)(msg)
