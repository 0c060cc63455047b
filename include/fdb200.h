/* fdb200.h -- C ABI of the B200-native engine for ForwardDiff.jl's chunked
 * dual-number sweep.
 *
 * This is the drop-in boundary: every entry point replaces one piece of the
 * reference's chunk-mode path (file:line under /root/reference given with each
 * declaration).  A Julia host reaches it through `ccall` (see INTEGRATION.md and
 * forwarddiff.jl_b200/julia/ForwardDiffB200.jl); the Python harness in
 * forwarddiff.jl_b200/fdb200/ binds the same symbols through ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types cross the boundary;
 *  - every function returns an fdb_status (0 = ok); the message of the last
 *    failure on a context is available from fdb_last_error();
 *  - indices follow the reference: `index` arguments are 1-based structural
 *    positions exactly as in src/apiutils.jl and src/gradient.jl; chunk ranges
 *    [chunk_lo, chunk_hi) are 0-based chunk numbers (chunk c seeds structural
 *    positions c*N+1 .. min((c+1)*N, n), src/gradient.jl:141-152);
 *  - device arrays are SoA: an array of n Dual{T,Float64,N} is stored as N+1
 *    planes of `ld` doubles (plane 0 = values, plane k = partial k); a nested
 *    Dual{T,Dual{T,Float64,N},N} array has (N+1)^2 planes ordered
 *    plane(a,b) = a*(N+1)+b with a the outer and b the inner index (0 = value),
 *    which is the reference's AoS field order (src/dual.jl:14-16,357-360);
 *  - host-side `*_aos` buffers use that AoS order: element e occupies
 *    doubles [e*P, (e+1)*P), P = number of planes;
 *  - one context per GPU.  A process may own several contexts (one host process driving all
 *    GPUs: fdb_multi_*), or there may be one process per GPU (fdb_exchange_* over CUDA IPC).
 *    Every entry point makes its context's device current for the duration of the call and
 *    restores the caller's.  Calls on a context are issued on its CUDA stream and are
 *    synchronous unless the context was put in async mode with fdb_set_async(ctx, 1);
 *  - a context is not re-entrant: one in-flight API call per context;
 *  - there is no CPU fallback: without a CUDA device fdb_init fails.
 */
#ifndef FDB200_H
#define FDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdb_ctx fdb_ctx;     /* opaque: device, stream, scratch, last error */
typedef struct fdb_array fdb_array; /* opaque: SoA device array (plain, Dual or nested Dual) */

typedef enum {
    FDB_OK = 0,
    FDB_ERR_BADARG = 1,      /* null pointer, bad enum, bad N (1..12; nested 1..8,12 unsupported above 8) */
    FDB_ERR_SHAPE = 2,       /* DimensionMismatch (src/gradient.jl:46,88-91; src/jacobian.jl:89,164) */
    FDB_ERR_CUDA = 3,
    FDB_ERR_OOM = 4,
    FDB_ERR_UNSUPPORTED = 5, /* op / function id not in the device catalogue */
    FDB_ERR_CHUNK = 6        /* ArgumentError "chunk size cannot be greater than structural_length(x)"
                                (src/gradient.jl:116-118, src/jacobian.jl:172-174) */
} fdb_status;

#define FDB_MAX_CHUNK 12 /* DEFAULT_CHUNK_THRESHOLD, src/prelude.jl:2 */

/* unary ops: src/dual.jl:476-490 (DiffRules loop), :519 (-), :587-597 (literal_pow), :709-717 (sin/cos) */
typedef enum {
    FDB_NEG = 1, FDB_EXP, FDB_LOG, FDB_SIN, FDB_COS, FDB_TANH, FDB_SQRT, FDB_ABS, FDB_ABS2, FDB_INV,
    FDB_POW0, FDB_POW1, FDB_POW2, FDB_POW3,
    FDB_TAN, FDB_EXP2, FDB_LOG2, FDB_LOG10, FDB_LOG1P, FDB_EXPM1, FDB_SINH, FDB_COSH, FDB_ASIN, FDB_ACOS, FDB_ATAN, FDB_CBRT
} fdb_op1;
/* binary ops: src/dual.jl:499-519 (+ -), DiffRules * max min, :532-544 (/), :549-585 (^) */
/* NaNMath (src/dual.jl:477,549): NaNMath.pow / log / sin / sqrt ... return NaN outside their domain where Base throws a DomainError.
 * Device math never throws -- out-of-domain arguments already give NaN -- so FDB_POW, FDB_LOG, FDB_SIN ... ARE the NaNMath variants
 * (pinned by test/DualTest.jl:514, test/DerivativeTest.jl:98-102).  NaNMath.max / NaNMath.min differ in value: they ignore a NaN
 * operand (FDB_NANMAX / FDB_NANMIN; derivative 1 for the argument returned). */
typedef enum { FDB_ADD = 1, FDB_SUB, FDB_MUL, FDB_DIV, FDB_POW, FDB_MAX, FDB_MIN, FDB_ATAN2, FDB_HYPOT, FDB_NANMAX, FDB_NANMIN } fdb_op2;
typedef enum { FDB_SUM = 1, FDB_PROD = 2 } fdb_redop;

/* scalar-valued catalogue targets for fdb_gradient_chunked / fdb_hessian_chunked */
typedef enum {
    FDB_FN_ROSENBROCK = 1, /* docs/src/user/advanced.md:36-44 (= DiffTests.rosenbrock_1) */
    FDB_FN_ACKLEY = 2,     /* DiffTests.ackley; C++ twin benchmarks/cpp/benchmarks.cpp:10-21 */
    FDB_FN_SUMSQ = 3,      /* sum of x[i]^2 */
    FDB_FN_PROD = 4        /* prod(x) */
} fdb_scalar_fn;
/* array-valued catalogue targets for fdb_jacobian_chunked */
typedef enum {
    FDB_JFN_TANH_AFFINE = 1, /* f(x) = tanh.(A*x .+ b) */
    FDB_JFN_BRUSSELATOR = 2, /* f!(dx, x): 1-D Brusselator RHS, x = [u; v] */
    FDB_JFN_JACTEST = 3      /* test/JacobianTest.jl:23-31 (n = 3, m = 4) */
} fdb_array_fn;

/* structural index sets of src/apiutils.jl:44-71 */
typedef enum { FDB_DENSE = 0, FDB_UPPER = 1, FDB_LOWER = 2, FDB_DIAGONAL = 3 } fdb_structure;

/* ---- lifecycle ---------------------------------------------------------- */
int fdb_device_count(int* count);
int fdb_init(int device, fdb_ctx** ctx);          /* binds the context to one GPU, creates its stream */
int fdb_shutdown(fdb_ctx* ctx);
const char* fdb_last_error(fdb_ctx* ctx);          /* ctx may be NULL: last error of a failed fdb_init */
int fdb_sync(fdb_ctx* ctx);
int fdb_set_async(fdb_ctx* ctx, int async);       /* 1: calls return after enqueueing (benchmarks) */
/* adopt a caller-owned cudaStream_t.  NULL returns to the context's own (non-blocking) stream -- it does NOT mean the
 * CUDA default stream: to run on the legacy default stream (what torch calls its default stream, handle 0) pass
 * cudaStreamLegacy = (void*)0x1, for the per-thread default stream cudaStreamPerThread = (void*)0x2. */
int fdb_set_stream(fdb_ctx* ctx, void* cuda_stream);
int fdb_get_stream(fdb_ctx* ctx, void** cuda_stream);
int fdb_set_nansafe(fdb_ctx* ctx, int enabled);   /* `nansafe_mode` preference, src/prelude.jl:1 */
/* 1 (default): terms that touch no seeded position are evaluated on one representative
 * partial lane (bit-identical results; see fdb_sweep.cu).  0: all N lanes for every term. */
int fdb_set_lane_sharing(fdb_ctx* ctx, int enabled);
/* Opt-in for fdb_gradient_chunked / fdb_gradient_host on catalogue targets with lane sharing on: a thread block sweeps `group`
 * consecutive chunks (1..64) and evaluates the terms that read no seeded position of any of them ONCE (their Dual{1} result is
 * the same for every chunk of the group).  Same gradient lane by lane; the FP64 work drops by ~group.  Default 1: the ceil(n/N)
 * chunk sweeps of src/gradient.jl:141-152 stay independent evaluations of f (what bench.py's headline measures). */
int fdb_set_chunk_grouping(fdb_ctx* ctx, int group);
/* 1 (default): A*xdual contractions run on FP64 tensor cores (DMMA) when shapes allow; 0: CUDA-core kernels */
int fdb_set_dmma(fdb_ctx* ctx, int enabled);
/* operand columns per thread-block tile of the DMMA contraction: 0 (default) = chosen per shape so that the equal-sized
 * tiles fill the SMs evenly, or 16 / 32 / 64 (measurement knob) */
int fdb_set_dmma_tile(fdb_ctx* ctx, int bn);
/* how the fused DMMA epilogue writes its 128 x 4N result blocks: 0 = every thread stores its own accumulator entries;
 * 2 = staged in shared memory and written by TMA tensor stores (local result + one store per peer arena); 3 = staged and copied
 * out by the block's threads in 256-byte runs; 1 (default) = staged when the result is mirrored to peers over NVLink, 0 otherwise */
int fdb_set_dmma_tma_store(fdb_ctx* ctx, int mode);
int fdb_kernel_launches(fdb_ctx* ctx, int64_t* count); /* kernels launched on this context so far */
const char* fdb_version(void);

/* ---- memory ------------------------------------------------------------------
 * fdb_array_alloc replaces `similar(x, Dual{T,V,N})` in the config constructors
 * (src/config.jl:122,159,185-186,225-226).  level 0 = Vector{Float64} (N ignored),
 * 1 = Vector{Dual{T,Float64,N}}, 2 = Vector{Dual{T,Dual{T,Float64,N},N}}. */
int fdb_array_alloc(fdb_ctx* ctx, int64_t n, int N, int level, fdb_array** out);
int fdb_array_wrap(fdb_ctx* ctx, double* device_ptr, int64_t n, int64_t ld, int N, int level, fdb_array** out);
int fdb_array_free(fdb_array* a);
int fdb_array_info(const fdb_array* a, int64_t* n, int* N, int* level, int64_t* ld, double** device_ptr);
int fdb_upload_values(fdb_array* a, const double* host);      /* plane 0 <- host[0..n) */
int fdb_download_values(const fdb_array* a, double* host);    /* host[0..n) <- plane 0 */
int fdb_upload_duals_aos(fdb_array* a, const double* host_aos);   /* AoS (src/dual.jl:357-360) -> SoA planes */
int fdb_download_duals_aos(const fdb_array* a, double* host_aos);
int fdb_copy(fdb_array* dst, const fdb_array* src);            /* same shape, device to device */

/* ---- host-side planning (pure host code, no GPU needed) ------------------ */
/* pickchunksize, src/prelude.jl:29-36 */
int64_t fdb_pickchunksize(int64_t input_length, int64_t threshold);
/* structural_length, src/prelude.jl:15-20 (rows = length for FDB_DENSE) */
int64_t fdb_structural_length(int structure, int64_t rows);
/* 0-based column-major storage index of structural position pos (0-based), src/apiutils.jl:44-71 */
int64_t fdb_structural_index(int structure, int64_t rows, int64_t pos);
/* chunk bookkeeping of src/gradient.jl:121-125 (= src/jacobian.jl:177-181) and the
 * contiguous chunk range owned by `rank` of `world`. */
typedef struct {
    int64_t xlen, remainder, lastchunksize, lastchunkindex; /* lastchunkindex is 1-based */
    int64_t middle_lo, middle_hi;                            /* middlechunks = middle_lo:middle_hi (1-based) */
    int64_t nchunks;
    int64_t chunk_lo, chunk_hi;                              /* this rank's chunks, 0-based half-open */
    int64_t elem_lo, elem_hi;                                /* structural positions they seed, 0-based half-open */
} fdb_chunk_plan;
int fdb_plan_chunks(int64_t xlen, int N, int rank, int world, fdb_chunk_plan* out);

/* ---- sweep utilities (bit-exact) --------------------------------------------
 * duals: level-1 (or level-2) array; x: array one level below holding the values. */
/* seed!(duals, x, index, seeds, chunksize), src/apiutils.jl:125-143; index<=0 selects
 * seed!(duals, x, seeds), src/apiutils.jl:107-123 (first N structural positions) */
int fdb_seed(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x, int64_t index, int64_t chunksize);
/* seed_zero_partials!(duals, x, index, count), src/apiutils.jl:83-87; index<=0 selects
 * seed_zero_partials!(duals, x), src/apiutils.jl:76-77 (every structural position) */
int fdb_seed_zero(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x, int64_t index, int64_t count);
/* the same on UpperTriangular / LowerTriangular / Diagonal storage (rows x rows, column-major) */
int fdb_seed_structured(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x, int structure, int64_t rows,
                        int64_t index, int64_t chunksize);
int fdb_seed_zero_structured(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x, int structure, int64_t rows,
                             int64_t index, int64_t count);
/* extract_gradient_chunk!(T, result, dual, index, chunksize), src/gradient.jl:74-81;
 * ydual is a dual array of length 1 */
int fdb_extract_gradient_chunk(fdb_ctx* ctx, fdb_array* result, const fdb_array* ydual, int64_t index,
                               int64_t chunksize);
/* the same with `structural_eachindex(result)` of an UpperTriangular / LowerTriangular / Diagonal result
 * (rows x rows column-major storage; test/GradientTest.jl:250-276) */
int fdb_extract_gradient_chunk_structured(fdb_ctx* ctx, fdb_array* result, const fdb_array* ydual, int structure, int64_t rows,
                                          int64_t index, int64_t chunksize);
/* fill!(a, 0.0) on every plane: off-structure entries of a structured Dual work buffer read as zero(Dual) */
int fdb_fill(fdb_ctx* ctx, fdb_array* a, double value);
/* extract_gradient!(T, result, dual), src/gradient.jl:66-72 (vector mode) */
int fdb_extract_gradient(fdb_ctx* ctx, fdb_array* result, const fdb_array* ydual);
/* extract_jacobian_chunk!(T, result, ydual, index, chunksize), src/jacobian.jl:109-118;
 * result is a plain array viewed as column-major with leading dimension ld_result */
int fdb_extract_jacobian_chunk(fdb_ctx* ctx, fdb_array* result, int64_t ld_result, const fdb_array* ydual,
                               int64_t index, int64_t chunksize);
/* extract_jacobian!(T, result, ydual, n), src/jacobian.jl:95-102 */
int fdb_extract_jacobian(fdb_ctx* ctx, fdb_array* result, int64_t ld_result, const fdb_array* ydual, int64_t n);
/* extract_value!: map!(d -> value(T,d), y, ydual), src/apiutils.jl:9-12 */
int fdb_extract_value(fdb_ctx* ctx, fdb_array* y, const fdb_array* ydual);

/* ---- the reference chunk loop as a replayed CUDA graph (fdb_graph.cu) ----------------------------------
 * The window {index, chunksize} of the current chunk lives in device memory (the context's chunk cursor);
 * the `_at_cursor` sweep utilities read it, fdb_cursor_advance moves it by one chunk (index += step,
 * chunksize = min(step, xlen - index + 1): the short last chunk of src/gradient.jl:121-125 falls out).
 * One loop iteration  seed -> f -> extract -> unseed -> advance  then has constant launch parameters: capture
 * it once between fdb_graph_begin / fdb_graph_end and replay it with fdb_graph_launch for the remaining chunks.
 * Arrays allocated during capture must stay alive until fdb_graph_destroy. */
int fdb_cursor_set(fdb_ctx* ctx, int64_t index, int64_t chunksize);
int fdb_cursor_advance(fdb_ctx* ctx, int64_t step, int64_t xlen);
int fdb_cursor_get(fdb_ctx* ctx, int64_t* index, int64_t* chunksize);
int fdb_seed_at_cursor(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x);
int fdb_seed_zero_at_cursor(fdb_ctx* ctx, fdb_array* duals, const fdb_array* x);
int fdb_extract_gradient_chunk_at_cursor(fdb_ctx* ctx, fdb_array* result, const fdb_array* ydual);
int fdb_extract_jacobian_chunk_at_cursor(fdb_ctx* ctx, fdb_array* result, int64_t ld_result, const fdb_array* ydual);
int fdb_graph_begin(fdb_ctx* ctx);
int fdb_graph_end(fdb_ctx* ctx, void** graph_exec);
int fdb_graph_launch(fdb_ctx* ctx, void* graph_exec, int64_t times);
int fdb_graph_destroy(void* graph_exec);

/* ---- array arithmetic on Dual arrays (what a user `f` is made of) ----------
 * Operands of fdb_map2 may each be a level-1 Dual array or a level-0 plain
 * array (a Vector{Float64} broadcast against duals = the Real (x)or Dual bodies
 * of src/dual.jl:150-210); at least one must be level 1.  A length-1 operand
 * broadcasts. */
int fdb_map1(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a);
int fdb_map2(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a, const fdb_array* b);
/* scalar_first = 0: a (op) s ; 1: s (op) a */
int fdb_map2_scalar(fdb_ctx* ctx, int op, fdb_array* out, const fdb_array* a, double s, int scalar_first);
/* fma, src/dual.jl:625-665.  kind 0: fma(a,b,c) all dual; 1: fma(a,b,s); 2: fma(a,s,c) */
int fdb_fma(fdb_ctx* ctx, int kind, fdb_array* out, const fdb_array* a, const fdb_array* b, const fdb_array* c,
            double s);
/* Dual matvec / matmul (Julia: A * xdual through LinearAlgebra.generic_matvecmul! on Vector{Dual}): a dense
 * contraction of A (m x n, plain, column-major, lda) against the planes [value | partials] of xdual, on FP64
 * tensor-core (DMMA) tiles staged by TMA bulk copies.  fdb_matmul is the same contraction on plain arrays:
 * C[m x ncols] = A[m x k] * B[k x ncols], all column-major. */
int fdb_dual_matvec(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t n, const fdb_array* xdual, fdb_array* ydual);
int fdb_matmul(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t k, const fdb_array* B, int64_t ldb, int64_t ncols,
               fdb_array* C, int64_t ldc);
/* C = A * B where BOTH matrices hold Duals (Julia: generic matrix product over the Dual*Dual body, src/dual.jl:260-300):
 * value(C) = value(A) value(B), partial_j(C) = value(A) partial_j(B) + partial_j(A) value(B): 2N+1 real FP64 tensor-core
 * contractions.  A (m x k), B (k x n), C (m x n): level-1 Dual arrays of column-major matrices, same chunk size. */
int fdb_dual_matmul(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t k, const fdb_array* B, int64_t ldb, int64_t n,
                    fdb_array* C, int64_t ldc);
/* sum / prod of a Dual array -> Dual array of length 1 (two-pass, deterministic) */
int fdb_reduce(fdb_ctx* ctx, int redop, fdb_array* out1, const fdb_array* a);
/* evaluate a catalogue target on a materialised dual array: out1 = f(xdual) */
int fdb_eval_scalar_fn(fdb_ctx* ctx, int fn, fdb_array* out1, const fdb_array* xdual);

/* ---- whole-path drivers ---------------------------------------------------------
 * Each replaces one chunk-mode loop of the reference and runs ALL requested chunk
 * sweeps batched on the device (seeds generated in registers, partials never leave
 * the SM, extraction fused).  x, grad, J, H, y are level-0 arrays.  Chunks
 * [chunk_lo, chunk_hi) are evaluated (chunk_hi < 0 = all) -- this is the sharding
 * hook: ranks own disjoint chunk ranges and therefore disjoint slices of grad / J / H.
 * N == xlen selects vector mode (src/gradient.jl:19-23). */
/* chunk_mode_gradient!, src/gradient.jl:114-159.  value_dev (level-0 array of length >= 1,
 * may be NULL) receives f(x) when the range includes the last chunk (gradient.jl:155). */
int fdb_gradient_chunked(fdb_ctx* ctx, int fn, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi,
                         fdb_array* grad, fdb_array* value_dev);
/* chunk_mode_jacobian(!), src/jacobian.jl:169-244.  J: column-major m x n with leading
 * dimension ldJ, columns of the requested chunks written; y (may be NULL) <- value.(ydual).
 * A (m x n column-major, lda), b (m) and alpha are the target's Float64 parameters. */
int fdb_jacobian_chunked(fdb_ctx* ctx, int fn, const fdb_array* x, int64_t m, int N, const fdb_array* A,
                         int64_t lda, const fdb_array* b, double alpha, int64_t chunk_lo, int64_t chunk_hi,
                         fdb_array* J, int64_t ldJ, fdb_array* y);
/* the two-layer target f(x) = tanh.(A2 * tanh.(A1*x .+ b1) .+ b2) (A1: h x n, A2: m x h): after the first layer
 * the partials are dense, so the second A*hdual is a dense FP64 tensor-core contraction against
 * [value | N partials] of every chunk.  Same chunk-range semantics as above. */
int fdb_jacobian_mlp2_chunked(fdb_ctx* ctx, const fdb_array* x, int64_t h, int64_t m, int N, const fdb_array* A1, int64_t lda1,
                              const fdb_array* b1, const fdb_array* A2, int64_t lda2, const fdb_array* b2, int64_t chunk_lo,
                              int64_t chunk_hi, fdb_array* J, int64_t ldJ, fdb_array* y);
/* hessian / hessian!(::DiffResult), src/hessian.jl:14-71: outer chunks [lo,hi) of the
 * Jacobian of the chunked gradient; H column-major n x n (ldH); grad, value_dev may be NULL. */
int fdb_hessian_chunked(fdb_ctx* ctx, int fn, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi,
                        fdb_array* H, int64_t ldH, fdb_array* grad, fdb_array* value_dev);

/* ---- op-programs: fused broadcast trees for an arbitrary user f -----------------------------
 * What Julia's `Broadcasted` tree + `sum` gives for a user function is flattened by the host into a small
 * program: instructions are 4 int32 words {op, dst, a, b} with op = kind*64 + code,
 *   kind 0 LOADX  dst <- seeded Dual of x[i + b]            (term part only; 0 <= b <= halo)
 *   kind 1 UNARY  dst <- code(a)                            code = fdb_op1
 *   kind 2 DD     dst <- a (code) b                         code = fdb_op2, both Dual registers
 *   kind 3 DC     dst <- a (code) consts[b]                 Dual (op) Real
 *   kind 4 CD     dst <- consts[a] (code) b                 Real (op) Dual
 *   kind 5 ACC    accumulator dst += a                      (`sum`; term part only)
 *   kind 6 MOV    dst <- a
 * and, executed by the run-time specialised kernels only (fdb_set_jit; the interpreter answers FDB_ERR_UNSUPPORTED):
 *   kind 7 DP     dst <- a (code) param[b][i]               Dual (op) data array (fdb_program_bind_params)
 *   kind 8 PD     dst <- param[a][i] (code) b               data array (op) Dual
 *   kind 9 GUARD  the next b instructions run only while i + a < n    (a reduction over a shorter view; every term
 *                 instruction of a program with guards belongs to a guarded group; terms run for i in [0, n - min a))
 *   kind 10 STAGE a = 1: the first-stage epilogue follows (scalar code on accumulators 0..K1-1, K1 = number of
 *                 accumulators used so far); a = 2: the second-stage terms follow.  term1 | STAGE 1 | epilogue1 |
 *                 STAGE 2 | term2: a reduced Dual re-enters a second broadcast (sum((x .- mean(x)).^2)); gradients only
 *   kind 11 LOADS dst <- S[a], S = the register file as the first-stage epilogue left it (term2 and final epilogue;
 *                 there registers K1..n_acc-1 are preloaded with the second-stage accumulators)
 * The term part runs for every i in [0, n - halo); the epilogue part runs once per chunk with registers
 * 0..n_acc-1 preloaded with the reduced accumulators; `result_reg` holds f(xdual).  PK_ACC with b = 1 is a PRODUCT
 * accumulator (`prod`: starts at one(Dual), combines by Dual*Dual; single-pass gradient programs, specialised kernels only).
 * Limits: 256 instructions per part, 64 constants, 32 registers, 8 accumulators for the run-time specialised kernels; the
 * interpreter takes 96 / 24 / 12 / 4 (FDB_ERR_UNSUPPORTED beyond).  Same chunk-range
 * semantics as fdb_gradient_chunked / fdb_jacobian_chunked.  The Jacobian form evaluates
 * y[i] = program(x[i .. i+halo]), i in [0, n - halo) (elementwise / stencil array functions); with the two-pass markers
 * (pass-1 sums, scalar epilogue, second broadcast whose result register is y[i]) it evaluates array functions that use reduced
 * values -- softmax exp.(x) ./ sum(exp.(x)) -- whose Jacobian is dense: one thread block per chunk, specialised kernels only. */
int fdb_program_gradient_chunked(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                 const double* consts, int n_consts, int n_acc, int halo, int result_reg, const fdb_array* x, int N,
                                 int64_t chunk_lo, int64_t chunk_hi, fdb_array* grad, fdb_array* value_dev);
/* the gradient form for UpperTriangular / LowerTriangular / Diagonal inputs (src/apiutils.jl:44-71, test/GradientTest.jl:250-276):
 * x and grad are the rows x rows column-major storage (off-structure entries of x must be zero; only structural entries of grad are
 * written), chunks run over the structural positions.  Single-pass programs; specialised kernels only. */
int fdb_program_gradient_chunked_structured(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                            const double* consts, int n_consts, int n_acc, int halo, int result_reg, int structure, int64_t rows,
                                            const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi, fdb_array* grad, fdb_array* value_dev);
int fdb_program_jacobian_chunked(fdb_ctx* ctx, const int32_t* prog, int n_instr, const double* consts, int n_consts, int halo,
                                 int result_reg, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi, fdb_array* J,
                                 int64_t ldJ, fdb_array* y);

/* jacobian(x -> g.(A * x), x, cfg) for a user function that contracts the input with a Float64 matrix A (m x xlen, plain,
 * column-major, lda) and applies an elementwise program g to the product: `A * xdual` through Julia's generic matvec over Dual
 * (src/dual.jl:476-490 bodies; caller src/jacobian.jl:220 `ydual = f(xdual)`).  The contraction of A against the chunk's lanes
 * runs on the FP64 tensor cores (DMMA), the program -- same instruction format, LOADX with offset 0 yields z_i = (A*x)_i, array
 * parameters are indexed by the row i -- runs on the product.  Same chunk-range semantics as fdb_jacobian_chunked. */
int fdb_program_matvec_jacobian_chunked(fdb_ctx* ctx, const int32_t* prog, int n_instr, const double* consts, int n_consts, int result_reg,
                                        const fdb_array* A, int64_t lda, int64_t m, const fdb_array* x, int N, int64_t chunk_lo,
                                        int64_t chunk_hi, fdb_array* J, int64_t ldJ, fdb_array* y);

/* hessian(f, x, cfg) of a flattened f (same program form as fdb_program_gradient_chunked; ops with a nested-dual rule:
 * + - * / neg, literal powers, exp log sin cos tanh sqrt abs abs2 inv).  Outer chunks [lo, hi).  f = one plain sum of terms
 * (n_acc = 1, no epilogue) runs on either executor; several sums and arithmetic after the reductions (e.g. Ackley) need the
 * run-time specialised kernel, which finalizes on full Dual{Dual{N},N} numbers (FDB_ERR_UNSUPPORTED otherwise). */
int fdb_program_hessian_chunked(fdb_ctx* ctx, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi,
                                const double* consts, int n_consts, int n_acc, int halo, int result_reg, const fdb_array* x, int N,
                                int64_t chunk_lo, int64_t chunk_hi, fdb_array* H, int64_t ldH, fdb_array* grad, fdb_array* value_dev);

/* Data arrays the user's f closes over (`f(x) = sum(w .* x.^2)`): up to FDB_MAX_PARAMS plain Float64 device arrays, read at the
 * term index i by the program instructions of kind 7 (Dual (x) param[b][i], Real-on-the-right rules of src/dual.jl) and 8
 * (param[a][i] (x) Dual).  Each must hold at least xlen - halo entries.  They stay bound until rebound; count = 0 unbinds.
 * Programs with array parameters run on the specialised kernels only. */
#define FDB_MAX_PARAMS 4
int fdb_program_bind_params(fdb_ctx* ctx, int count, const fdb_array* const* arrays);

/* Run-time specialisation (fdb_jit.cu).  Julia compiles f for Vector{Dual{T,Float64,N}} when gradient(f, x, cfg) first
 * runs; likewise the fdb_program_* drivers turn the op-program into a functor and instantiate the library's sweep kernel
 * templates for it with NVRTC (sm_100a, -fmad=false), caching the result per context.  1 (default): specialise when
 * libnvrtc is installed, otherwise (and with 0) the program is interpreted per thread.  Same results either way up to
 * the order of the reduction. */
int fdb_set_jit(fdb_ctx* ctx, int enabled);
int fdb_jit_stats(fdb_ctx* ctx, int64_t* compiled, int64_t* cache_hits, int64_t* failures);
const char* fdb_jit_last_log(fdb_ctx* ctx);   /* NVRTC log of the last failed specialisation ("" if none) */
/* host-only (no GPU, no context): generate and compile the specialised kernel of a program; kind 0 gradient,
 * 1 Jacobian, 2 Hessian, 3 Jacobian of g.(A * x), 4 two-pass array-valued Jacobian, 5 gradient of a structured input.  Returns 0 and the cubin size, or the NVRTC log in `log`. */
int fdb_jit_compile_check(int kind, const int32_t* term_prog, int n_term, const int32_t* epi_prog, int n_epi, const double* consts,
                          int n_consts, int n_acc, int halo, int result_reg, int n_params, int N, int nansafe, int lane_sharing,
                          char* log, int64_t log_capacity, int64_t* cubin_bytes);

/* ---- host-buffer entry points (what ForwardDiff.gradient!(out, f, x, cfg) with
 * host Vectors binds to): upload x, run the sweeps, download the result slice.  grad_host has
 * xlen entries; only the entries of the requested chunks are written. ------------ */
int fdb_gradient_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int N, int64_t chunk_lo,
                      int64_t chunk_hi, double* grad_host, double* value_host);
int fdb_hessian_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int N, int64_t chunk_lo,
                     int64_t chunk_hi, double* H_host, int64_t ldH, double* grad_host, double* value_host);
/* jacobian!(result, f, x, cfg) / jacobian!(result, f!, y, x, cfg) on host arrays (src/jacobian.jl:57-87): J_host is
 * column-major m x xlen (ldJ), only the columns of the requested chunks are written; y_host (may be NULL) receives
 * value.(ydual) when the range includes the last chunk.  A_host (m x xlen, lda) / b_host are the target's parameters
 * (NULL for targets without); reuse_params = 1: A and b are unchanged since the previous host call of the same shape
 * on this context and are not uploaded again. */
int fdb_jacobian_host(fdb_ctx* ctx, int fn, const double* x_host, int64_t xlen, int64_t m, int N, const double* A_host, int64_t lda,
                      const double* b_host, double alpha, int reuse_params, int64_t chunk_lo, int64_t chunk_hi, double* J_host,
                      int64_t ldJ, double* y_host);
/* page-lock a caller-owned host buffer (x, a result) so that the copies of the host entry points are asynchronous and,
 * with several devices, overlap (cudaHostRegister, portable) */
int fdb_host_register(void* ptr, int64_t bytes);
int fdb_host_unregister(void* ptr);

/* ---- one host process, several GPUs ("one context drives all visible devices") ----------------
 * What `ForwardDiff.gradient!(out, f, x, cfg)` (src/gradient.jl:35-44), `jacobian!` (src/jacobian.jl:57-87) and
 * `hessian!` (src/hessian.jl:31-37) called from ONE Julia session bind to.  fdb_multi_init creates one context per
 * device (ndev <= 0: all visible devices, at most FDB_MAX_RANKS; devices == NULL: 0..ndev-1).  Device i of n owns the
 * contiguous chunk range fdb_plan_chunks(xlen, N, i, n) gives.
 *   fdb_multi_*_host: inputs on the host, complete result on the host.  Every device's sweeps are launched before
 *       the host waits for any; each device copies its own result slice (gradient entries / column block) straight
 *       into the caller's buffer, so there is no inter-GPU traffic.
 *   fdb_multi_*_resident: inputs and results live in symmetric exchange arenas (fdb_multi_exchange_create: one arena
 *       per device, linked in-process with cudaDeviceEnablePeerAccess -- no CUDA IPC); arguments are offsets in
 *       doubles into the arena (the same on every device; -1 = not wanted, for optional outputs).  The sweep kernels
 *       store every extracted entry into every peer's arena and a device-side flag barrier follows: on return the
 *       COMPLETE result is resident on every device.  Replicated inputs go in with fdb_multi_upload.
 * Errors: the status of the first failing device; fdb_multi_last_error names the device. */
typedef struct fdb_multi fdb_multi;
int fdb_multi_init(int ndev, const int* devices, fdb_multi** out);
int fdb_multi_shutdown(fdb_multi* g);
int fdb_multi_size(const fdb_multi* g);
fdb_ctx* fdb_multi_ctx(fdb_multi* g, int i);        /* per-device context: fdb_set_nansafe etc., fdb_array_alloc for parameters */
const char* fdb_multi_last_error(fdb_multi* g);     /* g may be NULL: last error of a failed fdb_multi_init */
int fdb_multi_sync(fdb_multi* g);
int fdb_multi_gradient_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int N, double* grad_host, double* value_host);
int fdb_multi_jacobian_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int64_t m, int N, const double* A_host, int64_t lda,
                            const double* b_host, double alpha, int reuse_params, double* J_host, int64_t ldJ, double* y_host);
int fdb_multi_hessian_host(fdb_multi* g, int fn, const double* x_host, int64_t xlen, int N, double* H_host, int64_t ldH, double* grad_host,
                           double* value_host);
struct fdb_exchange;
int fdb_multi_exchange_create(fdb_multi* g, int64_t n_doubles, struct fdb_exchange** xs /* fdb_multi_size(g) entries */);
int fdb_multi_exchange_destroy(fdb_multi* g, struct fdb_exchange** xs);
int fdb_multi_barrier(fdb_multi* g, struct fdb_exchange* const* xs);
int fdb_multi_upload(fdb_multi* g, struct fdb_exchange* const* xs, int64_t offset, const double* host, int64_t count);
int fdb_multi_download(fdb_multi* g, struct fdb_exchange* const* xs, int device_index, int64_t offset, double* host, int64_t count);
int fdb_multi_gradient_resident(fdb_multi* g, struct fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int N, int64_t grad_off,
                                int64_t value_off);
/* A, b: one plain device array per device (allocated on fdb_multi_ctx(g, i)), or NULL */
int fdb_multi_jacobian_resident(fdb_multi* g, struct fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int64_t m, int N,
                                const fdb_array* const* A, int64_t lda, const fdb_array* const* b, double alpha, int64_t J_off, int64_t ldJ,
                                int64_t y_off);
int fdb_multi_hessian_resident(fdb_multi* g, struct fdb_exchange* const* xs, int fn, int64_t x_off, int64_t xlen, int N, int64_t H_off,
                               int64_t ldH, int64_t grad_off, int64_t value_off);

/* ---- multi-GPU result exchange over NVLink peer memory ------------
 * One process per GPU.  The reference's result (`result` of src/gradient.jl:141-152, the Jacobian of
 * src/jacobian.jl:109-118, H of src/hessian.jl:14-71) must end up complete on every rank although each rank
 * only sweeps its own chunk range.  Instead of an all-gather after the sweeps, every rank allocates a
 * symmetric *arena* (same size everywhere), the ranks exchange its 64-byte CUDA IPC handle through whatever
 * transport the host has (torch.distributed here, MPI.jl / NCCL.jl from Julia), and while an exchange is
 * enabled on a context the whole-path drivers store every extracted entry (gradient slice, Jacobian/Hessian
 * column block, y, value) into the local arena AND, from the same kernel, into every peer's arena at the same
 * offset: the extraction is the gather.  All result arrays handed to a driver must then be views into the
 * arena (fdb_exchange_view).  fdb_exchange_barrier is a device-side barrier over flag words in the arenas
 * (enqueued on the context's stream, no host sync): call it after the sweeps (all peers' stores have landed)
 * and before the next sweeps (all peers have consumed the previous result). */
typedef struct fdb_exchange fdb_exchange;
#define FDB_IPC_HANDLE_BYTES 64
#define FDB_MAX_RANKS 8
int fdb_exchange_create(fdb_ctx* ctx, int64_t n_doubles, int rank, int world, fdb_exchange** out);
int fdb_exchange_handle(fdb_exchange* x, unsigned char handle[FDB_IPC_HANDLE_BYTES]);
/* handles: world * FDB_IPC_HANDLE_BYTES bytes in rank order (this rank's own entry is ignored) */
int fdb_exchange_connect(fdb_exchange* x, const unsigned char* handles);
/* the same when all `world` exchanges belong to contexts of THIS process (xs[r] = rank r): peer access is enabled both
 * ways and the arenas are linked directly -- CUDA IPC cannot open a handle in the process that exported it */
int fdb_exchange_connect_local(fdb_exchange* const* xs, int world);
/* a barrier that waits longer than this for a peer gives up (default 20 s), counts a timeout and POISONS the context:
 * fdb_sync, every synchronous call and every download on it then fail -- results since the timeout are incomplete */
int fdb_exchange_set_timeout(fdb_exchange* x, double seconds);
/* level-0 view of n doubles (plane stride ld >= n) at `offset` doubles into the local arena */
int fdb_exchange_view(fdb_exchange* x, int64_t offset, int64_t n, int64_t ld, fdb_array** out);
int fdb_exchange_enable(fdb_ctx* ctx, fdb_exchange* x_or_null);
int fdb_exchange_barrier(fdb_exchange* x);
/* all-gather / broadcast of arena data the drivers did not mirror themselves: copy [offset, offset+count) doubles of the local
 * arena to the same place in every peer's arena (every rank pushing its slice + a barrier = all-gather; one rank = broadcast) */
int fdb_exchange_push(fdb_exchange* x, int64_t offset, int64_t count);
/* number of barrier waits that gave up after the timeout (a peer never arrived); synchronises the stream */
int fdb_exchange_timeouts(fdb_exchange* x, int64_t* count);
int fdb_exchange_destroy(fdb_exchange* x);
/* measurement helper (collective: every rank calls it): all ranks push the first n doubles of their arena into every
 * peer at once; returns this rank's outgoing GB/s.  variant 0: 8-byte stores repeated per peer (what the drivers do),
 * 1: 16-byte stores, 2: one peer per thread block.  rotate != 0: rank r visits peers r+1, r+2, ... */
int fdb_exchange_measure_store_bw(fdb_exchange* x, int64_t n_doubles, int variant, int rotate, double* gb_per_s);

/* ---- NVSwitch multicast form of the exchange (csrc/fdb_multicast.cu) --------------------------------------------------
 * With fdb_exchange_create the drivers repeat each result store once per peer (egress = (W-1) x the rank's slice).  NVSwitch
 * can replicate a write itself (NVLS): the ranks bind their arenas into one multicast object, and a store through its mapping
 * lands at the same offset of EVERY rank's arena -- each entry leaves the GPU once.  Such arenas are cuMemCreate allocations
 * shared as POSIX file descriptors; the handle that travels through the host transport is {pid, fd} (16 bytes) and the
 * importer duplicates the descriptor with pidfd_getfd(2).  Collective sequence:
 *   fdb_exchange_create_multicast;  fdb_exchange_share_handle -> all-gather;  rank 0: fdb_exchange_multicast_create -> broadcast;
 *   fdb_exchange_connect_multicast;  HOST BARRIER;  fdb_exchange_multicast_bind;  HOST BARRIER;  then as any exchange
 * (fdb_exchange_view / enable / barrier / push / destroy).  FDB_ERR_UNSUPPORTED when the device, driver or kernel (pidfd_getfd)
 * cannot do it: fall back to fdb_exchange_create. */
#define FDB_SHARE_HANDLE_BYTES 16
int fdb_multicast_supported(fdb_ctx* ctx, int* supported);
int fdb_exchange_create_multicast(fdb_ctx* ctx, int64_t n_doubles, int rank, int world, fdb_exchange** out);
int fdb_exchange_share_handle(fdb_exchange* x, unsigned char handle[FDB_SHARE_HANDLE_BYTES]);
int fdb_exchange_multicast_create(fdb_exchange* x_rank0, unsigned char handle[FDB_SHARE_HANDLE_BYTES]);
/* arena_handles: world * FDB_SHARE_HANDLE_BYTES bytes in rank order; mc_handle: rank 0's fdb_exchange_multicast_create output */
int fdb_exchange_connect_multicast(fdb_exchange* x, const unsigned char* arena_handles, const unsigned char* mc_handle);
int fdb_exchange_multicast_bind(fdb_exchange* x);
/* on = 0: keep the multicast mapping but store per peer again (comparison runs); on = 1: back to multicast */
int fdb_exchange_set_multicast(fdb_exchange* x, int on);

/* ---- measurement helpers (bench.py) ---------------------------------------- */
/* FP64 instruction-issue peak: a register-only chain of independent DMUL/DADD pairs;
 * returns achieved non-fused FP64 operations per second on this device. */
int fdb_measure_fp64_peak(fdb_ctx* ctx, double* dp_ops_per_s, double* dfma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* FDB200_H */
