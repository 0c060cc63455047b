// fdb_core.cu -- lifecycle, device arrays, AoS<->SoA conversion, host-side chunk
// planning, and the bit-exact sweep utilities (seed!/seed_zero_partials!/extract_*).
//
// Reference behaviour restated here (paths under /root/reference):
//   src/apiutils.jl:39-143   construct_seeds, structural_eachindex, seed!, seed_zero_partials!
//   src/gradient.jl:66-81    extract_gradient!, extract_gradient_chunk!
//   src/jacobian.jl:95-118   extract_jacobian!, extract_jacobian_chunk!
//   src/prelude.jl:15-36     structural_length, pickchunksize
//   src/gradient.jl:121-125  chunk bookkeeping
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "fdb_internal.h"
#include "fdb_dual.cuh"

namespace fdb {
thread_local std::string g_init_error;

int ensure_scratch(fdb_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return FDB_OK;
    if (ctx->scratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    size_t want = bytes + bytes / 4 + 4096;
    FDB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return FDB_OK;
}
int ensure_counter(fdb_ctx* ctx) {
    if (ctx->counter) return FDB_OK;
    FDB_CUDA(ctx, cudaMalloc(&ctx->counter, 256));
    FDB_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, 256, ctx->stream));
    return FDB_OK;
}
int ensure_pinned(fdb_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return FDB_OK;
    if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
    FDB_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return FDB_OK;
}

// 0-based storage index of structural position `pos` (src/apiutils.jl:44-71): fdb_dual.cuh, shared with the sweep kernels
__host__ __device__ inline int64_t structural_index(int structure, int64_t rows, int64_t pos) { return (int64_t)struct_index_of(structure, (long)rows, (long)pos); }

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// AoS (host wire order) <-> SoA planes.  One thread per (element, plane).
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, int64_t n, int64_t ld, int P) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * P) return;
    int64_t e = t / P; int q = (int)(t - e * P);
    soa[(int64_t)q * ld + e] = aos[t];
}
__global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, int64_t n, int64_t ld, int P) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * P) return;
    int64_t e = t / P; int q = (int)(t - e * P);
    aos[t] = soa[(int64_t)q * ld + e];
}

// seed!(duals, x, index, seeds, chunksize): slot i (1-based) of the window gets
// Dual(x[idx], single_seed(i)).  PB = planes of the value type (1, or N+1 when the
// value type is itself a Dual); outer plane a in 0..N.
//   plane(0, b)  <- x plane b
//   plane(a, b)  <- (a == i && b == 0) ? 1 : 0        a >= 1
__global__ void seed_kernel(double* __restrict__ d, int64_t ldd, const double* __restrict__ x, int64_t ldx, int N, int PB,
                            int structure, int64_t rows, int64_t nstruct, int64_t index, int64_t chunksize) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int P = (N + 1) * PB;
    if (t >= chunksize * P) return;
    int64_t i0 = t / P; int q = (int)(t - i0 * P);        // i0 = slot-1
    int64_t pos = index - 1 + i0;
    if (pos >= nstruct) return;                            // zip() stops at the shorter iterator
    int64_t idx = structural_index(structure, rows, pos);
    int a = q / PB, b = q - a * PB;
    double val = (a == 0) ? x[(int64_t)b * ldx + idx] : ((a == (int)i0 + 1 && b == 0) ? 1.0 : 0.0);
    d[(int64_t)q * ldd + idx] = val;
}
// seed_zero_partials!(duals, x, index, count): Dual(x[idx], zero(Partials))
__global__ void seed_zero_kernel(double* __restrict__ d, int64_t ldd, const double* __restrict__ x, int64_t ldx, int N, int PB,
                                 int structure, int64_t rows, int64_t nstruct, int64_t index, int64_t count) {
    int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // element within the window
    int q = blockIdx.y;                                              // plane
    if (i0 >= count) return;
    int64_t pos = index - 1 + i0;
    if (pos >= nstruct) return;                                      // Iterators.take clamps
    int64_t idx = structural_index(structure, rows, pos);
    int a = q / PB, b = q - a * PB;
    d[(int64_t)q * ldd + idx] = (a == 0) ? x[(int64_t)b * ldx + idx] : 0.0;
}
// dense windows of many elements (the O(n) calls: a fresh buffer, chunk mode's first-chunk tail):
// one thread owns two consecutive elements and writes all planes with 16-byte stores.
__global__ void __launch_bounds__(256) seed_zero_dense_kernel(double* __restrict__ d, int64_t ldd, const double* __restrict__ x, int64_t ldx,
                                                              int P, int PB, int64_t first, int64_t count, bool vec) {
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (e >= count) return;
    const bool pair = e + 1 < count;
    const int64_t idx = first + e;
    if (pair && vec) {
        for (int b = 0; b < PB; b++)
            *reinterpret_cast<double2*>(d + (int64_t)b * ldd + idx) = __ldg(reinterpret_cast<const double2*>(x + (int64_t)b * ldx + idx));
        const double2 z = make_double2(0.0, 0.0);
        for (int q = PB; q < P; q++) *reinterpret_cast<double2*>(d + (int64_t)q * ldd + idx) = z;
    } else {
        for (int b = 0; b < PB; b++) { d[(int64_t)b * ldd + idx] = x[(int64_t)b * ldx + idx]; if (pair) d[(int64_t)b * ldd + idx + 1] = x[(int64_t)b * ldx + idx + 1]; }
        for (int q = PB; q < P; q++) { d[(int64_t)q * ldd + idx] = 0.0; if (pair) d[(int64_t)q * ldd + idx + 1] = 0.0; }
    }
}
// extract_gradient_chunk!: result[idx(index-1+k)] = partials(ydual, k+1), k = 0..chunksize-1.
// result has PB planes (PB > 1 when the partials are themselves Duals).
__global__ void extract_gradient_chunk_kernel(double* __restrict__ res, int64_t ldr, const double* __restrict__ y, int64_t ldy,
                                              int PB, int structure, int64_t rows, int64_t nstruct, int64_t index, int64_t chunksize) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunksize * PB) return;
    int64_t k = t / PB; int b = (int)(t - k * PB);
    int64_t pos = index - 1 + k;
    if (pos >= nstruct) return;
    res[(int64_t)b * ldr + structural_index(structure, rows, pos)] = y[((k + 1) * PB + b) * ldy];   // structural_eachindex(result), gradient.jl:76
}
// extract_jacobian_chunk!: result[r, index-1+k] = partials(ydual[r], k+1)
__global__ void extract_jacobian_chunk_kernel(double* __restrict__ res, int64_t ldplane, int64_t ldres,
                                              const double* __restrict__ y, int64_t ldy, int PB, int64_t m,
                                              int64_t index, int64_t chunksize) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t k = blockIdx.y;
    int b = blockIdx.z;
    if (r >= m || k >= chunksize) return;
    res[(int64_t)b * ldplane + r + (index - 1 + k) * ldres] = y[((k + 1) * PB + b) * ldy + r];
}
__global__ void extract_value_kernel(double* __restrict__ out, int64_t ldo, const double* __restrict__ y, int64_t ldy, int PB, int64_t m) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (r >= m) return;
    out[(int64_t)b * ldo + r] = y[(int64_t)b * ldy + r];
}

}  // namespace fdb

using namespace fdb;

// ===========================================================================
// lifecycle
// ===========================================================================
extern "C" const char* fdb_version(void) { return "fdb200 0.1 (sm_100a)"; }

extern "C" int fdb_device_count(int* count) {
    if (!count) return FDB_ERR_BADARG;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; return fail(nullptr, FDB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = c;
    return FDB_OK;
}

extern "C" int fdb_init(int device, fdb_ctx** out) {
    if (!out) return FDB_ERR_BADARG;
    *out = nullptr;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess || c == 0)
        return fail(nullptr, FDB_ERR_CUDA, "no CUDA device available (%s); fdb200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= c) return fail(nullptr, FDB_ERR_BADARG, "device %d out of range (0..%d)", device, c - 1);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, FDB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    fdb_ctx* ctx = new fdb_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, FDB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return FDB_OK;
}

extern "C" int fdb_shutdown(fdb_ctx* ctx) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    jit_release(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->hostws) cudaFree(ctx->hostws);
    if (ctx->counter) cudaFree(ctx->counter);
    if (ctx->cursor) cudaFree(ctx->cursor);
    if (ctx->xch_timeout_flag) cudaFreeHost(ctx->xch_timeout_flag);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FDB_OK;
}

extern "C" const char* fdb_last_error(fdb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int fdb_sync(fdb_ctx* ctx) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, "fdb_sync");
}
extern "C" int fdb_set_async(fdb_ctx* ctx, int async) { if (!ctx) return FDB_ERR_BADARG; ctx->async = async != 0; return FDB_OK; }
extern "C" int fdb_set_nansafe(fdb_ctx* ctx, int en) { if (!ctx) return FDB_ERR_BADARG; ctx->nansafe = en != 0; return FDB_OK; }
extern "C" int fdb_set_lane_sharing(fdb_ctx* ctx, int en) { if (!ctx) return FDB_ERR_BADARG; ctx->lane_sharing = en != 0; return FDB_OK; }
extern "C" int fdb_set_chunk_grouping(fdb_ctx* ctx, int group) {
    if (!ctx || group < 1 || group > 64) return fail(ctx, FDB_ERR_BADARG, "fdb_set_chunk_grouping: group must be 1..64");
    ctx->grad_group = group;
    return FDB_OK;
}
extern "C" int fdb_set_dmma(fdb_ctx* ctx, int en) { if (!ctx) return FDB_ERR_BADARG; ctx->use_dmma = en != 0; return FDB_OK; }
extern "C" int fdb_set_dmma_tma_store(fdb_ctx* ctx, int mode) { if (!ctx || mode < 0 || mode > 3) return FDB_ERR_BADARG; ctx->dmma_tma_store = mode; return FDB_OK; }
extern "C" int fdb_set_dmma_tile(fdb_ctx* ctx, int bn) {
    if (!ctx || !(bn == 0 || bn == 16 || bn == 32 || bn == 64)) return fail(ctx, FDB_ERR_BADARG, "fdb_set_dmma_tile: 0 (auto), 16, 32 or 64");
    ctx->dmma_bn = bn; return FDB_OK;
}
extern "C" int fdb_set_stream(fdb_ctx* ctx, void* s) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    cudaStreamSynchronize(ctx->stream);
    if (s == nullptr) {
        if (!ctx->own_stream) { FDB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    } else {
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
        ctx->stream = (cudaStream_t)s; ctx->own_stream = false;
    }
    return FDB_OK;
}
extern "C" int fdb_get_stream(fdb_ctx* ctx, void** s) { if (!ctx || !s) return FDB_ERR_BADARG; *s = (void*)ctx->stream; return FDB_OK; }
extern "C" int fdb_kernel_launches(fdb_ctx* ctx, int64_t* count) { if (!ctx || !count) return FDB_ERR_BADARG; *count = ctx->launches; return FDB_OK; }

// ===========================================================================
// memory
// ===========================================================================
static int check_level(fdb_ctx* ctx, int N, int level) {
    if (level < 0 || level > 2) return fail(ctx, FDB_ERR_BADARG, "level %d outside 0..2", level);
    if (level > 0 && (N < 1 || N > FDB_MAX_CHUNK)) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    return FDB_OK;
}

extern "C" int fdb_array_alloc(fdb_ctx* ctx, int64_t n, int N, int level, fdb_array** out) {
    FDB_ENTER(ctx);
    if (!ctx || !out || n < 0) return fail(ctx, FDB_ERR_BADARG, "fdb_array_alloc: bad argument");
    int rc = check_level(ctx, N, level); if (rc) return rc;
    fdb_array* a = new fdb_array();
    a->ctx = ctx; a->n = n; a->N = level == 0 ? 0 : N; a->level = level; a->owned = true;
    a->ld = round_up(n > 0 ? n : 1, 32);            // 256-byte aligned planes: vector loads never straddle
    size_t bytes = (size_t)a->ld * a->planes() * sizeof(double);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(&a->data, bytes);
    if (e != cudaSuccess) { delete a; return fail(ctx, e == cudaErrorMemoryAllocation ? FDB_ERR_OOM : FDB_ERR_CUDA,
                                                  "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
    *out = a;
    return FDB_OK;
}
extern "C" int fdb_array_wrap(fdb_ctx* ctx, double* ptr, int64_t n, int64_t ld, int N, int level, fdb_array** out) {
    if (!ctx || !out || !ptr || n < 0 || ld < n) return fail(ctx, FDB_ERR_BADARG, "fdb_array_wrap: bad argument");
    int rc = check_level(ctx, N, level); if (rc) return rc;
    fdb_array* a = new fdb_array();
    a->ctx = ctx; a->data = ptr; a->n = n; a->ld = ld; a->N = level == 0 ? 0 : N; a->level = level; a->owned = false;
    *out = a;
    return FDB_OK;
}
extern "C" int fdb_array_free(fdb_array* a) {
    FDB_ENTER(a ? a->ctx : nullptr);
    if (!a) return FDB_ERR_BADARG;
    if (a->ctx->capturing && a->owned)     // freeing synchronises the stream, which is illegal during graph capture
        return fail(a->ctx, FDB_ERR_BADARG, "fdb_array_free: arrays allocated while a graph is captured must outlive the graph");
    if (a->owned && a->data) { cudaStreamSynchronize(a->ctx->stream); cudaFree(a->data); }
    delete a;
    return FDB_OK;
}
extern "C" int fdb_array_info(const fdb_array* a, int64_t* n, int* N, int* level, int64_t* ld, double** ptr) {
    if (!a) return FDB_ERR_BADARG;
    if (n) *n = a->n; if (N) *N = a->N; if (level) *level = a->level; if (ld) *ld = a->ld; if (ptr) *ptr = a->data;
    return FDB_OK;
}
extern "C" int fdb_upload_values(fdb_array* a, const double* host) {
    FDB_ENTER(a ? a->ctx : nullptr);
    if (!a || !host) return FDB_ERR_BADARG;
    fdb_ctx* ctx = a->ctx;
    FDB_CUDA(ctx, cudaMemcpyAsync(a->data, host, (size_t)a->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}
extern "C" int fdb_download_values(const fdb_array* a, double* host) {
    FDB_ENTER(a ? a->ctx : nullptr);
    if (!a || !host) return FDB_ERR_BADARG;
    fdb_ctx* ctx = a->ctx;
    FDB_CUDA(ctx, cudaMemcpyAsync(host, a->data, (size_t)a->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, "fdb_download_values");
}
extern "C" int fdb_upload_duals_aos(fdb_array* a, const double* host) {
    FDB_ENTER(a ? a->ctx : nullptr);
    if (!a || !host) return FDB_ERR_BADARG;
    fdb_ctx* ctx = a->ctx; int P = a->planes();
    if (a->n == 0) return FDB_OK;
    size_t bytes = (size_t)a->n * P * sizeof(double);
    int rc = ensure_scratch(ctx, bytes); if (rc) return rc;
    FDB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int64_t total = a->n * P;
    aos_to_soa_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(ctx->scratch, a->data, a->n, a->ld, P);
    rc = finish(ctx, 1, "aos_to_soa"); if (rc) return rc;
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}
extern "C" int fdb_download_duals_aos(const fdb_array* a, double* host) {
    FDB_ENTER(a ? a->ctx : nullptr);
    if (!a || !host) return FDB_ERR_BADARG;
    fdb_ctx* ctx = a->ctx; int P = a->planes();
    if (a->n == 0) return FDB_OK;
    size_t bytes = (size_t)a->n * P * sizeof(double);
    int rc = ensure_scratch(ctx, bytes); if (rc) return rc;
    int64_t total = a->n * P;
    soa_to_aos_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(a->data, ctx->scratch, a->n, a->ld, P);
    rc = finish(ctx, 1, "soa_to_aos"); if (rc) return rc;
    FDB_CUDA(ctx, cudaMemcpyAsync(host, ctx->scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_poison(ctx, "fdb_download_duals_aos");
}
extern "C" int fdb_copy(fdb_array* dst, const fdb_array* src) {
    FDB_ENTER(dst ? dst->ctx : nullptr);
    if (!dst || !src) return FDB_ERR_BADARG;
    fdb_ctx* ctx = dst->ctx;
    if (dst->n != src->n || dst->planes() != src->planes()) return fail(ctx, FDB_ERR_SHAPE, "fdb_copy: shape mismatch");
    FDB_CUDA(ctx, cudaMemcpy2DAsync(dst->data, (size_t)dst->ld * 8, src->data, (size_t)src->ld * 8, (size_t)src->n * 8,
                                    src->planes(), cudaMemcpyDeviceToDevice, ctx->stream));
    if (!ctx->async) FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}

// ===========================================================================
// host-side planning
// ===========================================================================
extern "C" int64_t fdb_pickchunksize(int64_t input_length, int64_t threshold) {
    // src/prelude.jl:29-36: round(Int, n / threshold, RoundUp) on Float64 quotients
    if (input_length <= threshold) return input_length;
    int64_t nchunks = (int64_t)ceil((double)input_length / (double)threshold);
    return (int64_t)ceil((double)input_length / (double)nchunks);
}
extern "C" int64_t fdb_structural_length(int structure, int64_t rows) {
    if (structure == FDB_UPPER || structure == FDB_LOWER) return (rows * (rows + 1)) >> 1;
    return rows;
}
extern "C" int64_t fdb_structural_index(int structure, int64_t rows, int64_t pos) {
    return structural_index(structure, rows, pos);
}
extern "C" int fdb_plan_chunks(int64_t xlen, int N, int rank, int world, fdb_chunk_plan* out) {
    if (!out || N < 1 || world < 1 || rank < 0 || rank >= world) return FDB_ERR_BADARG;
    if (xlen < N) return FDB_ERR_CHUNK;
    Plan p = make_plan(xlen, N);
    out->xlen = xlen; out->remainder = p.remainder; out->lastchunksize = p.lastchunksize;
    out->lastchunkindex = p.lastchunkindex; out->middle_lo = 2; out->middle_hi = p.middle_hi;
    out->nchunks = (xlen == N) ? 1 : p.nchunks;      // N == xlen is vector mode: a single sweep
    int64_t per = cdiv(out->nchunks, world);
    int64_t lo = (int64_t)rank * per, hi = lo + per;
    if (lo > out->nchunks) lo = out->nchunks;
    if (hi > out->nchunks) hi = out->nchunks;
    out->chunk_lo = lo; out->chunk_hi = hi;
    out->elem_lo = lo * N;
    out->elem_hi = hi * N < xlen ? hi * N : xlen;
    if (out->elem_lo > xlen) out->elem_lo = xlen;
    return FDB_OK;
}

// ===========================================================================
// sweep utilities
// ===========================================================================
static int check_seed_args(fdb_ctx* ctx, const fdb_array* d, const fdb_array* x, const char* what) {
    if (!ctx || !d || !x) return fail(ctx, FDB_ERR_BADARG, "%s: null argument", what);
    if (d->level < 1 || x->level != d->level - 1) return fail(ctx, FDB_ERR_BADARG, "%s: duals must be one dual level above x", what);
    if (x->level == 1 && x->N != d->N) return fail(ctx, FDB_ERR_BADARG, "%s: nested chunk sizes differ", what);
    if (d->n != x->n) return fail(ctx, FDB_ERR_SHAPE, "%s: DimensionMismatch (duals %lld, x %lld)", what, (long long)d->n, (long long)x->n);
    return FDB_OK;
}

static int seed_impl(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int structure, int64_t rows, int64_t nstruct,
                     int64_t index, int64_t chunksize) {
    int rc = check_seed_args(ctx, d, x, "fdb_seed"); if (rc) return rc;
    if (index <= 0) { index = 1; chunksize = d->N; }
    if (chunksize > d->N) return fail(ctx, FDB_ERR_BADARG, "fdb_seed: chunksize %lld > N=%d", (long long)chunksize, d->N);
    if (chunksize <= 0 || index > nstruct) return FDB_OK;
    int PB = x->planes(); int P = (d->N + 1) * PB;
    int64_t total = chunksize * P;
    seed_kernel<<<(unsigned)cdiv(total, 128), 128, 0, ctx->stream>>>(d->data, d->ld, x->data, x->ld, d->N, PB, structure, rows,
                                                                    nstruct, index, chunksize);
    return finish(ctx, 1, "fdb_seed");
}
static int seed_zero_impl(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int structure, int64_t rows, int64_t nstruct,
                          int64_t index, int64_t count) {
    int rc = check_seed_args(ctx, d, x, "fdb_seed_zero"); if (rc) return rc;
    if (index <= 0) { index = 1; count = nstruct; }
    if (count <= 0 || index > nstruct) return FDB_OK;     // zero-width window is a no-op (test/SeedTest.jl:74-79)
    if (index - 1 + count > nstruct) count = nstruct - (index - 1);
    int PB = x->planes(); int P = (d->N + 1) * PB;
    if (structure == FDB_DENSE && count >= 1024) {
        const int64_t first = index - 1;
        const bool vec = ((((uintptr_t)(d->data + first)) & 15) == 0) && ((((uintptr_t)(x->data + first)) & 15) == 0) &&
                         ((d->ld & 1) == 0) && ((x->ld & 1) == 0);
        seed_zero_dense_kernel<<<(unsigned)cdiv(cdiv(count, 2), 256), 256, 0, ctx->stream>>>(d->data, d->ld, x->data, x->ld, P, PB, first, count, vec);
        return finish(ctx, 1, "fdb_seed_zero");
    }
    dim3 grid((unsigned)cdiv(count, 256), (unsigned)P);
    seed_zero_kernel<<<grid, 256, 0, ctx->stream>>>(d->data, d->ld, x->data, x->ld, d->N, PB, structure, rows, nstruct, index, count);
    return finish(ctx, 1, "fdb_seed_zero");
}

extern "C" int fdb_seed(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int64_t index, int64_t chunksize) {
    FDB_ENTER(ctx);
    return seed_impl(ctx, d, x, FDB_DENSE, d ? d->n : 0, d ? d->n : 0, index, chunksize);
}
extern "C" int fdb_seed_zero(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int64_t index, int64_t count) {
    FDB_ENTER(ctx);
    return seed_zero_impl(ctx, d, x, FDB_DENSE, d ? d->n : 0, d ? d->n : 0, index, count);
}
static int check_structure(fdb_ctx* ctx, const fdb_array* d, int structure, int64_t rows) {
    if (structure < 0 || structure > 3) return fail(ctx, FDB_ERR_BADARG, "unknown structure %d", structure);
    if (structure != FDB_DENSE && d && d->n != rows * rows)
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: structured array needs rows*rows storage");
    return FDB_OK;
}
extern "C" int fdb_seed_structured(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int structure, int64_t rows,
                                   int64_t index, int64_t chunksize) {
    FDB_ENTER(ctx);
    int rc = check_structure(ctx, d, structure, rows); if (rc) return rc;
    return seed_impl(ctx, d, x, structure, rows, fdb_structural_length(structure, rows), index, chunksize);
}
extern "C" int fdb_seed_zero_structured(fdb_ctx* ctx, fdb_array* d, const fdb_array* x, int structure, int64_t rows,
                                        int64_t index, int64_t count) {
    FDB_ENTER(ctx);
    int rc = check_structure(ctx, d, structure, rows); if (rc) return rc;
    return seed_zero_impl(ctx, d, x, structure, rows, fdb_structural_length(structure, rows), index, count);
}

extern "C" int fdb_extract_gradient_chunk_structured(fdb_ctx* ctx, fdb_array* res, const fdb_array* y, int structure, int64_t rows,
                                                     int64_t index, int64_t chunksize) {
    FDB_ENTER(ctx);
    if (!ctx || !res || !y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient_chunk: null argument");
    if (y->level < 1 || res->level != y->level - 1) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient_chunk: result must be one level below ydual");
    if (y->n != 1)   // GRAD_ERROR, src/gradient.jl:88-91
        return fail(ctx, FDB_ERR_SHAPE, "gradient(f, x) expects that f(x) is a real number. Perhaps you meant jacobian(f, x)?");
    if (structure < 0 || structure > 3) return fail(ctx, FDB_ERR_BADARG, "unknown structure %d", structure);
    if (structure != FDB_DENSE && res->n != rows * rows) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: structured result needs rows*rows storage");
    if (chunksize > y->N) return fail(ctx, FDB_ERR_BADARG, "chunksize > N");
    if (index < 1) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient_chunk: index %lld < 1 (1-based structural position)", (long long)index);
    if (res->level == 1 && res->N != y->N)     // nested: the partials are Dual{T,V,N} themselves and must carry the same N (config.jl:198-201)
        return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient_chunk: nested chunk sizes differ (result N=%d, ydual N=%d)", res->N, y->N);
    if (chunksize <= 0) return FDB_OK;
    int PB = res->planes();
    int64_t total = chunksize * PB;
    const int64_t nstruct = structure == FDB_DENSE ? res->n : fdb_structural_length(structure, rows);
    extract_gradient_chunk_kernel<<<(unsigned)cdiv(total, 128), 128, 0, ctx->stream>>>(res->data, res->ld, y->data, y->ld, PB, structure, rows,
                                                                                      nstruct, index, chunksize);
    return finish(ctx, 1, "fdb_extract_gradient_chunk");
}
extern "C" int fdb_extract_gradient_chunk(fdb_ctx* ctx, fdb_array* res, const fdb_array* y, int64_t index, int64_t chunksize) {
    FDB_ENTER(ctx);
    return fdb_extract_gradient_chunk_structured(ctx, res, y, FDB_DENSE, res ? res->n : 0, index, chunksize);
}
extern "C" int fdb_fill(fdb_ctx* ctx, fdb_array* a, double value) {
    FDB_ENTER(ctx);
    // fill!(a, value) on every plane (used to zero the off-structure entries of structured work buffers)
    if (!ctx || !a) return fail(ctx, FDB_ERR_BADARG, "fdb_fill: null argument");
    if (value != 0.0) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_fill: only zero fill is implemented");
    if (a->n == 0) return FDB_OK;
    FDB_CUDA(ctx, cudaMemset2DAsync(a->data, (size_t)a->ld * sizeof(double), 0, (size_t)a->n * sizeof(double), (size_t)a->planes(), ctx->stream));
    if (!ctx->async) FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}
extern "C" int fdb_extract_gradient(fdb_ctx* ctx, fdb_array* res, const fdb_array* y) {
    FDB_ENTER(ctx);
    if (!y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient: null argument");
    return fdb_extract_gradient_chunk(ctx, res, y, 1, y->N);   // zip(1:npartials, idxs), src/gradient.jl:66-72
}
extern "C" int fdb_extract_jacobian_chunk(fdb_ctx* ctx, fdb_array* res, int64_t ldres, const fdb_array* y, int64_t index,
                                          int64_t chunksize) {
    FDB_ENTER(ctx);
    if (!ctx || !res || !y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_jacobian_chunk: null argument");
    if (y->level < 1 || res->level != y->level - 1) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_jacobian_chunk: result must be one level below ydual");
    if (chunksize > y->N || index < 1) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_jacobian_chunk: bad window");
    if (ldres < y->n || (index - 1 + chunksize) * ldres > res->n)
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Jacobian result too small for columns %lld..%lld",
                    (long long)index, (long long)(index - 1 + chunksize));
    if (chunksize <= 0 || y->n == 0) return FDB_OK;
    dim3 grid((unsigned)cdiv(y->n, 256), (unsigned)chunksize, (unsigned)res->planes());
    extract_jacobian_chunk_kernel<<<grid, 256, 0, ctx->stream>>>(res->data, res->ld, ldres, y->data, y->ld, res->planes(), y->n,
                                                                  index, chunksize);
    return finish(ctx, 1, "fdb_extract_jacobian_chunk");
}
extern "C" int fdb_extract_jacobian(fdb_ctx* ctx, fdb_array* res, int64_t ldres, const fdb_array* y, int64_t n) {
    FDB_ENTER(ctx);
    return fdb_extract_jacobian_chunk(ctx, res, ldres, y, 1, n);
}
extern "C" int fdb_extract_value(fdb_ctx* ctx, fdb_array* out, const fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !out || !y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_value: null argument");
    if (y->level < 1 || out->level != y->level - 1 || out->n != y->n) return fail(ctx, FDB_ERR_SHAPE, "fdb_extract_value: shape mismatch");
    if (y->n == 0) return FDB_OK;
    dim3 grid((unsigned)cdiv(y->n, 256), (unsigned)out->planes());
    extract_value_kernel<<<grid, 256, 0, ctx->stream>>>(out->data, out->ld, y->data, y->ld, out->planes(), y->n);
    return finish(ctx, 1, "fdb_extract_value");
}
