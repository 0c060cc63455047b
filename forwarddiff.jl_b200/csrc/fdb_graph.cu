// fdb_graph.cu -- the reference chunk loop as a replayed CUDA graph.
//
// Path (A) (the reference loop of src/gradient.jl:141-147 / src/jacobian.jl:199-205 on materialised device
// duals) is launch-bound: every middle chunk is the same short sequence of small kernels
//     seed!(xdual, x, i, seeds) -> f(xdual) -> extract_*_chunk!(result, ydual, i, N) -> seed_zero_partials!(xdual, x, i)
// in which only the window start `i` changes.  Here the window lives in device memory (a "chunk cursor" owned by
// the context), the sweep utilities have `_at_cursor` variants that read it, and a one-thread kernel advances
// it (i += N; chunksize = min(N, xlen - i + 1), which also produces the short last chunk of
// src/gradient.jl:121-125).  One iteration therefore has launch parameters that never change: it is captured
// ONCE into a CUDA graph (stream capture, relaxed mode) and replayed for the remaining chunks with one
// cudaGraphLaunch each instead of ~10 kernel launches driven from the host language.
#include "fdb_internal.h"

namespace fdb {

static int ensure_cursor(fdb_ctx* ctx) {
    if (ctx->cursor) return FDB_OK;
    FDB_CUDA(ctx, cudaMalloc(&ctx->cursor, 4 * sizeof(int64_t)));
    FDB_CUDA(ctx, cudaMemsetAsync(ctx->cursor, 0, 4 * sizeof(int64_t), ctx->stream));
    return FDB_OK;
}

// cursor = {index (1-based structural position), chunksize}
__global__ void cursor_set_kernel(int64_t* cur, int64_t index, int64_t chunksize) { cur[0] = index; cur[1] = chunksize; }
__global__ void cursor_advance_kernel(int64_t* cur, int64_t step, int64_t xlen) {
    const int64_t i = cur[0] + step;
    int64_t cs = xlen - i + 1;
    if (cs > step) cs = step;
    if (cs < 0) cs = 0;
    cur[0] = i; cur[1] = cs;
}
// seed!(duals, x, cursor.index, seeds, cursor.chunksize)   (src/apiutils.jl:125-143; dense indices)
__global__ void seed_at_cursor_kernel(double* __restrict__ d, int64_t ldd, const double* __restrict__ x, int64_t ldx, int N, int PB,
                                      int64_t n, const int64_t* __restrict__ cur) {
    const int64_t index = cur[0], chunksize = cur[1];
    const int P = (N + 1) * PB;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunksize * P) return;
    const int64_t i0 = t / P; const int q = t - (int)i0 * P;
    const int64_t idx = index - 1 + i0;
    if (idx >= n) return;
    const int a = q / PB, b = q - a * PB;
    d[(int64_t)q * ldd + idx] = (a == 0) ? x[(int64_t)b * ldx + idx] : ((a == (int)i0 + 1 && b == 0) ? 1.0 : 0.0);
}
// seed_zero_partials!(duals, x, cursor.index, cursor.chunksize)   (src/apiutils.jl:83-105)
__global__ void seed_zero_at_cursor_kernel(double* __restrict__ d, int64_t ldd, const double* __restrict__ x, int64_t ldx, int N, int PB,
                                           int64_t n, const int64_t* __restrict__ cur) {
    const int64_t index = cur[0], count = cur[1];
    const int P = (N + 1) * PB;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * P) return;
    const int64_t i0 = t / P; const int q = t - (int)i0 * P;
    const int64_t idx = index - 1 + i0;
    if (idx >= n) return;
    const int a = q / PB, b = q - a * PB;
    d[(int64_t)q * ldd + idx] = (a == 0) ? x[(int64_t)b * ldx + idx] : 0.0;
}
// extract_gradient_chunk!(T, result, ydual, cursor.index, cursor.chunksize)   (src/gradient.jl:74-81)
__global__ void extract_gradient_at_cursor_kernel(double* __restrict__ res, int64_t ldr, const double* __restrict__ y, int64_t ldy, int PB,
                                                  int64_t n, const int64_t* __restrict__ cur) {
    const int64_t index = cur[0], chunksize = cur[1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunksize * PB) return;
    const int64_t k = t / PB; const int b = t - (int)k * PB;
    const int64_t pos = index - 1 + k;
    if (pos >= n) return;
    res[(int64_t)b * ldr + pos] = y[((k + 1) * PB + b) * ldy];
}
// extract_jacobian_chunk!(T, result, ydual, cursor.index, cursor.chunksize)   (src/jacobian.jl:109-118)
__global__ void extract_jacobian_at_cursor_kernel(double* __restrict__ res, int64_t ldplane, int64_t ldres, const double* __restrict__ y,
                                                  int64_t ldy, int PB, int64_t m, const int64_t* __restrict__ cur) {
    const int64_t index = cur[0], chunksize = cur[1];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = blockIdx.y; const int b = blockIdx.z;
    if (r >= m || k >= chunksize) return;
    res[(int64_t)b * ldplane + r + (index - 1 + k) * ldres] = y[((k + 1) * PB + b) * ldy + r];
}

static int check_pair(fdb_ctx* ctx, const fdb_array* d, const fdb_array* x, const char* what) {
    if (!ctx || !d || !x) return fail(ctx, FDB_ERR_BADARG, "%s: null argument", what);
    if (d->level < 1 || x->level != d->level - 1) return fail(ctx, FDB_ERR_BADARG, "%s: duals must be one dual level above x", what);
    if (d->n != x->n) return fail(ctx, FDB_ERR_SHAPE, "%s: DimensionMismatch (duals %lld, x %lld)", what, (long long)d->n, (long long)x->n);
    return ensure_cursor(ctx);
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_cursor_set(fdb_ctx* ctx, int64_t index, int64_t chunksize) {
    FDB_ENTER(ctx);
    if (!ctx || index < 1 || chunksize < 0) return fail(ctx, FDB_ERR_BADARG, "fdb_cursor_set: bad argument");
    int rc = ensure_cursor(ctx); if (rc) return rc;
    cursor_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->cursor, index, chunksize);
    return finish(ctx, 1, "fdb_cursor_set");
}
extern "C" int fdb_cursor_advance(fdb_ctx* ctx, int64_t step, int64_t xlen) {
    FDB_ENTER(ctx);
    if (!ctx || step < 1) return fail(ctx, FDB_ERR_BADARG, "fdb_cursor_advance: bad argument");
    int rc = ensure_cursor(ctx); if (rc) return rc;
    cursor_advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->cursor, step, xlen);
    return finish(ctx, 1, "fdb_cursor_advance");
}
extern "C" int fdb_cursor_get(fdb_ctx* ctx, int64_t* index, int64_t* chunksize) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    int rc = ensure_cursor(ctx); if (rc) return rc;
    int64_t h[2];
    FDB_CUDA(ctx, cudaMemcpyAsync(h, ctx->cursor, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (index) *index = h[0]; if (chunksize) *chunksize = h[1];
    return FDB_OK;
}
extern "C" int fdb_seed_at_cursor(fdb_ctx* ctx, fdb_array* d, const fdb_array* x) {
    FDB_ENTER(ctx);
    int rc = check_pair(ctx, d, x, "fdb_seed_at_cursor"); if (rc) return rc;
    const int PB = x->planes(), P = (d->N + 1) * PB;
    seed_at_cursor_kernel<<<(unsigned)cdiv((int64_t)d->N * P, 128), 128, 0, ctx->stream>>>(d->data, d->ld, x->data, x->ld, d->N, PB, d->n, ctx->cursor);
    return finish(ctx, 1, "fdb_seed_at_cursor");
}
extern "C" int fdb_seed_zero_at_cursor(fdb_ctx* ctx, fdb_array* d, const fdb_array* x) {
    FDB_ENTER(ctx);
    int rc = check_pair(ctx, d, x, "fdb_seed_zero_at_cursor"); if (rc) return rc;
    const int PB = x->planes(), P = (d->N + 1) * PB;
    seed_zero_at_cursor_kernel<<<(unsigned)cdiv((int64_t)d->N * P, 128), 128, 0, ctx->stream>>>(d->data, d->ld, x->data, x->ld, d->N, PB, d->n, ctx->cursor);
    return finish(ctx, 1, "fdb_seed_zero_at_cursor");
}
extern "C" int fdb_extract_gradient_chunk_at_cursor(fdb_ctx* ctx, fdb_array* res, const fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !res || !y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_gradient_chunk_at_cursor: null argument");
    if (y->level < 1 || res->level != y->level - 1) return fail(ctx, FDB_ERR_BADARG, "result must be one level below ydual");
    if (y->n != 1) return fail(ctx, FDB_ERR_SHAPE, "gradient(f, x) expects that f(x) is a real number. Perhaps you meant jacobian(f, x)?");
    int rc = ensure_cursor(ctx); if (rc) return rc;
    const int PB = res->planes();
    extract_gradient_at_cursor_kernel<<<(unsigned)cdiv((int64_t)y->N * PB, 128), 128, 0, ctx->stream>>>(res->data, res->ld, y->data, y->ld, PB, res->n, ctx->cursor);
    return finish(ctx, 1, "fdb_extract_gradient_chunk_at_cursor");
}
extern "C" int fdb_extract_jacobian_chunk_at_cursor(fdb_ctx* ctx, fdb_array* res, int64_t ldres, const fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !res || !y) return fail(ctx, FDB_ERR_BADARG, "fdb_extract_jacobian_chunk_at_cursor: null argument");
    if (y->level < 1 || res->level != y->level - 1) return fail(ctx, FDB_ERR_BADARG, "result must be one level below ydual");
    if (ldres < y->n) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: leading dimension smaller than length(ydual)");
    if (y->n == 0) return FDB_OK;
    int rc = ensure_cursor(ctx); if (rc) return rc;
    dim3 grid((unsigned)cdiv(y->n, 256), (unsigned)y->N, (unsigned)res->planes());
    extract_jacobian_at_cursor_kernel<<<grid, 256, 0, ctx->stream>>>(res->data, res->ld, ldres, y->data, y->ld, res->planes(), y->n, ctx->cursor);
    return finish(ctx, 1, "fdb_extract_jacobian_chunk_at_cursor");
}

// ---- stream capture / replay ---------------------------------------------------------------------------------
extern "C" int fdb_graph_begin(fdb_ctx* ctx) {
    FDB_ENTER(ctx);
    if (!ctx) return FDB_ERR_BADARG;
    if (ctx->capturing) return fail(ctx, FDB_ERR_BADARG, "fdb_graph_begin: already capturing");
    int rc = ensure_cursor(ctx); if (rc) return rc;
    rc = ensure_counter(ctx); if (rc) return rc;
    FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->saved_async = ctx->async;
    ctx->async = true;                    // no stream synchronisation while capturing
    FDB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    return FDB_OK;
}
extern "C" int fdb_graph_end(fdb_ctx* ctx, void** exec_out) {
    FDB_ENTER(ctx);
    if (!ctx || !exec_out) return FDB_ERR_BADARG;
    if (!ctx->capturing) return fail(ctx, FDB_ERR_BADARG, "fdb_graph_end: not capturing");
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->capturing = false; ctx->async = ctx->saved_async;
    if (e != cudaSuccess || !graph) cudaGetLastError();      // an invalidated capture (f synchronised / read device data) is not sticky: the eager loop can follow
    if (e != cudaSuccess || !graph) return fail(ctx, FDB_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    *exec_out = (void*)exec;
    return FDB_OK;
}
extern "C" int fdb_graph_launch(fdb_ctx* ctx, void* exec, int64_t times) {
    FDB_ENTER(ctx);
    if (!ctx || !exec || times < 0) return FDB_ERR_BADARG;
    for (int64_t i = 0; i < times; i++) FDB_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)exec, ctx->stream));
    if (!ctx->async) FDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}
extern "C" int fdb_graph_destroy(void* exec) {
    if (!exec) return FDB_ERR_BADARG;
    cudaGraphExecDestroy((cudaGraphExec_t)exec);
    return FDB_OK;
}
