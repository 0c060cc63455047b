// fdb_jacobian.cu -- batched chunk-mode Jacobians: replaces the chunk loops of
// src/jacobian.jl:169-244 (chunk_mode_jacobian / chunk_mode_jacobian!, f and f! forms)
// and vector mode (:127-162) for the array-valued catalogue targets.
//
// Reference per chunk: seed!(xdual,x,i,seeds) -> ydual = f(xdual) | f!(seed_zero_partials!(ydual,y), xdual)
//   -> extract_jacobian_chunk!(result, ydual, i, N): result[:, i-1 .+ (1:N)] .= partials.(ydual, (1:N)')
// Here the grid covers (rows x chunks): each thread evaluates its rows of f on seeded
// Duals built in registers and stores the chunk's partials straight into the column
// block J[:, start .. start+chunksize) of the column-major result (fused extraction,
// coalesced along rows, streaming stores).  Bound: the 8*m*n bytes of J that must be
// written (HBM), DESIGN.md.
#include "fdb_internal.h"
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

FDB_DEV void chunk_window(int64_t c, int64_t nchunks_total, int lastchunksize, int64_t n, int N, int64_t& start, int& chunksize) {
    const bool last = (c == nchunks_total - 1);
    chunksize = last ? lastchunksize : N;                 // src/jacobian.jl:179
    start = last ? (n - lastchunksize) : c * (int64_t)N; // lastchunkindex - 1, src/jacobian.jl:180
}

// ---- 1-D Brusselator, f!(dx, x) form (definition: SURVEY.md section 8d) ---------------------
//   du_i = 1 + u_i^2 v_i - 4u_i + alpha (u_{i-1} - 2u_i + u_{i+1})
//   dv_i = 3u_i - u_i^2 v_i     + alpha (v_{i-1} - 2v_i + v_{i+1}),  u_0 = u_{M+1} = 1, v_0 = v_{M+1} = 3
// evaluated left to right on Duals.  L yields the Dual at a structural position.
template <class D, class L>
FDB_DEV void brusselator_rows(const L& X, int64_t i, int64_t M, double al, D& du, D& dv) {
    D u = X(i), v = X(M + i);
    D lapu, lapv;
    if (M == 1) { lapu = (1.0 - 2.0 * u) + 1.0; lapv = (3.0 - 2.0 * v) + 3.0; }
    else if (i == 0) { lapu = (1.0 - 2.0 * u) + X(i + 1); lapv = (3.0 - 2.0 * v) + X(M + i + 1); }
    else if (i == M - 1) { lapu = (X(i - 1) - 2.0 * u) + 1.0; lapv = (X(M + i - 1) - 2.0 * v) + 3.0; }
    else { lapu = (X(i - 1) - 2.0 * u) + X(i + 1); lapv = (X(M + i - 1) - 2.0 * v) + X(M + i + 1); }
    du = ((1.0 + pow2(u) * v) - 4.0 * u) + al * lapu;
    dv = (3.0 * u - pow2(u) * v) + al * lapv;
}

template <int N, bool NS, int THREADS, bool SHARE, bool MIR>
__global__ void __launch_bounds__(THREADS) brusselator_jacobian_kernel(const double* __restrict__ x, int64_t n, double alpha,
                                                                       int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                       double* __restrict__ J, int64_t ldJ, double* __restrict__ y,
                                                                       const unsigned nrb, const unsigned nch, const bool rowfast,
                                                                       const __grid_constant__ fdb_mirrors Mr) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t M = n / 2;
    // 1-D grid of nch chunks x nrb row blocks.  rowfast: blocks in flight fill a few column groups top to bottom, so
    // the write stream to the peers of an exchange walks memory in order (NVLink stores striding over the whole
    // matrix ran at a third of the link rate); otherwise chunk fastest, which spreads local HBM writes best.
    const int64_t c = chunk_lo + (rowfast ? blockIdx.x / nrb : blockIdx.x % nch);
    const unsigned rb = rowfast ? blockIdx.x % nrb : blockIdx.x / nch;
    int64_t start; int chunksize;
    chunk_window(c, nchunks_total, lastchunksize, n, N, start, chunksize);
    const int64_t i = (int64_t)rb * THREADS + threadIdx.x;               // grid point
    // warp-uniform: do rows w0..w0+31 read any seeded position?  they read u[w0-1..w0+32] and v[M + same]
    const int64_t w0 = i - (threadIdx.x & 31);
    const int64_t wend = start + chunksize;                               // window [start, wend)
    const bool hit_u = (w0 - 1 < wend) && (w0 + 32 >= start);
    const bool hit_v = (M + w0 - 1 < wend) && (M + w0 + 32 >= start);
    if (i >= M) return;
    const bool write_y = (y != nullptr) && (c == nchunks_total - 1);     // y <- value.(ydual) of the last sweep, jacobian.jl:229
    double* Ju = J + i + start * ldJ;
    double* Jv = Ju + M;
    if (SHARE && !hit_u && !hit_v) {
        D1 du, dv;
        brusselator_rows(ZeroLoader<1, NS>{x}, i, M, alpha, du, dv);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) { mstore_cs<MIR>(Mr, Ju + k * ldJ, du.p[0]); mstore_cs<MIR>(Mr, Jv + k * ldJ, dv.p[0]); }
        if (write_y) { mstore(Mr, &y[i], du.v); mstore(Mr, &y[M + i], dv.v); }
    } else {
        D du, dv;
        brusselator_rows(SeedLoader<N, NS>{x, start, chunksize}, i, M, alpha, du, dv);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) { mstore_cs<MIR>(Mr, Ju + k * ldJ, du.p[k]); mstore_cs<MIR>(Mr, Jv + k * ldJ, dv.p[k]); }   // extract_jacobian_chunk!, jacobian.jl:109-118
        if (write_y) { mstore(Mr, &y[i], du.v); mstore(Mr, &y[M + i], dv.v); }
    }
}

// ---- test/JacobianTest.jl:23-31 target (n = 3, m = 4): one thread per chunk -------------------
template <int N, bool NS>
__global__ void jactest_jacobian_kernel(const double* __restrict__ x, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                        double* __restrict__ J, int64_t ldJ, double* __restrict__ yout, const __grid_constant__ fdb_mirrors Mr) {
    typedef Dual<double, N, NS> D;
    const int64_t c = chunk_lo + blockIdx.x;
    int64_t start; int chunksize;
    chunk_window(c, nchunks_total, lastchunksize, 3, N, start, chunksize);
    SeedLoader<N, NS> X{x, start, chunksize};
    D y[4];
    y[0] = X(0) * X(1);
    y[0] = y[0] * sin_v(pow2(X(2)));
    y[1] = y[0] + X(2);
    y[2] = y[0] / y[1];
    y[3] = X(2);
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < N; k++) if (k < chunksize) mstore(Mr, &J[r + (start + k) * ldJ], y[r].p[k]);
        if (yout && c == nchunks_total - 1) mstore(Mr, &yout[r], y[r].v);
    }
}

// ---- f(x) = tanh.(A*x .+ b), CUDA-core version --------------------------------------------------
// One thread = one row of A, CH consecutive chunks.  `A*x` is the Real*Dual / Dual+Dual
// column sweep of LinearAlgebra.generic_matvecmul!; the summation order differs from the
// sequential reference (reduction tolerance).  Column j is "near" for a chunk when it lies
// in that chunk's window -- a block-uniform test since every thread walks the same j.
template <int N, bool NS, int THREADS, int CH, bool SHARE>
__global__ void __launch_bounds__(THREADS) tanh_affine_jacobian_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ b,
                                                                       const double* __restrict__ x, int64_t n, int64_t m,
                                                                       int64_t chunk_lo, int64_t chunk_hi, int64_t nchunks_total,
                                                                       int lastchunksize, double* __restrict__ J, int64_t ldJ,
                                                                       double* __restrict__ yout, const __grid_constant__ fdb_mirrors Mr) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t r = (int64_t)blockIdx.y * THREADS + threadIdx.x;
    const int64_t c0 = chunk_lo + (int64_t)blockIdx.x * CH;
    int64_t start[CH]; int csz[CH];
    D acc[CH]; D1 acc1[CH];
#pragma unroll
    for (int q = 0; q < CH; q++) {
        int64_t c = c0 + q; if (c >= chunk_hi) c = chunk_hi - 1;          // clamp: duplicates the last chunk, stores are skipped below
        chunk_window(c, nchunks_total, lastchunksize, n, N, start[q], csz[q]);
        acc[q].v = 0.0; acc1[q].v = 0.0; acc1[q].p[0] = 0.0;
#pragma unroll
        for (int k = 0; k < N; k++) acc[q].p[k] = 0.0;
    }
    if (r >= m) return;
    const double* Ar = A + r;
    for (int64_t j = 0; j < n; j++) {
        const double a = __ldg(Ar + j * lda);
        const double xj = __ldg(x + j);
#pragma unroll
        for (int q = 0; q < CH; q++) {
            const int64_t off = j - start[q];
            if (!SHARE || (off >= 0 && off < csz[q])) {
                D xd = seeded<N, NS>(xj, j, start[q], csz[q]);
                acc[q] = acc[q] + a * xd;                                  // C[i] += A[i,k] * b
            } else {
                D1 xd; xd.v = xj; xd.p[0] = 0.0;
                acc1[q] = acc1[q] + a * xd;
            }
        }
    }
    const double bi = b[r];
#pragma unroll
    for (int q = 0; q < CH; q++) {
        const int64_t c = c0 + q;
        if (c >= chunk_hi) break;
        D z = acc[q];
        if (SHARE) {
            D e; e.v = acc1[q].v;
#pragma unroll
            for (int k = 0; k < N; k++) e.p[k] = acc1[q].p[0];
            z = z + e;
        }
        D yv = tanh_v(z + bi);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < csz[q]) mstore_cs<true>(Mr, J + r + (start[q] + k) * ldJ, yv.p[k]);
        if (yout && c == nchunks_total - 1) mstore(Mr, &yout[r], yv.v);
    }
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_jacobian_chunked(fdb_ctx* ctx, int fn, const fdb_array* x, int64_t m, int N, const fdb_array* A, int64_t lda,
                                    const fdb_array* b, double alpha, int64_t chunk_lo, int64_t chunk_hi, fdb_array* J, int64_t ldJ,
                                    fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !J) return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_chunked: null argument");
    if (x->level != 0 || J->level != 0 || (y && y->level != 0)) return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_chunked: x, J, y must be plain arrays");
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N)   // src/jacobian.jl:172-174
        return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (ldJ < m || J->n < ldJ * (n - 1) + m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Jacobian result must hold %lld x %lld", (long long)m, (long long)n);
    if (y && y->n != m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: y has %lld entries, f(x) has %lld", (long long)y->n, (long long)m);
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;      // N == n: vector mode, src/jacobian.jl:21-22
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    const int lcs = (int)p.lastchunksize;
    double* yp = y ? y->data : nullptr;
    fdb_mirrors Mr;
    const fdb_array* outs[2] = {J, y};
    { int mrc = mirrors_for(ctx, &Mr, outs, 2, "fdb_jacobian_chunked"); if (mrc) return mrc; }
    constexpr int T = 256;
    switch (fn) {
        case FDB_JFN_BRUSSELATOR: {
            if (n % 2 != 0 || m != n) return fail(ctx, FDB_ERR_SHAPE, "brusselator needs x = [u; v] (even length) and m == n");
            const unsigned nrb = (unsigned)cdiv(n / 2, T);
            if ((chunk_hi - chunk_lo) * (int64_t)nrb > 0x7fffffffLL) return fail(ctx, FDB_ERR_UNSUPPORTED, "brusselator: chunk range too large for one launch");
            const unsigned nch = (unsigned)(chunk_hi - chunk_lo);
            const unsigned grid = nch * nrb;
            const bool rowfast = Mr.count > 0;      // measured: chunk-fastest 4.76 ms, row-fastest 4.99 ms locally (C4); 8.1 vs 2.8 ms over NVLink
#define FDB_BRUSS(SH, MI) FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe, \
                (brusselator_jacobian_kernel<NN, NS, T, SH, MI><<<grid, T, 0, ctx->stream>>>(x->data, n, alpha, chunk_lo, nchunks, lcs, J->data, ldJ, yp, nrb, nch, rowfast, Mr))))
            if (ctx->lane_sharing) { if (Mr.count) { FDB_BRUSS(true, true); } else { FDB_BRUSS(true, false); } }
            else { if (Mr.count) { FDB_BRUSS(false, true); } else { FDB_BRUSS(false, false); } }
#undef FDB_BRUSS
            return finish(ctx, 1, "fdb_jacobian_chunked(brusselator)");
        }
        case FDB_JFN_JACTEST: {
            if (n != 3 || m != 4) return fail(ctx, FDB_ERR_SHAPE, "jactest target needs n = 3, m = 4");
            if (N > 3) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x)");
            FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
                (jactest_jacobian_kernel<NN, NS><<<(unsigned)(chunk_hi - chunk_lo), 1, 0, ctx->stream>>>(x->data, chunk_lo, nchunks, lcs, J->data, ldJ, yp, Mr))));
            return finish(ctx, 1, "fdb_jacobian_chunked(jactest)");
        }
        case FDB_JFN_TANH_AFFINE: {
            if (!A || !b) return fail(ctx, FDB_ERR_BADARG, "tanh_affine needs A and b");
            if (A->level != 0 || b->level != 0 || lda < m || A->n < lda * (n - 1) + m || b->n != m)
                return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A must be %lld x %lld, b of length %lld", (long long)m, (long long)n, (long long)m);
            if (ctx->use_dmma && dmma_eligible(A->data, lda, m, n)) {
                // contraction on FP64 tensor cores (fdb_dmma.cu); all requested chunks in one GEMM
                return tanh_affine_dmma(ctx, A, lda, b, x, m, N, chunk_lo, chunk_hi, nchunks, lcs, J, ldJ, yp, nullptr, Mr);
            }
            constexpr int CH = 4; constexpr int TT = 128;
            dim3 grid((unsigned)cdiv(chunk_hi - chunk_lo, CH), (unsigned)cdiv(m, TT));
            if (ctx->lane_sharing) {
                FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
                    (tanh_affine_jacobian_kernel<NN, NS, TT, CH, true><<<grid, TT, 0, ctx->stream>>>(A->data, lda, b->data, x->data, n, m, chunk_lo, chunk_hi,
                                                                                                  nchunks, lcs, J->data, ldJ, yp, Mr))));
            } else {
                FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
                    (tanh_affine_jacobian_kernel<NN, NS, TT, CH, false><<<grid, TT, 0, ctx->stream>>>(A->data, lda, b->data, x->data, n, m, chunk_lo, chunk_hi,
                                                                                                   nchunks, lcs, J->data, ldJ, yp, Mr))));
            }
            return finish(ctx, 1, "fdb_jacobian_chunked(tanh_affine)");
        }
    }
    return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_jacobian_chunked: unknown target %d", fn);
}
