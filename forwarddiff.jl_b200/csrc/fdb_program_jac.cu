// fdb_program_jac.cu -- Jacobian chunk sweeps of a flattened elementwise / stencil user function
// (see fdb_program.cu; split into its own translation unit to keep build times down).
#include "fdb_program.cuh"
#include "fdb_jac_kernel.cuh"

namespace fdb {

// elementwise / stencil array targets: y[i] = program(x[i .. i+halo]) for i in [0, n - halo);
// J[i, start+k] = partials(y[i], k) -- the f(x)-form Jacobian chunk sweep (src/jacobian.jl:169-216)
template <int N, bool NS, int THREADS, bool MIR>
__global__ void __launch_bounds__(THREADS) program_jacobian_kernel(const __grid_constant__ Program P, const bool SHARE, const double* __restrict__ x, int64_t n,
                                                                   int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                   double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,
                                                                   const unsigned nrb, const unsigned nch, const bool rowfast,
                                                                   const __grid_constant__ fdb_mirrors M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    // 1-D grid; row block fastest when the stores also go to peers (see brusselator_jacobian_kernel)
    const int64_t c = chunk_lo + (rowfast ? blockIdx.x / nrb : blockIdx.x % nch);
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    const int64_t i = (int64_t)(rowfast ? blockIdx.x % nrb : blockIdx.x / nch) * THREADS + threadIdx.x;
    const int64_t m = n - P.halo;
    const int64_t w0 = i - (threadIdx.x & 31);
    const bool far = SHARE && ((w0 + 31 + P.halo < start) || (w0 >= start + chunksize));
    __shared__ double rf1[PG_MAXR][2][THREADS];          // Dual{1} register file of the lane-shared path
    if (i >= m) return;
    if (far) {
        D1 acc1[PG_MAXK];
        pg_run_shared1<NS, THREADS>(P.term, P.n_term, P.consts, ZeroLoader<1, NS>{x}, i, rf1, acc1);
        D1 y; y.v = rf1[P.result_reg][0][threadIdx.x]; y.p[0] = rf1[P.result_reg][1][threadIdx.x];
        for (int k = 0; k < chunksize; k++) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[0]);
        if (yout && last) mstore(M, &yout[i], y.v);
    } else {
        D R[PG_MAXR]; D acc[1];
        pg_run<D, N, NS>(P.term, P.n_term, P.consts, SeedLoader<N, NS>{x, start, chunksize}, i, R, acc);
        const D y = R[P.result_reg];
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[k]);
        if (yout && last) mstore(M, &yout[i], y.v);
    }
}

// y = g.(A * x): the interpreter twin of matvec_jacobian_body (fdb_jac_kernel.cuh); z_i comes from the tensor-core product
template <int N, bool NS, int THREADS, bool MIR>
__global__ void __launch_bounds__(THREADS) program_matvec_jacobian_kernel(const __grid_constant__ Program P, const bool SHARE, const double* __restrict__ Y,
                                                                          int64_t ldy, const double* __restrict__ A, int64_t lda, int64_t m, int64_t n,
                                                                          int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                          double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,
                                                                          const unsigned nrb, const unsigned nch, const bool rowfast,
                                                                          const __grid_constant__ fdb_mirrors M) {
    typedef Dual<double, N, NS> D;
    const int64_t cl = (rowfast ? blockIdx.x / nrb : blockIdx.x % nch), c = chunk_lo + cl;
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    const int64_t i = (int64_t)(rowfast ? blockIdx.x % nrb : blockIdx.x / nch) * THREADS + threadIdx.x;
    if (i >= m) return;
    ProductLoader<N, NS> z;
    z.ldy = ldy; z.lda = lda; z.chunksize = chunksize;
    if (SHARE) { z.Yv = Y + 2 * cl * ldy; z.Yl = z.Yv + ldy; z.Aw = A + start * lda; }
    else { z.Yv = Y; z.Yl = Y + (1 + start - chunk_lo * N) * ldy; z.Aw = nullptr; }
    D R[PG_MAXR]; D acc[1];
    pg_run<D, N, NS>(P.term, P.n_term, P.consts, z, i, R, acc);
    const D y = R[P.result_reg];
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[k]);
    if (yout && last) mstore(M, &yout[i], y.v);
}

}  // namespace fdb

using namespace fdb;

// jacobian(x -> g.(A * x), x, cfg) for a flattened elementwise g (LOADX 0 = z_i): contraction on the FP64 tensor cores, then the
// program on the product lanes.  A: m x n plain column-major (lda).
extern "C" int fdb_program_matvec_jacobian_chunked(fdb_ctx* ctx, const int32_t* prog, int n_instr, const double* consts, int n_consts, int result_reg,
                                                   const fdb_array* A, int64_t lda, int64_t m, const fdb_array* x, int N, int64_t chunk_lo,
                                                   int64_t chunk_hi, fdb_array* J, int64_t ldJ, fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !J || !A) return fail(ctx, FDB_ERR_BADARG, "fdb_program_matvec_jacobian_chunked: null argument");
    if (x->level != 0 || J->level != 0 || A->level != 0 || (y && y->level != 0)) return fail(ctx, FDB_ERR_BADARG, "A, x, J, y must be plain arrays");
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (m <= 0 || lda < m || A->n < lda * (n - 1) + m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A must be %lld x %lld", (long long)m, (long long)n);
    if (ldJ < m || J->n < ldJ * (n - 1) + m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Jacobian result must hold %lld x %lld", (long long)m, (long long)n);
    if (y && y->n != m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: y has %lld entries, f(x) has %lld", (long long)y->n, (long long)m);
    Program P;
    int rc = build_program(ctx, P, prog, n_instr, nullptr, 0, consts, n_consts, 0, 0, result_reg, true, ctx->n_prm); if (rc) return rc;
    rc = check_params(ctx, P, m, "fdb_program_matvec_jacobian_chunked"); if (rc) return rc;
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    fdb_mirrors M;
    const fdb_array* outs[2] = {J, y};
    rc = mirrors_for(ctx, &M, outs, 2, "fdb_program_matvec_jacobian_chunked"); if (rc) return rc;
    double* Y = nullptr; int64_t ldy = 0; int nl = 0;
    rc = matvec_product_lanes(ctx, A->data, lda, m, x, N, chunk_lo, chunk_hi, nchunks, &Y, &ldy, &nl); if (rc) return rc;
    constexpr int T = 128;
    const unsigned nrb = (unsigned)cdiv(m, T);
    if ((chunk_hi - chunk_lo) * (int64_t)nrb > 0x7fffffffLL) return fail(ctx, FDB_ERR_UNSUPPORTED, "chunk range too large for one launch");
    const unsigned nch = (unsigned)(chunk_hi - chunk_lo), grid = nch * nrb;
    double* yp = y ? y->data : nullptr;
    const bool share = ctx->lane_sharing, rowfast = M.count > 0;
    {
        void* fn = nullptr;
        if (jit_get(ctx, 3, P, N, ctx->nansafe, false, M.count > 0, &fn) == FDB_OK) {
            bool sh = share, rf = rowfast; const double* Yp = Y; int64_t ldy_ = ldy; const double* Ap = A->data; int64_t lda_ = lda, m_ = m, nn = n, lo = chunk_lo, nc = nchunks;
            int lcs = (int)p.lastchunksize; double* Jp = J->data; int64_t ld = ldJ; unsigned nrb_ = nrb, nch_ = nch;
            void* args[] = {&sh, &Yp, &ldy_, &Ap, &lda_, &m_, &nn, &lo, &nc, &lcs, &Jp, &ld, &yp, &nrb_, &nch_, &rf, &M};
            rc = jit_launch(ctx, fn, grid, 1, T, args, "fdb_program_matvec_jacobian_chunked(specialised)"); if (rc) return rc;
            return finish(ctx, nl + 1, "fdb_program_matvec_jacobian_chunked(specialised)");
        }
    }
    if (P.n_params || !P.fits_interpreter) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program with array parameters or beyond the interpreter's limits needs the run-time specialised kernel (%s)",
                                ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    if (M.count) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_matvec_jacobian_kernel<NN, NS, T, true><<<grid, T, 0, ctx->stream>>>(P, share, Y, ldy, A->data, lda, m, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp, nrb, nch, true, M))));
    } else {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_matvec_jacobian_kernel<NN, NS, T, false><<<grid, T, 0, ctx->stream>>>(P, share, Y, ldy, A->data, lda, m, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp, nrb, nch, false, M))));
    }
    return finish(ctx, nl + 1, "fdb_program_matvec_jacobian_chunked");
}

extern "C" int fdb_program_jacobian_chunked(fdb_ctx* ctx, const int32_t* prog, int n_instr, const double* consts, int n_consts, int halo,
                                            int result_reg, const fdb_array* x, int N, int64_t chunk_lo, int64_t chunk_hi, fdb_array* J,
                                            int64_t ldJ, fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !J) return fail(ctx, FDB_ERR_BADARG, "fdb_program_jacobian_chunked: null argument");
    if (x->level != 0 || J->level != 0 || (y && y->level != 0)) return fail(ctx, FDB_ERR_BADARG, "x, J, y must be plain arrays");
    const int64_t n = x->n, m = n - halo;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (m <= 0) return fail(ctx, FDB_ERR_SHAPE, "op-program Jacobian: empty output");
    if (ldJ < m || J->n < ldJ * (n - 1) + m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: Jacobian result must hold %lld x %lld", (long long)m, (long long)n);
    if (y && y->n != m) return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: y has %lld entries, f(x) has %lld", (long long)y->n, (long long)m);
    Program P;
    int n_acc = 0;                                   // pass-1 sums of a two-pass program (softmax ...): accumulators are numbered densely
    for (int i = 0; i < n_instr && prog; i++) if ((prog[4 * i] >> 6) == PK_ACC && prog[4 * i + 1] + 1 > n_acc) n_acc = prog[4 * i + 1] + 1;
    int rc = build_program(ctx, P, prog, n_instr, nullptr, 0, consts, n_consts, n_acc, halo, result_reg, true, ctx->n_prm); if (rc) return rc;
    rc = check_params(ctx, P, n, "fdb_program_jacobian_chunked"); if (rc) return rc;
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    constexpr int T = 128;
    const unsigned nrb = (unsigned)cdiv(m, T);
    if ((chunk_hi - chunk_lo) * (int64_t)nrb > 0x7fffffffLL) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program Jacobian: chunk range too large for one launch");
    const unsigned nch = (unsigned)(chunk_hi - chunk_lo);
    const unsigned grid = nch * nrb;
    double* yp = y ? y->data : nullptr;
    fdb_mirrors M;
    const fdb_array* outs[2] = {J, y};
    rc = mirrors_for(ctx, &M, outs, 2, "fdb_program_jacobian_chunked"); if (rc) return rc;
    if (P.two_pass) {
        // reductions inside an array-valued f: one block per chunk, pass 1 reduces, pass 2 writes the dense Jacobian rows
        void* fn = nullptr;
        if (jit_get(ctx, 4, P, N, ctx->nansafe, ctx->lane_sharing, M.count > 0, &fn) != FDB_OK)
            return fail(ctx, FDB_ERR_UNSUPPORTED, "a two-pass array-valued op-program needs the run-time specialised kernel (%s)",
                        ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
        const double* xp = x->data; int64_t nn = n, lo = chunk_lo, nc = nchunks, ld = ldJ; int lcs = (int)p.lastchunksize; double* Jp = J->data;
        void* args[] = {&xp, &nn, &lo, &nc, &lcs, &Jp, &ld, &yp, &M};
        rc = jit_launch(ctx, fn, nch, 1, 256, args, "fdb_program_jacobian_chunked(two-pass, specialised)"); if (rc) return rc;
        return finish(ctx, 1, "fdb_program_jacobian_chunked(two-pass, specialised)");
    }
    {   // specialised kernel for this program (fdb_jit.cu)
        void* fn = nullptr;
        if (jit_get(ctx, 1, P, N, ctx->nansafe, false, M.count > 0, &fn) == FDB_OK) {
            bool share = ctx->lane_sharing, rowfast = M.count > 0;
            const double* xp = x->data; int64_t nn = n, lo = chunk_lo, nc = nchunks, ld = ldJ; int lcs = (int)p.lastchunksize;
            double* Jp = J->data; unsigned nrb_ = nrb, nch_ = nch;
            void* args[] = {&share, &xp, &nn, &lo, &nc, &lcs, &Jp, &ld, &yp, &nrb_, &nch_, &rowfast, &M};
            rc = jit_launch(ctx, fn, grid, 1, T, args, "fdb_program_jacobian_chunked(specialised)"); if (rc) return rc;
            return finish(ctx, 1, "fdb_program_jacobian_chunked(specialised)");
        }
    }
    if (P.n_params || !P.fits_interpreter) return fail(ctx, FDB_ERR_UNSUPPORTED, "op-program with array parameters or beyond the interpreter's limits needs the run-time specialised kernel (%s)",
                                ctx->use_jit ? fdb_jit_last_log(ctx) : "disabled by fdb_set_jit");
    if (M.count) {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_jacobian_kernel<NN, NS, T, true><<<grid, T, 0, ctx->stream>>>(P, ctx->lane_sharing, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp, nrb, nch, true, M))));
    } else {
        FDB_DISPATCH_N(N, FDB_DISPATCH_NS(ctx->nansafe,
            (program_jacobian_kernel<NN, NS, T, false><<<grid, T, 0, ctx->stream>>>(P, ctx->lane_sharing, x->data, n, chunk_lo, nchunks, (int)p.lastchunksize, J->data, ldJ, yp, nrb, nch, false, M))));
    }
    return finish(ctx, 1, "fdb_program_jacobian_chunked");
}
