// fdb_hessian_kernel.cuh -- the batched nested-dual Hessian kernel (see fdb_hessian.cu), as a header so that the
// run-time specialiser (fdb_jit.cu) instantiates the same code for a user function's generated functor.
#pragma once
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

// far-from-both-windows loader: Dual{Dual{1},1} with all partials zero
template <bool NS> struct FFLoader {
    typedef Dual<Dual<double, 1, NS>, 1, NS> DO;
    const double* x;
    FDB_DEV DO operator()(int64_t i) const {
        DO d; d.v.v = x[i]; d.v.p[0] = 0.0; d.p[0].v = 0.0; d.p[0].p[0] = 0.0; return d;
    }
};
// a term that reads the inner window but no position of the outer window: every outer partial lane of every intermediate is the
// same "zero lane", so one representative outer lane is enough -- Dual{Dual{1},1} with the inner one-hot seed of lane k
template <bool NS> struct NearInnerLoader {
    typedef Dual<Dual<double, 1, NS>, 1, NS> DO;
    const double* x; int64_t istart; int ics; int k;
    FDB_DEV DO operator()(int64_t pos) const {
        DO d; d.v.v = x[pos]; d.v.p[0] = 0.0;
        d.p[0].v = ((pos - istart == k) && (k < ics)) ? 1.0 : 0.0; d.p[0].p[0] = 0.0;
        return d;
    }
};
// general loader for inner lane k: value = outer-seeded Dual{N}; the single inner partial lane
// carries one(Dual{N}) iff pos is slot k of the inner window (src/apiutils.jl:125-143 with
// Partials{N,Dual{T,V,N}} seeds, src/config.jl:225-226)
template <int N, bool NS> struct NearLoader {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DO;
    const double* x; int64_t ostart; int ocs; int64_t istart; int ics; int k;
    FDB_DEV DO operator()(int64_t pos) const {
        DO d; d.v = seeded<N, NS>(x[pos], pos, ostart, ocs);
        const bool hot = (pos - istart == k) && (k < ics);
        d.p[0].v = hot ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) d.p[0].p[j] = 0.0;
        return d;
    }
};

// A block handles HESS_GROUP consecutive inner chunks against one outer chunk.  A term that reads no seeded position has the
// same four shared lanes whatever the pair, so the terms far from ALL the group's inner windows and from the outer window are
// evaluated once per block (phase 1) instead of once per pair -- at C5 that is 2047 four-lane evaluations per eight pairs, where
// the first version spent three quarters of its FP64 instructions on them.  Terms inside the group's inner range are far for
// some pairs and near for others: their four-lane contribution is computed once (phase 1b) and added per pair when the term is
// not near that pair's window.  Near terms that read only ONE of the two windows share lanes on the other level (inner window only:
// all outer lanes are the same zero lane; outer window only: the inner partial is zero(Dual{N}) whatever the inner lane and pair),
// and the pairs whose windows do not meet -- all but the few around the diagonal -- go through phases 2 and 3 as one batch of work
// items with two barriers per block instead of two per pair.  Every term of every pair is still evaluated exactly as the reference's
// nested duals would see it (seeded lanes where it reads a seeded position, representative zero lanes elsewhere: 0 * Inf still
// poisons), only the order of the additions differs (a tree either way).
constexpr int HESS_GROUP = 8;     // measured at C5: 16 gains another 2 %, 8 keeps more blocks in flight for small n

template <class F, int N, bool NS, int THREADS>
FDB_DEV void hessian_sweep_body(const double* __restrict__ x, int64_t n, int64_t outer_lo,
                                                                int64_t nchunks_total, int lastchunksize, double* __restrict__ H,
                                                                int64_t ldH, double* __restrict__ grad, double* __restrict__ value_out,
                                                                const fdb_mirrors& M) {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DG;
    typedef Dual<Dual<double, 1, NS>, 1, NS> DF;
    constexpr int NA = N + F::HALO;                 // terms reading one window
    constexpr int MAXNEAR = 2 * NA;
    constexpr int MAXU = HESS_GROUP * N + F::HALO;
    // One buffer, two uses.  (1) pairs whose windows do not meet (all but the few around the diagonal), all at once:
    //     sa(g,t,k,0..1) = (value, representative outer lane) of partials(y,k) for term t of pair g's inner range;
    // (2) afterwards, pairs around the diagonal one at a time: sp(t,k,0..N) as described at phase 2 below.
    constexpr int SA_DOUBLES = HESS_GROUP * NA * N * 2, SP_DOUBLES = MAXNEAR * N * (N + 1);
    __shared__ double s_buf[SA_DOUBLES > SP_DOUBLES ? SA_DOUBLES : SP_DOUBLES];
    __shared__ double s_v[MAXNEAR];             // diagonal pairs, per near term: value(value(y))
    __shared__ double s_av[HESS_GROUP][NA];     // batched pairs, per inner-range term: value(value(y))
    __shared__ double s_b[NA][N + 1];           // terms of the outer window's range, inner partial zero(Dual{N}): partials(y,.) as Dual{N}, for every k and pair
    __shared__ double s_bv[NA];                 // ... their value(value(y))
    __shared__ double s_c4[MAXU][4];            // four shared lanes of every term of the group's inner range (zero for terms near the outer window)
    __shared__ double s_common[4];              // far from every window of the block
    __shared__ double s_ffg[HESS_GROUP][4];     // far sum of each pair
    __shared__ double s_red[THREADS / 32][4];
    auto sa = [&](int g, int t, int k, int q) -> double& { return s_buf[((g * NA + t) * N + k) * 2 + q]; };
    auto sp = [&](int t, int k, int j) -> double& { return s_buf[(t * N + k) * (N + 1) + j]; };

    const int64_t cg0 = (int64_t)blockIdx.x * HESS_GROUP, co = outer_lo + blockIdx.y;
    const int gcount = (int)((nchunks_total - cg0) < HESS_GROUP ? (nchunks_total - cg0) : HESS_GROUP);
    const bool olast = (co == nchunks_total - 1);
    const int ocs = olast ? lastchunksize : N;
    const int64_t ostart = olast ? n - lastchunksize : co * N;
    const int64_t nt = F::nterms(n);
    auto inner_start = [&](int64_t ci) { return ci == nchunks_total - 1 ? n - lastchunksize : ci * N; };
    auto inner_size = [&](int64_t ci) { return ci == nchunks_total - 1 ? lastchunksize : N; };
    // raw near range of inner chunk cg0 + g (terms reading a seeded position of its window)
    auto a_range = [&](int g, int64_t& r0, int64_t& r1) {
        const int64_t is = inner_start(cg0 + g);
        r0 = is - F::HALO; r1 = is + inner_size(cg0 + g);
        if (r0 < 0) r0 = 0; if (r1 > nt) r1 = nt; if (r1 < r0) r1 = r0;
    };

    // raw near ranges: B for the outer window, U for the union of the group's inner windows
    int64_t br0 = ostart - F::HALO, br1 = ostart + ocs;
    if (br0 < 0) br0 = 0; if (br1 > nt) br1 = nt; if (br1 < br0) br1 = br0;
    const int nbr = (int)(br1 - br0);
    int64_t u0 = inner_start(cg0) - F::HALO, u1 = inner_start(cg0 + gcount - 1) + inner_size(cg0 + gcount - 1);
    if (u0 < 0) u0 = 0; if (u1 > nt) u1 = nt; if (u1 < u0) u1 = u0;
    const int nu = (int)(u1 - u0);
    // a pair is "diagonal" when its two ranges meet: only then can a term read both windows
    auto diagonal = [&](int g) { int64_t r0, r1; a_range(g, r0, r1); return br0 <= r1 && r0 <= br1; };

    // ---- phase 1: terms far from every window of the block, four shared lanes ------------------------
    DF accff[1]; accff[0].v.v = 0.0; accff[0].v.p[0] = 0.0; accff[0].p[0].v = 0.0; accff[0].p[0].p[0] = 0.0;
    FFLoader<NS> xf{x};
    for (int64_t i = threadIdx.x; i < nt; i += THREADS) {
        const bool near = (i >= u0 && i < u1) || (i >= br0 && i < br1);
        if (!near) F::term(xf, i, n, accff);
    }
    {
        double r[4] = {accff[0].v.v, accff[0].v.p[0], accff[0].p[0].v, accff[0].p[0].p[0]};
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) r[q] = r[q] + shfl_down_d(r[q], off);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][q] = r[q];
        }
    }
    // ---- phase 1b: the terms of the inner range, each on its own (far for the pairs whose window it does not touch) ----
    for (int w = threadIdx.x; w < nu; w += THREADS) {
        const int64_t i = u0 + w;
        DF a[1]; a[0].v.v = 0.0; a[0].v.p[0] = 0.0; a[0].p[0].v = 0.0; a[0].p[0].p[0] = 0.0;
        if (!(i >= br0 && i < br1)) F::term(xf, i, n, a);            // near the outer window: never far, contributes nothing here
        s_c4[w][0] = a[0].v.v; s_c4[w][1] = a[0].v.p[0]; s_c4[w][2] = a[0].p[0].v; s_c4[w][3] = a[0].p[0].p[0];
    }
    // ---- phase 1c: the terms of the outer window's range with the inner partial zero(Dual{N}) -- what they are for every inner lane
    // of every pair whose inner window they do not read (threads from the top end: the low ones are busy with 1b) ----
    for (int w = THREADS - 1 - (int)threadIdx.x; w < nbr; w += THREADS) {
        NearLoader<N, NS> xn{x, ostart, ocs, 0, 0, 0};               // ics = 0: no inner lane is hot
        DG acc[1];
        acc[0].v.v = 0.0; acc[0].p[0].v = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) { acc[0].v.p[j] = 0.0; acc[0].p[0].p[j] = 0.0; }
        F::term(xn, br0 + w, n, acc);
        s_b[w][0] = acc[0].p[0].v;
#pragma unroll
        for (int j = 0; j < N; j++) s_b[w][j + 1] = acc[0].p[0].p[j];
        s_bv[w] = acc[0].v.v;
    }
    // ---- phase 2, pairs off the diagonal, all at once: one work item per (pair, term of its inner range, inner lane k).  Such a
    // term reads no position of the outer window: all outer lanes of every intermediate are the same zero lane, Dual{Dual{1},1} ----
    for (int w = threadIdx.x; w < gcount * NA * N; w += THREADS) {
        const int g = w / (NA * N), r = w - g * (NA * N), t = r / N, k = r - t * N;
        int64_t ar0, ar1; a_range(g, ar0, ar1);
        if (t >= (int)(ar1 - ar0) || (br0 <= ar1 && ar0 <= br1)) continue;
        NearInnerLoader<NS> xa{x, inner_start(cg0 + g), inner_size(cg0 + g), k};
        DF a[1]; a[0].v.v = 0.0; a[0].v.p[0] = 0.0; a[0].p[0].v = 0.0; a[0].p[0].p[0] = 0.0;
        F::term(xa, ar0 + t, n, a);
        sa(g, t, k, 0) = a[0].p[0].v; sa(g, t, k, 1) = a[0].p[0].p[0];
        if (k == 0) s_av[g][t] = a[0].v.v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int wq = 0; wq < THREADS / 32; wq++) s = s + s_red[wq][threadIdx.x];
        s_common[threadIdx.x] = s;
    }
    __syncthreads();
    // far sum of every pair: the block's common part + the inner-range terms outside the pair's own window
    for (int g = threadIdx.x >> 5; g < gcount; g += THREADS / 32) {
        const int l = threadIdx.x & 31;
        int64_t ar0, ar1; a_range(g, ar0, ar1);
        double r[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = l; w < nu; w += 32) {
            const int64_t i = u0 + w;
            if (i >= ar0 && i < ar1) continue;
#pragma unroll
            for (int q = 0; q < 4; q++) r[q] = r[q] + s_c4[w][q];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) r[q] = r[q] + shfl_down_d(r[q], off);
        }
        if (l == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) s_ffg[g][q] = r[q] + s_common[q];
        }
    }
    __syncthreads();
    // ---- phase 3, pairs off the diagonal: fold and store (fused extract_gradient_chunk! + extract_jacobian_chunk!) ----
    // far lanes: [0] = value.value, [1] = value.outer-partial, [2] = inner-partial.value, [3] = inner-partial.outer-partial
    for (int e = threadIdx.x; e < gcount * N * (N + 1); e += THREADS) {
        const int g = e / (N * (N + 1)), r = e - g * (N * (N + 1)), k = r / (N + 1), j = r - k * (N + 1);
        int64_t ar0, ar1; a_range(g, ar0, ar1);
        if (br0 <= ar1 && ar0 <= br1) continue;
        const int64_t ci = cg0 + g, istart = inner_start(ci);
        const int ics = inner_size(ci), na = (int)(ar1 - ar0);
        if (k < ics) {
            double s = 0.0;
            if (ar0 < br0) {                                         // term order: the lower range first
                for (int t = 0; t < na; t++) s = s + sa(g, t, k, j == 0 ? 0 : 1);
                for (int t = 0; t < nbr; t++) s = s + s_b[t][j];
            } else {
                for (int t = 0; t < nbr; t++) s = s + s_b[t][j];
                for (int t = 0; t < na; t++) s = s + sa(g, t, k, j == 0 ? 0 : 1);
            }
            s = s + (j == 0 ? s_ffg[g][2] : s_ffg[g][3]);
            if (j == 0) { if (olast && grad) mstore(M, &grad[istart + k], s); }
            else if (j - 1 < ocs && H) mstore(M, &H[(istart + k) + (ostart + j - 1) * ldH], s);
        }
        if (r == 0 && olast && ci == nchunks_total - 1 && value_out) {
            double s = 0.0;
            for (int t = 0; t < na; t++) s = s + s_av[g][t];
            for (int t = 0; t < nbr; t++) s = s + s_bv[t];
            mstore(M, value_out, s + s_ffg[g][0]);
        }
    }

    // ---- pairs around the diagonal, one at a time (at most three per block row of the pair grid) ----
    for (int g = 0; g < gcount; g++) {
        if (!diagonal(g)) continue;                              // block-uniform
        __syncthreads();                                         // the buffer is free: the batched fold / the previous pair is done
        const int64_t ci = cg0 + g;
        const bool ilast = (ci == nchunks_total - 1);
        const int ics = ilast ? lastchunksize : N;
        const int64_t istart = ilast ? n - lastchunksize : ci * N;
        int64_t ar0, ar1; a_range(g, ar0, ar1);
        int64_t a0 = ar0, a1 = ar1, b0 = br0, b1 = br1;
        a0 = a0 < b0 ? a0 : b0; a1 = a1 > b1 ? a1 : b1; b0 = b1 = a1;      // the ranges meet: one merged range
        const int na = (int)(a1 - a0), nnear = na;
        // phase 2: (term, inner lane k).  A term reading both windows: Dual{Dual{N},1} per k (the full nested evaluation) -> sp(t,k,0..N);
        // inner window only: Dual{Dual{1},1} per k -> sp(t,k,0..1); outer window only: Dual{Dual{N},1} once -> sp(t,0,0..N)
        for (int w = threadIdx.x; w < na * N; w += THREADS) {
            const int t = w / N, k = w - t * N;
            const int64_t i = a0 + t;
            const bool inA = (i >= ar0 && i < ar1), inB = (i >= br0 && i < br1);
            if (inA && !inB) {
                NearInnerLoader<NS> xa{x, istart, ics, k};
                DF a[1]; a[0].v.v = 0.0; a[0].v.p[0] = 0.0; a[0].p[0].v = 0.0; a[0].p[0].p[0] = 0.0;
                F::term(xa, i, n, a);
                sp(t, k, 0) = a[0].p[0].v; sp(t, k, 1) = a[0].p[0].p[0];
                if (k == 0) s_v[t] = a[0].v.v;
            } else if (inA || k == 0) {
                NearLoader<N, NS> xn{x, ostart, ocs, istart, ics, k};
                DG acc[1];
                acc[0].v.v = 0.0; acc[0].p[0].v = 0.0;
#pragma unroll
                for (int j = 0; j < N; j++) { acc[0].v.p[j] = 0.0; acc[0].p[0].p[j] = 0.0; }
                F::term(xn, i, n, acc);
                sp(t, k, 0) = acc[0].p[0].v;
#pragma unroll
                for (int j = 0; j < N; j++) sp(t, k, j + 1) = acc[0].p[0].p[j];
                if (k == 0) s_v[t] = acc[0].v.v;
            }
        }
        __syncthreads();
        // phase 3
        for (int e = threadIdx.x; e < N * (N + 1); e += THREADS) {
            const int k = e / (N + 1), j = e - k * (N + 1);     // j = 0: value of partials(y,k); j >= 1: outer partial j
            if (k >= ics) continue;
            double s = 0.0;
            for (int t = 0; t < nnear; t++) {
                const int64_t i = a0 + t;
                const bool inA = (i >= ar0 && i < ar1), inB = (i >= br0 && i < br1);
                s = s + ((inA && inB) ? sp(t, k, j) : (inA ? sp(t, k, j == 0 ? 0 : 1) : sp(t, 0, j)));
            }
            s = s + (j == 0 ? s_ffg[g][2] : s_ffg[g][3]);
            if (j == 0) { if (olast && grad) mstore(M, &grad[istart + k], s); }
            else if (j - 1 < ocs && H) mstore(M, &H[(istart + k) + (ostart + j - 1) * ldH], s);
        }
        if (threadIdx.x == 0 && olast && ilast && value_out) {
            double s = 0.0;
            for (int t = 0; t < nnear; t++) s = s + s_v[t];
            mstore(M, value_out, s + s_ffg[g][0]);
        }
    }
}

// ---- general form: K sum accumulators and arithmetic after the reductions (F::finalize), e.g. Ackley ---------------
// Same two-level lane sharing as above, but the block first assembles the K reduced accumulators as FULL nested duals
// Dual{Dual{N},N} -- far terms contribute their four shared lanes, near terms their per-inner-lane Dual{Dual{N},1}
// results, folded in term order -- and then runs F::finalize on them (scalar nested-dual arithmetic, once per block,
// one thread per inner partial lane), the Dual{T,Dual{T,V,N},N} evaluation of the tail of f in the reference
// (src/hessian.jl:14-19).
// Near terms are processed in rounds of R terms so that the staging buffer stays under 40 KB for any K.
template <class F, int N, bool NS, int THREADS>
FDB_DEV void hessian_general_body(const double* __restrict__ x, int64_t n, int64_t outer_lo, int64_t nchunks_total, int lastchunksize,
                                  double* __restrict__ H, int64_t ldH, double* __restrict__ grad, double* __restrict__ value_out,
                                  const fdb_mirrors& M) {
    typedef Dual<double, N, NS> DI;
    typedef Dual<DI, 1, NS> DG;
    typedef Dual<Dual<double, 1, NS>, 1, NS> DF;
    constexpr int K = F::K;
    constexpr int LANES = (N + 1) * (N + 1);                       // [a][b]: a = inner index (0 = value), b = outer index
    constexpr int RCAP = (40 * 1024) / (K * (N + 1) * (N + 1) * 8);
    constexpr int RT = THREADS / N;
    constexpr int R = RCAP < 1 ? 1 : (RCAP < RT ? RCAP : RT);      // near terms per round
    __shared__ double s_stage[K][R][N + 1][N + 1];                  // [kk][t][a][b] of one round (a = 0 row written by lane k = 0)
    __shared__ double s_acc[K][N + 1][N + 1];                       // running fold, later the assembled accumulators
    __shared__ double s_ff[K][4];
    __shared__ double s_red[THREADS / 32][K][4];

    const int64_t ci = blockIdx.x, co = outer_lo + blockIdx.y;
    const bool ilast = (ci == nchunks_total - 1), olast = (co == nchunks_total - 1);
    const int ics = ilast ? lastchunksize : N, ocs = olast ? lastchunksize : N;
    const int64_t istart = ilast ? n - lastchunksize : ci * N;
    const int64_t ostart = olast ? n - lastchunksize : co * N;
    const int64_t nt = F::nterms(n);
    int64_t a0 = istart - F::HALO, a1 = istart + ics, b0 = ostart - F::HALO, b1 = ostart + ocs;
    if (a0 < 0) a0 = 0; if (b0 < 0) b0 = 0; if (a1 > nt) a1 = nt; if (b1 > nt) b1 = nt;
    if (a1 < a0) a1 = a0; if (b1 < b0) b1 = b0;
    if (b0 <= a1 && a0 <= b1) { a0 = a0 < b0 ? a0 : b0; a1 = a1 > b1 ? a1 : b1; b0 = b1 = a1; }
    const int na = (int)(a1 - a0), nb = (int)(b1 - b0), nnear = na + nb;

    // phase 1: far terms on four shared lanes per accumulator
    DF accff[K];
#pragma unroll
    for (int kk = 0; kk < K; kk++) { accff[kk].v.v = 0.0; accff[kk].v.p[0] = 0.0; accff[kk].p[0].v = 0.0; accff[kk].p[0].p[0] = 0.0; }
    FFLoader<NS> xf{x};
    for (int64_t i = threadIdx.x; i < nt; i += THREADS) {
        const bool near = (i >= a0 && i < a1) || (i >= b0 && i < b1);
        if (!near) F::term(xf, i, n, accff);
    }
#pragma unroll
    for (int kk = 0; kk < K; kk++) {
        double r[4] = {accff[kk].v.v, accff[kk].v.p[0], accff[kk].p[0].v, accff[kk].p[0].p[0]};
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) r[q] = r[q] + shfl_down_d(r[q], off);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][kk][q] = r[q];
        }
    }
    for (int e = threadIdx.x; e < K * LANES; e += THREADS) (&s_acc[0][0][0])[e] = 0.0;
    __syncthreads();
    if (threadIdx.x < K * 4) {
        const int kk = threadIdx.x / 4, q = threadIdx.x % 4;
        double s = 0.0;
        for (int wq = 0; wq < THREADS / 32; wq++) s = s + s_red[wq][kk][q];
        s_ff[kk][q] = s;
    }
    // phase 2: near terms, R per round; work item = (term of the round, inner lane k)
    for (int t0 = 0; t0 < nnear; t0 += R) {
        const int rt = (nnear - t0 < R) ? nnear - t0 : R;
        if (threadIdx.x < rt * N) {
            const int tl = threadIdx.x / N, k = threadIdx.x - tl * N, t = t0 + tl;
            const int64_t i = t < na ? a0 + t : b0 + (t - na);
            NearLoader<N, NS> xn{x, ostart, ocs, istart, ics, k};
            DG acc[K];
#pragma unroll
            for (int kk = 0; kk < K; kk++) {
                acc[kk].v.v = 0.0; acc[kk].p[0].v = 0.0;
#pragma unroll
                for (int j = 0; j < N; j++) { acc[kk].v.p[j] = 0.0; acc[kk].p[0].p[j] = 0.0; }
            }
            F::term(xn, i, n, acc);
#pragma unroll
            for (int kk = 0; kk < K; kk++) {
                s_stage[kk][tl][k + 1][0] = acc[kk].p[0].v;
#pragma unroll
                for (int j = 0; j < N; j++) s_stage[kk][tl][k + 1][j + 1] = acc[kk].p[0].p[j];
                if (k == 0) {
                    s_stage[kk][tl][0][0] = acc[kk].v.v;
#pragma unroll
                    for (int j = 0; j < N; j++) s_stage[kk][tl][0][j + 1] = acc[kk].v.p[j];
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < K * LANES; e += THREADS) {        // fold this round in term order
            const int kk = e / LANES, ab = e - kk * LANES;
            double s = (&s_acc[kk][0][0])[ab];
            for (int tl = 0; tl < rt; tl++) s = s + (&s_stage[kk][tl][0][0])[ab];
            (&s_acc[kk][0][0])[ab] = s;
        }
        __syncthreads();
    }
    __syncthreads();
    // phase 3: add the shared far lanes, finalize, store.  A Dual{Dual{N},N} is N independent Dual{Dual{N},1} numbers
    // sharing their value part (the inner partial lanes never interact), so thread k runs F::finalize on the
    // accumulators restricted to inner lane k -- N threads in parallel on 2(N+1)-lane numbers instead of one thread on
    // (N+1)^2 lanes -- and stores row istart+k of the block: value(partials(y,k)) -> gradient, partials(partials(y,k),j) -> H.
    if (threadIdx.x < N) {
        const int k = threadIdx.x;
        DG accg[K];
#pragma unroll
        for (int kk = 0; kk < K; kk++) {
            accg[kk].v.v = s_acc[kk][0][0] + s_ff[kk][0];
            accg[kk].p[0].v = s_acc[kk][k + 1][0] + s_ff[kk][2];
#pragma unroll
            for (int j = 0; j < N; j++) {
                accg[kk].v.p[j] = s_acc[kk][0][j + 1] + s_ff[kk][1];
                accg[kk].p[0].p[j] = s_acc[kk][k + 1][j + 1] + s_ff[kk][3];
            }
        }
        const DG y = F::finalize(accg, n);
        if (k < ics) {
            if (olast && grad) mstore(M, &grad[istart + k], y.p[0].v);
            if (H) {
#pragma unroll
                for (int j = 0; j < N; j++)
                    if (j < ocs) mstore(M, &H[(istart + k) + (ostart + j) * ldH], y.p[0].p[j]);
            }
        }
        if (k == 0 && olast && ilast && value_out) mstore(M, value_out, y.v.v);
    }
}

template <class F, int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) hessian_sweep_kernel(const double* __restrict__ x, int64_t n, int64_t outer_lo,
                                                                int64_t nchunks_total, int lastchunksize, double* __restrict__ H,
                                                                int64_t ldH, double* __restrict__ grad, double* __restrict__ value_out,
                                                                const __grid_constant__ fdb_mirrors M) {
    if constexpr (F::PLAIN_SUM) hessian_sweep_body<F, N, NS, THREADS>(x, n, outer_lo, nchunks_total, lastchunksize, H, ldH, grad, value_out, M);
    else hessian_general_body<F, N, NS, THREADS>(x, n, outer_lo, nchunks_total, lastchunksize, H, ldH, grad, value_out, M);
}

}  // namespace fdb
