// fdb_sweep_kernel.cuh -- the batched chunk-mode gradient kernel (see fdb_sweep.cu), as a header so that the
// run-time specialiser (fdb_jit.cu) instantiates the very same code for a user function's generated functor.
#pragma once
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

// SHARE = lane sharing.  A term whose inputs all lie outside the seeded window sees
// Duals with N identical (zero) partial lanes, so its N partial lanes are N copies of
// one IEEE computation.  With SHARE the kernel evaluates such terms on Dual{1}
// (value + one representative partial lane) into a separate accumulator and adds the
// representative lane to all N lanes at the end; results are bit-identical lane by lane
// (including NaN/Inf propagation, since the representative lane is really computed),
// only the summation order changes, which the reduction tolerance already covers.
// Every term of f is still evaluated in every chunk.  SHARE=false evaluates all N
// lanes for every term (bench.py reports both).
// A term's inputs held in registers: the far loops below fetch x for the NEXT group of terms before they evaluate the
// current group, so the L2 latency of the x stream (x is L2-resident, every block re-reads all of it: ncu showed as many
// long-scoreboard stalls as FP64-pipe stalls) is hidden behind FP64 work of the same warp.  The functor is called with the
// literal term index 0, so x(0 + b) becomes register v[b]; only functors whose term depends on i through x(i + b) alone
// (F::PURE_X) take this path.
template <int H, bool NS> struct RegLoader {
    double v[H + 1];
    FDB_DEV void fetch(const double* __restrict__ x, int64_t i) {
#pragma unroll
        for (int b = 0; b <= H; b++) v[b] = x[i + b];
    }
    FDB_DEV Dual<double, 1, NS> operator()(int64_t j) const {
        Dual<double, 1, NS> d; d.v = v[j]; d.p[0] = 0.0; return d;
    }
};

// far terms of iterations [it0, it1) (term index = it * THREADS + threadIdx.x, all valid), Dual{1} accumulators
template <class F, bool NS, int THREADS>
FDB_DEV void far_range(const double* __restrict__ x, int64_t n, int64_t it0, int64_t it1, Dual<double, 1, NS> (&acc1)[F::K]) {
    constexpr int U = 4;
    const int64_t t = threadIdx.x;
    int64_t it = it0;
    if constexpr (F::PURE_X) {
        RegLoader<F::HALO, NS> a[U], b[U];
        if (it + U <= it1) {
#pragma unroll
            for (int u = 0; u < U; u++) a[u].fetch(x, (it + u) * THREADS + t);
        }
        // two groups per trip: while group A is evaluated group B's loads are in flight, and vice versa
        for (; it + 3 * U <= it1; it += 2 * U) {
#pragma unroll
            for (int u = 0; u < U; u++) b[u].fetch(x, (it + U + u) * THREADS + t);
#pragma unroll
            for (int u = 0; u < U; u++) F::term(a[u], (int64_t)0, n, acc1);
#pragma unroll
            for (int u = 0; u < U; u++) a[u].fetch(x, (it + 2 * U + u) * THREADS + t);
#pragma unroll
            for (int u = 0; u < U; u++) F::term(b[u], (int64_t)0, n, acc1);
        }
        if (it + U <= it1) {                       // group A is loaded and not yet evaluated
#pragma unroll
            for (int u = 0; u < U; u++) F::term(a[u], (int64_t)0, n, acc1);
            it += U;
        }
    }
    ZeroLoader<1, NS> xz{x};
    if constexpr (!F::PURE_X) {
#pragma unroll 4
        for (; it < it1; it++) F::term(xz, it * THREADS + t, n, acc1);
    } else {
        for (; it < it1; it++) F::term(xz, it * THREADS + t, n, acc1);
    }
}

// pass over all terms of f for the chunk window [start, start+chunksize): this thread's partial accumulators
template <class F, int N, bool NS, int THREADS, bool SHARE>
FDB_DEV void sweep_terms(const double* __restrict__ x, int64_t n, int64_t start, int chunksize, Dual<double, N, NS> (&acc)[F::K]) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    SeedLoader<N, NS> xs{x, start, chunksize};
    ZeroLoader<1, NS> xz{x};
    D1 acc1[F::K];
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
    D1 like1; like1.v = 0.0; like1.p[0] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) { acc[k] = F::init(k, like); acc1[k] = F::init(k, like1); }
    const int64_t nt = F::nterms(n);
    if (SHARE) {
        // iterations [0,itA) and [itB,nit) touch no seeded position anywhere in the block
        const int64_t nit = (nt + THREADS - 1) / THREADS;
        int64_t itA = (start - F::HALO) / THREADS; if (itA < 0) itA = 0; if (itA > nit) itA = nit;
        int64_t itB = (start + chunksize + THREADS - 1) / THREADS; if (itB > nit) itB = nit; if (itB < itA) itB = itA;
        far_range<F, NS, THREADS>(x, n, 0, itA, acc1);
        for (int64_t it = itA; it < itB; it++) {
            const int64_t i = it * THREADS + threadIdx.x;
            if (i < nt) F::term(xs, i, n, acc);
        }
        const int64_t full = nt / THREADS;      // iterations whose every lane is a valid term
        far_range<F, NS, THREADS>(x, n, itB, full, acc1);
        for (int64_t it = (itB > full ? itB : full); it < nit; it++) {
            const int64_t i = it * THREADS + threadIdx.x;
            if (i < nt) F::term(xz, i, n, acc1);
        }
        // fold the shared accumulator into all N lanes
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D e; e.v = acc1[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) e.p[j] = acc1[k].p[0];
            // for K > 1 or non-additive combines the neutral value lane must not be counted twice:
            // acc[k] started from init, acc1[k] too; combine(init, z) == z for sum and prod.
            acc[k] = F::combine(k, acc[k], e);
        }
    } else {
        for (int64_t i = threadIdx.x; i < nt; i += THREADS) F::term(xs, i, n, acc);
    }
}

template <class F, int N, bool NS, int THREADS, bool SHARE>
FDB_DEV void gradient_sweep_body(const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                 double* __restrict__ grad, double* __restrict__ value_out, const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c = chunk_lo + blockIdx.x;                       // 0-based chunk number
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;               // src/gradient.jl:123
    const int64_t start = last ? (n - lastchunksize) : c * N;     // lastchunkindex-1, src/gradient.jl:124
    D acc[F::K];
    sweep_terms<F, N, NS, THREADS, SHARE>(x, n, start, chunksize, acc);
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize(acc, n);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore(M, &grad[start + k], y.p[k]);   // extract_gradient_chunk!, src/gradient.jl:74-81 (+ peers)
        if (last && value_out) mstore(M, value_out, y.v);          // extract_value!, src/gradient.jl:155
    }
}

// Opt-in variant (fdb_set_chunk_grouping): a block sweeps `group` consecutive chunks.  A term that reads no seeded position of ANY
// of them has the same Dual{1} result (value, representative zero lane) for every chunk of the group, so the far iterations are
// evaluated once per block and folded into each chunk's accumulators; the iterations that hold the group's windows are evaluated
// per chunk on all N lanes with that chunk's seeds, exactly as above.  The gradient is the same lane by lane (the far lanes are
// really computed, 0 * Inf still poisons), the FP64 work drops by ~group.  The default stays one chunk per block: the ceil(n/N)
// sweeps of src/gradient.jl:141-152 are then independent evaluations of f, which is what BASELINE.json's C2 number measures.
template <class F, int N, bool NS, int THREADS>
FDB_DEV void gradient_sweep_group_body(const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t chunk_hi, int64_t nchunks_total,
                                       int lastchunksize, int group, double* __restrict__ grad, double* __restrict__ value_out,
                                       const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c0 = chunk_lo + (int64_t)blockIdx.x * group;
    const int gcount = (int)((chunk_hi - c0) < group ? (chunk_hi - c0) : group);
    auto cstart = [&](int64_t c) { return c == nchunks_total - 1 ? n - lastchunksize : c * N; };
    auto csize = [&](int64_t c) { return c == nchunks_total - 1 ? lastchunksize : N; };
    const int64_t ustart = cstart(c0), uend = cstart(c0 + gcount - 1) + csize(c0 + gcount - 1);
    const int64_t nt = F::nterms(n);
    const int64_t nit = (nt + THREADS - 1) / THREADS;
    int64_t itA = (ustart - F::HALO) / THREADS; if (itA < 0) itA = 0; if (itA > nit) itA = nit;
    int64_t itB = (uend + THREADS - 1) / THREADS; if (itB > nit) itB = nit; if (itB < itA) itB = itA;
    ZeroLoader<1, NS> xz{x};
    D1 acc1[F::K];
    D1 like1; like1.v = 0.0; like1.p[0] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) acc1[k] = F::init(k, like1);
    // iterations [0,itA) and [itB,nit) touch no seeded position of any chunk of the group
    far_range<F, NS, THREADS>(x, n, 0, itA, acc1);
    const int64_t full = nt / THREADS;
    far_range<F, NS, THREADS>(x, n, itB, full, acc1);
    for (int64_t it = (itB > full ? itB : full); it < nit; it++) {
        const int64_t i = it * THREADS + threadIdx.x;
        if (i < nt) F::term(xz, i, n, acc1);
    }
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
    for (int g = 0; g < gcount; g++) {
        if (g) __syncthreads();                                    // block_combine's staging is reused
        const int64_t c = c0 + g;
        const bool last = (c == nchunks_total - 1);
        const int chunksize = last ? lastchunksize : N;
        const int64_t start = last ? (n - lastchunksize) : c * N;
        SeedLoader<N, NS> xs{x, start, chunksize};
        D acc[F::K];
#pragma unroll
        for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
        for (int64_t it = itA; it < itB; it++) {
            const int64_t i = it * THREADS + threadIdx.x;
            if (i < nt) F::term(xs, i, n, acc);
        }
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D e; e.v = acc1[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) e.p[j] = acc1[k].p[0];
            acc[k] = F::combine(k, acc[k], e);
        }
        block_combine<F, D, THREADS>(acc);
        if (threadIdx.x == 0) {
            D y = F::finalize(acc, n);
#pragma unroll
            for (int k = 0; k < N; k++)
                if (k < chunksize) mstore(M, &grad[start + k], y.p[k]);
            if (last && value_out) mstore(M, value_out, y.v);
        }
    }
}

// Two-pass form: f uses a reduced value inside a second broadcast -- sum((x .- mean(x)).^2), log(sum(exp.(x .- m))) ...
// Pass 1 is the sweep above (lane sharing included) over the first-stage terms; thread 0 runs F::finalize1 on the reduced
// Duals and publishes the resulting scalar Duals S (dense partials: every lane of the window contributes) in shared
// memory; pass 2 evaluates the second-stage terms, which read S, on all N lanes for every term (no sharing: S is not a
// zero-partial operand); F::finalize2 closes.  Used by specialised op-programs only (fdb_jit.cu).
template <class F, int N, bool NS, int THREADS, bool SHARE>
FDB_DEV void gradient_sweep2_body(const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                  double* __restrict__ grad, double* __restrict__ value_out, const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c = chunk_lo + blockIdx.x;                       // 0-based chunk number
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;               // src/gradient.jl:123
    const int64_t start = last ? (n - lastchunksize) : c * N;     // lastchunkindex-1, src/gradient.jl:124
    __shared__ D s_S[F::NSCAL];
    D acc[F::K];
    sweep_terms<F, N, NS, THREADS, SHARE>(x, n, start, chunksize, acc);
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) F::finalize1(acc, s_S);
    __syncthreads();
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) acc[k] = F::init(k, like);
    SeedLoader<N, NS> xs{x, start, chunksize};
    const int64_t nt = F::nterms(n);
    for (int64_t i = threadIdx.x; i < nt; i += THREADS) F::term2(xs, i, n, s_S, acc);
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize2(acc, s_S, n);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore(M, &grad[start + k], y.p[k]);
        if (last && value_out) mstore(M, value_out, y.v);
    }
}

// ---- structured inputs: gradient(f, x::UpperTriangular / LowerTriangular / Diagonal) on the batched sweep --------------------------
// Only structurally non-zero positions are seeded and extracted (src/apiutils.jl:44-71, test/GradientTest.jl:250-276); f sees the
// rows x rows column-major storage whose off-structure entries are zero(Dual).  Chunks run over STRUCTURAL positions; the seeded
// window of a chunk is the storage span [idx(start), idx(start+chunksize-1)], which is what the far / near split works on.
template <int N, bool NS> struct StructSeedLoader {
    const double* x; long start; int chunksize; int kind; long rows;
    FDB_DEV Dual<double, N, NS> operator()(int64_t i) const {
        const long pos = struct_pos_of(kind, rows, (long)i);
        return seeded<N, NS>(x[i], pos < 0 ? -1000000 : pos, start, chunksize);       // off-structure: Dual(x[i] = 0, zero partials)
    }
};
template <class F, int N, bool NS, int THREADS, bool SHARE>
FDB_DEV void gradient_sweep_struct_body(const double* __restrict__ x, int64_t nx, int kind, int64_t rows, int64_t nstruct, int64_t chunk_lo,
                                        int64_t nchunks_total, int lastchunksize, double* __restrict__ grad, double* __restrict__ value_out,
                                        const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    const int64_t c = chunk_lo + blockIdx.x;
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (nstruct - lastchunksize) : c * N;               // structural position of the window
    const int64_t w0 = struct_index_of(kind, rows, start), w1 = struct_index_of(kind, rows, start + chunksize - 1) + 1;   // storage span
    StructSeedLoader<N, NS> xs{x, (long)start, chunksize, kind, (long)rows};
    ZeroLoader<1, NS> xz{x};
    D acc[F::K]; D1 acc1[F::K];
    D like; like.v = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) like.p[k] = 0.0;
    D1 like1; like1.v = 0.0; like1.p[0] = 0.0;
#pragma unroll
    for (int k = 0; k < F::K; k++) { acc[k] = F::init(k, like); acc1[k] = F::init(k, like1); }
    const int64_t nt = F::nterms(nx);
    if (SHARE) {
        const int64_t nit = (nt + THREADS - 1) / THREADS;
        int64_t itA = (w0 - F::HALO) / THREADS; if (itA < 0) itA = 0; if (itA > nit) itA = nit;
        int64_t itB = (w1 + THREADS - 1) / THREADS; if (itB > nit) itB = nit; if (itB < itA) itB = itA;
        for (int64_t it = 0; it < nit; it++) {
            const int64_t i = it * THREADS + threadIdx.x;
            if (i >= nt) continue;
            if (it >= itA && it < itB) F::term(xs, i, nx, acc); else F::term(xz, i, nx, acc1);
        }
#pragma unroll
        for (int k = 0; k < F::K; k++) {
            D e; e.v = acc1[k].v;
#pragma unroll
            for (int j = 0; j < N; j++) e.p[j] = acc1[k].p[0];
            acc[k] = F::combine(k, acc[k], e);
        }
    } else {
        for (int64_t i = threadIdx.x; i < nt; i += THREADS) F::term(xs, i, nx, acc);
    }
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) {
        D y = F::finalize(acc, nx);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore(M, &grad[struct_index_of(kind, rows, start + k)], y.p[k]);    // structural_eachindex(result), gradient.jl:76
        if (last && value_out) mstore(M, value_out, y.v);
    }
}

// Array-valued two-pass form: y = g.(x, reduced values), e.g. softmax exp.(x) ./ sum(exp.(x)) or (x .- mean(x)) ./ std.  The
// Jacobian is dense (every output depends on every input through the reduced Duals), so one block sweeps one chunk: pass 1
// reduces (lane sharing included), finalize1 publishes the scalar Duals S, pass 2 evaluates y[i] on all N lanes for every row
// and stores J[i, start+k] = partials(y[i], k) straight from registers (extract_jacobian_chunk!, src/jacobian.jl:109-118).
template <class F, int N, bool NS, int THREADS, bool SHARE, bool MIR>
FDB_DEV void jacobian_sweep2_body(const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                  double* __restrict__ J, int64_t ldJ, double* __restrict__ yout, const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    const int64_t c = chunk_lo + blockIdx.x;
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    __shared__ D s_S[F::NSCAL];
    D acc[F::K];
    sweep_terms<F, N, NS, THREADS, SHARE>(x, n, start, chunksize, acc);
    block_combine<F, D, THREADS>(acc);
    if (threadIdx.x == 0) F::finalize1(acc, s_S);
    __syncthreads();
    SeedLoader<N, NS> xs{x, start, chunksize};
    const int64_t m = n - F::HALO;
    for (int64_t i = threadIdx.x; i < m; i += THREADS) {
        const D y = F::template apply2<D>(xs, i, n, s_S);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[k]);
        if (yout && last) mstore(M, &yout[i], y.v);
    }
}

template <class F, int N, bool NS, int THREADS, bool SHARE>
__global__ void __launch_bounds__(THREADS) gradient_sweep_kernel(const double* __restrict__ x, int64_t n, int64_t chunk_lo,
                                                                 int64_t nchunks_total, int lastchunksize,
                                                                 double* __restrict__ grad, double* __restrict__ value_out,
                                                                 const __grid_constant__ fdb_mirrors M) {
    gradient_sweep_body<F, N, NS, THREADS, SHARE>(x, n, chunk_lo, nchunks_total, lastchunksize, grad, value_out, M);
}

template <class F, int N, bool NS, int THREADS>
__global__ void __launch_bounds__(THREADS) gradient_sweep_group_kernel(const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t chunk_hi,
                                                                       int64_t nchunks_total, int lastchunksize, int group,
                                                                       double* __restrict__ grad, double* __restrict__ value_out,
                                                                       const __grid_constant__ fdb_mirrors M) {
    gradient_sweep_group_body<F, N, NS, THREADS>(x, n, chunk_lo, chunk_hi, nchunks_total, lastchunksize, group, grad, value_out, M);
}

}  // namespace fdb
