// fdb_jac_kernel.cuh -- Jacobian chunk sweeps of an elementwise / stencil array function
//     y[i] = F::apply(x[i .. i+HALO]),  i in [0, n - HALO)         (the f(x)-form sweep, src/jacobian.jl:169-216)
// with J[i, start+k] = partials(y[i], k) stored straight from registers (extract_jacobian_chunk!,
// src/jacobian.jl:109-118).  Instantiated by the run-time specialiser (fdb_jit.cu) for a user function's
// generated functor; the interpreter twin is program_jacobian_kernel (fdb_program_jac.cu).
#pragma once
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

template <class F, int N, bool NS, int THREADS, bool MIR>
FDB_DEV void elementwise_jacobian_body(const bool SHARE, const double* __restrict__ x, int64_t n, int64_t chunk_lo, int64_t nchunks_total,
                                       int lastchunksize, double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,
                                       const unsigned nrb, const unsigned nch, const bool rowfast, const fdb_mirrors& M) {
    typedef Dual<double, N, NS> D;
    typedef Dual<double, 1, NS> D1;
    // 1-D grid; row block fastest when the stores also go to peers (see brusselator_jacobian_kernel)
    const int64_t c = chunk_lo + (rowfast ? blockIdx.x / nrb : blockIdx.x % nch);
    const bool last = (c == nchunks_total - 1);
    const int chunksize = last ? lastchunksize : N;
    const int64_t start = last ? (n - lastchunksize) : c * N;
    const int64_t i = (int64_t)(rowfast ? blockIdx.x % nrb : blockIdx.x / nch) * THREADS + threadIdx.x;
    const int64_t m = n - F::HALO;
    const int64_t w0 = i - (threadIdx.x & 31);
    // warp-uniform: rows w0 .. w0+31 read no seeded position -> one representative partial lane (lane sharing)
    const bool far = SHARE && ((w0 + 31 + F::HALO < start) || (w0 >= start + chunksize));
    if (i >= m) return;
    if (far) {
        const D1 y = F::template apply<D1>(ZeroLoader<1, NS>{x}, i);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[0]);
        if (yout && last) mstore(M, &yout[i], y.v);
    } else {
        const D y = F::template apply<D>(SeedLoader<N, NS>{x, start, chunksize}, i);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[k]);
        if (yout && last) mstore(M, &yout[i], y.v);
    }
}

// ---- elementwise function of a matrix-vector product: y = g.(A * x) -----------------------------------------------------------
// `A * xdual` with one-hot seeds is a Dual vector z whose value is (A*x)[i] and whose partial k is A[i, start+k]
// (Real*Dual then Dual+Dual per element, src/dual.jl:476-490).  The contraction runs on the FP64 tensor cores (fdb_dmma.cu) and
// leaves per chunk either  [value lane | representative zero-partial lane]  (lane sharing: the seeded window column of A is added
// here, as the reference's last `+=` would)  or  [value lane | one column per seeded position]  (all lanes).  ProductLoader
// rebuilds z_i from that product; the user's elementwise program g then runs on it exactly like F::apply on a seeded x.
template <int N, bool NS> struct ProductLoader {
    const double* Yv;      // value lane of this chunk's product
    const double* Yl;      // lane sharing: the representative zero lane; all lanes: product column of the chunk's first seeded position
    const double* Aw;      // lane sharing: A(:, start) (window columns of A, stride lda); all lanes: nullptr
    int64_t ldy, lda; int chunksize;
    FDB_DEV Dual<double, N, NS> operator()(int64_t i) const {
        Dual<double, N, NS> d; d.v = Yv[i];
        if (Aw) {
            const double zl = Yl[i];
#pragma unroll
            for (int k = 0; k < N; k++) d.p[k] = (k < chunksize) ? zl + Aw[i + (int64_t)k * lda] * 1.0 : zl;
        } else {
#pragma unroll
            for (int k = 0; k < N; k++) d.p[k] = (k < chunksize) ? Yl[i + (int64_t)k * ldy] : 0.0;
        }
        return d;
    }
};

// value/zero lanes: Y column 2*cl, 2*cl+1 (SHARE) ; value: Y column 0, partial lanes: Y column 1 + (start - c0) + k (all lanes)
template <class F, int N, bool NS, int THREADS, bool MIR>
FDB_DEV void matvec_jacobian_body(const bool SHARE, const double* __restrict__ Y, int64_t ldy, const double* __restrict__ A, int64_t lda,
                                  int64_t m, int64_t n, int64_t chunk_lo, int64_t nchunks_total, int lastchunksize, double* __restrict__ J,
                                  int64_t ldJ, double* __restrict__ yout, const unsigned nrb, const unsigned nch, const bool rowfast,
                                  const fdb_mirrors& M) {
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
    const D y = F::template apply<D>(z, i);
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < chunksize) mstore_cs<MIR>(M, J + i + (start + k) * ldJ, y.p[k]);
    if (yout && last) mstore(M, &yout[i], y.v);
}

}  // namespace fdb
