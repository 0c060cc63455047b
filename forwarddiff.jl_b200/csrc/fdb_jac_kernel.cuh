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

}  // namespace fdb
