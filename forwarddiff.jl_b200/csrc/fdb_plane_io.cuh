// fdb_plane_io.cuh -- SoA plane IO shared by the elementwise and reduction kernels: one thread owns two
// consecutive elements and reads / writes each of the N+1 planes with one 16-byte access.
#pragma once
#include "fdb_dual.cuh"

namespace fdb {

// ---- plane IO: two consecutive elements per thread -----------------------------------
template <int N, bool NS> struct Dual2 { Dual<double, N, NS> a, b; };

// `bcast` arrays have a single element that every position reads
template <int N, bool NS>
FDB_DEV Dual2<N, NS> load2(const double* __restrict__ base, int64_t ld, int64_t e, bool pair, bool bcast, bool vec) {
    Dual2<N, NS> r;
    if (bcast) {
        r.a.v = base[0];
#pragma unroll
        for (int k = 0; k < N; k++) r.a.p[k] = base[(int64_t)(k + 1) * ld];
        r.b = r.a;
    } else if (pair && vec) {
        double2 t = __ldg(reinterpret_cast<const double2*>(base + e));
        r.a.v = t.x; r.b.v = t.y;
#pragma unroll
        for (int k = 0; k < N; k++) {
            double2 u = __ldg(reinterpret_cast<const double2*>(base + (int64_t)(k + 1) * ld + e));
            r.a.p[k] = u.x; r.b.p[k] = u.y;
        }
    } else {
        r.a.v = base[e];
#pragma unroll
        for (int k = 0; k < N; k++) r.a.p[k] = base[(int64_t)(k + 1) * ld + e];
        r.b = r.a;
        if (pair) {            // unaligned view: scalar loads for the second element too
            r.b.v = base[e + 1];
#pragma unroll
            for (int k = 0; k < N; k++) r.b.p[k] = base[(int64_t)(k + 1) * ld + e + 1];
        }
    }
    return r;
}
FDB_DEV void loadreal2(const double* __restrict__ base, int64_t e, bool pair, bool bcast, bool vec, double& a, double& b) {
    if (bcast) { a = b = base[0]; }
    else if (pair && vec) { double2 t = __ldg(reinterpret_cast<const double2*>(base + e)); a = t.x; b = t.y; }
    else { a = b = base[e]; if (pair) b = base[e + 1]; }
}
template <int N, bool NS>
FDB_DEV void store2(double* __restrict__ base, int64_t ld, int64_t e, bool pair, bool vec, const Dual<double, N, NS>& a, const Dual<double, N, NS>& b) {
    if (pair && vec) {
        *reinterpret_cast<double2*>(base + e) = make_double2(a.v, b.v);
#pragma unroll
        for (int k = 0; k < N; k++) *reinterpret_cast<double2*>(base + (int64_t)(k + 1) * ld + e) = make_double2(a.p[k], b.p[k]);
    } else {
        base[e] = a.v;
#pragma unroll
        for (int k = 0; k < N; k++) base[(int64_t)(k + 1) * ld + e] = a.p[k];
        if (pair) {
            base[e + 1] = b.v;
#pragma unroll
            for (int k = 0; k < N; k++) base[(int64_t)(k + 1) * ld + e + 1] = b.p[k];
        }
    }
}

}  // namespace fdb
