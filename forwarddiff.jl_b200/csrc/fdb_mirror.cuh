// fdb_mirror.cuh -- result stores of the whole-path drivers.  With a peer-memory exchange enabled
// (fdb_exchange.cu) every extracted entry is written to the local arena and, by the same thread, to the same
// offset of every peer's arena over NVLink: extract_gradient_chunk! / extract_jacobian_chunk!
// (src/gradient.jl:74-81, src/jacobian.jl:109-118) double as the all-gather.
#pragma once
#ifndef FDB_MAX_RANKS
#define FDB_MAX_RANKS 8      /* include/fdb200.h */
#endif

// byte offsets from the local exchange arena to every peer's arena; passed by value to the driver kernels,
// which repeat each result store at p + delta[r].  count == 0: single rank / exchange off.
// mc != 0 (NVSwitch multicast, fdb_multicast.cu): count == 1 and delta[0] leads to the MULTICAST mapping of all arenas -- one
// store there is replicated by the switch into every rank's arena, this rank's included, so the local store is dropped.
// (`multimem.st` is an ordinary STG in SASS: what makes the store a multicast is the mapping behind the address, so kernels
// that know nothing about `mc` -- they store locally and at delta[0] -- stay correct, they only write the local copy twice.)
struct fdb_mirrors {
    int count = 0;
    int mc = 0;
    long long delta[FDB_MAX_RANKS - 1] = {0, 0, 0, 0, 0, 0, 0};
};

namespace fdb {

// The mirrors travel as a __grid_constant__ kernel parameter, so delta[r] is an indexed constant-bank read
// (no local-memory copy); the peer loops are not unrolled (they wait on NVLink, not on issue).
__device__ __forceinline__ void multimem_st(double* p, double v) {
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void mstore(const fdb_mirrors& M, double* p, double v) {
    if (M.mc) { multimem_st(reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[0]), v); return; }
    *p = v;
#pragma unroll 1
    for (int r = 0; r < M.count; r++) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[r]) = v;
}
// streaming (evict-first) local store for write-once result matrices.  MIR is a template flag in the
// write-bound Jacobian kernels so that the single-GPU instantiation is exactly the plain store.
template <bool MIR>
__device__ __forceinline__ void mstore_cs(const fdb_mirrors& M, double* p, double v) {
    if (MIR && M.mc) { multimem_st(reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[0]), v); return; }
    __stcs(p, v);
    if (MIR) {
#pragma unroll 1
        for (int r = 0; r < M.count; r++) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[r]) = v;
    }
}

}  // namespace fdb
