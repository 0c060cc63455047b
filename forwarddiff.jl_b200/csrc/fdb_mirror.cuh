// fdb_mirror.cuh -- result stores of the whole-path drivers.  With a peer-memory exchange enabled
// (fdb_exchange.cu) every extracted entry is written to the local arena and, by the same thread, to the same
// offset of every peer's arena over NVLink: extract_gradient_chunk! / extract_jacobian_chunk!
// (src/gradient.jl:74-81, src/jacobian.jl:109-118) double as the all-gather.
#pragma once
#include "fdb_internal.h"

namespace fdb {

__device__ __forceinline__ void mstore(const fdb_mirrors& M, double* p, double v) {
    *p = v;
    for (int r = 0; r < M.count; r++) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[r]) = v;
}
// streaming (evict-first) local store for write-once result matrices
__device__ __forceinline__ void mstore_cs(const fdb_mirrors& M, double* p, double v) {
    __stcs(p, v);
    for (int r = 0; r < M.count; r++) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + M.delta[r]) = v;
}
__device__ __forceinline__ void mstore2(const fdb_mirrors& M, double* p, double2 v) {
    *reinterpret_cast<double2*>(p) = v;
    for (int r = 0; r < M.count; r++) *reinterpret_cast<double2*>(reinterpret_cast<char*>(p) + M.delta[r]) = v;
}

}  // namespace fdb
