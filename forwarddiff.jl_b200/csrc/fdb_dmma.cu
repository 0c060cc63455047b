// fdb_dmma.cu -- FP64 tensor-core (DMMA) contraction for Dual matvec / matmul.
//
// `A * xdual` on a Vector{Dual{T,Float64,N}} (Julia: LinearAlgebra.generic_matvecmul!, Real*Dual then
// Dual+Dual per element) is a dense contraction of A against the planes [value | partials]:
//      Y[m x P] = A[m x k] * X[k x P],   P = N+1 columns per chunk, batched over chunks.
// This is the only tensor-core kernel in the engine.  FP64 has no tcgen05 kind; the FP64 tensor path
// on sm_100a is `mma.sync.aligned.m8n8k4.f64` (SASS DMMA).  Operand tiles are staged into shared
// memory by the TMA engine with bulk asynchronous copies (`cp.async.bulk ... mbarrier::complete_tx`,
// SASS UBLKCP) issued by one producer thread into a 4-stage mbarrier ring; eight consumer warps each own
// a 32x32 accumulator tile in registers.  Rows of the shared tiles are padded (A: 132 doubles per
// k-row, B: 36 doubles per column) so the m8n8k4 fragment loads are bank-conflict free.
//
// Arithmetic note: DMMA accumulates with fused multiply-adds in k order, the reference with separate
// multiply and add; both are summation-order variants of the same dot product (tolerance 1e-12
// relative, BASELINE.json north_star).
#include <cuda.h>

#include "fdb_internal.h"
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"

namespace fdb {

constexpr int DM_BM = 128, DM_BN = 64, DM_KT = 32, DM_STAGES = 4;
constexpr int DM_BS = DM_KT + 4;          // padded column of the packed B tile (doubles): stride = 8 mod 32 words
constexpr int DM_A_BYTES = DM_KT * DM_BM * 8;            // dense [KT][BM] box written by one tensor copy
constexpr int DM_B_ELEMS = DM_BN * DM_BS, DM_B_BYTES = DM_B_ELEMS * 8;
constexpr int DM_STAGE_BYTES = DM_A_BYTES + DM_B_BYTES;
constexpr int DM_SMEM = DM_STAGES * DM_STAGE_BYTES + 2 * DM_STAGES * 8 + 128;

FDB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FDB_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FDB_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FDB_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FDB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (1-D), completion counted in bytes on an mbarrier
FDB_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA tensor copy (2-D box of the column-major A described by a CUtensorMap) global -> shared
FDB_DEV void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
FDB_DEV void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Pack B (k x ncols, column-major) into the tile order the kernel consumes: tile (j, s) holds columns
// [64j, 64j+64) x k-rows [32s, 32s+32) as 64 padded columns of 36 doubles, contiguous, so that one
// bulk copy stages it.  Out-of-range entries are zero.  Column c of B is read from column c % bmod when
// bmod > 0 (a batch of chunks sharing identical operand columns).
__global__ void pack_b_kernel(const double* __restrict__ B, int64_t ldb, int64_t k, int64_t ncols, int bmod, int nslab,
                              double* __restrict__ Bp) {
    const int s = blockIdx.x, j = blockIdx.y;
    double* tile = Bp + ((int64_t)j * nslab + s) * DM_B_ELEMS;
    for (int e = threadIdx.x; e < DM_B_ELEMS; e += blockDim.x) {
        const int c = e / DM_BS, kk = e - c * DM_BS;
        const int64_t col = (int64_t)j * DM_BN + c, row = (int64_t)s * DM_KT + kk;
        double v = 0.0;
        if (kk < DM_KT && col < ncols && row < k) v = B[row + (bmod > 0 ? col % bmod : col) * ldb];
        tile[e] = v;
    }
}

// C[m x ncols] = A[m x k] * B[k x ncols]; A through a tensor map (box 128 x 32, zero fill out of bounds),
// B pre-packed (pack_b_kernel); btiles = number of distinct packed column tiles (1 when every tile is identical).
__global__ void __launch_bounds__(288, 1) dmma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const double* __restrict__ Bp, int btiles,
                                                          double* __restrict__ C, int64_t ldc, int64_t m, int64_t k, int64_t ncols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DM_STAGES * DM_STAGE_BYTES);
    uint64_t* empty = full + DM_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * DM_BN, m0 = (int64_t)blockIdx.y * DM_BM;
    const int nslab = (int)((k + DM_KT - 1) / DM_KT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < DM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ---- producer: one thread, two TMA copies per k-slab -----------------------------------
        if (lane == 0) {
            const double* Bt = Bp + (int64_t)(blockIdx.x % btiles) * nslab * DM_B_ELEMS;
            for (int s = 0; s < nslab; s++) {
                const int st = s % DM_STAGES;
                if (s >= DM_STAGES) mbar_wait(&empty[st], ((s / DM_STAGES) - 1) & 1);
                unsigned char* stage = smem + st * DM_STAGE_BYTES;
                mbar_expect_tx(&full[st], DM_STAGE_BYTES);
                tma_load_2d(stage, &tmA, (int)m0, s * DM_KT, &full[st]);                         // A tile: [32 k][128 m]
                bulk_g2s(stage + DM_A_BYTES, Bt + (int64_t)s * DM_B_ELEMS, DM_B_BYTES, &full[st]);   // packed B tile
            }
        }
    } else {
        // ---- consumers: 8 warps, 4 (M) x 2 (N), 32x32 accumulators each -----------------------
        const int wm = warp & 3, wn = warp >> 2;
        const int g = lane >> 2, t = lane & 3;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int s = 0; s < nslab; s++) {
            const int st = s % DM_STAGES;
            mbar_wait(&full[st], (s / DM_STAGES) & 1);
            const double* As = reinterpret_cast<const double*>(smem + st * DM_STAGE_BYTES);
            const double* Bs = As + DM_KT * DM_BM;
#pragma unroll
            for (int k4 = 0; k4 < DM_KT / 4; k4++) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[(k4 * 4 + t) * DM_BM + wm * 32 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bs[(wn * 32 + j * 8 + g) * DM_BS + k4 * 4 + t];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        // ---- epilogue: accumulator fragments straight to C (thread holds C[g][2t], C[g][2t+1]) ---
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = m0 + wm * 32 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t c = n0 + wn * 32 + j * 8 + 2 * t;
                if (r < m) {
                    if (c < ncols) C[r + c * ldc] = acc[i][j][0];
                    if (c + 1 < ncols) C[r + (c + 1) * ldc] = acc[i][j][1];
                }
            }
        }
    }
}

// shape-generic fallback (any alignment): one thread per output row, columns in registers
template <int PC>
__global__ void __launch_bounds__(128) gemm_fallback_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                                                            double* __restrict__ C, int64_t ldc, int64_t m, int64_t k, int64_t ncols, int bmod) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.y * PC;
    if (r >= m) return;
    double acc[PC];
#pragma unroll
    for (int j = 0; j < PC; j++) acc[j] = 0.0;
    for (int64_t kk = 0; kk < k; kk++) {
        const double a = __ldg(A + r + kk * lda);
#pragma unroll
        for (int j = 0; j < PC; j++) {
            int64_t c = c0 + j; if (c >= ncols) c = ncols - 1;
            const int64_t bc = bmod > 0 ? c % bmod : c;
            acc[j] = acc[j] + a * __ldg(B + kk + bc * ldb);
        }
    }
#pragma unroll
    for (int j = 0; j < PC; j++) if (c0 + j < ncols) C[r + (c0 + j) * ldc] = acc[j];
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool dmma_eligible(const double* A, int64_t lda, int64_t m, int64_t k) {
    return (lda % 2 == 0) && ((((uintptr_t)A) & 15) == 0) && m >= DM_BM && k >= DM_KT && encode_tiled_fn() != nullptr;
}
size_t gemm_workspace_doubles(int64_t k, int64_t ncols, int bmod) {
    const int64_t nslab = cdiv(k, DM_KT);
    const int64_t btiles = (bmod > 0 && DM_BN % bmod == 0) ? 1 : cdiv(ncols, DM_BN);
    return (size_t)(btiles * nslab * DM_B_ELEMS);
}

// ws: gemm_workspace_doubles() doubles of device scratch for the packed B operand
int launch_gemm(fdb_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t m, int64_t k,
                int64_t ncols, int bmod, double* ws, bool* used_dmma, int* nlaunch) {
    if (m <= 0 || ncols <= 0) { if (nlaunch) *nlaunch = 0; return FDB_OK; }
    EncodeTiledFn enc = encode_tiled_fn();
    const bool ok = enc != nullptr && (lda % 2 == 0) && ((((uintptr_t)A) & 15) == 0) && k > 0 && ws != nullptr;
    if (used_dmma) *used_dmma = ok;
    if (!ok) {
        constexpr int PC = 13;
        dim3 grid((unsigned)cdiv(m, 128), (unsigned)cdiv(ncols, PC));
        gemm_fallback_kernel<PC><<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, k, ncols, bmod);
        if (nlaunch) *nlaunch = 1;
        return FDB_OK;
    }
    if (!ctx->dmma_configured) {       // a per-device attribute: every context (GPU) of the process sets it once
        cudaError_t e = cudaFuncSetAttribute(dmma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(dmma_gemm_kernel): %s", cudaGetErrorString(e));
        ctx->dmma_configured = true;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)m, (cuuint64_t)k};
    const cuuint64_t gstride[1] = {(cuuint64_t)lda * 8};
    const cuuint32_t box[2] = {DM_BM, DM_KT};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ctx, FDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int nslab = (int)cdiv(k, DM_KT);
    const int btiles = (bmod > 0 && DM_BN % bmod == 0) ? 1 : (int)cdiv(ncols, DM_BN);
    dim3 pg((unsigned)nslab, (unsigned)btiles);
    pack_b_kernel<<<pg, 256, 0, ctx->stream>>>(B, ldb, k, ncols, bmod, nslab, ws);
    dim3 grid((unsigned)cdiv(ncols, DM_BN), (unsigned)cdiv(m, DM_BM));
    dmma_gemm_kernel<<<grid, 288, DM_SMEM, ctx->stream>>>(tm, ws, btiles, C, ldc, m, k, ncols);
    if (nlaunch) *nlaunch = 2;
    return FDB_OK;
}

// ---- batched chunk sweeps of f(x) = tanh.(A*x .+ b) on the DMMA contraction ------------------------------
// Operand columns for the contraction.  SHARE: two columns per chunk, [x | 0] -- the value lane and the
// representative zero-partial lane (identical for every chunk, so the GEMM reads them with bmod = 2).
// Dense: N+1 columns per chunk, [x | one-hot seeds of the chunk] (src/apiutils.jl:125-143).
__global__ void build_operand_shared_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ B, int64_t ldb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    B[j] = x[j];
    B[ldb + j] = 0.0;
}
__global__ void build_operand_dense_kernel(const double* __restrict__ x, int64_t n, int N, int64_t chunk_lo, int64_t nchunks_total,
                                           int lastchunksize, double* __restrict__ B, int64_t ldb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c = chunk_lo + blockIdx.y;
    if (j >= n) return;
    const bool last = (c == nchunks_total - 1);
    const int cs = last ? lastchunksize : N;
    const int64_t start = last ? n - lastchunksize : c * (int64_t)N;
    double* col = B + (int64_t)blockIdx.y * (N + 1) * ldb + j;
    col[0] = x[j];
    for (int kk = 0; kk < N; kk++) col[(int64_t)(kk + 1) * ldb] = (j - start == kk && kk < cs) ? 1.0 : 0.0;
}
// epilogue: z = A*xdual .+ b ; y = tanh.(z) (DiffRules: val = tanh(v), deriv = 1 - val^2) ; J[:, start+k] = partials(y, k)
template <bool NS, bool SHARE, bool MIR>
__global__ void __launch_bounds__(256) tanh_affine_epilogue_kernel(const double* __restrict__ Y, int64_t ldy, int P, const double* __restrict__ A,
                                                                  int64_t lda, const double* __restrict__ b, int64_t m, int64_t n, int N,
                                                                  int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                  double* __restrict__ J, int64_t ldJ, double* __restrict__ yout,
                                                                  const __grid_constant__ fdb_mirrors Mr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t cl = blockIdx.y, c = chunk_lo + cl;
    if (r >= m) return;
    const bool last = (c == nchunks_total - 1);
    const int cs = last ? lastchunksize : N;
    const int64_t start = last ? n - lastchunksize : c * (int64_t)N;
    const double* Yc = Y + cl * P * ldy + r;
    const double zv = Yc[0] + b[r];                      // Dual + Real: value only
    const double val = tanh(zv);
    const double deriv = 1.0 - val * val;
    for (int kk = 0; kk < cs; kk++) {
        double p;
        if (SHARE) p = Yc[ldy] + A[r + (start + kk) * lda] * 1.0;   // shared zero lane + the one seeded column of the window
        else p = Yc[(int64_t)(kk + 1) * ldy];
        mstore_cs<MIR>(Mr, J + r + (start + kk) * ldJ, mulp<NS>(p, deriv));
    }
    if (yout && last) mstore(Mr, &yout[r], val);
}

int tanh_affine_dmma(fdb_ctx* ctx, const fdb_array* A, int64_t lda, const fdb_array* b, const fdb_array* x, int64_t m, int N, int64_t lo, int64_t hi,
                     int64_t nchunks, int lcs, fdb_array* J, int64_t ldJ, double* yout, bool* used_dmma, const fdb_mirrors& Mr) {
    const int64_t n = x->n, nc = hi - lo;
    const bool share = ctx->lane_sharing;
    const int P = share ? 2 : N + 1;
    const int64_t ldb = round_up(n, 32), ldy = round_up(m, 32);
    // scratch: operand columns + product columns
    const int64_t bcols = share ? 2 : nc * P;
    const size_t wsd = gemm_workspace_doubles(n, nc * P, share ? 2 : 0);
    size_t need = ((size_t)(ldb * bcols + ldy * nc * P) + wsd + 64) * sizeof(double);
    int rc = ensure_scratch(ctx, need); if (rc) return rc;
    double* Bop = ctx->scratch; double* Y = ctx->scratch + ldb * bcols;
    double* ws = Y + ldy * nc * P; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
    if (share) build_operand_shared_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x->data, n, Bop, ldb);
    else {
        dim3 g((unsigned)cdiv(n, 256), (unsigned)nc);
        build_operand_dense_kernel<<<g, 256, 0, ctx->stream>>>(x->data, n, N, lo, nchunks, lcs, Bop, ldb);
    }
    int ngemm = 0;
    rc = launch_gemm(ctx, A->data, lda, Bop, ldb, Y, ldy, m, n, nc * P, share ? 2 : 0, ws, used_dmma, &ngemm); if (rc) return rc;
    dim3 ge((unsigned)cdiv(m, 256), (unsigned)nc);
    if (ctx->nansafe) {
        if (share) { if (Mr.count) tanh_affine_epilogue_kernel<true, true, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<true, true, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
        else { if (Mr.count) tanh_affine_epilogue_kernel<true, false, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<true, false, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
    } else {
        if (share) { if (Mr.count) tanh_affine_epilogue_kernel<false, true, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<false, true, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
        else { if (Mr.count) tanh_affine_epilogue_kernel<false, false, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<false, false, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
    }
    return finish(ctx, 2 + ngemm, "fdb_jacobian_chunked(tanh_affine, dmma)");
}

// ---- two layers: f(x) = tanh.(A2 * tanh.(A1*x .+ b1) .+ b2) -------------------------------------------------
// After the first layer the partials are dense, so the second `A * hdual` is a genuine dense contraction of A2
// against [value | N partials] of every chunk: one DMMA GEMM m x h x (N+1)*nchunks.
// Layer-1 epilogue that keeps the hidden Duals (value + N partial columns per chunk) instead of extracting:
template <bool NS, bool SHARE>
__global__ void __launch_bounds__(256) tanh_affine_hidden_kernel(const double* __restrict__ Y, int64_t ldy, int P, const double* __restrict__ A,
                                                                int64_t lda, const double* __restrict__ b, int64_t h, int64_t n, int N,
                                                                int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                double* __restrict__ H, int64_t ldh) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t cl = blockIdx.y, c = chunk_lo + cl;
    if (r >= h) return;
    const bool last = (c == nchunks_total - 1);
    const int cs = last ? lastchunksize : N;
    const int64_t start = last ? n - lastchunksize : c * (int64_t)N;
    const double* Yc = Y + cl * P * ldy + r;
    double* Hc = H + cl * (N + 1) * ldh + r;
    const double val = tanh(Yc[0] + b[r]);
    const double deriv = 1.0 - val * val;
    Hc[0] = val;
    for (int kk = 0; kk < N; kk++) {
        double p;
        if (kk >= cs) p = SHARE ? Yc[ldy] : Yc[(int64_t)(kk + 1) * ldy];      // unseeded slots of the short last chunk: the zero lane
        else if (SHARE) p = Yc[ldy] + A[r + (start + kk) * lda] * 1.0;
        else p = Yc[(int64_t)(kk + 1) * ldy];
        Hc[(int64_t)(kk + 1) * ldh] = mulp<NS>(p, deriv);
    }
}

int mlp2_dmma(fdb_ctx* ctx, const fdb_array* A1, int64_t lda1, const fdb_array* b1, const fdb_array* A2, int64_t lda2, const fdb_array* b2,
              const fdb_array* x, int64_t h, int64_t m, int N, int64_t lo, int64_t hi, int64_t nchunks, int lcs, fdb_array* J, int64_t ldJ,
              double* yout, const fdb_mirrors& Mr) {
    const int64_t n = x->n, nc = hi - lo;
    const bool share = ctx->lane_sharing;
    const int P1 = share ? 2 : N + 1, P = N + 1;
    const int64_t ldb = round_up(n, 32), ldy = round_up(h, 32), ldh = round_up(h, 32), ldz = round_up(m, 32);
    const int64_t bcols = share ? 2 : nc * P1;
    const size_t ws1 = gemm_workspace_doubles(n, nc * P1, share ? 2 : 0), ws2 = gemm_workspace_doubles(h, nc * P, 0);
    const size_t wsd = ws1 > ws2 ? ws1 : ws2;
    size_t need = ((size_t)(ldb * bcols + ldy * nc * P1 + ldh * nc * P + ldz * nc * P) + wsd + 64) * sizeof(double);
    int rc = ensure_scratch(ctx, need); if (rc) return rc;
    double* Bop = ctx->scratch; double* Y1 = Bop + ldb * bcols; double* H1 = Y1 + ldy * nc * P1; double* Z2 = H1 + ldh * nc * P;
    double* ws = Z2 + ldz * nc * P; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
    if (share) build_operand_shared_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x->data, n, Bop, ldb);
    else {
        dim3 g((unsigned)cdiv(n, 256), (unsigned)nc);
        build_operand_dense_kernel<<<g, 256, 0, ctx->stream>>>(x->data, n, N, lo, nchunks, lcs, Bop, ldb);
    }
    int n1 = 0, n2 = 0;
    rc = launch_gemm(ctx, A1->data, lda1, Bop, ldb, Y1, ldy, h, n, nc * P1, share ? 2 : 0, ws, nullptr, &n1); if (rc) return rc;
    dim3 gh((unsigned)cdiv(h, 256), (unsigned)nc);
    if (ctx->nansafe) {
        if (share) tanh_affine_hidden_kernel<true, true><<<gh, 256, 0, ctx->stream>>>(Y1, ldy, P1, A1->data, lda1, b1->data, h, n, N, lo, nchunks, lcs, H1, ldh);
        else tanh_affine_hidden_kernel<true, false><<<gh, 256, 0, ctx->stream>>>(Y1, ldy, P1, A1->data, lda1, b1->data, h, n, N, lo, nchunks, lcs, H1, ldh);
    } else {
        if (share) tanh_affine_hidden_kernel<false, true><<<gh, 256, 0, ctx->stream>>>(Y1, ldy, P1, A1->data, lda1, b1->data, h, n, N, lo, nchunks, lcs, H1, ldh);
        else tanh_affine_hidden_kernel<false, false><<<gh, 256, 0, ctx->stream>>>(Y1, ldy, P1, A1->data, lda1, b1->data, h, n, N, lo, nchunks, lcs, H1, ldh);
    }
    // second layer: dense partials -> DMMA contraction of A2 against all hidden planes of all chunks
    rc = launch_gemm(ctx, A2->data, lda2, H1, ldh, Z2, ldz, m, h, nc * P, 0, ws, nullptr, &n2); if (rc) return rc;
    dim3 ge((unsigned)cdiv(m, 256), (unsigned)nc);
    if (ctx->nansafe) { if (Mr.count) tanh_affine_epilogue_kernel<true, false, true><<<ge, 256, 0, ctx->stream>>>(Z2, ldz, P, A2->data, lda2, b2->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<true, false, false><<<ge, 256, 0, ctx->stream>>>(Z2, ldz, P, A2->data, lda2, b2->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
    else { if (Mr.count) tanh_affine_epilogue_kernel<false, false, true><<<ge, 256, 0, ctx->stream>>>(Z2, ldz, P, A2->data, lda2, b2->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); else tanh_affine_epilogue_kernel<false, false, false><<<ge, 256, 0, ctx->stream>>>(Z2, ldz, P, A2->data, lda2, b2->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout, Mr); }
    return finish(ctx, 3 + n1 + n2, "fdb_jacobian_mlp2_chunked");
}

}  // namespace fdb

using namespace fdb;

// jacobian(x -> tanh.(A2*tanh.(A1*x .+ b1) .+ b2), x, JacobianConfig(f, x, Chunk{N}())): chunk sweeps [lo, hi)
extern "C" int fdb_jacobian_mlp2_chunked(fdb_ctx* ctx, const fdb_array* x, int64_t h, int64_t m, int N, const fdb_array* A1, int64_t lda1,
                                         const fdb_array* b1, const fdb_array* A2, int64_t lda2, const fdb_array* b2, int64_t chunk_lo,
                                         int64_t chunk_hi, fdb_array* J, int64_t ldJ, fdb_array* y) {
    FDB_ENTER(ctx);
    if (!ctx || !x || !A1 || !b1 || !A2 || !b2 || !J) return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_mlp2_chunked: null argument");
    const fdb_array* all[] = {x, A1, b1, A2, b2, J};
    for (const fdb_array* a : all) if (a->level != 0) return fail(ctx, FDB_ERR_BADARG, "fdb_jacobian_mlp2_chunked: operands must be plain arrays");
    const int64_t n = x->n;
    if (N < 1 || N > FDB_MAX_CHUNK) return fail(ctx, FDB_ERR_BADARG, "chunk size N=%d outside 1..12", N);
    if (n < N) return fail(ctx, FDB_ERR_CHUNK, "chunk size cannot be greater than ForwardDiff.structural_length(x) (%d > %lld)", N, (long long)n);
    if (h <= 0 || m <= 0 || lda1 < h || lda2 < m || A1->n < lda1 * (n - 1) + h || A2->n < lda2 * (h - 1) + m || b1->n != h || b2->n != m ||
        ldJ < m || J->n < ldJ * (n - 1) + m || (y && (y->level != 0 || y->n != m)))
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A1 must be %lld x %lld, A2 %lld x %lld, J %lld x %lld", (long long)h, (long long)n,
                    (long long)m, (long long)h, (long long)m, (long long)n);
    Plan p = make_plan(n, N);
    const int64_t nchunks = (n == N) ? 1 : p.nchunks;
    if (chunk_hi < 0) chunk_hi = nchunks;
    if (chunk_lo < 0 || chunk_lo > chunk_hi || chunk_hi > nchunks) return fail(ctx, FDB_ERR_BADARG, "bad chunk range");
    if (chunk_lo == chunk_hi) return FDB_OK;
    fdb_mirrors Mr;
    const fdb_array* outs[2] = {J, y};
    int rc = mirrors_for(ctx, &Mr, outs, 2, "fdb_jacobian_mlp2_chunked"); if (rc) return rc;
    return mlp2_dmma(ctx, A1, lda1, b1, A2, lda2, b2, x, h, m, N, chunk_lo, chunk_hi, nchunks, (int)p.lastchunksize, J, ldJ, y ? y->data : nullptr, Mr);
}

// C[m x ncols] = A[m x k] * B[k x ncols] on plain column-major arrays
extern "C" int fdb_matmul(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t k, const fdb_array* B, int64_t ldb, int64_t ncols,
                          fdb_array* C, int64_t ldc) {
    FDB_ENTER(ctx);
    if (!ctx || !A || !B || !C) return fail(ctx, FDB_ERR_BADARG, "fdb_matmul: null argument");
    if (A->level != 0 || B->level != 0 || C->level != 0) return fail(ctx, FDB_ERR_BADARG, "fdb_matmul: operands must be plain arrays");
    if (lda < m || ldb < k || ldc < m || A->n < lda * (k - 1) + m || B->n < ldb * (ncols - 1) + k || C->n < ldc * (ncols - 1) + m)
        return fail(ctx, FDB_ERR_SHAPE, "fdb_matmul: DimensionMismatch");
    int rc = ensure_scratch(ctx, (gemm_workspace_doubles(k, ncols, 0) + 64) * sizeof(double)); if (rc) return rc;
    double* ws = (double*)((((uintptr_t)ctx->scratch) + 127) & ~(uintptr_t)127);
    int nl = 0;
    rc = launch_gemm(ctx, A->data, lda, B->data, ldb, C->data, ldc, m, k, ncols, 0, ws, nullptr, &nl); if (rc) return rc;
    return finish(ctx, nl, "fdb_matmul");
}

// ydual = A * xdual  (A: m x n plain column-major; xdual: n Duals; ydual: m Duals): the planes are the columns
extern "C" int fdb_dual_matvec(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t n, const fdb_array* xdual, fdb_array* ydual) {
    FDB_ENTER(ctx);
    if (!ctx || !A || !xdual || !ydual) return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matvec: null argument");
    if (A->level != 0 || xdual->level < 1 || ydual->level != xdual->level || xdual->N != ydual->N)
        return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matvec: A plain, xdual/ydual Dual arrays of the same type");
    if (xdual->n != n || ydual->n != m || lda < m || A->n < lda * (n - 1) + m)
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A is %lld x %lld, x has %lld, y has %lld", (long long)m, (long long)n, (long long)xdual->n, (long long)ydual->n);
    int rc = ensure_scratch(ctx, (gemm_workspace_doubles(n, xdual->planes(), 0) + 64) * sizeof(double)); if (rc) return rc;
    double* ws = (double*)((((uintptr_t)ctx->scratch) + 127) & ~(uintptr_t)127);
    int nl = 0;
    rc = launch_gemm(ctx, A->data, lda, xdual->data, xdual->ld, ydual->data, ydual->ld, m, n, xdual->planes(), 0, ws, nullptr, &nl); if (rc) return rc;
    return finish(ctx, nl, "fdb_dual_matvec");
}
