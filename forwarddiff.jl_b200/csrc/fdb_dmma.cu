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
#include "fdb_internal.h"
#include "fdb_targets.cuh"

namespace fdb {

constexpr int DM_BM = 128, DM_BN = 64, DM_KT = 32, DM_STAGES = 4;
constexpr int DM_AS = DM_BM + 4;          // padded k-row of the A tile (doubles): stride = 8 mod 32 words
constexpr int DM_BS = DM_KT + 4;          // padded column of the B tile (doubles): stride = 8 mod 32 words
constexpr int DM_A_BYTES = DM_KT * DM_AS * 8, DM_B_BYTES = DM_BN * DM_BS * 8;
constexpr int DM_STAGE_BYTES = DM_A_BYTES + DM_B_BYTES;
constexpr int DM_SMEM = DM_STAGES * DM_STAGE_BYTES + 2 * DM_STAGES * 8 + 16;

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
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
FDB_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
FDB_DEV void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[m x ncols] = A[m x k] * B[k x ncols]; all column-major.  Column c of B is read from column
// (bmod > 0 ? c % bmod : c) -- bmod lets a batch of chunks share identical operand columns.
// Requirements (checked by the host): k % 32 == 0, m % 2 == 0, lda/ldb even, 16-byte aligned bases.
__global__ void __launch_bounds__(288, 1) dmma_gemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                                                          double* __restrict__ C, int64_t ldc, int64_t m, int64_t k, int64_t ncols, int bmod) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DM_STAGES * DM_STAGE_BYTES);
    uint64_t* empty = full + DM_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * DM_BN, m0 = (int64_t)blockIdx.y * DM_BM;
    const int rows = (int)(m - m0 < DM_BM ? m - m0 : DM_BM);          // valid rows of this tile (even)
    const int cols = (int)(ncols - n0 < DM_BN ? ncols - n0 : DM_BN);
    const int nslab = (int)(k / DM_KT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < DM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rows / columns beyond the matrix edge are never copied: clear them once so DMMA reads zeros
    if (rows < DM_BM || cols < DM_BN) {
        double* z = reinterpret_cast<double*>(smem);
        for (int i = threadIdx.x; i < DM_STAGES * DM_STAGE_BYTES / 8; i += blockDim.x) z[i] = 0.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ---- producer: one thread drives the TMA engine ----------------------------------------
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(DM_KT * rows * 8 + cols * DM_KT * 8);
            for (int s = 0; s < nslab; s++) {
                const int st = s % DM_STAGES;
                if (s >= DM_STAGES) mbar_wait(&empty[st], ((s / DM_STAGES) - 1) & 1);
                double* As = reinterpret_cast<double*>(smem + st * DM_STAGE_BYTES);
                double* Bs = As + DM_KT * DM_AS;
                mbar_expect_tx(&full[st], tx);
                const int64_t k0 = (int64_t)s * DM_KT;
                for (int kk = 0; kk < DM_KT; kk++)               // one k-row of the A tile: `rows` contiguous doubles of column k0+kk
                    bulk_g2s(As + kk * DM_AS, A + m0 + (k0 + kk) * lda, (uint32_t)rows * 8, &full[st]);
                for (int c = 0; c < cols; c++) {                 // one column of the B tile: KT contiguous doubles
                    const int64_t bc = bmod > 0 ? (n0 + c) % bmod : (n0 + c);
                    bulk_g2s(Bs + c * DM_BS, B + k0 + bc * ldb, DM_KT * 8, &full[st]);
                }
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
            const double* Bs = As + DM_KT * DM_AS;
#pragma unroll
            for (int k4 = 0; k4 < DM_KT / 4; k4++) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[(k4 * 4 + t) * DM_AS + wm * 32 + i * 8 + g];
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

// shape-generic fallback (any m, k, alignment): one thread per output row, columns in registers
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

bool dmma_eligible(const double* A, int64_t lda, int64_t m, int64_t k) {
    return (k % DM_KT == 0) && (m % 2 == 0) && (lda % 2 == 0) && ((((uintptr_t)A) & 15) == 0) && m >= DM_BM && k >= DM_KT;
}

int launch_gemm(fdb_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t m, int64_t k,
                int64_t ncols, int bmod, bool* used_dmma) {
    const bool ok = (k % DM_KT == 0) && (m % 2 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) && ((((uintptr_t)A) & 15) == 0) &&
                    ((((uintptr_t)B) & 15) == 0) && m > 0 && ncols > 0 && k > 0;
    if (used_dmma) *used_dmma = ok;
    if (ok) {
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(dmma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
            if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(dmma_gemm_kernel): %s", cudaGetErrorString(e));
            configured = true;
        }
        dim3 grid((unsigned)cdiv(ncols, DM_BN), (unsigned)cdiv(m, DM_BM));
        dmma_gemm_kernel<<<grid, 288, DM_SMEM, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, k, ncols, bmod);
    } else if (m > 0 && ncols > 0) {
        constexpr int PC = 13;
        dim3 grid((unsigned)cdiv(m, 128), (unsigned)cdiv(ncols, PC));
        gemm_fallback_kernel<PC><<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, k, ncols, bmod);
    }
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
template <bool NS, bool SHARE>
__global__ void __launch_bounds__(256) tanh_affine_epilogue_kernel(const double* __restrict__ Y, int64_t ldy, int P, const double* __restrict__ A,
                                                                  int64_t lda, const double* __restrict__ b, int64_t m, int64_t n, int N,
                                                                  int64_t chunk_lo, int64_t nchunks_total, int lastchunksize,
                                                                  double* __restrict__ J, int64_t ldJ, double* __restrict__ yout) {
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
        __stcs(J + r + (start + kk) * ldJ, mulp<NS>(p, deriv));
    }
    if (yout && last) yout[r] = val;
}

int tanh_affine_dmma(fdb_ctx* ctx, const fdb_array* A, int64_t lda, const fdb_array* b, const fdb_array* x, int64_t m, int N, int64_t lo, int64_t hi,
                     int64_t nchunks, int lcs, fdb_array* J, int64_t ldJ, double* yout, bool* used_dmma) {
    const int64_t n = x->n, nc = hi - lo;
    const bool share = ctx->lane_sharing;
    const int P = share ? 2 : N + 1;
    const int64_t ldb = round_up(n, 32), ldy = round_up(m, 32);
    // scratch: operand columns + product columns
    const int64_t bcols = share ? 2 : nc * P;
    size_t need = (size_t)(ldb * bcols + ldy * nc * P) * sizeof(double);
    int rc = ensure_scratch(ctx, need); if (rc) return rc;
    double* Bop = ctx->scratch; double* Y = ctx->scratch + ldb * bcols;
    if (share) build_operand_shared_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x->data, n, Bop, ldb);
    else {
        dim3 g((unsigned)cdiv(n, 256), (unsigned)nc);
        build_operand_dense_kernel<<<g, 256, 0, ctx->stream>>>(x->data, n, N, lo, nchunks, lcs, Bop, ldb);
    }
    rc = launch_gemm(ctx, A->data, lda, Bop, ldb, Y, ldy, m, n, nc * P, share ? 2 : 0, used_dmma); if (rc) return rc;
    dim3 ge((unsigned)cdiv(m, 256), (unsigned)nc);
    if (ctx->nansafe) {
        if (share) tanh_affine_epilogue_kernel<true, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout);
        else tanh_affine_epilogue_kernel<true, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout);
    } else {
        if (share) tanh_affine_epilogue_kernel<false, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout);
        else tanh_affine_epilogue_kernel<false, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, P, A->data, lda, b->data, m, n, N, lo, nchunks, lcs, J->data, ldJ, yout);
    }
    return finish(ctx, 3, "fdb_jacobian_chunked(tanh_affine, dmma)");
}

}  // namespace fdb

using namespace fdb;

// C[m x ncols] = A[m x k] * B[k x ncols] on plain column-major arrays
extern "C" int fdb_matmul(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t k, const fdb_array* B, int64_t ldb, int64_t ncols,
                          fdb_array* C, int64_t ldc) {
    if (!ctx || !A || !B || !C) return fail(ctx, FDB_ERR_BADARG, "fdb_matmul: null argument");
    if (A->level != 0 || B->level != 0 || C->level != 0) return fail(ctx, FDB_ERR_BADARG, "fdb_matmul: operands must be plain arrays");
    if (lda < m || ldb < k || ldc < m || A->n < lda * (k - 1) + m || B->n < ldb * (ncols - 1) + k || C->n < ldc * (ncols - 1) + m)
        return fail(ctx, FDB_ERR_SHAPE, "fdb_matmul: DimensionMismatch");
    int rc = launch_gemm(ctx, A->data, lda, B->data, ldb, C->data, ldc, m, k, ncols, 0, nullptr); if (rc) return rc;
    return finish(ctx, 1, "fdb_matmul");
}

// ydual = A * xdual  (A: m x n plain column-major; xdual: n Duals; ydual: m Duals): the planes are the columns
extern "C" int fdb_dual_matvec(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t n, const fdb_array* xdual, fdb_array* ydual) {
    if (!ctx || !A || !xdual || !ydual) return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matvec: null argument");
    if (A->level != 0 || xdual->level < 1 || ydual->level != xdual->level || xdual->N != ydual->N)
        return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matvec: A plain, xdual/ydual Dual arrays of the same type");
    if (xdual->n != n || ydual->n != m || lda < m || A->n < lda * (n - 1) + m)
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A is %lld x %lld, x has %lld, y has %lld", (long long)m, (long long)n, (long long)xdual->n, (long long)ydual->n);
    int rc = launch_gemm(ctx, A->data, lda, xdual->data, xdual->ld, ydual->data, ydual->ld, m, n, xdual->planes(), 0, nullptr); if (rc) return rc;
    return finish(ctx, 1, "fdb_dual_matvec");
}
