// fdb_dmma.cu -- FP64 tensor-core (DMMA) contraction for Dual matvec / matmul.
//
// `A * xdual` on a Vector{Dual{T,Float64,N}} (Julia: LinearAlgebra.generic_matvecmul!, Real*Dual then
// Dual+Dual per element) is a dense contraction of A against the planes [value | partials]:
//      Y[m x P] = A[m x k] * X[k x P],   P = N+1 columns per chunk, batched over chunks.
// This is the only tensor-core kernel in the engine.  FP64 has no tcgen05 kind; the FP64 tensor path
// on sm_100a is `mma.sync.aligned.m8n8k4.f64` (SASS DMMA).
//
// Kernel shape (round 2).  CTA tile 128 x BN x 16 (BN = 64 / 32 / 16 operand columns), eight DMMA warps (4 along m x 2
// along n; the TMA issue rotates over them), a ring of mbarrier-guarded stages filled by TMA, TWO CTAs resident per SM (<= 111 KB of
// shared memory and <= 112 registers each): while one CTA sits in its epilogue -- result stores, over NVLink when a peer
// exchange is enabled -- the other keeps the DMMA pipe busy, so the stores overlap the math without any explicit
// software pipelining between tiles.
//   * A tile (128 m x 16 k of the column-major matrix): eight `cp.async.bulk.tensor.2d` boxes of 16 m x 16 k with
//     CU_TENSOR_MAP_SWIZZLE_128B.  Each box is 16 rows (k) of 128 bytes (16 m); the swizzle XORs the 16-byte chunk index
//     with k % 8.  An m8n8k4 step lets lane (g, t) pick ANY four k of the slab as long as the B fragment uses the same
//     ones, so step s4 uses k = 8*(s4/2) + 2t + (s4%2): the XOR term 2t + e then spreads the sixteen lanes of a
//     half-warp over all sixteen 8-byte banks -- conflict-free (round 1: dense [k][128 m] box, 4-way replays on every A
//     fragment load, 3.5e9 excess wavefronts per C3 launch).  Out-of-range rows / k are zero-filled by the TMA unit.
//   * B operand: pre-packed by pack_b_kernel into tiles of BN padded columns (16 + 4 doubles, bank-conflict free) whose
//     k positions are already in fragment order (position s4*4 + t), staged by ONE bulk copy.
//   * Thread blocks are rasterised in bands of eight m-tiles (column tile next, m-tile fastest): the blocks in flight share
//     eight A row-blocks (L2-resident across the band) and a sliding window of B tiles, instead of re-reading all of B
//     from DRAM for every m-tile row (round 1: 40.7 GB of DRAM reads against 1.7 GB algorithmic).
//   * Epilogue: plain store of the accumulator fragments, or -- for the lane-shared chunk sweeps of tanh.(A*x .+ b) --
//     the whole rest of the sweep fused in: z + b, tanh, 1 - val^2, the seeded window column, extraction into the
//     Jacobian (local and mirrored into every peer's arena).  The product never goes to HBM.
//
// Arithmetic note: DMMA accumulates with fused multiply-adds in a permuted k order, the reference with separate
// multiply and add in ascending k; both are summation-order variants of the same dot product (tolerance 1e-12
// relative, BASELINE.json north_star).  One-hot operand columns come out exact.
#include <cuda.h>

#include "fdb_internal.h"
#include "fdb_mirror.cuh"
#include "fdb_targets.cuh"
#include "fdb_tma.cuh"

namespace fdb {

using tma::smem_u32; using tma::mbar_init; using tma::mbar_expect_tx; using tma::mbar_arrive; using tma::mbar_wait; using tma::bulk_g2s;

constexpr int DM_BM = 128, DM_BK = 16, DM_BS = DM_BK + 4;      // BS: padded column of the packed B tile (stride = 8 mod 32 words)
constexpr int DM_A_BYTES = DM_BK * DM_BM * 8;                  // eight swizzled [16 k][16 m] boxes of 2 KB
constexpr int DM_THREADS = 256;
constexpr int DM_BAND = 8;                                     // m-tiles per rasterisation band
constexpr int DM_STG_BYTES = 4 * FDB_MAX_CHUNK * DM_BM * 8;    // staging buffer of the fused epilogue: 128 rows x 4N columns
template <int BN> struct DmCfg {
    static constexpr int B_ELEMS = BN * DM_BS, B_BYTES = B_ELEMS * 8;
    static constexpr int STAGE_BYTES = (DM_A_BYTES + B_BYTES + 1023) / 1024 * 1024;      // A tiles stay 1024-byte aligned (swizzle atom)
    static constexpr int STAGES = (108 * 1024) / STAGE_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;            // + alignment slack
    // persistent variant (one block per SM): ring + two staging buffers of one 128 x 4N result block each
    static constexpr int STAGES_P0 = (227 * 1024 - 1024 - 256 - 2 * DM_STG_BYTES) / STAGE_BYTES;
    static constexpr int STAGES_P = STAGES_P0 > 8 ? 8 : STAGES_P0;
    static constexpr int SMEM_P = STAGES_P * STAGE_BYTES + 2 * DM_STG_BYTES + 2 * STAGES_P * 8 + 1024;
    static constexpr int WJ = BN / 16;                                                   // 8-column DMMA tiles per warp (warps: 4 m x 2 n)
};

// TMA tensor copy (2-D box of the column-major A described by a CUtensorMap) global -> shared
FDB_DEV void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
FDB_DEV void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// slab position of logical k offset kk (0..15): DMMA step s4 = 2*(kk/8) + kk%2 reads k = 8*(s4/2) + 2t + s4%2 in lane t
__host__ __device__ inline int dm_kpos(int kk) { return (kk >> 3) * 8 + (kk & 1) * 4 + ((kk & 7) >> 1); }

// Pack B (k x ncols, column-major) into the tile order the kernel consumes: tile (j, s) holds columns
// [BN*j, BN*j+BN) x k-rows [16s, 16s+16) as BN padded columns of 20 doubles in fragment order, contiguous, so that one
// bulk copy stages it.  Out-of-range entries are zero.  Column c of B is read from column c % bmod when
// bmod > 0 (a batch of chunks sharing identical operand columns).
// onehot_c0 >= 0: the operand is not read from memory but generated: column 0 = x (= B), column 1 + q = the one-hot seed of
// structural position onehot_c0 + q (src/apiutils.jl:125-143) -- the all-lanes operand of the tanh.(A*x .+ b) sweeps.
__global__ void pack_b_kernel(const double* __restrict__ B, int64_t ldb, int64_t k, int64_t ncols, int bmod, int nslab, int BN,
                              int64_t onehot_c0, double* __restrict__ Bp) {
    const int s = blockIdx.x, j = blockIdx.y;
    const int belems = BN * DM_BS;
    double* tile = Bp + ((int64_t)j * nslab + s) * belems;
    for (int e = threadIdx.x; e < belems; e += blockDim.x) {
        const int c = e / DM_BS, kk = e - c * DM_BS;
        const int64_t col = (int64_t)j * BN + c, row = (int64_t)s * DM_BK + kk;
        double v = 0.0;
        if (kk < DM_BK && col < ncols && row < k) {
            if (onehot_c0 >= 0) v = col == 0 ? B[row] : (row == onehot_c0 + col - 1 ? 1.0 : 0.0);
            else v = B[row + (bmod > 0 ? col % bmod : col) * ldb];
        }
        tile[c * DM_BS + (kk < DM_BK ? dm_kpos(kk) : kk)] = v;
    }
}

// Fused epilogue of the lane-shared tanh.(A*x .+ b) chunk sweeps: operand columns come in pairs
// (2c, 2c+1) = (value lane, representative zero-partial lane) of chunk chunk_lo + c, and the m8n8k4 accumulator
// fragment gives thread (g, t) exactly such a pair for row g.
struct TanhEpi {
    const double* A; int64_t lda; const double* b; int64_t n; int N; int64_t chunk_lo, nchunks_total; int lastchunksize;
    int64_t nc; double* J; int64_t ldJ; double* yout;
};

enum { DM_EPI_STORE = 0, DM_EPI_TANH = 1 };

// Tensor maps of the Jacobian for the tile epilogue's TMA stores: [0] = the local result, [1 .. count-1] = the same matrix in
// every peer's arena (fdb_mirror.cuh deltas).  Box = 128 rows x 4N columns = the four chunks one DMMA column group carries.
// mode 1: TMA tensor stores from the staging buffer; mode 2: the block's threads copy the staged block out themselves, a warp
// per 32 consecutive rows of a column (256-byte runs: the pattern that reaches the link rate in fdb_exchange_measure_store_bw)
struct StoreMaps { CUtensorMap map[FDB_MAX_RANKS]; int count; int mode; };

// Measured at 8 ranks (C3, 172 operand columns per rank; profiles/r02_bench_8gpu_*.json), sweep + exchange:
//   threads store their own accumulator entries (64-byte runs)                         1.79 ms
//   staged, TMA tensor stores, two blocks per SM (wait for the engine to read)          1.45 ms
//   persistent, one block per SM, two staging buffers, stores overlap the next tile    2.16 ms  (BN = 16 leaves two warps per
//       scheduler with four independent DMMA chains each: the tensor pipe starves; BN = 64 tiles quantise to two waves)
// so the persistent variant is compiled out.
constexpr bool DM_PERSIST_MIRROR = false;
// 16-column tiles with the fused epilogue: THREE blocks per SM (3-stage ring, <= 80 registers).  One block's 8 warps issue 20
// shared loads per 16 DMMAs and cannot keep the pipe busy through a neighbour's epilogue (A reads + J stores, ~10 us of a
// ~350 us block life) or its NVLink stores; measured at C3 with BN = 16 on one B200: 5.98 ms against 6.52 ms with two blocks.
__host__ __device__ constexpr bool DM_BN16_THREE(int EPI, int BN) { return EPI == DM_EPI_TANH && BN == 16; }
constexpr int DM_MIRROR_STORE_MODE = 2;      // default when mirroring: 1 = TMA tensor stores, 2 = staged thread stores

FDB_DEV void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}

// C[m x ncols] = A[m x k] * B[k x ncols]; A through a tensor map (boxes 16 x 16, swizzle 128B, zero fill out of bounds),
// B pre-packed (pack_b_kernel); btiles = number of distinct packed column tiles (1 when every tile is identical).
// A thread block walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... (one tile per block unless PERSIST); the slab ring and
// its mbarrier phases run on across tiles, so the first slabs of the next tile are in flight during the epilogue of this one.
// PERSIST (fused epilogue + peer mirroring): one block per SM, and the epilogue's TMA stores leave from two dedicated staging
// buffers, so the DMMA warps go straight on to the next tile while the TMA engine pushes the finished block to the peers over
// NVLink; they wait for a staging buffer only when it is needed again, two column groups later.
template <int BN, int EPI, bool NS, bool MIR>
__global__ void __launch_bounds__(DM_THREADS, (EPI == DM_EPI_TANH && MIR && DM_PERSIST_MIRROR) ? 1 : (DM_BN16_THREE(EPI, BN) ? 3 : 2))
dmma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const double* __restrict__ Bp, int btiles, double* __restrict__ C, int64_t ldc,
                 int64_t m, int64_t k, int64_t ncols, int tiles_m, int tiles_n, const int accumulate, const __grid_constant__ TanhEpi E,
                 const __grid_constant__ fdb_mirrors Mr, const __grid_constant__ StoreMaps SM) {
    using Cfg = DmCfg<BN>;
    constexpr int WJ = Cfg::WJ;
    constexpr bool PERSIST = (EPI == DM_EPI_TANH && MIR && DM_PERSIST_MIRROR);
    constexpr int STAGES = PERSIST ? Cfg::STAGES_P : (DM_BN16_THREE(EPI, BN) ? 3 : Cfg::STAGES);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* stg_base = smem + STAGES * Cfg::STAGE_BYTES;                       // PERSIST: two staging buffers behind the ring
    uint64_t* full = reinterpret_cast<uint64_t*>(stg_base + (PERSIST ? 2 * DM_STG_BYTES : 0));
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nslab = (int)((k + DM_BK - 1) / DM_BK);
    const int ntiles = tiles_m * tiles_n;
    const int my_tiles = PERSIST ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 1;
    const int total = my_tiles * nslab;                                               // slabs this block consumes (< 2^31: tiles per block x k/16)
    // band rasterisation: DM_BAND m-tiles x all column tiles per band, m-tile fastest
    auto tile_coords = [&](int tile, int& tm, int& tn) {
        const int per_band = DM_BAND * tiles_n, band = tile / per_band, within = tile - band * per_band;
        const int rows = min(DM_BAND, tiles_m - band * DM_BAND);
        tn = within / rows; tm = band * DM_BAND + (within - tn * rows);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        tma::mbar_fence_init();
    }
    __syncthreads();

    // ---- eight DMMA warps, 4 (M) x 2 (N), 32 x (BN/2) accumulators each.  There is no dedicated producer warp (a ninth warp
    // would cap the kernel at 96 registers with two CTAs per SM and spill accumulators): lane 0 of warp (slab mod 8) issues
    // the TMA copies of the slab STAGES - 1 ahead before that warp starts on its own slab, so the issue cost rotates over the warps.
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    // slab gp of this block's tile sequence = slab sp of the tile with coordinates (ptm, ptn)
    auto produce = [&](int gp, int sp, int pm0, const double* pbt) {    // called by one thread; pm0 = first row of the tile, pbt = its packed B tiles
        const int st = gp % STAGES;
        if (gp >= STAGES) mbar_wait(&empty[st], (uint32_t)(((gp / STAGES) - 1) & 1));
        unsigned char* stage = smem + st * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[st], DM_A_BYTES + Cfg::B_BYTES);
#pragma unroll
        for (int bx = 0; bx < DM_BM / 16; bx++)
            tma_load_2d(stage + bx * 2048, &tmA, pm0 + 16 * bx, sp * DM_BK, &full[st]);             // [16 k][16 m], 128B-swizzled
        bulk_g2s(stage + DM_A_BYTES, pbt + (int64_t)sp * Cfg::B_ELEMS, Cfg::B_BYTES, &full[st]);       // packed B tile
    };
    if (threadIdx.x == 0)
        for (int gp = 0; gp < STAGES - 1 && gp < total; gp++) {
            int ptm, ptn; tile_coords((int)blockIdx.x + (gp / nslab) * (int)gridDim.x, ptm, ptn);
            produce(gp, gp % nslab, ptm * DM_BM, Bp + (int64_t)(ptn % btiles) * nslab * Cfg::B_ELEMS);
        }
    // warp w issues the slabs w + STAGES - 1, w + STAGES - 1 + 8, ...: it tracks (tile, slab in tile) of its next one incrementally,
    // so the tile coordinates (integer divisions) are recomputed only when that slab enters a new tile
    int p_gp = warp + STAGES - 1, p_ti = p_gp / nslab, p_sp = p_gp - p_ti * nslab, p_m0 = 0;
    const double* p_bt = Bp;
    if (p_gp < total) {
        int ptm, ptn; tile_coords((int)blockIdx.x + p_ti * (int)gridDim.x, ptm, ptn);
        p_m0 = ptm * DM_BM; p_bt = Bp + (int64_t)(ptn % btiles) * nslab * Cfg::B_ELEMS;
    }
    // swizzled A fragment address (in doubles): box (wm*2 + i/2) * 256 + k * 16 + ((chunk ^ k%8) << 1) + (g & 1),
    // chunk = (i & 1) * 4 + g / 2  ==>  base + krow + (((g/2) ^ x) << 1) ^ ((i & 1) << 3), + 256 for i >= 2
    const int a_base = wm * 512 + (g & 1), gh = g >> 1;
    const int b_off = (wn * (BN / 2) + g) * DM_BS + t;
    int stg_sel = 0;                                                              // PERSIST: staging buffer of the next column group
    int gc = 0;                                                                   // slabs consumed so far
    for (int ti = 0; ti < my_tiles; ti++) {
        int tm, tn; tile_coords((int)blockIdx.x + ti * (int)gridDim.x, tm, tn);
        const int64_t n0 = (int64_t)tn * BN, m0 = (int64_t)tm * DM_BM;
        const double* bt0 = Bp + (int64_t)(tn % btiles) * nslab * Cfg::B_ELEMS;
        double acc[4][WJ][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < WJ; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        for (int s = 0; s < nslab; s++, gc++) {
            if (warp == (gc & 7)) {
                if constexpr (!PERSIST) {                             // one tile per block: the slab ahead belongs to this tile (block-uniform operands)
                    if (lane == 0 && s + STAGES - 1 < nslab) produce(s + STAGES - 1, s + STAGES - 1, (int)m0, bt0);
                } else if (p_gp < total) {
                    if (lane == 0) produce(p_gp, p_sp, p_m0, p_bt);
                    p_gp += 8; p_sp += 8;
                    if (p_sp >= nslab && p_gp < total) {
                        do { p_sp -= nslab; p_ti++; } while (p_sp >= nslab);
                        int ptm, ptn; tile_coords((int)blockIdx.x + p_ti * (int)gridDim.x, ptm, ptn);
                        p_m0 = ptm * DM_BM; p_bt = Bp + (int64_t)(ptn % btiles) * nslab * Cfg::B_ELEMS;
                    }
                }
                __syncwarp();
            }
            const int st = gc % STAGES;
            mbar_wait(&full[st], (uint32_t)((gc / STAGES) & 1));
            const double* As = reinterpret_cast<const double*>(smem + st * Cfg::STAGE_BYTES) + a_base;
            const double* Bs = reinterpret_cast<const double*>(smem + st * Cfg::STAGE_BYTES + DM_A_BYTES) + b_off;
#pragma unroll
            for (int s4 = 0; s4 < DM_BK / 4; s4++) {
                const int x = 2 * t + (s4 & 1);                            // k % 8 of this lane's k = 8*(s4/2) + x
                const int o0 = ((s4 >> 1) * 8 + x) * 16 + ((gh ^ x) << 1), o1 = o0 ^ 8;
                double a[4], b[WJ];
                a[0] = As[o0]; a[1] = As[o1]; a[2] = As[o0 + 256]; a[3] = As[o1 + 256];
#pragma unroll
                for (int j = 0; j < WJ; j++) b[j] = Bs[j * 8 * DM_BS + s4 * 4];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < WJ; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        if (EPI == DM_EPI_STORE) {
            // ---- accumulator fragments straight to C (thread holds C[g][2t], C[g][2t+1]) ----------------------
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int64_t r = m0 + wm * 32 + i * 8 + g;
#pragma unroll
                for (int j = 0; j < WJ; j++) {
                    const int64_t c = n0 + wn * (BN / 2) + j * 8 + 2 * t;
                    if (r < m) {
                        if (accumulate) {        // C += A*B: the second product of a Dual*Dual partial (fdb_dual_matmul)
                            if (c < ncols) C[r + c * ldc] = C[r + c * ldc] + acc[i][j][0];
                            if (c + 1 < ncols) C[r + (c + 1) * ldc] = C[r + (c + 1) * ldc] + acc[i][j][1];
                        } else {
                            if (c < ncols) C[r + c * ldc] = acc[i][j][0];
                            if (c + 1 < ncols) C[r + (c + 1) * ldc] = acc[i][j][1];
                        }
                    }
                }
            }
        } else {
            // ---- z = A*xdual .+ b ; y = tanh.(z) (DiffRules: val = tanh(v), deriv = 1 - val^2) ; J[:, start+kk] = partials(y, kk) ----
            // A DMMA column group (8 operand columns) carries four chunks: lane t of the owning warps holds (value lane, zero lane)
            // of chunk 4*gi + t for its rows.  Full groups are staged in shared memory as the dense 128 x 4N block of J they are
            // and written by the TMA engine -- one tensor store for the local result and one per peer arena (1 KB contiguous per
            // column: round 1 measured 661 vs 235 GB/s over NVLink between long and short store runs).  Groups cut by the end of
            // the chunk range (or a result the TMA cannot address) are stored by the threads themselves.
            if (!PERSIST && SM.count > 0) __syncthreads();            // every warp is done with the ring: it doubles as the staging buffer
            const int N = E.N;
#pragma unroll
            for (int gi = 0; gi < BN / 8; gi++) {
                const int gw = gi / WJ, j = gi % WJ;                  // owning warp column and its accumulator index (compile time)
                const int64_t cl0 = n0 / 2 + gi * 4;                  // first local chunk of the group
                if (cl0 >= E.nc) break;
                const bool staged = SM.count > 0 && cl0 + 4 <= E.nc;  // block-uniform
                double* stg = reinterpret_cast<double*>(PERSIST ? stg_base + stg_sel * DM_STG_BYTES : smem);     // [4N columns][128 rows]
                if (PERSIST && staged) {
                    // the buffer was handed to the TMA engine two groups ago: wait until that store has read it
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncthreads();
                }
                if (wn == gw) {
                    const int64_t cl = cl0 + t;
                    if (cl < E.nc) {
                        const int64_t c = E.chunk_lo + cl;
                        const bool last = (c == E.nchunks_total - 1);
                        const int cs = last ? E.lastchunksize : N;
                        const int64_t start = last ? E.n - E.lastchunksize : c * (int64_t)N;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int rl = wm * 32 + i * 8 + g;
                            const int64_t r = m0 + rl;
                            if (r >= m) continue;
                            const double zv = acc[i][j][0] + E.b[r];                            // Dual + Real: value only
                            const double val = tanh(zv);
                            const double deriv = 1.0 - val * val;
                            const double zero_lane = acc[i][j][1];
                            const double* Ac = E.A + r + start * E.lda;
                            if (staged) {
                                double* sc = stg + (t * N) * DM_BM + rl;
                                for (int kk = 0; kk < cs; kk++) {
                                    const double p = zero_lane + __ldg(Ac + (int64_t)kk * E.lda) * 1.0;   // shared zero lane + the one seeded column of the window
                                    sc[kk * DM_BM] = mulp<NS>(p, deriv);
                                }
                            } else {
                                double* Jc = E.J + r + start * E.ldJ;
                                for (int kk = 0; kk < cs; kk++) {
                                    const double p = zero_lane + __ldg(Ac + (int64_t)kk * E.lda) * 1.0;
                                    mstore_cs<MIR>(Mr, Jc + (int64_t)kk * E.ldJ, mulp<NS>(p, deriv));
                                }
                            }
                            if (E.yout && last) mstore(Mr, &E.yout[r], val);
                        }
                    }
                }
                if (staged && SM.mode == 2) {
                    // the block copies the staged 128 x 4N block out itself: consecutive lanes take consecutive rows of a column, so
                    // every warp store is one 256-byte run -- locally and at the same offset of every peer arena.  The stores are
                    // posted: nobody waits for them, the block goes on (or retires) while they drain over NVLink.
                    __syncthreads();
                    const int64_t col0 = (E.chunk_lo + cl0) * N;
                    for (int idx = threadIdx.x; idx < 4 * N * DM_BM; idx += DM_THREADS) {
                        const int col = idx >> 7, row = idx & (DM_BM - 1);
                        if (m0 + row < m && col0 + col < E.n) mstore_cs<MIR>(Mr, E.J + (m0 + row) + (col0 + col) * E.ldJ, stg[idx]);
                    }
                    __syncthreads();                                                  // the staging buffer is free again
                } else if (staged) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async-proxy (TMA) reads
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const int col0 = (int)((E.chunk_lo + cl0) * N);               // columns past n (short last chunk) and rows past m are clipped
                        for (int d = 0; d < SM.count; d++) tma_store_2d(&SM.map[d], stg, (int)m0, col0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        if (!PERSIST) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    if (PERSIST) stg_sel ^= 1;
                    else __syncthreads();                                             // the ring (= staging buffer) is free again
                }
            }
        }
    }
    // the block retires only when its tensor stores have been performed (not just read from shared memory): whatever follows
    // this kernel on the stream -- the exchange's flag barrier -- is ordered behind them
    if (EPI == DM_EPI_TANH && threadIdx.x == 0 && SM.count > 0 && SM.mode == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// shape-generic fallback (any alignment): one thread per output row, columns in registers
template <int PC>
__global__ void __launch_bounds__(128) gemm_fallback_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                                                            double* __restrict__ C, int64_t ldc, int64_t m, int64_t k, int64_t ncols, int bmod,
                                                            bool accumulate) {
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
    for (int j = 0; j < PC; j++) if (c0 + j < ncols) C[r + (c0 + j) * ldc] = accumulate ? C[r + (c0 + j) * ldc] + acc[j] : acc[j];
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
    return (lda % 2 == 0) && ((((uintptr_t)A) & 15) == 0) && m >= DM_BM && k >= DM_BK && encode_tiled_fn() != nullptr;
}

// Column-tile width for a contraction with `ncols` operand columns: the CTA tiles are equal-sized units dealt to the SMs
// (two resident per SM sharing the DMMA pipe), so the run time is ceil(units / SMs) * BN up to the extra L2 -> SM traffic
// of narrow tiles (the A tile is re-read once per column tile).
static int pick_bn(fdb_ctx* ctx, int64_t m, int64_t ncols) {
    if (ctx->dmma_bn == 16 || ctx->dmma_bn == 32 || ctx->dmma_bn == 64) return ctx->dmma_bn;
    const int64_t tm = cdiv(m, DM_BM);
    int best = 64; double best_cost = 1e300;
    const int cand[3] = {64, 32, 16};
    const double penalty[3] = {1.0, 1.04, 1.12};
    for (int q = 0; q < 3; q++) {
        const int64_t units = tm * cdiv(ncols, cand[q]);
        const double cost = (double)cdiv(units, ctx->sm_count) * cand[q] * penalty[q];
        if (cost < best_cost) { best_cost = cost; best = cand[q]; }
    }
    return best;
}
size_t gemm_workspace_doubles(int64_t k, int64_t ncols, int bmod) {
    // sized for the narrowest tile (most padding): 16-column tiles of 20-double columns
    const int64_t nslab = cdiv(k, DM_BK);
    const int64_t cols = (bmod > 0 && 16 % bmod == 0) ? 64 : round_up(ncols, 64);
    return (size_t)(cols * DM_BS * nslab);
}

template <int BN, int EPI, bool NS, bool MIR>
static int launch_gemm_bn(fdb_ctx* ctx, const CUtensorMap& tm, const double* ws, int btiles, double* C, int64_t ldc, int64_t m, int64_t k, int64_t ncols,
                          int accumulate, const TanhEpi& E, const fdb_mirrors& Mr, const StoreMaps& SM) {
    constexpr bool PERSIST = (EPI == DM_EPI_TANH && MIR && DM_PERSIST_MIRROR);
    constexpr int SMEM = PERSIST ? DmCfg<BN>::SMEM_P : (DM_BN16_THREE(EPI, BN) ? 3 * DmCfg<BN>::STAGE_BYTES + 2 * 3 * 8 + 1024 : DmCfg<BN>::SMEM);
    static unsigned long long configured = 0;
    if (first_use_on_device(configured, ctx->device)) {
        cudaError_t e = cudaFuncSetAttribute(dmma_gemm_kernel<BN, EPI, NS, MIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return fail(ctx, FDB_ERR_CUDA, "cudaFuncSetAttribute(dmma_gemm_kernel): %s", cudaGetErrorString(e));
    }
    const int tiles_m = (int)cdiv(m, DM_BM), tiles_n = (int)cdiv(ncols, BN);
    const int ntiles = tiles_m * tiles_n;
    const int grid = PERSIST ? (ntiles < ctx->sm_count ? ntiles : ctx->sm_count) : ntiles;
    dmma_gemm_kernel<BN, EPI, NS, MIR><<<(unsigned)grid, DM_THREADS, SMEM, ctx->stream>>>(tm, ws, btiles, C, ldc, m, k, ncols, tiles_m, tiles_n, accumulate, E,
                                                                                          Mr, SM);
    return FDB_OK;
}

// ws: gemm_workspace_doubles() doubles of device scratch for the packed B operand.  epi != null: the fused tanh epilogue
// (C unused); otherwise the product is stored to C.
int launch_gemm(fdb_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t m, int64_t k,
                int64_t ncols, int bmod, double* ws, bool* used_dmma, int* nlaunch, const TanhEpi* epi = nullptr, const fdb_mirrors* mirrors = nullptr,
                bool accumulate = false, int64_t onehot_c0 = -1) {
    if (m <= 0 || ncols <= 0) { if (nlaunch) *nlaunch = 0; return FDB_OK; }
    EncodeTiledFn enc = encode_tiled_fn();
    const bool ok = enc != nullptr && (lda % 2 == 0) && ((((uintptr_t)A) & 15) == 0) && k > 0 && ws != nullptr;
    if (used_dmma) *used_dmma = ok;
    if (!ok) {
        if (epi || onehot_c0 >= 0) return fail(ctx, FDB_ERR_UNSUPPORTED, "launch_gemm: the fused epilogue / generated operand need the tensor-core path");
        constexpr int PC = 13;
        dim3 grid((unsigned)cdiv(m, 128), (unsigned)cdiv(ncols, PC));
        gemm_fallback_kernel<PC><<<grid, 128, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, k, ncols, bmod, accumulate);
        if (nlaunch) *nlaunch = 1;
        return FDB_OK;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)m, (cuuint64_t)k};
    const cuuint64_t gstride[1] = {(cuuint64_t)lda * 8};
    const cuuint32_t box[2] = {16, DM_BK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ctx, FDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int BN = pick_bn(ctx, m, ncols);
    const int nslab = (int)cdiv(k, DM_BK);
    const int btiles = (bmod > 0 && BN % bmod == 0) ? 1 : (int)cdiv(ncols, BN);
    dim3 pg((unsigned)nslab, (unsigned)btiles);
    pack_b_kernel<<<pg, 128, 0, ctx->stream>>>(B, ldb, k, ncols, bmod, nslab, BN, onehot_c0, ws);
    TanhEpi E0 = {}; fdb_mirrors M0;
    const TanhEpi& E = epi ? *epi : E0;
    const fdb_mirrors& Mr = mirrors ? *mirrors : M0;
    thread_local static StoreMaps SM;    // 1 KB: kept out of the stack frame; filled per launch (copied into the launch's parameters)
    SM.count = 0; SM.mode = 0;
    // Staging pays off when the blocks also go to peers over NVLink (long contiguous runs); locally the threads' own streaming
    // stores are 1 % faster (5.49 vs 5.56 ms at C3).  ctx->dmma_tma_store: 0 = never stage, 1 = default (stage when mirroring),
    // 2 = always TMA tensor stores, 3 = always staged thread stores.
    const int want = ctx->dmma_tma_store == 1 ? (Mr.count > 0 ? DM_MIRROR_STORE_MODE : 0) : (ctx->dmma_tma_store == 2 ? 1 : (ctx->dmma_tma_store == 3 ? 2 : 0));
    if (epi && want == 2) { SM.count = 1; SM.mode = 2; }
    if (epi && want == 1 && (E.ldJ % 2 == 0) && ((((uintptr_t)E.J) & 15) == 0) && E.N * 4 <= 256) {
        // the Jacobian as a tensor: 128-row x 4N-column boxes for the epilogue's TMA stores, one map per destination arena
        const cuuint64_t jdim[2] = {(cuuint64_t)m, (cuuint64_t)E.n};
        const cuuint64_t jstride[1] = {(cuuint64_t)E.ldJ * 8};
        const cuuint32_t jbox[2] = {DM_BM, (cuuint32_t)(4 * E.N)};
        bool good = true;
        int nmaps = 0;
        for (int d = Mr.mc ? 1 : 0; d <= Mr.count && good; d++) {      // multicast: the one mapping at delta[0] reaches every arena, this rank's included
            char* base = reinterpret_cast<char*>(E.J) + (d == 0 ? 0 : Mr.delta[d - 1]);
            good = enc(&SM.map[nmaps++], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, jdim, jstride, jbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        SM.count = good ? nmaps : 0;
        SM.mode = 1;
    }
    int rc;
#define FDB_GEMM_BN(EPI_, NS_, MIR_)                                                                              \
    (BN == 64 ? launch_gemm_bn<64, EPI_, NS_, MIR_>(ctx, tm, ws, btiles, C, ldc, m, k, ncols, accumulate ? 1 : 0, E, Mr, SM)                \
     : BN == 32 ? launch_gemm_bn<32, EPI_, NS_, MIR_>(ctx, tm, ws, btiles, C, ldc, m, k, ncols, accumulate ? 1 : 0, E, Mr, SM)              \
                : launch_gemm_bn<16, EPI_, NS_, MIR_>(ctx, tm, ws, btiles, C, ldc, m, k, ncols, accumulate ? 1 : 0, E, Mr, SM))
    if (!epi) rc = FDB_GEMM_BN(DM_EPI_STORE, false, false);
    else if (ctx->nansafe) rc = Mr.count ? FDB_GEMM_BN(DM_EPI_TANH, true, true) : FDB_GEMM_BN(DM_EPI_TANH, true, false);
    else rc = Mr.count ? FDB_GEMM_BN(DM_EPI_TANH, false, true) : FDB_GEMM_BN(DM_EPI_TANH, false, false);
#undef FDB_GEMM_BN
    if (rc) return rc;
    if (nlaunch) *nlaunch = 2;
    return FDB_OK;
}

// ---- batched chunk sweeps of f(x) = tanh.(A*x .+ b) on the DMMA contraction ------------------------------
// Operand columns for the contraction.  SHARE: two columns per chunk, [x | 0] -- the value lane and the
// representative zero-partial lane (identical for every chunk, so the GEMM reads them with bmod = 2).
// All lanes: [x | one-hot seed of every requested structural position] (src/apiutils.jl:125-143): every partial lane of
// every chunk is a column of the contraction; the value lane, identical in every chunk, is contracted once
// (2*m*n*(n+1) flops for the whole Jacobian).
__global__ void build_operand_shared_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ B, int64_t ldb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    B[j] = x[j];
    B[ldb + j] = 0.0;
}
// column 0 = x; column 1 + q = one-hot seed of structural position c0 + q
__global__ void build_operand_onehot_kernel(const double* __restrict__ x, int64_t n, int64_t c0, int64_t ncolsJ, double* __restrict__ B, int64_t ldb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t q = blockIdx.y;                       // 0: value column; 1..: blocks of 16 one-hot columns
    if (j >= n) return;
    if (q == 0) { B[j] = x[j]; return; }
    for (int64_t col = (q - 1) * 16; col < min(q * 16, ncolsJ); col++) B[(1 + col) * ldb + j] = (j == c0 + col) ? 1.0 : 0.0;
}
// dense, per-chunk operand for consumers that need [value | N partial lanes] of each chunk (the two-layer target)
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
// epilogue on a stored product: z = A*xdual .+ b ; y = tanh.(z) ; J[:, start+k] = partials(y, k)
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
// the same for the one-hot operand layout: Y[:, 0] = A*x, Y[:, 1 + q] = partial lane of structural position c0 + q;
// block (row block, strip of 16 columns): one tanh per row and strip
template <bool NS, bool MIR>
__global__ void __launch_bounds__(256) tanh_affine_onehot_epilogue_kernel(const double* __restrict__ Y, int64_t ldy, const double* __restrict__ b, int64_t m,
                                                                         int64_t c0, int64_t ncolsJ, bool has_last, double* __restrict__ J, int64_t ldJ,
                                                                         double* __restrict__ yout, const __grid_constant__ fdb_mirrors Mr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double val = tanh(Y[r] + b[r]);
    const double deriv = 1.0 - val * val;
    const int64_t q0 = (int64_t)blockIdx.y * 16, q1 = min(q0 + 16, ncolsJ);
    for (int64_t q = q0; q < q1; q++) mstore_cs<MIR>(Mr, J + r + (c0 + q) * ldJ, mulp<NS>(Y[r + (1 + q) * ldy], deriv));
    if (yout && has_last && blockIdx.y == 0) mstore(Mr, &yout[r], val);
}

int tanh_affine_dmma(fdb_ctx* ctx, const fdb_array* A, int64_t lda, const fdb_array* b, const fdb_array* x, int64_t m, int N, int64_t lo, int64_t hi,
                     int64_t nchunks, int lcs, fdb_array* J, int64_t ldJ, double* yout, bool* used_dmma, const fdb_mirrors& Mr) {
    const int64_t n = x->n, nc = hi - lo;
    const int64_t ldb = round_up(n, 32), ldy = round_up(m, 32);
    if (ctx->lane_sharing) {
        // value lane + representative zero lane of every chunk through the contraction; everything after it fused into the tile epilogue
        const size_t wsd = gemm_workspace_doubles(n, nc * 2, 2);
        int rc = ensure_scratch(ctx, ((size_t)(ldb * 2) + wsd + 64) * sizeof(double)); if (rc) return rc;
        double* Bop = ctx->scratch;
        double* ws = Bop + ldb * 2; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
        build_operand_shared_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x->data, n, Bop, ldb);
        TanhEpi E = {A->data, lda, b->data, n, N, lo, nchunks, lcs, nc, J->data, ldJ, yout};
        int ngemm = 0;
        rc = launch_gemm(ctx, A->data, lda, Bop, ldb, nullptr, 0, m, n, nc * 2, 2, ws, used_dmma, &ngemm, &E, &Mr); if (rc) return rc;
        return finish(ctx, 1 + ngemm, "fdb_jacobian_chunked(tanh_affine, dmma, fused epilogue)");
    }
    // all lanes: [x | one-hot columns of the requested structural positions]
    const int64_t c0 = lo * N, c1 = (hi == nchunks) ? n : hi * N, ncolsJ = c1 - c0, P = 1 + ncolsJ;
    // the operand [x | one-hot seeds] is generated straight into the packed tiles (no materialised operand matrix)
    const size_t wsd = gemm_workspace_doubles(n, P, 0);
    int rc = ensure_scratch(ctx, ((size_t)(ldy * P) + wsd + 64) * sizeof(double)); if (rc) return rc;
    double* Y = ctx->scratch;
    double* ws = Y + ldy * P; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
    (void)ldb;
    int ngemm = 0;
    rc = launch_gemm(ctx, A->data, lda, x->data, n, Y, ldy, m, n, P, 0, ws, used_dmma, &ngemm, nullptr, nullptr, false, c0); if (rc) return rc;
    dim3 ge((unsigned)cdiv(m, 256), (unsigned)cdiv(ncolsJ, 16));
    const bool has_last = (hi == nchunks);
    if (ctx->nansafe) {
        if (Mr.count) tanh_affine_onehot_epilogue_kernel<true, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, b->data, m, c0, ncolsJ, has_last, J->data, ldJ, yout, Mr);
        else tanh_affine_onehot_epilogue_kernel<true, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, b->data, m, c0, ncolsJ, has_last, J->data, ldJ, yout, Mr);
    } else {
        if (Mr.count) tanh_affine_onehot_epilogue_kernel<false, true><<<ge, 256, 0, ctx->stream>>>(Y, ldy, b->data, m, c0, ncolsJ, has_last, J->data, ldJ, yout, Mr);
        else tanh_affine_onehot_epilogue_kernel<false, false><<<ge, 256, 0, ctx->stream>>>(Y, ldy, b->data, m, c0, ncolsJ, has_last, J->data, ldJ, yout, Mr);
    }
    return finish(ctx, 1 + ngemm, "fdb_jacobian_chunked(tanh_affine, dmma, all lanes)");
}

// The product lanes of `A * xdual` for chunks [lo, hi), left in the context's scratch for a consumer kernel that applies
// an arbitrary elementwise program to z = A*x (fdb_program_matvec_jacobian_chunked).  Lane sharing: 2 columns per chunk
// [value | zero lane]; all lanes: column 0 = value, column 1 + q = one-hot product of structural position lo*N + q.
int matvec_product_lanes(fdb_ctx* ctx, const double* A, int64_t lda, int64_t m, const fdb_array* x, int N, int64_t lo, int64_t hi, int64_t nchunks,
                         double** Y_out, int64_t* ldy_out, int* nlaunch) {
    const int64_t n = x->n, nc = hi - lo;
    const int64_t ldb = round_up(n, 32), ldy = round_up(m, 32);
    int ngemm = 0, rc;
    if (ctx->lane_sharing) {
        const size_t wsd = gemm_workspace_doubles(n, nc * 2, 2);
        rc = ensure_scratch(ctx, ((size_t)(ldb * 2 + ldy * nc * 2) + wsd + 64) * sizeof(double)); if (rc) return rc;
        double* Bop = ctx->scratch; double* Y = Bop + ldb * 2;
        double* ws = Y + ldy * nc * 2; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
        build_operand_shared_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x->data, n, Bop, ldb);
        rc = launch_gemm(ctx, A, lda, Bop, ldb, Y, ldy, m, n, nc * 2, 2, ws, nullptr, &ngemm); if (rc) return rc;
        *Y_out = Y;
    } else {
        const int64_t c0 = lo * N, c1 = (hi == nchunks) ? n : hi * N, ncolsJ = c1 - c0, P = 1 + ncolsJ;
        const size_t wsd = gemm_workspace_doubles(n, P, 0);
        rc = ensure_scratch(ctx, ((size_t)(ldb * P + ldy * P) + wsd + 64) * sizeof(double)); if (rc) return rc;
        double* Bop = ctx->scratch; double* Y = Bop + ldb * P;
        double* ws = Y + ldy * P; ws = (double*)((((uintptr_t)ws) + 127) & ~(uintptr_t)127);
        dim3 gb((unsigned)cdiv(n, 256), (unsigned)(1 + cdiv(ncolsJ, 16)));
        build_operand_onehot_kernel<<<gb, 256, 0, ctx->stream>>>(x->data, n, c0, ncolsJ, Bop, ldb);
        rc = launch_gemm(ctx, A, lda, Bop, ldb, Y, ldy, m, n, P, 0, ws, nullptr, &ngemm); if (rc) return rc;
        *Y_out = Y;
    }
    *ldy_out = ldy;
    if (nlaunch) *nlaunch = 1 + ngemm;
    return FDB_OK;
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

// C = A * B on matrices of Duals (both operands carry partials): the Dual*Dual body of src/dual.jl:260-300 summed over the inner
// index, i.e.  value(C) = value(A)*value(B),  partial_j(C) = value(A)*partial_j(B) + partial_j(A)*value(B):  2N+1 real FP64
// tensor-core contractions, the second product of every partial accumulated onto the first.  A, B, C are level-1 arrays
// holding column-major m x k, k x n and m x n matrices (leading dimensions lda, ldb, ldc within each plane).
extern "C" int fdb_dual_matmul(fdb_ctx* ctx, const fdb_array* A, int64_t lda, int64_t m, int64_t k, const fdb_array* B, int64_t ldb, int64_t n,
                               fdb_array* C, int64_t ldc) {
    FDB_ENTER(ctx);
    if (!ctx || !A || !B || !C) return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matmul: null argument");
    if (A->level != 1 || B->level != 1 || C->level != 1 || A->N != B->N || A->N != C->N)
        return fail(ctx, FDB_ERR_BADARG, "fdb_dual_matmul: A, B, C must be level-1 Dual arrays with the same chunk size (one tag, src/dual.jl:150-210)");
    if (m <= 0 || k <= 0 || n <= 0 || lda < m || ldb < k || ldc < m || A->n < lda * (k - 1) + m || B->n < ldb * (n - 1) + k || C->n < ldc * (n - 1) + m)
        return fail(ctx, FDB_ERR_SHAPE, "DimensionMismatch: A is %lld x %lld, B %lld x %lld, C %lld x %lld", (long long)m, (long long)k, (long long)k,
                    (long long)n, (long long)m, (long long)n);
    int rc = ensure_scratch(ctx, (gemm_workspace_doubles(k, n, 0) + 64) * sizeof(double)); if (rc) return rc;
    double* ws = (double*)((((uintptr_t)ctx->scratch) + 127) & ~(uintptr_t)127);
    const int N = A->N;
    int total = 0, nl = 0;
    for (int q = 0; q <= N; q++) {            // value(A) * plane q of B  ->  plane q of C
        rc = launch_gemm(ctx, A->data, lda, B->data + (int64_t)q * B->ld, ldb, C->data + (int64_t)q * C->ld, ldc, m, k, n, 0, ws, nullptr, &nl); if (rc) return rc;
        total += nl;
    }
    for (int q = 1; q <= N; q++) {            // partial q of A * value(B), accumulated
        rc = launch_gemm(ctx, A->data + (int64_t)q * A->ld, lda, B->data, ldb, C->data + (int64_t)q * C->ld, ldc, m, k, n, 0, ws, nullptr, &nl, nullptr, nullptr, true);
        if (rc) return rc;
        total += nl;
    }
    return finish(ctx, total, "fdb_dual_matmul");
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
