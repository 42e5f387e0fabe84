// smallk_b200 — FP64 tensor-core (DMMA.8x8x4) "skinny" GEMM for the NMF contractions.
//
// Every dense product of an NMF outer iteration has one small dimension (k, the
// rank) and streams the big matrix exactly once:
//
//   WtA  (k x n) = Wt (k x m) * A (m x n)            NN   reduction over m
//   HAt  (k x m) = H  (k x n) * A' (n x m)           NT   reduction over n
//   WtW  (k x k) = Wt * Wt',  HHt (k x k) = H * H'   NT   (B := the same matrix)
//   gradH(k x n) = WtW (k x k) * H (k x n) - WtA     NN   with the subtraction fused
//
// These replace the El::Gemm -> dgemm calls of the reference
// (common/include/dense_matrix_ops.hpp:255-270; call sites
// common/include/nmf_solver_bpp.hpp:330-374, nmf_solver_hals.hpp:155-196,
// nmf_solver_mu.hpp:110-161, nmf_solver_rank2.hpp:342-452).
//
// Shape of the kernel: C[M x N] = Aop[M x R] * Bop[R x N] with M = k small.
//   * Aop is column-major M x R (contiguous along M).
//   * NN: B is column-major R x N (contiguous along the reduction).
//     NT: B is column-major N x R (contiguous along N), used transposed.
//   * CTA tile 64 x 128, reduction chunk 32, 8 warps as 2 x 4, warp tile 32 x 32
//     = 4 x 4 DMMA.8x8x4 accumulators (32 doubles / lane).
//   * cp.async 3-stage shared-memory pipeline, zero-filled tails; smem leading
//     dimensions are = 4 (mod 16) doubles so every fragment LDS.64 is conflict free.
//   * Split over the reduction ("split-R") so the grid fills 148 SMs whatever the
//     shape. The splits of one output tile are neighbours in launch order (they run
//     in the same wave), each stores its partial tile in the workspace (fragment
//     order: fully coalesced, L2-resident) and takes a ticket; the CTA that arrives
//     LAST adds the partials in ASCENDING SPLIT ORDER — a fixed order, so results do
//     not depend on scheduling — and writes the tile. No second kernel, no DRAM round
//     trip of the partials (r01: reduce_partials_kernel, 0.31 GB per big product).
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"
#include "peer_device.cuh"

namespace smk {

namespace {

constexpr int BM = 64, BN = 128, BK = 32;
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr int LDA_S = BM + 4;    // As[BK][LDA_S]
constexpr int LDB_NN = BK + 4;   // Bs[BN][LDB_NN]
constexpr int LDB_NT = BN + 4;   // Bs[BK][LDB_NT]
constexpr int A_STAGE = BK * LDA_S;
constexpr int B_STAGE_NN = BN * LDB_NN;
constexpr int B_STAGE_NT = BK * LDB_NT;

template <bool NT>
__host__ __device__ constexpr int stage_doubles() { return A_STAGE + (NT ? B_STAGE_NT : B_STAGE_NN); }

struct GemmParams
{
    const double* A; long long lda;
    const double* B; long long ldb;
    double* C; long long ldc;          // used when splits == 1
    double* partial;                   // [splits][M*N], ld = M, used when splits > 1
    const double* D; long long ldd;    // optional: C = acc - D (splits == 1 only)
    int M, N, R;
    int splits, rchunk;                // reduction range per split, multiple of BK
    int to_partial;                    // write the tile to `partial` even when splits == 1 (the caller reduces / forwards it)
    int fixup;                         // splits > 1: in-kernel reduction by the last-arriving CTA of each tile (1-D grid)
    int ntn;                           // tiles along N
    unsigned int* tickets;             // one arrival counter per tile, zero on entry, left zero
    double* slots;                     // [tile][split][32][256] partial tiles in fragment order
    // multi-GPU H*A': the finished tile goes straight to the receive slot of the rank that owns its columns (peer.cu layout),
    // and the CTA that finishes the LAST tile publishes the epoch to every peer: product, reduction and NVLink transfer in one kernel
    GemmScatter sc;
};

// Copy `nvec` vectors of `len` contiguous doubles (a tile) into shared memory.
// vec v lives at g + v*ldg, goes to s + v*lds; only the first vvalid vectors and
// the first lvalid doubles of each exist, the rest is zero-filled.
template <int LEN, int NVEC, int LDS, int VEC>
__device__ __forceinline__ void load_tile(double* s, const double* g, long long ldg, int vvalid, int lvalid, int tid)
{
    constexpr int CH_PER_VEC = LEN / VEC;
    constexpr int CHUNKS = CH_PER_VEC * NVEC;
#pragma unroll
    for (int c0 = 0; c0 < CHUNKS; c0 += THREADS)
    {
        int c = c0 + tid;
        if (CHUNKS % THREADS != 0 && c >= CHUNKS) break;
        int v = c / CH_PER_VEC;
        int off = (c % CH_PER_VEC) * VEC;
        int rem = (v < vvalid) ? (lvalid - off) : 0;
        rem = rem < 0 ? 0 : (rem > VEC ? VEC : rem);
        const double* src = (rem > 0) ? (g + static_cast<long long>(v) * ldg + off) : g;
        if (VEC == 2) cp_async16(s + v * LDS + off, src, rem * 8);
        else          cp_async8(s + v * LDS + off, src, rem * 8);
    }
}

// What every GEMM kernel of this file does with its 64 x 128 accumulator tile: (FIX) the in-kernel split-R reduction by the
// last-arriving CTA of the tile, then the store — to C (minus D), to the split's partial tile, or (multi-GPU H*A') straight into
// the receive slot of the rank that owns the columns. All threads of the CTA call it (it synchronises); threads with
// `active` false (the TMA kernel's producer warp) own no accumulators and only take part in the barriers.
template <bool FIX>
__device__ __forceinline__ void gemm_finish(const GemmParams& p, double (&acc)[4][4][2], int tile_m, int tile_n, int split, int tid, bool active)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, t4 = lane & 3;
    const int n0 = tile_n * BN;
    const int m0 = tile_m * BM;
    if (FIX && p.splits > 1)
    {
        // ---- split-R fix-up: partial tile to the workspace, ticket, the last arriver sums in split order
        __shared__ bool s_last;
        const int tile = tile_m * p.ntn + tile_n;
        double* slot0 = p.slots + static_cast<long long>(tile) * p.splits * (BM * BN);
        double* mine = slot0 + static_cast<long long>(split) * (BM * BN) + tid;
        if (active)
        {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    mine[((i * 4 + j) * 2 + 0) * THREADS] = acc[i][j][0];
                    mine[((i * 4 + j) * 2 + 1) * THREADS] = acc[i][j][1];
                }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.tickets + tile, 1u) == static_cast<unsigned int>(p.splits - 1));
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (tid == 0) p.tickets[tile] = 0u;                 // ready for the next launch
        for (int z = 0; active && z < p.splits; ++z)
        {
            const double* src = slot0 + static_cast<long long>(z) * (BM * BN) + tid;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const double v0 = __ldcg(src + ((i * 4 + j) * 2 + 0) * THREADS), v1 = __ldcg(src + ((i * 4 + j) * 2 + 1) * THREADS);
                    if (z == 0) { acc[i][j][0] = v0; acc[i][j][1] = v1; }
                    else { acc[i][j][0] += v0; acc[i][j][1] += v1; }
                }
        }
    }

    // epilogue
    double* out;
    long long ldo;
    if (!FIX && (p.splits > 1 || p.to_partial)) { out = p.partial + static_cast<long long>(split) * p.M * p.N; ldo = p.M; }
    else              { out = p.C; ldo = p.ldc; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= p.M || !active) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
                const int col = n0 + wn * 32 + j * 8 + 2 * t4 + e;
                if (col >= p.N) continue;
                double v = acc[i][j][e];
                if ((FIX || (p.splits == 1 && !p.to_partial)) && p.D) v -= p.D[static_cast<long long>(col) * p.ldd + row];
                if (FIX && p.sc.nranks > 0)
                {
                    const int g = col / p.sc.cols_per_rank;
                    double* dst = reinterpret_cast<double*>(p.sc.table.base[g] + p.sc.recv_off) + static_cast<long long>(p.sc.rank) * p.sc.piece +
                                  static_cast<long long>(col - g * p.sc.cols_per_rank) * p.M + row;
                    *dst = v;
                }
                else out[static_cast<long long>(col) * ldo + row] = v;
            }
        }
    }
    if (FIX && p.sc.nranks > 0)
    {
        // all tiles stored -> publish. Every CTA that wrote a final tile counts; the last one signals the peers.
        __shared__ bool s_pub;
        __threadfence_system();
        __syncthreads();
        if (tid == 0)
        {
            const unsigned int t = atomicAdd(p.sc.done, 1u);
            s_pub = (t == static_cast<unsigned int>(p.sc.ntiles - 1));
            if (s_pub) *p.sc.done = 0u;
        }
        __syncthreads();
        if (s_pub && tid < p.sc.nranks)
            st_release_sys(reinterpret_cast<unsigned long long*>(p.sc.table.base[tid] + kPeerFlagOffset) + kFlagScatter * kPeerMaxRanks + p.sc.rank,
                           p.sc.epoch);
    }
}

// FIX = false: the plain kernel (grid = tiles x splits, tile or split-R partial written as is).
// FIX = true : 1-D grid, in-kernel split-R reduction by the last-arriving CTA of each tile and, optionally, the multi-GPU scatter
//              epilogue. A separate instantiation so that the plain kernel's code (registers, schedule) is untouched by it.
template <bool NT, int VEC, bool FIX>
__global__ void __launch_bounds__(THREADS, 2) gemm_skinny_kernel(GemmParams p)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;       // 2 x 4 warps
    const int g = lane >> 2, t4 = lane & 3;

    int tile_n, tile_m, split_;
    // split-major launch order in both forms: concurrently running CTAs work on the same reduction range of neighbouring tiles
    // (measured r02: with the splits of one tile as neighbours the same kernel is 6 % slower)
    if (FIX) { const int ntiles = p.ntn * ((p.M + BM - 1) / BM); split_ = blockIdx.x / ntiles; const int tile = blockIdx.x % ntiles; tile_n = tile % p.ntn; tile_m = tile / p.ntn; }
    else { tile_n = blockIdx.x; tile_m = blockIdx.y; split_ = blockIdx.z; }
    const int n0 = tile_n * BN;
    const int m0 = tile_m * BM;
    const int split = split_;
    const int r_begin = split * p.rchunk;
    const int r_end = min(p.R, r_begin + p.rchunk);
    const int nchunks = (r_end > r_begin) ? (r_end - r_begin + BK - 1) / BK : 0;

    const int mvalid = p.M - m0;   // rows of this tile that exist (may exceed BM)
    const int nvalid = p.N - n0;

    auto issue = [&](int chunk, int slot) {
        double* As = smem + slot * stage_doubles<NT>();
        double* Bs = As + A_STAGE;
        const int r0 = r_begin + chunk * BK;
        const int rvalid = r_end - r0;
        // Aop tile: BK vectors (reduction index) of BM contiguous rows
        load_tile<BM, BK, LDA_S, VEC>(As, p.A + static_cast<long long>(r0) * p.lda + m0, p.lda, rvalid, mvalid, tid);
        if (NT)   // BK vectors (reduction index) of BN contiguous columns-of-C
            load_tile<BN, BK, LDB_NT, VEC>(Bs, p.B + static_cast<long long>(r0) * p.ldb + n0, p.ldb, rvalid, nvalid, tid);
        else      // BN vectors (column of C) of BK contiguous reduction entries
            load_tile<BK, BN, LDB_NN, VEC>(Bs, p.B + static_cast<long long>(n0) * p.ldb + r0, p.ldb, nvalid, rvalid, tid);
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s)
    {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }

    for (int it = 0; it < nchunks; ++it)
    {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = it + STAGES - 1;
            if (nxt < nchunks) issue(nxt, nxt % STAGES);
            cp_async_commit();
        }
        const double* As = smem + (it % STAGES) * stage_doubles<NT>();
        const double* Bs = As + A_STAGE;
        const double* a_ptr = As + t4 * LDA_S + wm * 32 + g;
        const double* b_ptr = NT ? (Bs + t4 * LDB_NT + wn * 32 + g) : (Bs + (wn * 32 + g) * LDB_NN + t4);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4)
        {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = a_ptr[kk * LDA_S + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = NT ? b_ptr[kk * LDB_NT + j * 8] : b_ptr[j * 8 * LDB_NN + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    gemm_finish<FIX>(p, acc, tile_m, tile_n, split, tid, true);
}

// ---------------------------------------------------------------------------
// The same product with the tiles moved by TMA (cp.async.bulk.tensor + mbarrier), for the A-sized contractions.
//
//   * One producer warp (one elected lane) issues the tensor copies of a reduction chunk — the 64 x 32 tile of Aop and the
//     128 x 32 tile of Bop, 12 (NN) or 24 (NT) boxes — into a two-stage ring; eight consumer warps wait on the stage's `full` mbarrier, run their
//     8 x 16 DMMA.8x8x4 and arrive on its `empty` mbarrier. No block-wide barrier in the main loop, no per-thread address
//     arithmetic or copy instructions in the math warps (the cp.async form spends 12 LDGSTS + their addresses per thread
//     and chunk and one __syncthreads per chunk).
//   * Layout. A TMA box is written densely, so the padded leading dimensions that keep the cp.async form free of bank
//     conflicts are not available; the 64-byte swizzle is used instead. Every box has an innermost extent of 8 doubles
//     (64 bytes) of the CONTIGUOUS global dimension:
//         Aop  (M x R, contiguous in M)      : map (M, R), box (8, 32),  8 boxes per tile  -> smem [m/8][r][m%8]
//         B NT (N x R, contiguous in N)      : map (N, R), box (8, 32), 16 boxes per tile  -> smem [n/8][r][n%8]
//         B NN (R x N, contiguous in R)      : map (R, N), box (8, 128), 4 boxes per tile  -> smem [r/8][n][r%8]
//     64-byte rows XOR-swizzled by address bits 7-8 are conflict-free for the m8n8k4 fragments PROVIDED the four reduction
//     indices a quad of lanes holds are {0, 1, 4, 5} (+2 for the second step) of each group of eight —
//     rho(t4, s) = (t4 & 1) + 4 (t4 >> 1) + 2 s — instead of four consecutive ones: then the 16 lanes of a half-warp hit 16
//     different 8-byte banks in all three layouts (bank bits = {r & 1 | n & 1, 16-byte chunk ^ row bits, low bit of the
//     contiguous index}). The order in which the reduction indices of a group of eight are added is therefore
//     0,1,4,5 | 2,3,6,7: fixed, but not the cp.async kernel's.
//   * Out-of-range parts of a box (ragged M or N, the tail of the reduction) are zero-filled by TMA.
// Needs even leading dimensions and 16-byte aligned bases; gemm_f64 falls back to the cp.async kernel otherwise.
// ---------------------------------------------------------------------------
constexpr int TMA_STAGES = 2;
constexpr int TMA_THREADS = THREADS + 32;                 // 8 math warps + the producer warp
constexpr int TMA_A_BYTES = BM * BK * 8, TMA_B_BYTES = BN * BK * 8;
constexpr int TMA_STAGE_BYTES = TMA_A_BYTES + TMA_B_BYTES;
constexpr int TMA_SMEM_BYTES = TMA_STAGES * TMA_STAGE_BYTES + 1024 /* alignment slack */ + 64 /* mbarriers */;

__device__ __forceinline__ void mbar_init(unsigned int bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned int dst, const CUtensorMap* map, int c0, int c1, unsigned int bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <bool NT, bool FIX>
__global__ void __launch_bounds__(TMA_THREADS, 2)
gemm_tma_kernel(GemmParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB)
{
    extern __shared__ unsigned char tma_smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool math = warp < THREADS / 32;
    const int wm = warp & 1, wn = (warp >> 1) & 3;     // 2 x 4 math warps
    const int g = lane >> 2, t4 = lane & 3;

    int tile_n, tile_m, split;
    if (FIX) { const int ntiles = p.ntn * ((p.M + BM - 1) / BM); split = blockIdx.x / ntiles; const int tile = blockIdx.x % ntiles; tile_n = tile % p.ntn; tile_m = tile / p.ntn; }
    else { tile_n = blockIdx.x; tile_m = blockIdx.y; split = blockIdx.z; }
    const int n0 = tile_n * BN;
    const int m0 = tile_m * BM;
    const int r_begin = split * p.rchunk;
    const int r_end = min(p.R, r_begin + p.rchunk);
    const int nchunks = (r_end > r_begin) ? (r_end - r_begin + BK - 1) / BK : 0;

    // stage ring, 1024-byte aligned (the swizzle pattern is a function of the shared-memory address bits), then the mbarriers
    const unsigned int raw = static_cast<unsigned int>(__cvta_generic_to_shared(tma_smem_raw));
    const unsigned int base = (raw + 1023u) & ~1023u;
    const unsigned int bars = base + TMA_STAGES * TMA_STAGE_BYTES;       // full[s] at bars + 8 s, empty[s] at bars + 8 (TMA_STAGES + s)
    const unsigned char* stage0 = tma_smem_raw + (base - raw);
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (TMA_STAGES + s), THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (!math)
    {
        if (lane == 0)
        {
            for (int it = 0; it < nchunks; ++it)
            {
                const int s = it % TMA_STAGES;
                if (it >= TMA_STAGES) mbar_wait(bars + 8 * (TMA_STAGES + s), ((it / TMA_STAGES) - 1) & 1);
                const unsigned int full = bars + 8 * s;
                const unsigned int As = base + s * TMA_STAGE_BYTES, Bs = As + TMA_A_BYTES;
                const int r0 = r_begin + it * BK;
                mbar_expect_tx(full, TMA_STAGE_BYTES);
#pragma unroll
                for (int b = 0; b < BM / 8; ++b) tma_load_2d(As + b * (8 * BK * 8), &mapA, m0 + 8 * b, r0, full);
                if (NT)
                {
#pragma unroll
                    for (int b = 0; b < BN / 8; ++b) tma_load_2d(Bs + b * (8 * BK * 8), &mapB, n0 + 8 * b, r0, full);
                }
                else
                {
#pragma unroll
                    for (int b = 0; b < BK / 8; ++b) tma_load_2d(Bs + b * (8 * BN * 8), &mapB, r0 + 8 * b, n0, full);
                }
            }
        }
    }
    else
    {
        // byte offsets of this lane's fragment elements inside a stage, for the two reduction steps s2 of a group of eight
        unsigned int offA[2], offB[2];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
        {
            const int rho = (t4 & 1) + 4 * (t4 >> 1) + 2 * s2;
            const int sw = (rho >> 1) & 3;              // row bits 1-2 of a [..][r][8] tile = the swizzle term
            offA[s2] = static_cast<unsigned int>(((wm * 4) * 32 + rho) * 64 + (((g >> 1) ^ sw) << 4) + ((g & 1) << 3));
            if (NT) offB[s2] = static_cast<unsigned int>(((wn * 4) * 32 + rho) * 64 + (((g >> 1) ^ sw) << 4) + ((g & 1) << 3));
            else    offB[s2] = static_cast<unsigned int>((wn * 32 + g) * 64 + (((rho >> 1) ^ ((g >> 1) & 3)) << 4) + ((rho & 1) << 3));
        }
        for (int it = 0; it < nchunks; ++it)
        {
            const int s = it % TMA_STAGES;
            mbar_wait(bars + 8 * s, (it / TMA_STAGES) & 1);
            const unsigned char* As = stage0 + s * TMA_STAGE_BYTES;
            const unsigned char* Bs = As + TMA_A_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < BK / 8; ++k8)
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2)
                {
                    double a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double*>(As + offA[s2] + i * 2048 + k8 * 512);
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double*>(Bs + offB[s2] + (NT ? j * 2048 + k8 * 512 : j * 512 + k8 * 8192));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (TMA_STAGES + s));
        }
    }
    gemm_finish<FIX>(p, acc, tile_m, tile_n, split, tid, math);
}

// C = sum_s partial[s] (- D), partials added in ascending split order.
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int splits, int M, int N,
                                       double* __restrict__ C, long long ldc,
                                       const double* __restrict__ D, long long ldd)
{
    const long long total = static_cast<long long>(M) * N;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        double s = partial[i];
        for (int z = 1; z < splits; ++z) s += partial[static_cast<long long>(z) * total + i];
        const int row = static_cast<int>(i % M);
        const long long col = i / M;
        if (D) s -= D[col * ldd + row];
        C[col * ldc + row] = s;
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// 2-D map of a column-major d0 x d1 matrix of doubles (leading dimension ld), boxes of b0 x b1, 64-byte swizzle
bool make_map_2d(CUtensorMap* map, const double* base, long long d0, long long d1, long long ld, int b0, int b1)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(double)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(b0), static_cast<cuuint32_t>(b1)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace

size_t gemm_workspace_bytes(int M, int N, int max_splits)
{
    return static_cast<size_t>(max_splits) * M * N * sizeof(double);
}

// The first kGemmTicketBytes of every split-R workspace hold the per-tile arrival counters of the in-kernel fix-up; they
// must be zero before the first product (gemm_workspace_prepare) and every launch leaves them zero.
constexpr size_t kGemmTicketBytes = 64 * 1024;
constexpr long long kGemmMaxTickets = kGemmTicketBytes / sizeof(unsigned int);

void gemm_workspace_prepare(cudaStream_t stream, double* workspace, size_t workspace_bytes)
{
    if (workspace && workspace_bytes >= kGemmTicketBytes) SMK_CUDA(cudaMemsetAsync(workspace, 0, kGemmTicketBytes, stream));
}

// Pick the split count that best fills `slots` CTAs-in-flight.
int gemm_pick_splits(int M, int N, int R, int num_sms, size_t workspace_bytes, bool whole_tiles)
{
    const long long tiles = static_cast<long long>(ceil_div(N, BN)) * ceil_div(M, BM);
    const int max_by_r = ceil_div(R, BK);
    // the in-kernel reduction stores partial tiles whole (64 x 128 slots, whatever part of them exists); the reduction kernel's
    // layout is [split][M x N]
    const size_t usable = workspace_bytes > kGemmTicketBytes ? workspace_bytes - kGemmTicketBytes : 0;
    const size_t per_split = whole_tiles ? static_cast<size_t>(BM * BN) * static_cast<size_t>(tiles) : static_cast<size_t>(M) * N;
    long long max_by_ws = static_cast<long long>(usable / (sizeof(double) * per_split));
    const int chunks = ceil_div(R, BK);
    const bool few_tiles = tiles < 8;                      // Gram matrices: one tile, the reduction kernel adds hundreds of partials grid-wide
    int smax = static_cast<int>(std::min<long long>(std::min<long long>(max_by_r, max_by_ws), few_tiles ? 4LL * num_sms : 32LL));
    if (smax < 1) smax = 1;
    if (tiles >= 4LL * num_sms) smax = std::min(smax, 1);   // plenty of tiles already
    int best = 1;
    if (few_tiles)
    {
        // the split that fills the waves best (ties: fewer splits)
        double best_eff = -1.0;
        for (int s = 1; s <= smax; ++s)
        {
            const int per = ceil_div(chunks, s);               // chunks per CTA (rchunk is a multiple of BK)
            const int eff_s = ceil_div(chunks, per);
            const long long ctas = tiles * eff_s;
            const long long waves = (ctas + num_sms - 1) / num_sms;
            const double eff = static_cast<double>(ctas) / (static_cast<double>(waves) * num_sms);
            const double score = eff - 1e-4 * s;
            if (score > best_eff + 1e-12) { best_eff = score; best = s; }
        }
        return best;
    }
    // The big products: a cost model in units of one reduction chunk, fitted to a sweep of s = 1..32 on the column shards of C2
    // at 1, 2, 4 and 8 GPUs (tools/sweep_splits.py, profiles/sweep_r02_splits_c2.json; its pick is within 1.3 % of the best
    // measured split on average, 5.6 % at worst). A CTA costs (chunks + kOverhead): pipeline prologue, epilogue, partial tile.
    // `num_sms` CTAs run at a time (two per SM); a last wave that leaves every SM at most ONE CTA costs only kLoneWave of a
    // full wave, because a CTA that has its SM to itself runs almost twice as fast as two that share the FP64 pipe.
    constexpr double kOverhead = 2.0, kLoneWave = 0.55;
    double best_cost = 1e300;
    for (int s = 1; s <= smax; ++s)
    {
        const int per = ceil_div(chunks, s);
        const int eff_s = ceil_div(chunks, per);
        const long long ctas = tiles * eff_s;
        const long long full = ctas / num_sms, rem = ctas % num_sms;
        const double unit = per + kOverhead;
        const double cost = unit * (static_cast<double>(full) + (rem == 0 ? 0.0 : (rem <= num_sms / 2 ? kLoneWave : 1.0)));
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    return best;
}

void gemm_f64(cudaStream_t stream, bool nt, int M, int N, int R,
              const double* A, long long lda, const double* B, long long ldb,
              double* C, long long ldc, const double* D, long long ldd,
              double* workspace, size_t workspace_bytes, int num_sms, int* partials_only, const GemmScatter* scatter)
{
    if (M <= 0 || N <= 0) { if (partials_only) *partials_only = 0; return; }
    GemmParams p;
    if (scatter) p.sc = *scatter; else p.sc.nranks = 0;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.D = D; p.ldd = ldd;
    p.partial = workspace;
    p.M = M; p.N = N; p.R = R;
    const char* fx = getenv("SMK_GEMM_FIXUP");              // =1: in-kernel reduction on one GPU too (measurements)
    const bool want_fixup = scatter || (!partials_only && fx && atoi(fx) == 1 && M * static_cast<long long>(N) > 65536);
    int splits = (workspace && R > 0) ? gemm_pick_splits(M, N, R, 2 * num_sms, workspace_bytes, want_fixup) : 1;   // 2 CTAs per SM
    if (workspace && R > 0 && M * static_cast<long long>(N) > 65536)
    {
        // SMK_GEMM_SPLITS_NN / SMK_GEMM_SPLITS_NT = n force the split count of the big products W'A / H A' (measurements)
        const char* e = getenv(nt ? "SMK_GEMM_SPLITS_NT" : "SMK_GEMM_SPLITS_NN");
        const int forced = e ? atoi(e) : 0;
        if (forced > 0)
        {
            const long long tl = static_cast<long long>(ceil_div(N, BN)) * ceil_div(M, BM);
            const size_t usable = workspace_bytes > kGemmTicketBytes ? workspace_bytes - kGemmTicketBytes : 0;
            const long long cap = static_cast<long long>(usable / (sizeof(double) * static_cast<size_t>(BM * BN) * static_cast<size_t>(tl)));
            splits = static_cast<int>(std::max<long long>(1, std::min<long long>(forced, cap)));
        }
    }
    int rchunk = ceil_div(ceil_div(R > 0 ? R : 1, splits), BK) * BK;
    splits = R > 0 ? ceil_div(R, rchunk) : 1;
    p.splits = splits; p.rchunk = rchunk;
    p.to_partial = partials_only ? 1 : 0;
    p.ntn = ceil_div(N, BN);
    const long long tiles = static_cast<long long>(p.ntn) * ceil_div(M, BM);
    // In-kernel reduction (the last-arriving CTA of a tile adds its partials): used where the finished tile has to leave through
    // the kernel's own epilogue, i.e. the multi-GPU scatter. On one GPU the separate grid-wide reduction kernel is FASTER
    // (measured r02, C2: 1.66 ms product + reduction kernel vs 1.71 ms with the fix-up: the tile-major launch order and the
    // serial additions of the last arriver cost more than the 75 us kernel they replace), and the Gram matrices (one tile,
    // hundreds of splits) need the grid-wide reduction anyway.
    p.fixup = (want_fixup && splits > 1 && splits <= 32 && tiles <= kGemmMaxTickets) ? 1 : 0;
    if (scatter)
    {
        if (splits > 1 && !p.fixup) throw std::string("gemm_f64: the scatter epilogue needs the in-kernel reduction");
        p.sc.ntiles = static_cast<int>(tiles);
    }
    p.tickets = reinterpret_cast<unsigned int*>(workspace);
    p.slots = workspace ? workspace + kGemmTicketBytes / sizeof(double) : nullptr;
    if (!p.fixup && workspace) p.partial = workspace + kGemmTicketBytes / sizeof(double);     // [splits][M x N] layout behind the tickets

    const bool vec2 = aligned16(A) && aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0);
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    const bool fix = p.fixup || scatter;          // the scatter epilogue lives in the FIX instantiation also when splits == 1
    if (fix) grid = dim3(static_cast<unsigned int>(tiles * splits), 1, 1);

    // the A-sized contractions: tiles by TMA (SMK_GEMM_TMA=0 keeps the cp.async kernel: measurements)
    const char* tma_env = getenv("SMK_GEMM_TMA");                  // read per call: the tests run both kernels in one process
    const bool tma_on = !(tma_env && atoi(tma_env) == 0);
    bool launched = false;
    const char* tma_small_env = getenv("SMK_GEMM_TMA_SHORT");      // =0: products with a short reduction and no split (G * X - R) stay on the cp.async kernel
    const bool tma_short = !(tma_small_env && atoi(tma_small_env) == 0);
    if (tma_on && vec2 && (workspace || tma_short) && R >= 4 * BK && M * static_cast<long long>(N) > 65536 && lda < (1LL << 36) && ldb < (1LL << 36))
    {
        CUtensorMap mapA, mapB;
        const bool ok = make_map_2d(&mapA, A, M, R, lda, 8, BK) &&
                        (nt ? make_map_2d(&mapB, B, N, R, ldb, 8, BK) : make_map_2d(&mapB, B, R, N, ldb, 8, BN));
        if (ok)
        {
            auto launch_tma = [&](auto kern) {
                SMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
                kern<<<grid, TMA_THREADS, TMA_SMEM_BYTES, stream>>>(p, mapA, mapB);
                SMK_LAUNCH_CHECK();
            };
            if (nt) { if (fix) launch_tma(gemm_tma_kernel<true, true>); else launch_tma(gemm_tma_kernel<true, false>); }
            else    { if (fix) launch_tma(gemm_tma_kernel<false, true>); else launch_tma(gemm_tma_kernel<false, false>); }
            launched = true;
        }
    }
    const size_t smem = static_cast<size_t>(STAGES) * (nt ? stage_doubles<true>() : stage_doubles<false>()) * sizeof(double);

    auto launch = [&](auto kern) {
        SMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, THREADS, smem, stream>>>(p);
        SMK_LAUNCH_CHECK();
    };
    if (launched) {}
    else if (fix)
    {
        if (nt) { if (vec2) launch(gemm_skinny_kernel<true, 2, true>); else launch(gemm_skinny_kernel<true, 1, true>); }
        else    { if (vec2) launch(gemm_skinny_kernel<false, 2, true>); else launch(gemm_skinny_kernel<false, 1, true>); }
    }
    else
    {
        if (nt) { if (vec2) launch(gemm_skinny_kernel<true, 2, false>); else launch(gemm_skinny_kernel<true, 1, false>); }
        else    { if (vec2) launch(gemm_skinny_kernel<false, 2, false>); else launch(gemm_skinny_kernel<false, 1, false>); }
    }

    if (partials_only) { *partials_only = splits; return; }     // gemm_partials(workspace) = [splits][M x N] tiles, ld = M; the caller sums them
    if (splits > 1 && !p.fixup)
    {
        const long long total = static_cast<long long>(M) * N;
        int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * num_sms));
        reduce_partials_kernel<<<blocks, 256, 0, stream>>>(p.partial, splits, M, N, C, ldc, D, ldd);
        SMK_LAUNCH_CHECK();
    }
}

const double* gemm_partials(const double* workspace) { return workspace + kGemmTicketBytes / sizeof(double); }

} // namespace smk
