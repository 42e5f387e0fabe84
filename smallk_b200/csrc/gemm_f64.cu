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
//     shape; partial tiles go to a workspace and are summed in a FIXED order by
//     reduce_partials_kernel, so results do not depend on scheduling.
#include "common.cuh"
#include "kernels.h"

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

template <bool NT, int VEC>
__global__ void __launch_bounds__(THREADS, 2) gemm_skinny_kernel(GemmParams p)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;       // 2 x 4 warps
    const int g = lane >> 2, t4 = lane & 3;

    const int n0 = blockIdx.x * BN;
    const int m0 = blockIdx.y * BM;
    const int split = blockIdx.z;
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

    // epilogue
    double* out;
    long long ldo;
    if (p.splits > 1 || p.to_partial) { out = p.partial + static_cast<long long>(split) * p.M * p.N; ldo = p.M; }
    else              { out = p.C; ldo = p.ldc; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
                const int col = n0 + wn * 32 + j * 8 + 2 * t4 + e;
                if (col >= p.N) continue;
                double v = acc[i][j][e];
                if (p.splits == 1 && !p.to_partial && p.D) v -= p.D[static_cast<long long>(col) * p.ldd + row];
                out[static_cast<long long>(col) * ldo + row] = v;
            }
        }
    }
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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace

size_t gemm_workspace_bytes(int M, int N, int max_splits)
{
    return static_cast<size_t>(max_splits) * M * N * sizeof(double);
}

// Pick the split count that best fills `slots` CTAs-in-flight.
int gemm_pick_splits(int M, int N, int R, int num_sms, size_t workspace_bytes)
{
    const long long tiles = static_cast<long long>(ceil_div(N, BN)) * ceil_div(M, BM);
    const int max_by_r = ceil_div(R, BK);
    long long max_by_ws = static_cast<long long>(workspace_bytes / (sizeof(double) * static_cast<size_t>(M) * N));
    int smax = static_cast<int>(std::min<long long>(std::min<long long>(max_by_r, max_by_ws), 4LL * num_sms));
    if (smax < 1) smax = 1;
    if (tiles >= 4LL * num_sms) smax = std::min(smax, 1);   // plenty of tiles already
    double best_eff = -1.0;
    int best = 1;
    for (int s = 1; s <= smax; ++s)
    {
        // real chunking: rchunk is a multiple of BK, so the effective split count may be lower
        int rchunk = ceil_div(ceil_div(R, s), BK) * BK;
        int eff_s = ceil_div(R, rchunk);
        long long ctas = tiles * eff_s;
        long long waves = (ctas + num_sms - 1) / num_sms;
        double eff = static_cast<double>(ctas) / (static_cast<double>(waves) * num_sms);
        // prefer fewer splits on ties (less partial traffic); small penalty per split
        double score = eff - 1e-4 * s;
        if (score > best_eff + 1e-12) { best_eff = score; best = s; }
    }
    return best;
}

void gemm_f64(cudaStream_t stream, bool nt, int M, int N, int R,
              const double* A, long long lda, const double* B, long long ldb,
              double* C, long long ldc, const double* D, long long ldd,
              double* workspace, size_t workspace_bytes, int num_sms, int* partials_only)
{
    if (M <= 0 || N <= 0) { if (partials_only) *partials_only = 0; return; }
    GemmParams p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.D = D; p.ldd = ldd;
    p.partial = workspace;
    p.M = M; p.N = N; p.R = R;
    int splits = (workspace && R > 0) ? gemm_pick_splits(M, N, R, 2 * num_sms, workspace_bytes) : 1;   // 2 CTAs per SM
    int rchunk = ceil_div(ceil_div(R > 0 ? R : 1, splits), BK) * BK;
    splits = R > 0 ? ceil_div(R, rchunk) : 1;
    p.splits = splits; p.rchunk = rchunk;
    p.to_partial = partials_only ? 1 : 0;

    const bool vec2 = aligned16(A) && aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0);
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    const size_t smem = static_cast<size_t>(STAGES) * (nt ? stage_doubles<true>() : stage_doubles<false>()) * sizeof(double);

    auto launch = [&](auto kern) {
        SMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, THREADS, smem, stream>>>(p);
        SMK_LAUNCH_CHECK();
    };
    if (nt) { if (vec2) launch(gemm_skinny_kernel<true, 2>); else launch(gemm_skinny_kernel<true, 1>); }
    else    { if (vec2) launch(gemm_skinny_kernel<false, 2>); else launch(gemm_skinny_kernel<false, 1>); }

    if (partials_only) { *partials_only = splits; return; }     // workspace = [splits][M x N] tiles, ld = M; the caller sums them
    if (splits > 1)
    {
        const long long total = static_cast<long long>(M) * N;
        int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * num_sms));
        reduce_partials_kernel<<<blocks, 256, 0, stream>>>(workspace, splits, M, N, C, ldc, D, ldd);
        SMK_LAUNCH_CHECK();
    }
}

} // namespace smk
