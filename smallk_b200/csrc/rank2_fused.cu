// smallk_b200 — one Rank2 NMF iteration on a sparse matrix in three kernels.
//
// hierclust factors ~200 column subsets of the graph with Solver_Generic_Rank2 (common/include/nmf_solver_rank2.hpp:
// 321-461), some 14 000 outer iterations for 64 leaves at DBLP scale. With k = 2 every operand of an iteration is a
// few megabytes and lives in L2; what the generic kernel sequence (solver.cu: ~20 launches per iteration, k-agnostic
// GEMM / SpMM / reduction kernels) spends its time on is launch latency and one-lane-pair-per-row walks. Here the
// iteration is regrouped around its three sweeps over the data, every Gram matrix and norm riding on the sweep that
// produces its operand:
//
//   rank2_h_kernel     columns j:  H(:,j) <- solve(W'W, (W'A)(:,j));           block partials of H H'
//   rank2_w_kernel     rows i:     (H A')(:,i) by a CSR walk;  W(i,:) <- solve(H H', (H A')(:,i))  (unnormalised);
//                                  block partials of W'W; the last block turns them into the column norms s of W
//                                  (NormalizeAndScale, normalize.hpp:118-161), the normalised W'W and the rescaled
//                                  H H' (nmf_solver_rank2.hpp:420-441)
//   rank2_grad_kernel  columns j:  H(:,j) *= s;  (W'A)(:,j) by a CSC walk over W / s;  gradH = W'W H - W'A
//                      rows i:     W(i,:) /= s;  (H A')(:,i) *= s;  gradW = H H' W - H A'
//                                  block partials of both projected-gradient sums (projected_gradient.hpp:125-171),
//                                  which the last block adds up for ProgressEst::Update
//
// A compressed row / column is walked by 4 lanes (up to 64 entries) or a warp (up to kSpmmSeg) that load a batch of
// entries at a time and add them in storage order (the order of the reference and of the generic SpMM: hierclust's
// priority score ranks the entries of W, so rounding-level reordering shows up in it); those with more than kSpmmSeg
// entries (hubs: the multi-segment list of the segment table; the generic SpMM segments them too, so no order to keep)
// get a whole CTA, entries strided over its threads and a fixed tree. The Gram / norm / PG
// reductions are fixed-shape trees over a grid whose size depends only on the matrix: reproducible run to run. The state left behind (H, Wt, WtW, HHt, WtA, HAt, gradH,
// gradWt) is what the generic sequence leaves.
#include <cfloat>
#include "context.h"
#include "solver.h"
#include "rank2_math.cuh"

namespace smk {

namespace {

constexpr int kThreads = 256;

// sum over the block, returned to thread 0 (fixed shape: xor butterfly inside the warp, warps added in order)
__device__ __forceinline__ double block_sum0(double v, double* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                            // red may still be read from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (kThreads >> 5); ++w) s += red[w];
    return s;
}

// true in exactly one block: the one that arrives last. Called after thread 0 has written this block's partials (they
// are the only data other blocks of this launch read, so thread 0 alone fences). The ticket resets itself.
__device__ __forceinline__ bool last_block(unsigned int* ticket, bool* flag)
{
    if (threadIdx.x == 0)
    {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        *flag = (t == gridDim.x - 1);
        if (*flag) { *ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    return *flag;
}

// three sums over the block at once, returned to thread 0 (same fixed shape as block_sum0, one barrier pair)
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red3)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    __syncthreads();
    if (lane == 0) { red3[warp] = a; red3[8 + warp] = b; red3[16 + warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        a = 0.0; b = 0.0; c = 0.0;
        for (int w = 0; w < (kThreads >> 5); ++w) { a += red3[w]; b += red3[8 + w]; c += red3[16 + w]; }
    }
}

// totals of the per-block partials (ncomp <= 3 per block), by the calling block; valid in thread 0
__device__ __forceinline__ void partial_totals(const double* __restrict__ partial, int nblocks, int ncomp, double& t0, double& t1, double& t2,
                                               double* red3)
{
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads)
    {
        const double* p = partial + static_cast<size_t>(i) * ncomp;
        a += __ldcg(p); b += __ldcg(p + 1);
        if (ncomp > 2) c += __ldcg(p + 2);
    }
    block_sum3(a, b, c, red3);
    t0 = a; t1 = b; t2 = c;
}

// Sum over the stored entries [beg, end) of val * (X(:, idx) .* f), added ONE AFTER THE OTHER in storage order with a
// single accumulator per component — the order of the reference's loops (sparse_gemm_ab_impl.hpp / _ba_impl.hpp) and
// of the generic SpMM — while the loads run kGroup wide: lane g of the group fetches entry o + g (coalesced index /
// value reads two batches ahead, the gather they feed one batch ahead), parks value and operand in the warp's
// shared-memory stage (sv: 32 doubles, sx: 32 double2), and every lane of the group replays the batch in order from
// broadcast reads: one LDS.64 + one LDS.128 + two DFMA per entry (handing the entries round by shuffles cost ~10
// instructions per entry and made the kernels issue-bound). Every lane of the group returns the same sum. The groups
// of a warp run different trip counts: each synchronises only its own lanes.
template <int kGroup>
__device__ __forceinline__ double2 walk_in_order(const unsigned int* __restrict__ idx, const double* __restrict__ val,
                                                 const double* __restrict__ X, unsigned int beg, unsigned int end,
                                                 const double f0, const double f1, double* sv, double2* sx)
{
    const int lane = threadIdx.x & 31, g = lane & (kGroup - 1), gb = lane & ~(kGroup - 1);
    const unsigned int mask = kGroup == 32 ? 0xFFFFFFFFu : (((1u << (kGroup & 31)) - 1u) << gb);
    double a0 = 0.0, a1 = 0.0;
    unsigned int o = beg;
    unsigned int i2 = 0;
    double v1 = 0.0, v2 = 0.0, x10 = 0.0, x11 = 0.0;
    bool ok2 = o + kGroup + g < end;
    if (o + g < end)
    {
        v1 = val[o + g];
        const double2 x = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(idx[o + g]));
        x10 = x.x * f0; x11 = x.y * f1;
    }
    if (ok2) { i2 = idx[o + kGroup + g]; v2 = val[o + kGroup + g]; }
    while (o < end)
    {
        __syncwarp(mask);                       // the previous batch has been read
        sv[lane] = v1; sx[lane] = make_double2(x10, x11);
        __syncwarp(mask);
        v1 = v2; x10 = 0.0; x11 = 0.0;
        if (ok2)
        {
            const double2 x = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(i2));
            x10 = x.x * f0; x11 = x.y * f1;
        }
        o += kGroup;
        ok2 = o + kGroup + g < end;
        i2 = 0; v2 = 0.0;
        if (ok2) { i2 = idx[o + kGroup + g]; v2 = val[o + kGroup + g]; }
        // entries past the end carry v = x = 0: fma(0, 0, a) == a
#pragma unroll
        for (int t = 0; t < kGroup; ++t)
        {
            const double vt = sv[gb + t];
            const double2 xt = sx[gb + t];
            a0 = fma(vt, xt.x, a0);
            a1 = fma(vt, xt.y, a1);
        }
    }
    return make_double2(a0, a1);
}

// A hub (more than kSpmmSeg entries), by the whole CTA. The generic SpMM cuts such a row into segments and adds segment
// sums, so there is no reference order to keep: entries strided over the 256 threads, fixed tree. Valid in thread 0.
__device__ __forceinline__ double2 walk_hub(const unsigned int* __restrict__ idx, const double* __restrict__ val,
                                            const double* __restrict__ X, unsigned int beg, unsigned int end,
                                            const double f0, const double f1, double* red)
{
    double a0 = 0.0, a1 = 0.0;
    unsigned int o = beg + threadIdx.x;
    for (; o + 3 * kThreads < end; o += 4 * kThreads)
    {
        const unsigned int i0 = idx[o], i1 = idx[o + kThreads], i2 = idx[o + 2 * kThreads], i3 = idx[o + 3 * kThreads];
        const double v0 = val[o], v1 = val[o + kThreads], v2 = val[o + 2 * kThreads], v3 = val[o + 3 * kThreads];
        const double2 x0 = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(i0));
        const double2 x1 = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(i1));
        const double2 x2 = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(i2));
        const double2 x3 = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(i3));
        a0 += v0 * (x0.x * f0); a1 += v0 * (x0.y * f1);
        a0 += v1 * (x1.x * f0); a1 += v1 * (x1.y * f1);
        a0 += v2 * (x2.x * f0); a1 += v2 * (x2.y * f1);
        a0 += v3 * (x3.x * f0); a1 += v3 * (x3.y * f1);
    }
    for (; o < end; o += kThreads)
    {
        const double v0 = val[o];
        const double2 x0 = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(idx[o]));
        a0 += v0 * (x0.x * f0); a1 += v0 * (x0.y * f1);
    }
    a0 = block_sum0(a0, red);
    a1 = block_sum0(a1, red);
    return make_double2(a0, a1);
}

// The compressed rows [0, count) that are not hubs, by the warps of the first `blocks` CTAs: a warp takes 32 consecutive
// rows (rows_per_warp: 4..32, fewer when there are few rows, so that the launch still fills the machine), walks the short ones (<= kShortRow entries) eight at a time with 4-lane groups and the others one at a time with
// all 32 lanes, and hands every sum to finish(row, sum) in one lane.
constexpr unsigned int kShortRow = 64;
constexpr int kShortGroup = 4;                  // lanes per short row (mean row length at C4: 12)
constexpr int kRowsPerRound = 32 / kShortGroup;

template <typename Finish>
__device__ __forceinline__ void walk_rows(int count, int blocks, int rows_per_warp, const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ idx,
                                          const double* __restrict__ val, const double* __restrict__ X, const double f0, const double f1,
                                          double* stage_all, Finish&& finish)
{
    constexpr unsigned int kFull = 0xFFFFFFFFu;
    double* sv = stage_all + 96 * (threadIdx.x >> 5);
    double2* sx = reinterpret_cast<double2*>(sv + 32);
    const int lane = threadIdx.x & 31;
    const long long nwarps = static_cast<long long>(blocks) * (kThreads / 32);
    for (long long base = (static_cast<long long>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5)) * rows_per_warp; base < count;
         base += nwarps * rows_per_warp)
    {
        const long long mine = base + lane;
        unsigned int my_beg = 0, my_len = 0;
        if (lane < rows_per_warp && mine < count) { my_beg = ptr[mine]; my_len = ptr[mine + 1] - my_beg; }
        unsigned int wide = __ballot_sync(kFull, my_len > kShortRow && my_len <= static_cast<unsigned int>(kSpmmSeg));
#pragma unroll 1
        for (int r = 0; kRowsPerRound * r < rows_per_warp; ++r)
        {
            const int src = kRowsPerRound * r + lane / kShortGroup;
            const unsigned int beg = __shfl_sync(kFull, my_beg, src), len = __shfl_sync(kFull, my_len, src);
            if (src < rows_per_warp && base + src < count && len <= kShortRow)
            {
                const double2 sum = walk_in_order<kShortGroup>(idx, val, X, beg, beg + len, f0, f1, sv, sx);
                if ((lane & (kShortGroup - 1)) == 0) finish(static_cast<unsigned int>(base + src), sum);
            }
        }
        while (wide)
        {
            const int src = __ffs(wide) - 1;
            wide &= wide - 1;
            const unsigned int beg = __shfl_sync(kFull, my_beg, src), len = __shfl_sync(kFull, my_len, src);
            const double2 sum = walk_in_order<32>(idx, val, X, beg, beg + len, f0, f1, sv, sx);
            if (lane == 0) finish(static_cast<unsigned int>(base + src), sum);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rank2_h_kernel(int n, double* __restrict__ H, const double* __restrict__ WtW, const double* __restrict__ WtA,
               double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ HHt,
               int* __restrict__ status, int outer_iter)
{
    __shared__ double red[kThreads / 32];
    __shared__ double red3[24];
    __shared__ bool is_last;
    Rank2Solve sol;
    sol.init(WtW, false);
    if (sol.fail)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        return;
    }
    double s00 = 0.0, s01 = 0.0, s11 = 0.0;
    for (int j = blockIdx.x * kThreads + threadIdx.x; j < n; j += gridDim.x * kThreads)
    {
        const double2 b = *reinterpret_cast<const double2*>(WtA + 2 * static_cast<size_t>(j));
        const double2 x = sol.apply(b.x, b.y);
        *reinterpret_cast<double2*>(H + 2 * static_cast<size_t>(j)) = x;
        s00 += x.x * x.x; s01 += x.x * x.y; s11 += x.y * x.y;
    }
    block_sum3(s00, s01, s11, red3);
    if (threadIdx.x == 0)
    {
        double* p = partial + 3 * static_cast<size_t>(blockIdx.x);
        __stcg(p, s00); __stcg(p + 1, s01); __stcg(p + 2, s11);
    }
    if (!last_block(ticket, &is_last)) return;
    double t00, t01, t11;
    partial_totals(partial, gridDim.x, 3, t00, t01, t11, red3);
    if (threadIdx.x == 0) { HHt[0] = t00; HHt[1] = t01; HHt[2] = t01; HHt[3] = t11; }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rank2_w_kernel(int m, int light_blocks, int rows_per_warp, const unsigned int* __restrict__ rowptr, const unsigned int* __restrict__ colidx,
               const double* __restrict__ valr, int nheavy, const unsigned int* __restrict__ heavy,
               const double* __restrict__ H, double* __restrict__ HHt, double* __restrict__ HAt, double* __restrict__ Wu,
               double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ norms,
               double* __restrict__ WtW, int* __restrict__ status, int outer_iter)
{
    __shared__ double red[kThreads / 32];
    __shared__ double red3[24];
    __shared__ __align__(16) double stage[96 * (kThreads / 32)];
    __shared__ bool is_last;
    Rank2Solve sol;
    sol.init(HHt, true);
    if (sol.fail)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        return;
    }
    double s00 = 0.0, s01 = 0.0, s11 = 0.0;
    if (static_cast<int>(blockIdx.x) < light_blocks)
    {
        walk_rows(m, light_blocks, rows_per_warp, rowptr, colidx, valr, H, 1.0, 1.0, stage, [&](const unsigned int i, const double2 b) {
            const double2 x = sol.apply(b.x, b.y);
            *reinterpret_cast<double2*>(HAt + 2 * static_cast<size_t>(i)) = b;
            *reinterpret_cast<double2*>(Wu + 2 * static_cast<size_t>(i)) = x;
            s00 += x.x * x.x; s01 += x.x * x.y; s11 += x.y * x.y;
        });
    }
    else
    {
        for (int h = blockIdx.x - light_blocks; h < nheavy; h += gridDim.x - light_blocks)
        {
            const unsigned int i = heavy[h];
            const double2 b = walk_hub(colidx, valr, H, rowptr[i], rowptr[i + 1], 1.0, 1.0, red);
            if (threadIdx.x == 0)
            {
                const double2 x = sol.apply(b.x, b.y);
                *reinterpret_cast<double2*>(HAt + 2 * static_cast<size_t>(i)) = b;
                *reinterpret_cast<double2*>(Wu + 2 * static_cast<size_t>(i)) = x;
                s00 += x.x * x.x; s01 += x.x * x.y; s11 += x.y * x.y;
            }
        }
    }
    block_sum3(s00, s01, s11, red3);
    if (threadIdx.x == 0)
    {
        double* p = partial + 3 * static_cast<size_t>(blockIdx.x);
        __stcg(p, s00); __stcg(p + 1, s01); __stcg(p + 2, s11);
    }
    if (!last_block(ticket, &is_last)) return;
    double t00, t01, t11;
    partial_totals(partial, gridDim.x, 3, t00, t01, t11, red3);
    if (threadIdx.x == 0)
    {
        // NormalizeAndScale: column norms of W (normalize.hpp:118-138; < eps throws, :47-48)
        const double n0 = sqrt(t00), n1 = sqrt(t11);
        if (fabs(n0) < DBL_EPSILON || fabs(n1) < DBL_EPSILON) status[ST_NORM_EPS] = 1;
        norms[0] = n0; norms[1] = n1;
        const double r0 = 1.0 / n0, r1 = 1.0 / n1;
        // W'W of the normalised W
        WtW[0] = t00 * r0 * r0; WtW[1] = t01 * r0 * r1; WtW[2] = t01 * r0 * r1; WtW[3] = t11 * r1 * r1;
        // H H' of the rescaled H (nmf_solver_rank2.hpp:425-434)
        const double e00 = HHt[0], e01 = HHt[2], e11 = HHt[3];
        HHt[0] = e00 * n0 * n0; HHt[2] = e01 * n0 * n1; HHt[1] = e01 * n0 * n1; HHt[3] = e11 * n1 * n1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rank2_grad_kernel(int n, int m, int col_blocks, int heavy_blocks, int rows_per_warp, const unsigned int* __restrict__ colptr,
                  const unsigned int* __restrict__ rowidx, const double* __restrict__ val, int nheavy,
                  const unsigned int* __restrict__ heavy, const double* __restrict__ Wu, const double* __restrict__ norms,
                  const double* __restrict__ WtW, const double* __restrict__ HHt, double* __restrict__ H,
                  double* __restrict__ WtA, double* __restrict__ gradH, double* __restrict__ Wt, double* __restrict__ HAt,
                  double* __restrict__ gradWt, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                  double* __restrict__ acc)
{
    __shared__ double red[kThreads / 32];
    __shared__ double red3[24];
    __shared__ __align__(16) double stage[96 * (kThreads / 32)];
    __shared__ bool is_last;
    const double n0 = norms[0], n1 = norms[1];
    const double r0 = 1.0 / n0, r1 = 1.0 / n1;
    double pg_w = 0.0, pg_h = 0.0;
    const int b = blockIdx.x;
    if (b < col_blocks + heavy_blocks)
    {
        const double g00 = WtW[0], g10 = WtW[1], g01 = WtW[2], g11 = WtW[3];
        auto finish_col = [&](const unsigned int j, const double a0, const double a1) {
            const double2 h = *reinterpret_cast<const double2*>(H + 2 * static_cast<size_t>(j));
            const double h0 = h.x * n0, h1 = h.y * n1;
            const double q0 = (g00 * h0 + g01 * h1) - a0, q1 = (g10 * h0 + g11 * h1) - a1;
            *reinterpret_cast<double2*>(H + 2 * static_cast<size_t>(j)) = make_double2(h0, h1);
            *reinterpret_cast<double2*>(WtA + 2 * static_cast<size_t>(j)) = make_double2(a0, a1);
            *reinterpret_cast<double2*>(gradH + 2 * static_cast<size_t>(j)) = make_double2(q0, q1);
            if (q0 < 0.0 || h0 > 0.0) pg_h += q0 * q0;
            if (q1 < 0.0 || h1 > 0.0) pg_h += q1 * q1;
        };
        if (b < col_blocks)
        {
            walk_rows(n, col_blocks, rows_per_warp, colptr, rowidx, val, Wu, r0, r1, stage, [&](const unsigned int j, const double2 a) { finish_col(j, a.x, a.y); });
        }
        else
        {
            for (int h = b - col_blocks; h < nheavy; h += heavy_blocks)
            {
                const unsigned int j = heavy[h];
                const double2 a = walk_hub(rowidx, val, Wu, colptr[j], colptr[j + 1], r0, r1, red);
                if (threadIdx.x == 0) finish_col(j, a.x, a.y);
            }
        }
    }
    else
    {
        const double e00 = HHt[0], e10 = HHt[1], e01 = HHt[2], e11 = HHt[3];
        const int rb = b - col_blocks - heavy_blocks, row_blocks = gridDim.x - col_blocks - heavy_blocks;
        for (int i = rb * kThreads + threadIdx.x; i < m; i += row_blocks * kThreads)
        {
            const double2 wu = *reinterpret_cast<const double2*>(Wu + 2 * static_cast<size_t>(i));
            const double2 ha = *reinterpret_cast<const double2*>(HAt + 2 * static_cast<size_t>(i));
            const double w0 = wu.x * r0, w1 = wu.y * r1;
            const double a0 = ha.x * n0, a1 = ha.y * n1;
            const double q0 = (e00 * w0 + e01 * w1) - a0, q1 = (e10 * w0 + e11 * w1) - a1;
            *reinterpret_cast<double2*>(Wt + 2 * static_cast<size_t>(i)) = make_double2(w0, w1);
            *reinterpret_cast<double2*>(HAt + 2 * static_cast<size_t>(i)) = make_double2(a0, a1);
            *reinterpret_cast<double2*>(gradWt + 2 * static_cast<size_t>(i)) = make_double2(q0, q1);
            if (q0 < 0.0 || w0 > 0.0) pg_w += q0 * q0;
            if (q1 < 0.0 || w1 > 0.0) pg_w += q1 * q1;
        }
    }
    double unused = 0.0;
    block_sum3(pg_w, pg_h, unused, red3);
    if (threadIdx.x == 0)
    {
        double* p = partial + 2 * static_cast<size_t>(blockIdx.x);
        __stcg(p, pg_w); __stcg(p + 1, pg_h);
    }
    if (!last_block(ticket, &is_last)) return;
    double tw, th;
    partial_totals(partial, gridDim.x, 2, tw, th, unused, red3);
    if (threadIdx.x == 0) { acc[0] = tw; acc[1] = th; }
}

// rows a warp takes per trip: as few as keeps every warp of `cap` CTAs busy once (short dependent chains), at most 32
int rows_per_warp_for(int count, int cap)
{
    int rpw = kRowsPerRound;
    while (rpw < 32 && static_cast<long long>(cap) * (kThreads / 32) * rpw < count) rpw *= 2;
    return rpw;
}

} // namespace

// One outer iteration of Solver_Generic_Rank2::operator() (nmf_solver_rank2.hpp:353-455) on the active sparse matrix,
// plus both projected-gradient sums (left in c->acc[0], c->acc[1]). Wu: 2 * m doubles of scratch.
void rank2_fused_step(smk_ctx* c, double* Wu)
{
    const SparseDev& S = *c->Sa;
    const int m = c->m, n = c->n, sms = c->num_sms;
    const int cap = 4 * sms;
    const int hb = std::max(1, std::min(ceil_div(n, kThreads), cap));
    rank2_h_kernel<<<hb, kThreads, 0, c->stream>>>(n, c->H.p, c->WtW.p, c->WtA.p, c->partial.p, c->ticket.p, c->HHt.p,
                                                   c->status.p, c->steps_done);
    SMK_LAUNCH_CHECK();
    {
        const int rpw = rows_per_warp_for(m, cap);
        const int light = std::max(1, std::min(ceil_div(m, rpw * (kThreads / 32)), cap));
        const int heavy = S.seg_rows.nmulti > 0 ? std::min(S.seg_rows.nmulti, sms) : 0;
        rank2_w_kernel<<<light + heavy, kThreads, 0, c->stream>>>(m, light, rpw, S.rowptr.p, S.colidx.p, S.valr.p, S.seg_rows.nmulti,
                                                                  S.seg_rows.multi_col.p, c->H.p, c->HHt.p, c->HAt.p, Wu, c->partial.p,
                                                                  c->ticket.p + 1, c->norms.p, c->WtW.p, c->status.p, c->steps_done);
        SMK_LAUNCH_CHECK();
    }
    {
        const int rpw = rows_per_warp_for(n, cap);
        const int cols = std::max(1, std::min(ceil_div(n, rpw * (kThreads / 32)), cap));
        const int heavy = S.seg_cols.nmulti > 0 ? std::min(S.seg_cols.nmulti, sms) : 0;
        const int rows = std::max(1, std::min(ceil_div(m, kThreads), 2 * sms));
        rank2_grad_kernel<<<cols + heavy + rows, kThreads, 0, c->stream>>>(n, m, cols, heavy, rpw, S.colptr.p, S.rowidx.p, S.val.p,
                                                                          S.seg_cols.nmulti, S.seg_cols.multi_col.p, Wu, c->norms.p,
                                                                          c->WtW.p, c->HHt.p, c->H.p, c->WtA.p, c->gradH.p, c->Wt.p,
                                                                          c->HAt.p, c->gradWt.p, c->partial.p, c->ticket.p + 2, c->acc.p);
        SMK_LAUNCH_CHECK();
    }
}

} // namespace smk
