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
// A compressed row / column is walked by 8 lanes that load 8 entries at a time and add them in storage order (the
// order of the reference and of the generic SpMM: hierclust's priority score ranks the entries of W, so rounding-level
// reordering shows up in it); those with more than kSpmmSeg entries (hubs: the multi-segment list of the segment
// table) get a whole CTA, one group per kSpmmSeg-entry segment, segment sums added in order. The Gram / norm / PG
// reductions are fixed-shape trees over a grid whose size depends only on the matrix: reproducible run to run. The state left behind (H, Wt, WtW, HHt, WtA, HAt, gradH,
// gradWt) is what the generic sequence leaves.
#include <cfloat>
#include "context.h"
#include "solver.h"
#include "rank2_math.cuh"

namespace smk {

namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 8;                       // lanes per compressed row / column
constexpr int kGroupsPerBlock = kThreads / kGroup;

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

// true in exactly one block: the one that arrives last. The ticket resets itself for the next launch.
__device__ __forceinline__ bool last_block(unsigned int* ticket, bool* flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned int t = atomicAdd(ticket, 1u);
        *flag = (t == gridDim.x - 1);
        if (*flag) *ticket = 0u;
    }
    __syncthreads();
    if (*flag) __threadfence();
    return *flag;
}

// total of column c of the per-block partials (ncomp per block), by the calling block; valid in thread 0
__device__ __forceinline__ double partial_total(const double* __restrict__ partial, int nblocks, int ncomp, int c, double* red)
{
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += kThreads) s += __ldcg(partial + static_cast<size_t>(b) * ncomp + c);
    return block_sum0(s, red);
}

// Sum over the stored entries [beg, end) of val * (X(:, idx) .* f), added ONE AFTER THE OTHER in storage order with a
// single accumulator per component — the order of the reference's loops (sparse_gemm_ab_impl.hpp / _ba_impl.hpp) and
// of the generic SpMM — while the loads run 8 wide: lane g of the group fetches entry o + g (coalesced index / value
// reads, one gather each, the next batch already in flight), then all lanes replay the batch in order from shuffles.
// Every lane of the group returns the same sum. The groups of a warp run different trip counts: each names only its
// own lanes in the shuffles.
__device__ __forceinline__ double2 walk_in_order(const unsigned int* __restrict__ idx, const double* __restrict__ val,
                                                 const double* __restrict__ X, unsigned int beg, unsigned int end,
                                                 const double f0, const double f1)
{
    const int lane = threadIdx.x & 31, g = lane & (kGroup - 1);
    const unsigned int mask = 0xFFu << (lane & ~(kGroup - 1));
    double a0 = 0.0, a1 = 0.0;
    double nv = 0.0, nx0 = 0.0, nx1 = 0.0;
    unsigned int o = beg;
    if (o + g < end)
    {
        const double2 x = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(idx[o + g]));
        nv = val[o + g]; nx0 = x.x * f0; nx1 = x.y * f1;
    }
    while (o < end)
    {
        const double v = nv, x0 = nx0, x1 = nx1;
        o += kGroup;
        nv = 0.0; nx0 = 0.0; nx1 = 0.0;
        if (o + g < end)
        {
            const double2 x = *reinterpret_cast<const double2*>(X + 2 * static_cast<size_t>(idx[o + g]));
            nv = val[o + g]; nx0 = x.x * f0; nx1 = x.y * f1;
        }
        // entries past the end carry v = x = 0: fma(0, 0, a) == a
#pragma unroll
        for (int t = 0; t < kGroup; ++t)
        {
            const double vt = __shfl_sync(mask, v, t, kGroup);
            a0 = fma(vt, __shfl_sync(mask, x0, t, kGroup), a0);
            a1 = fma(vt, __shfl_sync(mask, x1, t, kGroup), a1);
        }
    }
    return make_double2(a0, a1);
}

// A hub (more than kSpmmSeg entries), by the whole CTA: segments of kSpmmSeg entries as in the generic SpMM, one group
// per segment, each summed in order, then the segment sums added in segment order. Valid in thread 0.
__device__ __forceinline__ double2 walk_hub(const unsigned int* __restrict__ idx, const double* __restrict__ val,
                                            const double* __restrict__ X, unsigned int beg, unsigned int end,
                                            const double f0, const double f1, double2* part)
{
    const int group = threadIdx.x / kGroup;
    const unsigned int nseg = (end - beg + kSpmmSeg - 1) / kSpmmSeg;
    double a0 = 0.0, a1 = 0.0;
    for (unsigned int base = 0; base < nseg; base += kGroupsPerBlock)
    {
        const unsigned int sg = base + group;
        if (sg < nseg)
        {
            const unsigned int b = beg + sg * kSpmmSeg;
            const double2 p = walk_in_order(idx, val, X, b, min(end, b + kSpmmSeg), f0, f1);
            if ((threadIdx.x & (kGroup - 1)) == 0) part[group] = p;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned int cnt = min(static_cast<unsigned int>(kGroupsPerBlock), nseg - base);
            for (unsigned int q = 0; q < cnt; ++q) { a0 += part[q].x; a1 += part[q].y; }
        }
        __syncthreads();
    }
    return make_double2(a0, a1);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rank2_h_kernel(int n, double* __restrict__ H, const double* __restrict__ WtW, const double* __restrict__ WtA,
               double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ HHt,
               int* __restrict__ status, int outer_iter)
{
    __shared__ double red[kThreads / 32];
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
    s00 = block_sum0(s00, red); s01 = block_sum0(s01, red); s11 = block_sum0(s11, red);
    if (threadIdx.x == 0)
    {
        double* p = partial + 3 * static_cast<size_t>(blockIdx.x);
        __stcg(p, s00); __stcg(p + 1, s01); __stcg(p + 2, s11);
    }
    if (!last_block(ticket, &is_last)) return;
    const double t00 = partial_total(partial, gridDim.x, 3, 0, red);
    const double t01 = partial_total(partial, gridDim.x, 3, 1, red);
    const double t11 = partial_total(partial, gridDim.x, 3, 2, red);
    if (threadIdx.x == 0) { HHt[0] = t00; HHt[1] = t01; HHt[2] = t01; HHt[3] = t11; }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rank2_w_kernel(int m, int light_blocks, const unsigned int* __restrict__ rowptr, const unsigned int* __restrict__ colidx,
               const double* __restrict__ valr, int nheavy, const unsigned int* __restrict__ heavy,
               const double* __restrict__ H, double* __restrict__ HHt, double* __restrict__ HAt, double* __restrict__ Wu,
               double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ norms,
               double* __restrict__ WtW, int* __restrict__ status, int outer_iter)
{
    __shared__ double red[kThreads / 32];
    __shared__ double2 part[kGroupsPerBlock];
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
        const int g = threadIdx.x & (kGroup - 1);
        for (int i = blockIdx.x * kGroupsPerBlock + threadIdx.x / kGroup; i < m; i += light_blocks * kGroupsPerBlock)
        {
            const unsigned int beg = rowptr[i], end = rowptr[i + 1];
            if (end - beg > static_cast<unsigned int>(kSpmmSeg)) continue;         // a hub: a whole CTA walks it below
            const double2 b = walk_in_order(colidx, valr, H, beg, end, 1.0, 1.0);
            if (g == 0)
            {
                const double2 x = sol.apply(b.x, b.y);
                *reinterpret_cast<double2*>(HAt + 2 * static_cast<size_t>(i)) = b;
                *reinterpret_cast<double2*>(Wu + 2 * static_cast<size_t>(i)) = x;
                s00 += x.x * x.x; s01 += x.x * x.y; s11 += x.y * x.y;
            }
        }
    }
    else
    {
        for (int h = blockIdx.x - light_blocks; h < nheavy; h += gridDim.x - light_blocks)
        {
            const unsigned int i = heavy[h];
            const double2 b = walk_hub(colidx, valr, H, rowptr[i], rowptr[i + 1], 1.0, 1.0, part);
            if (threadIdx.x == 0)
            {
                const double2 x = sol.apply(b.x, b.y);
                *reinterpret_cast<double2*>(HAt + 2 * static_cast<size_t>(i)) = b;
                *reinterpret_cast<double2*>(Wu + 2 * static_cast<size_t>(i)) = x;
                s00 += x.x * x.x; s01 += x.x * x.y; s11 += x.y * x.y;
            }
        }
    }
    s00 = block_sum0(s00, red); s01 = block_sum0(s01, red); s11 = block_sum0(s11, red);
    if (threadIdx.x == 0)
    {
        double* p = partial + 3 * static_cast<size_t>(blockIdx.x);
        __stcg(p, s00); __stcg(p + 1, s01); __stcg(p + 2, s11);
    }
    if (!last_block(ticket, &is_last)) return;
    const double t00 = partial_total(partial, gridDim.x, 3, 0, red);
    const double t01 = partial_total(partial, gridDim.x, 3, 1, red);
    const double t11 = partial_total(partial, gridDim.x, 3, 2, red);
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
rank2_grad_kernel(int n, int m, int col_blocks, int heavy_blocks, const unsigned int* __restrict__ colptr,
                  const unsigned int* __restrict__ rowidx, const double* __restrict__ val, int nheavy,
                  const unsigned int* __restrict__ heavy, const double* __restrict__ Wu, const double* __restrict__ norms,
                  const double* __restrict__ WtW, const double* __restrict__ HHt, double* __restrict__ H,
                  double* __restrict__ WtA, double* __restrict__ gradH, double* __restrict__ Wt, double* __restrict__ HAt,
                  double* __restrict__ gradWt, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                  double* __restrict__ acc)
{
    __shared__ double red[kThreads / 32];
    __shared__ double2 part[kGroupsPerBlock];
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
            const int g = threadIdx.x & (kGroup - 1);
            for (int j = b * kGroupsPerBlock + threadIdx.x / kGroup; j < n; j += col_blocks * kGroupsPerBlock)
            {
                const unsigned int beg = colptr[j], end = colptr[j + 1];
                if (end - beg > static_cast<unsigned int>(kSpmmSeg)) continue;
                const double2 a = walk_in_order(rowidx, val, Wu, beg, end, r0, r1);
                if (g == 0) finish_col(j, a.x, a.y);
            }
        }
        else
        {
            for (int h = b - col_blocks; h < nheavy; h += heavy_blocks)
            {
                const unsigned int j = heavy[h];
                const double2 a = walk_hub(rowidx, val, Wu, colptr[j], colptr[j + 1], r0, r1, part);
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
    pg_w = block_sum0(pg_w, red); pg_h = block_sum0(pg_h, red);
    if (threadIdx.x == 0)
    {
        double* p = partial + 2 * static_cast<size_t>(blockIdx.x);
        __stcg(p, pg_w); __stcg(p + 1, pg_h);
    }
    if (!last_block(ticket, &is_last)) return;
    const double tw = partial_total(partial, gridDim.x, 2, 0, red);
    const double th = partial_total(partial, gridDim.x, 2, 1, red);
    if (threadIdx.x == 0) { acc[0] = tw; acc[1] = th; }
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
        const int light = std::max(1, std::min(ceil_div(m, kGroupsPerBlock), cap));
        const int heavy = S.seg_rows.nmulti > 0 ? std::min(S.seg_rows.nmulti, sms) : 0;
        rank2_w_kernel<<<light + heavy, kThreads, 0, c->stream>>>(m, light, S.rowptr.p, S.colidx.p, S.valr.p, S.seg_rows.nmulti,
                                                                  S.seg_rows.multi_col.p, c->H.p, c->HHt.p, c->HAt.p, Wu, c->partial.p,
                                                                  c->ticket.p + 1, c->norms.p, c->WtW.p, c->status.p, c->steps_done);
        SMK_LAUNCH_CHECK();
    }
    {
        const int cols = std::max(1, std::min(ceil_div(n, kGroupsPerBlock), cap));
        const int heavy = S.seg_cols.nmulti > 0 ? std::min(S.seg_cols.nmulti, sms) : 0;
        const int rows = std::max(1, std::min(ceil_div(m, kThreads), 2 * sms));
        rank2_grad_kernel<<<cols + heavy + rows, kThreads, 0, c->stream>>>(n, m, cols, heavy, S.colptr.p, S.rowidx.p, S.val.p,
                                                                          S.seg_cols.nmulti, S.seg_cols.multi_col.p, Wu, c->norms.p,
                                                                          c->WtW.p, c->HHt.p, c->H.p, c->WtA.p, c->gradH.p, c->Wt.p,
                                                                          c->HAt.p, c->gradWt.p, c->partial.p, c->ticket.p + 2, c->acc.p);
        SMK_LAUNCH_CHECK();
    }
}

} // namespace smk
