// smallk_b200 — fused factor-update kernels: HALS sweeps, rank-2 solve + active set,
// NormalizeAndScale. All operate on "k x big" column-major factors (H, or Wt = W').
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace smk {

namespace {

constexpr int kMaxKPL = 8;          // rows per lane: k <= 256
constexpr int kSweepBlocks = 2048;  // upper bound on the grid of the per-row sweep kernels

__device__ __forceinline__ double block_sum_f(double v, double* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (warp == 0) v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
    __syncthreads();
    v = red[0];
    __syncthreads();
    return v;
}

// ---------------------------------------------------------------------------
// HALS, H side (no cross-column coupling): one warp owns a column for the whole
// sweep; the column sits in registers (k/32 entries per lane), row r of G is read
// through L1 (every warp reads the same k x k matrix).
// ---------------------------------------------------------------------------
template <int KPL>
__global__ void hals_sweep_cols_kernel(int k, int q, double* __restrict__ X, const double* __restrict__ G,
                                       const double* __restrict__ R)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long j = blockIdx.x * static_cast<long long>(warps_per_block) + (threadIdx.x >> 5); j < q;
         j += static_cast<long long>(gridDim.x) * warps_per_block)
    {
        double* xcol = X + j * k;
        const double* rcol = R + j * k;
        double x[KPL];
#pragma unroll
        for (int e = 0; e < KPL; ++e) { const int p = lane + 32 * e; x[e] = (p < k) ? xcol[p] : 0.0; }
        for (int r = 0; r < k; ++r)
        {
            const double* gcol = G + static_cast<long long>(r) * k;      // G is symmetric: row r is read as column r (coalesced)
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < KPL; ++e) { const int p = lane + 32 * e; if (p < k) s += __ldg(gcol + p) * x[e]; }
            s = warp_sum(s);
            const double grr = __ldg(G + r + static_cast<long long>(r) * k);
            const double rr = __ldg(rcol + r);
            const int re = r >> 5, rl = r & 31;
            double xr = 0.0;
#pragma unroll
            for (int e = 0; e < KPL; ++e) if (e == re) xr = x[e];
            xr = __shfl_sync(0xffffffffu, xr, rl);
            double h = xr + (rr - s) / grr;
            if (isnan(h) || h < 0.0) h = 0.0;
#pragma unroll
            for (int e = 0; e < KPL; ++e) if (e == re && lane == rl) x[e] = h;
        }
#pragma unroll
        for (int e = 0; e < KPL; ++e) { const int p = lane + 32 * e; if (p < k) xcol[p] = x[e]; }
    }
}

// ---------------------------------------------------------------------------
// HALS, W side: the unit-norm scaling of column r of W couples all rows of W after
// every step, so the sweep is k dependent passes. Pass r
//   (1) finishes pass r-1: norm from the block partials, entry r-1 of every column scaled
//       (or set to eps/norm when the whole column was clamped to zero),
//   (2) updates entry r of every column and writes this block's partial sum of squares
//       and zero count.
// partial layout: [parity][0..kSweepBlocks) sums, [parity][kSweepBlocks..2*kSweepBlocks) zero counts.
// ---------------------------------------------------------------------------
template <int KPL>
__global__ void hals_sweep_row_kernel(int k, int q, int r, double* __restrict__ X, const double* __restrict__ G,
                                      const double* __restrict__ R, double* __restrict__ partial, int nblocks_prev,
                                      double* __restrict__ norms)
{
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;

    // (1) finish row r-1
    double inv_prev = 1.0;
    bool fill_prev = false;
    if (r > 0)
    {
        const double* pp = partial + ((r - 1) & 1) * 2 * kSweepBlocks;
        double s = 0.0, z = 0.0;
        for (int i = threadIdx.x; i < nblocks_prev; i += blockDim.x) { s += pp[i]; z += pp[kSweepBlocks + i]; }
        s = block_sum_f(s, red);
        z = block_sum_f(z, red);
        double norm;
        if (z == static_cast<double>(q)) { fill_prev = true; norm = DBL_EPSILON * sqrt(static_cast<double>(q)); }
        else norm = sqrt(s);
        inv_prev = 1.0 / norm;
        if (blockIdx.x == 0 && threadIdx.x == 0) norms[r - 1] = norm;
    }
    if (r == k)
    {
        // final pass: only the scaling of row k-1 remains
        for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < q;
             j += static_cast<long long>(gridDim.x) * blockDim.x)
        {
            double* px = X + j * k + (k - 1);
            *px = (fill_prev ? DBL_EPSILON : *px) * inv_prev;
        }
        return;
    }

    double sumsq = 0.0, zeros = 0.0;
    const double grr = __ldg(G + r + static_cast<long long>(r) * k);
    for (long long j = blockIdx.x * static_cast<long long>(warps_per_block) + (threadIdx.x >> 5); j < q;
         j += static_cast<long long>(gridDim.x) * warps_per_block)
    {
        double* xcol = X + j * k;
        double s = 0.0, xr = 0.0;
#pragma unroll
        for (int e = 0; e < KPL; ++e)
        {
            const int p = lane + 32 * e;
            if (p < k)
            {
                double xv = xcol[p];
                if (p == r - 1) { xv = (fill_prev ? DBL_EPSILON : xv) * inv_prev; xcol[p] = xv; }
                if (p == r) xr = xv;
                s += __ldg(G + static_cast<long long>(r) * k + p) * xv;
            }
        }
        s = warp_sum(s);
        xr = __shfl_sync(0xffffffffu, xr, r & 31);   // only the owning lane had it; others hold 0 -> take owner's
        if (lane == 0)
        {
            double w = xr + (__ldg(R + j * k + r) - s) / grr;
            if (isnan(w) || w < 0.0) { w = 0.0; zeros += 1.0; }
            xcol[r] = w;
            sumsq += w * w;
        }
    }
    sumsq = block_sum_f(sumsq, red);
    zeros = block_sum_f(zeros, red);
    if (threadIdx.x == 0)
    {
        double* pc = partial + (r & 1) * 2 * kSweepBlocks;
        pc[blockIdx.x] = sumsq;
        pc[kSweepBlocks + blockIdx.x] = zeros;
    }
}

// ---------------------------------------------------------------------------
// HALS, W side, blocked: the same k dependent steps, but the k-long dot products are not re-read from memory at
// every step. Columns are taken in blocks of kHalsB. For a block [c0, c0+B):
//   phase A (hals_block_outer_kernel): Q(l, j) = sum over p OUTSIDE the block of X(p,j) G(p,c0+l) - R(c0+l,j)
//            — one pass over X per block (k/B passes per sweep instead of k), a 16 x k x q GEMM on the FP64 tensor pipe;
//   phase B (hals_block_step_kernel), B dependent steps: step l needs only the block's own B entries of each column
//            and Q(l, j):  x <- max(0, x - (Q + sum_{p in block} X(p,j) G(p,c)) / G(c,c)),  NaN -> 0,
//            then the grid-wide unit-norm scaling of row c is applied by the NEXT kernel (as in hals_sweep_row_kernel).
// Traffic per sweep and column of X: k*k/B + k*(B+3) doubles instead of k*k, all of it streamed (the B steps of a
// block run on a compact q x B copy of the block, not on 128-byte pieces of 8k-byte-strided columns).
// ---------------------------------------------------------------------------
constexpr int kHalsB = 16;

// norm of row `prev` from the block partials of the step that produced it; returns 1/norm and whether the row is refilled with eps
__device__ __forceinline__ double finish_prev_row(const double* __restrict__ partial, int prev, int nblocks_prev, int q,
                                                  double* __restrict__ norms, double* red, bool& fill_prev)
{
    const double* pp = partial + (prev & 1) * 2 * kSweepBlocks;
    double s = 0.0, z = 0.0;
    for (int i = threadIdx.x; i < nblocks_prev; i += blockDim.x) { s += pp[i]; z += pp[kSweepBlocks + i]; }
    s = block_sum_f(s, red);
    z = block_sum_f(z, red);
    double norm;
    fill_prev = (z == static_cast<double>(q));
    if (fill_prev) norm = DBL_EPSILON * sqrt(static_cast<double>(q));
    else norm = sqrt(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) norms[prev] = norm;
    return 1.0 / norm;
}

// Phase A on the FP64 tensor pipe: D (16 block columns x 8 columns of X) = GmT (16 x k) * X (k x 8), two DMMA.8x8x4 per
// four rows of X. One warp per 8 columns of X; the B fragments come straight from global memory (each lane reads
// X(4s + (lane & 3), j0 + (lane >> 2)): eight fully used 32-byte sectors per warp load, no shared-memory tile, no
// barriers), the A fragments of the masked Gram block sit in shared memory in fragment order.
// The block's own entries are copied to a COMPACT buffer Xb (q x 16) on which the B step kernels then stream at full
// bandwidth; the previous block's finished entries are taken from its compact buffer and written back to X here.
__global__ void __launch_bounds__(256)
hals_block_outer_kernel(int k, int q, int c0, double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ R,
                        double* __restrict__ Q, double* __restrict__ Xb_cur, const double* __restrict__ Xb_prev,
                        const double* __restrict__ partial, int nblocks_prev, double* __restrict__ norms)
{
    extern __shared__ __align__(16) double sA[];       // [ksteps][2 m-tiles][32 lanes]
    __shared__ double red[32];
    const int ksteps = (k + 3) >> 2;
    const int nb = min(kHalsB, k - c0);
    bool fill_prev = false;
    double inv_prev = 1.0;
    if (c0 > 0) inv_prev = finish_prev_row(partial, c0 - 1, nblocks_prev, q, norms, red, fill_prev);
    for (int e = threadIdx.x; e < ksteps * 64; e += blockDim.x)
    {
        const int ln = e & 31, mt = (e >> 5) & 1, st = e >> 6;
        const int row = mt * 8 + (ln >> 2), p = 4 * st + (ln & 3);
        const bool inside = (p >= c0 && p < c0 + nb);
        sA[e] = (row < nb && p < k && !inside) ? G[static_cast<long long>(c0 + row) * k + p] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n = lane >> 2, kk = lane & 3;
    constexpr int BS = kHalsB / 4;                      // k-steps per block
    const int st_cur = c0 >> 2;                         // first k-step of this block
    const int st_prev = st_cur - BS;                    // first k-step of the previous block (c0 > 0)
    for (long long j0 = (static_cast<long long>(blockIdx.x) * wpb + warp) * 8; j0 < q; j0 += static_cast<long long>(gridDim.x) * wpb * 8)
    {
        const long long j = j0 + n;
        const bool livecol = j < q;
        const long long jc = livecol ? j : static_cast<long long>(q) - 1;
        const double* xcol = X + jc * k;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        double wb[BS], cur[BS];
        // rows before the previous block, and rows after this block: read from X
        auto plain = [&](int st) {
            const int p = 4 * st + kk;
            double b = xcol[min(p, k - 1)];
            if (p >= k || !livecol) b = 0.0;
            const double a0 = sA[(st * 2) * 32 + lane], a1 = sA[(st * 2 + 1) * 32 + lane];
            dmma884(c00, c01, a0, b);
            dmma884(c10, c11, a1, b);
        };
#pragma unroll 8
        for (int st = 0; st < max(st_prev, 0); ++st) plain(st);
        if (c0 > 0)
        {
            // the previous block: finished values from its compact buffer (row c0-1 still needs its scaling)
#pragma unroll
            for (int u = 0; u < BS; ++u)
            {
                const int st = st_prev + u, t = 4 * u + kk;
                double b = Xb_prev[jc * kHalsB + t];
                if (t == kHalsB - 1) b = (fill_prev ? DBL_EPSILON : b) * inv_prev;
                wb[u] = b;
                if (!livecol) b = 0.0;
                const double a0 = sA[(st * 2) * 32 + lane], a1 = sA[(st * 2 + 1) * 32 + lane];
                dmma884(c00, c01, a0, b);
                dmma884(c10, c11, a1, b);
            }
        }
        // this block: no contribution (masked), entries go to the compact buffer
#pragma unroll
        for (int u = 0; u < BS; ++u)
        {
            const int p = c0 + 4 * u + kk;
            cur[u] = xcol[min(p, k - 1)];
        }
#pragma unroll 8
        for (int st = st_cur + BS; st < ksteps; ++st) plain(st);

        const int row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
        {
            const long long jj = j0 + col + cc;
            if (jj < q)
            {
                const double* rj = R + jj * k + c0;
                if (row < nb) Q[static_cast<long long>(row) * q + jj] = (cc ? c01 : c00) - rj[row];
                if (row + 8 < nb) Q[static_cast<long long>(row + 8) * q + jj] = (cc ? c11 : c10) - rj[row + 8];
            }
        }
        if (livecol)
        {
#pragma unroll
            for (int u = 0; u < BS; ++u)
            {
                const int t = 4 * u + kk;
                if (c0 > 0) X[j * k + (c0 - kHalsB) + t] = wb[u];
                Xb_cur[j * kHalsB + t] = (t < nb) ? cur[u] : 0.0;
            }
        }
    }
}

// One step of phase B on the compact block buffer Xb (q x 16): half a warp per column of X.
__global__ void __launch_bounds__(256)
hals_block_step_kernel(int k, int q, int c0, int l, double* __restrict__ Xb, const double* __restrict__ G,
                       const double* __restrict__ Q, double* __restrict__ partial, int nblocks_prev, double* __restrict__ norms)
{
    __shared__ double red[32];
    __shared__ double sg[kHalsB];
    const int nb = min(kHalsB, k - c0);
    const int c = c0 + l;
    bool fill_prev = false;
    double inv_prev = 1.0;
    if (l > 0) inv_prev = finish_prev_row(partial, c - 1, nblocks_prev, q, norms, red, fill_prev);
    if (threadIdx.x < kHalsB) sg[threadIdx.x] = (threadIdx.x < nb) ? G[static_cast<long long>(c) * k + c0 + threadIdx.x] : 0.0;
    __syncthreads();
    const double gcc = sg[l];
    const double* ql = Q + static_cast<long long>(l) * q;
    double sumsq = 0.0, zeros = 0.0;
    const int t = threadIdx.x & (kHalsB - 1);
    const long long halves = (static_cast<long long>(gridDim.x) * blockDim.x) / kHalsB;
    const long long h = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) / kHalsB;
    constexpr int U = 4;                                 // columns in flight per half-warp
    for (long long base = 0; base < q; base += U * halves)
    {
        double v[U], dot[U], qv[U];
        bool live[U];
        // branch-free loads from clamped addresses: all 2U loads of a trip are in flight together
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long j = base + u * halves + h;
            live[u] = j < q;
            const long long jc = live[u] ? j : q - 1;
            v[u] = Xb[jc * kHalsB + t];
            qv[u] = ql[jc];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (!live[u]) v[u] = 0.0;
            if (l > 0 && t == l - 1 && live[u])
            {
                v[u] = (fill_prev ? DBL_EPSILON : v[u]) * inv_prev;
                Xb[(base + u * halves + h) * kHalsB + t] = v[u];
            }
            dot[u] = v[u] * sg[t];
        }
#pragma unroll
        for (int o = kHalsB / 2; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < U; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (live[u] && t == l)
            {
                const long long j = base + u * halves + h;
                double w = v[u] - (qv[u] + dot[u]) / gcc;
                if (isnan(w) || w < 0.0) { w = 0.0; zeros += 1.0; }
                Xb[j * kHalsB + l] = w;
                sumsq += w * w;
            }
        }
    }
    sumsq = block_sum_f(sumsq, red);
    zeros = block_sum_f(zeros, red);
    if (threadIdx.x == 0)
    {
        double* pc = partial + (c & 1) * 2 * kSweepBlocks;
        pc[blockIdx.x] = sumsq;
        pc[kSweepBlocks + blockIdx.x] = zeros;
    }
}

// End of the sweep: the last block goes back to X, its last row (row k-1) scaled to unit norm.
__global__ void __launch_bounds__(256)
hals_block_writeback_kernel(int k, int q, int c0, double* __restrict__ X, const double* __restrict__ Xb,
                            const double* __restrict__ partial, int nblocks_prev, double* __restrict__ norms)
{
    __shared__ double red[32];
    const int nb = k - c0;
    bool fill_prev = false;
    const double inv_prev = finish_prev_row(partial, k - 1, nblocks_prev, q, norms, red, fill_prev);
    const long long total = static_cast<long long>(q) * kHalsB;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const long long j = e / kHalsB;
        const int t = static_cast<int>(e % kHalsB);
        if (t < nb)
        {
            double v = Xb[e];
            if (t == nb - 1) v = (fill_prev ? DBL_EPSILON : v) * inv_prev;
            X[j * k + c0 + t] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// rank 2
// ---------------------------------------------------------------------------
__global__ void rank2_update_kernel(int q, double* __restrict__ X, const double* __restrict__ G,
                                    const double* __restrict__ B, int w_side, int* __restrict__ status, int outer_iter)
{
    // G is 2 x 2 column-major: G[0]=(0,0) G[1]=(1,0) G[2]=(0,1) G[3]=(1,1)
    const double a00 = G[0], a10 = G[1], a01 = G[2], a11 = G[3];
    const double eps = DBL_EPSILON;
    bool fail = (fabs(a00) < eps) && (fabs(a01) < eps);
    double t = 0.0, a2 = 1.0, b2 = 0.0, d2 = 1.0;
    const bool cosine = fabs(a00) >= fabs(a01);
    if (!fail)
    {
        if (!w_side)
        {   // SystemSolveH, nmf_solver_rank2.hpp:81-131
            if (cosine) { t = -a10 / a00; a2 = a00 - t * a10; b2 = a01 - t * a11; d2 = a11 + t * a01; }
            else        { t = -a00 / a10; a2 = -a10 + t * a00; b2 = -a11 + t * a01; d2 = a01 + t * a11; }
        }
        else
        {   // SystemSolveW, nmf_solver_rank2.hpp:157-211
            if (cosine) { t = a01 / a00; a2 = a00 + t * a01; b2 = a10 + t * a11; d2 = a11 - t * a10; }
            else        { t = a00 / a01; a2 = -a01 - t * a00; b2 = -a11 - t * a10; d2 = a10 - t * a11; }
        }
        if (fabs(d2 / a2) < eps) fail = true;
    }
    if (fail)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        return;
    }
    const double inv_a2 = 1.0 / a2, inv_d2 = 1.0 / d2;
    // OptimalActiveSet{H,W}: :218-318
    const double inv00 = 1.0 / a00, inv11 = 1.0 / a11, sq00 = sqrt(a00), sq11 = sqrt(a11);

    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < q;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const double2 bb = *reinterpret_cast<const double2*>(B + 2 * i);
        const double b0 = bb.x, b1 = bb.y;
        double e2, f2;
        if (!w_side)
        {
            if (cosine) { e2 = b0 - t * b1;  f2 = b1 + t * b0; }
            else        { e2 = -b1 + t * b0; f2 = b0 + t * b1; }
        }
        else
        {
            if (cosine) { e2 = b0 + t * b1;  f2 = b1 - t * b0; }
            else        { e2 = -b1 - t * b0; f2 = b0 - t * b1; }
        }
        double x1 = f2 * inv_d2;
        double x0 = (e2 - b2 * x1) * inv_a2;
        if (x0 <= 0.0 || x1 <= 0.0)
        {
            double v1 = b0 * inv00, v2 = b1 * inv11;
            const double vv1 = v1 * sq00, vv2 = v2 * sq11;
            if (vv1 >= vv2) v2 = 0.0; else v1 = 0.0;
            x0 = v1; x1 = v2;
        }
        *reinterpret_cast<double2*>(X + 2 * i) = make_double2(x0, x1);
    }
}

// ---------------------------------------------------------------------------
// NormalizeAndScale
// ---------------------------------------------------------------------------
// partial[block][k] = sum over this block's columns of X(r,j)^2
template <int KPL>
__global__ void row_sumsq_kernel(int k, int q, const double* __restrict__ X, double* __restrict__ partial)
{
    extern __shared__ double sh[];     // [warps][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    double acc[KPL];
#pragma unroll
    for (int e = 0; e < KPL; ++e) acc[e] = 0.0;
    for (long long j = blockIdx.x * static_cast<long long>(warps_per_block) + warp; j < q;
         j += static_cast<long long>(gridDim.x) * warps_per_block)
    {
        const double* xcol = X + j * k;
#pragma unroll
        for (int e = 0; e < KPL; ++e) { const int p = lane + 32 * e; if (p < k) { const double v = xcol[p]; acc[e] += v * v; } }
    }
#pragma unroll
    for (int e = 0; e < KPL; ++e) { const int p = lane + 32 * e; if (p < k) sh[warp * k + p] = acc[e]; }
    __syncthreads();
    for (int p = threadIdx.x; p < k; p += blockDim.x)
    {
        double s = 0.0;
        for (int w = 0; w < warps_per_block; ++w) s += sh[w * k + p];
        partial[static_cast<long long>(blockIdx.x) * k + p] = s;
    }
}

__global__ void norms_finalize_kernel(int k, int nblocks, const double* __restrict__ partial, double* __restrict__ norms,
                                      int* __restrict__ status, double* __restrict__ HHt)
{
    for (int p = threadIdx.x; p < k; p += blockDim.x)
    {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partial[static_cast<long long>(b) * k + p];
        const double nr = sqrt(s);
        if (fabs(nr) < DBL_EPSILON) status[ST_NORM_EPS] = 1;    // normalize.hpp:47-48 throws
        norms[p] = nr;
    }
    __syncthreads();
    if (HHt && threadIdx.x == 0)
    {   // nmf_solver_rank2.hpp:425-434 (k == 2)
        const double s0 = norms[0], s1 = norms[1];
        const double e00 = HHt[0], e01 = HHt[2], e11 = HHt[3];
        HHt[0] = e00 * s0 * s0;
        HHt[2] = e01 * s0 * s1;
        HHt[1] = e01 * s0 * s1;
        HHt[3] = e11 * s1 * s1;
    }
}

// X(r,j) *= (invert ? 1/norms[r] : norms[r])
__global__ void scale_rows_kernel(int k, long long q, double* __restrict__ X, const double* __restrict__ norms, int invert)
{
    const long long total = static_cast<long long>(k) * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const double nr = norms[i % k];
        const double f = invert ? (1.0 / nr) : nr;
        X[i] *= f;
    }
}

template <typename F>
void dispatch_kpl(int k, F&& f)
{
    if (k <= 32) f(std::integral_constant<int, 1>());
    else if (k <= 64) f(std::integral_constant<int, 2>());
    else if (k <= 128) f(std::integral_constant<int, 4>());
    else if (k <= 256) f(std::integral_constant<int, 8>());
    else throw std::string("k > 256 is not supported");
}

int ew_blocks(long long total, int num_sms) { return static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 8LL * num_sms))); }

} // namespace

size_t hals_sweep_scratch_doubles(int q) { return 3 * static_cast<size_t>(kHalsB) * q; }   // Q + two compact block buffers

void hals_sweep(cudaStream_t stream, int k, int q, double* X, const double* G, const double* R,
                bool normalize_rows, double* norms, double* partial, int num_sms, double* scratch)
{
    if (q <= 0) return;
    const int threads = 256, wpb = threads / 32;
    if (normalize_rows && scratch && k >= 8)
    {
        // blocked sweep (see hals_block_outer_kernel)
        const size_t smem = static_cast<size_t>((k + 3) / 4) * 64 * sizeof(double);
        SMK_CUDA(cudaFuncSetAttribute(hals_block_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int outer_blocks = std::max(1, std::min(ceil_div(q, 64), 8 * num_sms));
        const int step_blocks = std::max(1, std::min(std::min(ceil_div(static_cast<long long>(q) * kHalsB, threads), 8 * num_sms), kSweepBlocks));
        double* Qs = scratch;
        double* Xb[2] = {scratch + static_cast<size_t>(kHalsB) * q, scratch + 2 * static_cast<size_t>(kHalsB) * q};
        int cur = 0, c_last = 0;
        for (int c0 = 0; c0 < k; c0 += kHalsB, cur ^= 1)
        {
            hals_block_outer_kernel<<<outer_blocks, threads, smem, stream>>>(k, q, c0, X, G, R, Qs, Xb[cur], c0 > 0 ? Xb[cur ^ 1] : nullptr,
                                                                            partial, step_blocks, norms);
            SMK_LAUNCH_CHECK();
            const int nb = std::min(kHalsB, k - c0);
            for (int l = 0; l < nb; ++l)
            {
                hals_block_step_kernel<<<step_blocks, threads, 0, stream>>>(k, q, c0, l, Xb[cur], G, Qs, partial, step_blocks, norms);
                SMK_LAUNCH_CHECK();
            }
            c_last = c0;
        }
        hals_block_writeback_kernel<<<step_blocks, threads, 0, stream>>>(k, q, c_last, X, Xb[cur ^ 1], partial, step_blocks, norms);
        SMK_LAUNCH_CHECK();
        return;
    }
    dispatch_kpl(k, [&](auto kpl) {
        constexpr int KPL = decltype(kpl)::value;
        if (!normalize_rows)
        {
            int blocks = std::max(1, std::min(ceil_div(q, wpb), 8 * num_sms));
            hals_sweep_cols_kernel<KPL><<<blocks, threads, 0, stream>>>(k, q, X, G, R);
            SMK_LAUNCH_CHECK();
        }
        else
        {
            int blocks = std::max(1, std::min(std::min(ceil_div(q, wpb), 4 * num_sms), kSweepBlocks));
            for (int r = 0; r <= k; ++r)
            {
                hals_sweep_row_kernel<KPL><<<blocks, threads, 0, stream>>>(k, q, r, X, G, R, partial, blocks, norms);
                SMK_LAUNCH_CHECK();
            }
        }
    });
}

void rank2_update(cudaStream_t stream, int q, double* X, const double* G, const double* B, bool w_side,
                  int* status, int outer_iter)
{
    if (q <= 0) return;
    int blocks = std::max(1, std::min(ceil_div(q, 256), 148 * 8));
    rank2_update_kernel<<<blocks, 256, 0, stream>>>(q, X, G, B, w_side ? 1 : 0, status, outer_iter);
    SMK_LAUNCH_CHECK();
}

void normalize_and_scale(cudaStream_t stream, int k, int m, int n, double* Wt, double* H, double* norms,
                         int* status, double* partial, int num_sms, double* HHt, double* HAt)
{
    const int threads = 256, wpb = threads / 32;
    // partial must hold blocks*k doubles: the context allocates 4096 + 512*k, so cap the grid at 512
    int blocks = std::max(1, std::min(std::min(ceil_div(m, wpb), 2 * num_sms), 512));
    dispatch_kpl(k, [&](auto kpl) {
        constexpr int KPL = decltype(kpl)::value;
        row_sumsq_kernel<KPL><<<blocks, threads, wpb * k * sizeof(double), stream>>>(k, m, Wt, partial);
        SMK_LAUNCH_CHECK();
    });
    norms_finalize_kernel<<<1, 256, 0, stream>>>(k, blocks, partial, norms, status, HHt);
    SMK_LAUNCH_CHECK();
    scale_rows_kernel<<<ew_blocks(static_cast<long long>(k) * m, num_sms), 256, 0, stream>>>(k, m, Wt, norms, 1);
    SMK_LAUNCH_CHECK();
    scale_rows_kernel<<<ew_blocks(static_cast<long long>(k) * n, num_sms), 256, 0, stream>>>(k, n, H, norms, 0);
    SMK_LAUNCH_CHECK();
    if (HAt)
    {
        scale_rows_kernel<<<ew_blocks(static_cast<long long>(k) * m, num_sms), 256, 0, stream>>>(k, m, HAt, norms, 0);
        SMK_LAUNCH_CHECK();
    }
}

} // namespace smk
