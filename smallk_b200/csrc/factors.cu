// smallk_b200 — fused factor-update kernels: HALS sweeps, rank-2 solve + active set,
// NormalizeAndScale. All operate on "k x big" column-major factors (H, or Wt = W').
#include <cfloat>
#include "common.cuh"
#include "kernels.h"
#include "rank2_math.cuh"

namespace smk {

namespace {

constexpr int kMaxKPL = 8;          // rows per lane of the register kernels: k <= 256 (larger k: the *_big fallback kernels)
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
// k > 256 (no k-vector fits the registers of a warp): the same two sweeps with runtime loops over the rows. Fallback
// forms — correct for any k, not tuned: the H side walks a column in place (it stays in L1 for the k steps of its warp),
// the W side is k + 1 dependent launches, each of which re-reads X.
// ---------------------------------------------------------------------------
__global__ void hals_sweep_cols_big_kernel(int k, int q, double* X, const double* __restrict__ G, const double* __restrict__ R)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long j = blockIdx.x * static_cast<long long>(warps_per_block) + (threadIdx.x >> 5); j < q;
         j += static_cast<long long>(gridDim.x) * warps_per_block)
    {
        double* xcol = X + j * k;
        const double* rcol = R + j * k;
        for (int r = 0; r < k; ++r)
        {
            const double* gcol = G + static_cast<long long>(r) * k;
            double s = 0.0;
            for (int p = lane; p < k; p += 32) s += __ldg(gcol + p) * xcol[p];
            s = warp_sum(s);
            double h = xcol[r] + (__ldg(rcol + r) - s) / __ldg(gcol + r);
            if (isnan(h) || h < 0.0) h = 0.0;
            __syncwarp();                       // every lane has read the old column
            if (lane == 0) xcol[r] = h;
            __syncwarp();                       // ... and sees the new entry in the next step
        }
    }
}

// pass r of the W-side sweep: same two halves and the same partial layout as hals_sweep_row_kernel
__global__ void hals_sweep_row_big_kernel(int k, int q, int r, double* X, const double* __restrict__ G,
                                          const double* __restrict__ R, double* __restrict__ partial, int nblocks_prev,
                                          double* __restrict__ norms)
{
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
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
        for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < q;
             j += static_cast<long long>(gridDim.x) * blockDim.x)
        {
            double* px = X + j * k + (k - 1);
            *px = (fill_prev ? DBL_EPSILON : *px) * inv_prev;
        }
        return;
    }
    double sumsq = 0.0, zeros = 0.0;
    const double* grow = G + static_cast<long long>(r) * k;
    const double grr = __ldg(grow + r);
    for (long long j = blockIdx.x * static_cast<long long>(warps_per_block) + (threadIdx.x >> 5); j < q;
         j += static_cast<long long>(gridDim.x) * warps_per_block)
    {
        double* xcol = X + j * k;
        double s = 0.0, xr = 0.0;
        for (int p = lane; p < k; p += 32)
        {
            double xv = xcol[p];
            if (p == r - 1) { xv = (fill_prev ? DBL_EPSILON : xv) * inv_prev; xcol[p] = xv; }
            if (p == r) xr = xv;
            s += __ldg(grow + p) * xv;
        }
        s = warp_sum(s);
        xr = __shfl_sync(0xffffffffu, xr, r & 31);
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
// every step. Columns of W (rows of X = Wt) are taken in blocks of kHalsB. For a block [c0, c0+B):
//   phase A (hals_block_outer_kernel): for row l of the block
//            Q(l, j) = sum over p NOT in [c0, c0+l) of X(p,j) G(p,c0+l) - R(c0+l,j)
//            i.e. everything the update of row c0+l needs except the rows of this block that are updated before it
//            (the triangular mask keeps the block's own not-yet-updated rows, whose old values are still in X) —
//            one pass over X per block (k/B passes per sweep instead of k), a 16 x k x q GEMM on the FP64 tensor pipe;
//   phase B (hals_block_step_kernel<L>), B dependent steps on a COMPACT row-major copy Xb (B x q) of the block:
//            x_l <- max(0, x_l - (Q(l,j) + sum_{t<l} Xb(t,j) G(c0+t,c)) / G(c,c)),  NaN -> 0,
//            one thread per column of X, every load a fully coalesced 8-byte-per-lane stream; step l reads l + 2
//            rows and writes 2 (its own row and the unit-norm scaling of row l-1, which needs the grid-wide norm
//            of the previous step). The block that finishes last reduces the per-block partial sums in fixed order
//            and publishes 1/norm for the next launch (deterministic; no second reduction pass per launch).
// Traffic per sweep and column of X: k*k/B + k*(B/2 + 4) doubles instead of k*k.
// ---------------------------------------------------------------------------
constexpr int kHalsB = 16;

// rowinfo[2*c] = 1 / norm of row c, rowinfo[2*c+1] = 1.0 when the whole row was clamped (refill with eps)
__device__ __forceinline__ void publish_row_norm(double sumsq, double zeros, int c, int q, double* __restrict__ partial,
                                                 double* __restrict__ rowinfo, unsigned int* __restrict__ ticket,
                                                 double* __restrict__ norms, double* red)
{
    __shared__ bool s_last;
    sumsq = block_sum_f(sumsq, red);
    zeros = block_sum_f(zeros, red);
    if (threadIdx.x == 0)
    {
        partial[blockIdx.x] = sumsq;
        partial[kSweepBlocks + blockIdx.x] = zeros;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0, z = 0.0;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x)
    {
        s += __ldcg(partial + i);
        z += __ldcg(partial + kSweepBlocks + i);
    }
    s = block_sum_f(s, red);
    z = block_sum_f(z, red);
    if (threadIdx.x == 0)
    {
        const bool fill = (z == static_cast<double>(q));
        const double norm = fill ? DBL_EPSILON * sqrt(static_cast<double>(q)) : sqrt(s);
        norms[c] = norm;
        rowinfo[2 * c] = 1.0 / norm;
        rowinfo[2 * c + 1] = fill ? 1.0 : 0.0;
        *ticket = 0u;
    }
}

// Phase A on the FP64 tensor pipe: D (16 block rows x 8 columns of X) = GmT (16 x k) * X (k x 8), two DMMA.8x8x4 per
// four rows of X. One warp per 8 columns of X; the B fragments come straight from global memory (each lane reads
// X(4s + (lane & 3), j0 + (lane >> 2)): eight fully used 32-byte sectors per warp load, no shared-memory tile, no
// barriers), the A fragments of the masked Gram block sit in shared memory in fragment order.
// The block's own entries are copied to the compact buffer Xb_cur (16 x q); the previous block's finished entries are
// taken from its compact buffer and written back to X here.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
hals_block_outer_kernel(int k, int q, int c0, double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ R,
                        double* __restrict__ Q, double* __restrict__ Xb_cur, const double* __restrict__ Xb_prev,
                        const double* __restrict__ rowinfo)
{
    extern __shared__ __align__(16) double sA[];       // [ksteps][2 m-tiles][32 lanes]
    const int ksteps = (k + 3) >> 2;
    const int nb = min(kHalsB, k - c0);
    bool fill_prev = false;
    double inv_prev = 1.0;
    if (c0 > 0) { inv_prev = rowinfo[2 * (c0 - 1)]; fill_prev = rowinfo[2 * (c0 - 1) + 1] != 0.0; }
    for (int e = threadIdx.x; e < ksteps * 64; e += blockDim.x)
    {
        const int ln = e & 31, mt = (e >> 5) & 1, st = e >> 6;
        const int row = mt * 8 + (ln >> 2), p = 4 * st + (ln & 3);
        const bool updated_before = (p >= c0 && p < c0 + row);      // rows of this block updated before row c0+row
        sA[e] = (row < nb && p < k && !updated_before) ? G[static_cast<long long>(c0 + row) * k + p] : 0.0;
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
        auto mma_step = [&](int st, double b) {
            const double a0 = sA[(st * 2) * 32 + lane], a1 = sA[(st * 2 + 1) * 32 + lane];
            dmma884(c00, c01, a0, b);
            dmma884(c10, c11, a1, b);
        };
        // rows before the previous block, and rows after this block: read from X. Nothing selects on a loaded value here (a
        // select behind each load made the warp wait for it before issuing the next one: 79 % of the stall samples, r02):
        // a row index past k is clamped and meets a zero column of the masked Gram block, a column index past q is clamped
        // and only feeds accumulator columns that are never stored.
        // eight loads are issued before the first DMMA consumes one (volatile asm keeps that order: left to itself the compiler
        // re-used two registers and every DMMA waited for the load just before it — 86 % of the stall samples)
        auto plain8 = [&](int st0, int st_end) {
            double b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                const int p = 4 * (st0 + u) + kk;
                asm volatile("ld.global.f64 %0, [%1];" : "=d"(b[u]) : "l"(xcol + min(p, k - 1)));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (st0 + u < st_end) mma_step(st0 + u, b[u]);
        };
        for (int st = 0; st < max(st_prev, 0); st += 8) plain8(st, max(st_prev, 0));
        if (c0 > 0)
        {
            // the previous block: finished values from its compact buffer (row c0-1 still needs its scaling)
#pragma unroll
            for (int u = 0; u < BS; ++u)
            {
                const int t = 4 * u + kk;
                double b = Xb_prev[static_cast<long long>(t) * q + jc];
                if (t == kHalsB - 1) b = (fill_prev ? DBL_EPSILON : b) * inv_prev;
                wb[u] = b;
                mma_step(st_prev + u, b);
            }
        }
        // this block: old values (triangular mask in sA), and a copy to the compact buffer
#pragma unroll
        for (int u = 0; u < BS; ++u)
        {
            const int p = c0 + 4 * u + kk;
            cur[u] = xcol[min(p, k - 1)];
            if (st_cur + u < ksteps) mma_step(st_cur + u, cur[u]);
        }
        for (int st = st_cur + BS; st < ksteps; st += 8) plain8(st, ksteps);

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
                Xb_cur[static_cast<long long>(t) * q + j] = (c0 + t < k) ? cur[u] : 0.0;
            }
        }
    }
}

// Step L of phase B on the compact block buffer Xb (16 x q, row-major): one thread per column of X.
template <int L>
__global__ void __launch_bounds__(256)
hals_block_step_kernel(int k, int q, int c0, double* __restrict__ Xb, const double* __restrict__ G, const double* __restrict__ Q,
                       double* __restrict__ partial, double* __restrict__ rowinfo, unsigned int* __restrict__ ticket,
                       double* __restrict__ norms)
{
    __shared__ double red[32];
    __shared__ double sg[kHalsB];
    const int c = c0 + L;
    if (threadIdx.x < kHalsB) sg[threadIdx.x] = (threadIdx.x < L) ? G[static_cast<long long>(c) * k + c0 + threadIdx.x] : 0.0;
    __syncthreads();
    bool fill_prev = false;
    double inv_prev = 1.0;
    if (L > 0) { inv_prev = rowinfo[2 * (c - 1)]; fill_prev = rowinfo[2 * (c - 1) + 1] != 0.0; }
    const double gcc = G[static_cast<long long>(c) * k + c];
    double g[L > 0 ? L : 1];
#pragma unroll
    for (int t = 0; t < L; ++t) g[t] = sg[t];
    const double* ql = Q + static_cast<long long>(L) * q;
    double* xl = Xb + static_cast<long long>(L) * q;
    double sumsq = 0.0, zeros = 0.0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < q; j += stride)
    {
        double y[L > 0 ? L : 1];
#pragma unroll
        for (int t = 0; t < L; ++t) y[t] = Xb[static_cast<long long>(t) * q + j];
        const double x = xl[j];
        const double qv = __ldcs(ql + j);
        if (L > 0)
        {
            y[L > 0 ? L - 1 : 0] = (fill_prev ? DBL_EPSILON : y[L > 0 ? L - 1 : 0]) * inv_prev;
            Xb[static_cast<long long>(L > 0 ? L - 1 : 0) * q + j] = y[L > 0 ? L - 1 : 0];
        }
        double dot = 0.0;
#pragma unroll
        for (int t = 0; t < L; ++t) dot += g[t] * y[t];
        double w = x - (qv + dot) / gcc;
        if (isnan(w) || w < 0.0) { w = 0.0; zeros += 1.0; }
        xl[j] = w;
        sumsq += w * w;
    }
    publish_row_norm(sumsq, zeros, c, q, partial, rowinfo, ticket, norms, red);
}

// ---------------------------------------------------------------------------
// Phase B of a block as ONE cooperative kernel (replaces the 16 step launches when the columns of X fit): the grid-wide
// coupling of a step is a single scalar, the norm of the row just updated, so every thread keeps its columns for the whole
// block and the steps are separated by grid barriers instead of kernel boundaries. The 16 rows are taken in sub-blocks of
// kSubB = 4 whose running right-hand sides V(r, j) = Q(l, j) + sum_{t < l} G(c0+t, c0+l) y_t(j) stay in SHARED memory
// (4 x q doubles over the grid: 216 KB per SM at q = 1e6), so a step reads one row of the block (x_l) and writes one (y_l-1)
// instead of re-reading all finished rows: 72 row passes per block instead of 184.
//   per column j and step l:  y_{l-1} = scale(w_{l-1});  V(r', j) += G(c0+l-1, c0+r') y_{l-1} for the sub-block's later rows;
//                             w_l = max(0, x_l - V(l, j) / G(c, c)), NaN -> 0;  partial sums of w_l^2 and of the zero count
//   barrier; every CTA adds the per-CTA partials in CTA order (same value everywhere, independent of scheduling).
// The block's last row stays unscaled in Xb with its 1/norm in rowinfo, as the step kernels leave it.
// ---------------------------------------------------------------------------
constexpr int kSubB = 4;
constexpr int kSweepThreads = 1024;

// grid barrier state of the cooperative launch: bar[0] = arrival count, bar[1] = generation (both zero before the first use)
constexpr int kSweepColsMax = 8;       // columns per thread (two groups of four): 8 x 1024 >= the 6 912 columns whose four rows fill the shared memory of an SM

// sum over the CTA of two values at once; every warp ends up with the totals (one barrier, fixed order)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();                      // the previous use of red is over
    if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    a = warp_sum(lane < nw ? red[lane] : 0.0);
    b = warp_sum(lane < nw ? red[32 + lane] : 0.0);
}

__global__ void __launch_bounds__(kSweepThreads, 1)
hals_block_sweep_kernel(int k, int q, int c0, int nb, int cpc, double* __restrict__ Xb, const double* __restrict__ G,
                        const double* __restrict__ Q, double* __restrict__ partial, double* __restrict__ rowinfo,
                        unsigned int* __restrict__ bar, double* __restrict__ norms)
{
    extern __shared__ __align__(16) double sV[];           // [kSubB][cpc]
    __shared__ double sg[kHalsB][kHalsB + 1];              // sg[l][t] = G(c0 + l, c0 + t)
    __shared__ double red[64];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < kHalsB * kHalsB)
    {
        const int l = tid >> 4, t = tid & 15;
        sg[l][t] = (l < nb && t < nb) ? G[static_cast<long long>(c0 + l) * k + c0 + t] : 0.0;
    }
    __syncthreads();
    const long long j_begin = static_cast<long long>(blockIdx.x) * cpc;
    const int ncol = static_cast<int>(max(0LL, min(static_cast<long long>(cpc), q - j_begin)));
    const long long qq = q;
    double* xb0 = Xb + j_begin;                            // column jl of this CTA in row t: xb0[t * qq + jl]
    const double* q0 = Q + j_begin;
    unsigned int parity = 0;
    double inv_prev = 1.0;
    bool fill_prev = false;
    for (int l0 = 0; l0 < nb; l0 += kSubB)
    {
        const int ns = min(kSubB, nb - l0);
        // right-hand sides of the sub-block from everything before it (and the scaling of row l0 - 1, whose norm is known now);
        // two columns at a time: their 8 + 2 (l0 - 1) loads are in flight together
        for (int jl = tid; jl < ncol; jl += 2 * kSweepThreads)
        {
            const int jl2 = jl + kSweepThreads;
            const bool two = jl2 < ncol;
            const int jb = two ? jl2 : jl;
            double va[kSubB], vb[kSubB];
#pragma unroll
            for (int r = 0; r < kSubB; ++r)
            {
                const long long row = static_cast<long long>(l0 + min(r, ns - 1)) * qq;
                va[r] = q0[row + jl];
                vb[r] = q0[row + jb];
            }
            double ya = 0.0, yb = 0.0;
            if (l0 > 0)
            {
                ya = (fill_prev ? DBL_EPSILON : sV[(kSubB - 1) * cpc + jl]) * inv_prev;
                yb = (fill_prev ? DBL_EPSILON : sV[(kSubB - 1) * cpc + jb]) * inv_prev;
                xb0[static_cast<long long>(l0 - 1) * qq + jl] = ya;
                if (two) xb0[static_cast<long long>(l0 - 1) * qq + jb] = yb;
            }
            for (int t = 0; t < l0 - 1; ++t)
            {
                const double y1 = xb0[static_cast<long long>(t) * qq + jl], y2 = xb0[static_cast<long long>(t) * qq + jb];
#pragma unroll
                for (int r = 0; r < kSubB; ++r) { va[r] = fma(sg[l0 + r][t], y1, va[r]); vb[r] = fma(sg[l0 + r][t], y2, vb[r]); }
            }
            if (l0 > 0)
            {
#pragma unroll
                for (int r = 0; r < kSubB; ++r) { va[r] = fma(sg[l0 + r][l0 - 1], ya, va[r]); vb[r] = fma(sg[l0 + r][l0 - 1], yb, vb[r]); }
            }
#pragma unroll
            for (int r = 0; r < kSubB; ++r)
            {
                sV[r * cpc + jl] = va[r];
                if (two) sV[r * cpc + jb] = vb[r];
            }
        }
        for (int r = 0; r < ns; ++r)
        {
            const int l = l0 + r;
            const double gcc = sg[l][l];
            double sumsq = 0.0, zeros = 0.0;
            double* xrow = xb0 + static_cast<long long>(l) * qq;             // row l of the block (input, old values)
            double* yrow = xb0 + static_cast<long long>(l - 1) * qq;         // row l - 1 (output, scaled); unused when r == 0
            const double* vcur = sV + r * cpc;
            // G(c0 + l - 1, c0 + l0 + r2) for the sub-block's rows from r on (zero past the sub-block's end: nothing is added)
            double gp[kSubB];
#pragma unroll
            for (int r2 = 0; r2 < kSubB; ++r2) gp[r2] = (r > 0 && r2 >= r && r2 < ns) ? sg[l0 + r2][l - 1] : 0.0;
            const double scale_prev = inv_prev;
            const bool fill = fill_prev;
            // this thread's entries of row l in two groups (4 + 3 columns); a group's loads are all requested before its first
            // value is used (one DRAM latency per group, not per column; the whole row at once does not fit 64 registers)
#pragma unroll 1
            for (int c0g = 0; c0g < kSweepColsMax; c0g += 4)
            {
                double xv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    const int jl = tid + (c0g + c) * kSweepThreads;
                    xv[c] = xrow[ncol > 0 ? static_cast<long long>(min(jl, ncol - 1)) : -j_begin];     // always a valid address
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    const int jl = tid + (c0g + c) * kSweepThreads;
                    if (jl < ncol)
                    {
                        double v = vcur[jl];
                        if (r > 0)
                        {
                            const double wp = sV[(r - 1) * cpc + jl];
                            const double y = (fill ? DBL_EPSILON : wp) * scale_prev;
                            yrow[jl] = y;
                            v = fma(gp[r], y, v);
#pragma unroll
                            for (int r2 = 1; r2 < kSubB; ++r2)
                                if (r2 > r && r2 < ns) sV[r2 * cpc + jl] = fma(gp[r2], y, sV[r2 * cpc + jl]);
                        }
                        double w = xv[c] - v / gcc;
                        if (!(w >= 0.0)) { w = 0.0; zeros += 1.0; }           // NaN or negative (-0.0 passes, as in the reference's test)
                        sV[r * cpc + jl] = w;
                        if (l == nb - 1) xrow[jl] = w;      // the block's last row stays unscaled in Xb
                        sumsq += w * w;
                    }
                }
            }
            block_sum2(sumsq, zeros, red);
            double* pp = partial + parity * 2 * kSweepBlocks;
            if (tid == 0) { pp[blockIdx.x] = sumsq; pp[kSweepBlocks + blockIdx.x] = zeros; }
            // grid barrier; then ONE warp per CTA adds the per-CTA partials (in CTA order: the same norm on every SM) and hands
            // the totals to the others through shared memory — all 32 warps of all CTAs reading the same ten cache lines
            // serialised on their L2 slices (a third of the kernel's stall samples)
            __syncthreads();
            if (tid < 32)
            {
                if (tid == 0)
                {
                    volatile unsigned int* gen = bar + 1;
                    const unsigned int g0 = *gen;
                    __threadfence();
                    if (atomicAdd(bar, 1u) == gridDim.x - 1)
                    {
                        bar[0] = 0u;
                        __threadfence();
                        atomicExch(bar + 1, g0 + 1u);
                    }
                    else
                        while (*gen == g0) { }
                    __threadfence();
                }
                __syncwarp();
                double s1 = 0.0, z1 = 0.0;
                for (int i = lane; i < static_cast<int>(gridDim.x); i += 32) { s1 += __ldcg(pp + i); z1 += __ldcg(pp + kSweepBlocks + i); }
                s1 = warp_sum(s1);
                z1 = warp_sum(z1);
                if (lane == 0) { red[0] = s1; red[1] = z1; }
            }
            __syncthreads();
            const double s = red[0], z = red[1];
            fill_prev = (z == static_cast<double>(q));
            const double norm = fill_prev ? DBL_EPSILON * sqrt(static_cast<double>(q)) : sqrt(s);
            inv_prev = 1.0 / norm;
            if (blockIdx.x == 0 && tid == 0)
            {
                norms[c0 + l] = norm;
                rowinfo[2 * (c0 + l)] = inv_prev;
                rowinfo[2 * (c0 + l) + 1] = fill_prev ? 1.0 : 0.0;
            }
            parity ^= 1u;
        }
    }
}

// End of the sweep: the last block goes back to X, its last row (row k-1) scaled to unit norm.
__global__ void __launch_bounds__(256)
hals_block_writeback_kernel(int k, int q, int c0, double* __restrict__ X, const double* __restrict__ Xb,
                            const double* __restrict__ rowinfo)
{
    const int nb = k - c0;
    const double inv_prev = rowinfo[2 * (k - 1)];
    const bool fill_prev = rowinfo[2 * (k - 1) + 1] != 0.0;
    // thread e: column j = e / 16, entry t = e % 16: the writes to X are 128-byte pieces, the reads 8 rows of Xb
    const long long total = static_cast<long long>(q) * kHalsB;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const long long j = e / kHalsB;
        const int t = static_cast<int>(e % kHalsB);
        if (t < nb)
        {
            double v = Xb[static_cast<long long>(t) * q + j];
            if (t == nb - 1) v = (fill_prev ? DBL_EPSILON : v) * inv_prev;
            X[j * k + c0 + t] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// HALS, H side, fused and blocked: no coupling between the columns of X, so a whole sweep of a column is done in ONE
// pass, by the same blocking as the W side. A warp owns 8 columns of X in shared memory (k x 8 tile, XOR-swizzled so
// the DMMA B fragments are conflict-free). For each block of 16 rows:
//   phase A: D (16 x 8) = masked GmT (16 x k) * tile (k x 8) on the FP64 tensor pipe (two DMMA.8x8x4 per four rows),
//            A fragments of the block's Gram slab in shared memory (one slab per CTA and block, triangular mask as on
//            the W side);
//   phase B: lanes 0..7 run the 16 dependent steps of their column in registers,
//            x_l <- max(0, x_l - (D_l - R_l + sum_{t<l} x_t G(c0+t,c)) / G(c,c)),  NaN -> 0.
// X is read once and written once; G slabs come from L2 (k*k*8 bytes per 64 columns).
// ---------------------------------------------------------------------------
constexpr int kColsWarps = 8;

__device__ __forceinline__ int xs_index(int p, int n) { return p * 8 + (n ^ (((p >> 1) & 1) << 2)); }

__global__ void __launch_bounds__(kColsWarps * 32)
hals_cols_fused_kernel(int k, int q, double* __restrict__ X, const double* __restrict__ G, const double* __restrict__ R)
{
    extern __shared__ __align__(16) double smem[];
    const int ksteps = (k + 3) >> 2;
    const int kpad = ksteps * 4;
    double* sA = smem;                                   // [ksteps][2][32]
    double* gd = sA + ksteps * 64;                       // [16][16] diagonal block G(c0+t, c0+l) at t*16+l
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = gd + 256 + warp * (kpad * 8 + 256);     // [kpad][8] swizzled
    double* qs = xs + kpad * 8;                          // [16][8]  D
    double* rs = qs + 128;                               // [16][8]  R
    const int n = lane >> 2, kk = lane & 3;
    const long long nchunks = (static_cast<long long>(q) + 8 * kColsWarps - 1) / (8 * kColsWarps);
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
    {
        const long long j0 = (ch * kColsWarps + warp) * 8;
        const bool livewarp = j0 < q;
        // tile load: lane -> column (lane & 7), rows p0 + (lane >> 3)
        {
            const long long j = j0 + (lane & 7);
            const bool live = j < q;
            const double* xcol = X + (live ? j : 0) * k;
#pragma unroll 8
            for (int p0 = 0; p0 < kpad; p0 += 4)
            {
                const int p = p0 + (lane >> 3);
                xs[xs_index(p, lane & 7)] = (live && p < k) ? xcol[p] : 0.0;
            }
        }
        for (int c0 = 0; c0 < k; c0 += kHalsB)
        {
            const int nb = min(kHalsB, k - c0);
            __syncthreads();                             // previous slab no longer in use; tiles visible
            for (int e = threadIdx.x; e < ksteps * 64; e += blockDim.x)
            {
                const int ln = e & 31, mt = (e >> 5) & 1, st = e >> 6;
                const int row = mt * 8 + (ln >> 2), p = 4 * st + (ln & 3);
                const bool updated_before = (p >= c0 && p < c0 + row);
                sA[e] = (row < nb && p < k && !updated_before) ? G[static_cast<long long>(c0 + row) * k + p] : 0.0;
            }
            if (threadIdx.x < 256)
            {
                const int t = threadIdx.x >> 4, l = threadIdx.x & 15;
                gd[threadIdx.x] = (t < nb && l < nb) ? G[static_cast<long long>(c0 + l) * k + c0 + t] : (t == l ? 1.0 : 0.0);
            }
            // R tile of this block: 8 columns x 16 rows
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const int e = lane + 32 * i, col = e >> 4, row = e & 15;
                const long long j = j0 + col;
                rs[row * 8 + col] = (j < q && row < nb) ? R[j * k + c0 + row] : 0.0;
            }
            __syncthreads();
            if (!livewarp) continue;
            double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 8
            for (int st = 0; st < ksteps; ++st)
            {
                const double b = xs[xs_index(4 * st + kk, n)];
                const double a0 = sA[(st * 2) * 32 + lane], a1 = sA[(st * 2 + 1) * 32 + lane];
                dmma884(c00, c01, a0, b);
                dmma884(c10, c11, a1, b);
            }
            {
                const int row = lane >> 2, col = 2 * (lane & 3);
                qs[row * 8 + col] = c00; qs[row * 8 + col + 1] = c01;
                qs[(row + 8) * 8 + col] = c10; qs[(row + 8) * 8 + col + 1] = c11;
            }
            __syncwarp();
            if (lane < 8)
            {
                double y[kHalsB];
#pragma unroll
                for (int l = 0; l < kHalsB; ++l)
                {
                    if (l < nb)
                    {
                        const int xi = xs_index(c0 + l, lane);
                        double dot = qs[l * 8 + lane] - rs[l * 8 + lane];
#pragma unroll
                        for (int t = 0; t < l; ++t) dot += gd[t * 16 + l] * y[t];
                        double w = xs[xi] - dot / gd[l * 16 + l];
                        if (isnan(w) || w < 0.0) w = 0.0;
                        y[l] = w;
                        xs[xi] = w;
                    }
                    else y[l] = 0.0;
                }
            }
            __syncwarp();
        }
        // tile store
        {
            const long long j = j0 + (lane & 7);
            if (j < q)
            {
                double* xcol = X + j * k;
#pragma unroll 8
                for (int p0 = 0; p0 < kpad; p0 += 4)
                {
                    const int p = p0 + (lane >> 3);
                    if (p < k) xcol[p] = xs[xs_index(p, lane & 7)];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// rank 2
// ---------------------------------------------------------------------------
__global__ void rank2_update_kernel(int q, double* __restrict__ X, const double* __restrict__ G,
                                    const double* __restrict__ B, int w_side, int* __restrict__ status, int outer_iter)
{
    Rank2Solve sol;
    sol.init(G, w_side != 0);
    if (sol.fail)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        return;
    }
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < q;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const double2 bb = *reinterpret_cast<const double2*>(B + 2 * i);
        *reinterpret_cast<double2*>(X + 2 * i) = sol.apply(bb.x, bb.y);
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

// the same partial sums for k > 256: a thread per row, the block's share of the columns in column order
__global__ void row_sumsq_big_kernel(int k, int q, const double* __restrict__ X, double* __restrict__ partial)
{
    const long long per = (static_cast<long long>(q) + gridDim.x - 1) / gridDim.x;
    const long long j0 = blockIdx.x * per, j1 = (j0 + per < q) ? j0 + per : q;
    for (int p = threadIdx.x; p < k; p += blockDim.x)
    {
        double s = 0.0;
        for (long long j = j0; j < j1; ++j) { const double v = X[j * k + p]; s += v * v; }
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
    else throw std::string("dispatch_kpl: k > 256 takes the fallback kernels");
}

int ew_blocks(long long total, int num_sms) { return static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 8LL * num_sms))); }

} // namespace

// Q + two compact block buffers + the per-row norm records (2 x 256) + the finish ticket
size_t hals_sweep_scratch_doubles(int q) { return 3 * static_cast<size_t>(kHalsB) * q + 2 * 256 + 8; }

namespace {
template <int L>
void launch_block_step(cudaStream_t stream, int blocks, int k, int q, int c0, double* Xb, const double* G, const double* Q,
                       double* partial, double* rowinfo, unsigned int* ticket, double* norms)
{
    hals_block_step_kernel<L><<<blocks, 256, 0, stream>>>(k, q, c0, Xb, G, Q, partial, rowinfo, ticket, norms);
    SMK_LAUNCH_CHECK();
}
} // namespace

void hals_sweep(cudaStream_t stream, int k, int q, double* X, const double* G, const double* R,
                bool normalize_rows, double* norms, double* partial, int num_sms, double* scratch)
{
    if (q <= 0) return;
    const int threads = 256, wpb = threads / 32;
    const char* dbg = getenv("SMK_HALS_DEBUG");
    const int dbgv = dbg ? atoi(dbg) : 0;
    if (k > kMaxKPL * 32)
    {
        // any k: the fallback kernels (runtime loops over the rows of a column)
        if (!normalize_rows)
        {
            const int blocks = std::max(1, std::min(ceil_div(q, wpb), 8 * num_sms));
            hals_sweep_cols_big_kernel<<<blocks, threads, 0, stream>>>(k, q, X, G, R);
            SMK_LAUNCH_CHECK();
        }
        else
        {
            const int blocks = std::max(1, std::min(std::min(ceil_div(q, wpb), 4 * num_sms), kSweepBlocks));
            for (int r = 0; r <= k; ++r)
            {
                hals_sweep_row_big_kernel<<<blocks, threads, 0, stream>>>(k, q, r, X, G, R, partial, blocks, norms);
                SMK_LAUNCH_CHECK();
            }
        }
        return;
    }
    if (normalize_rows && scratch && k >= 8 && !(dbgv & 1))
    {
        // blocked sweep (see hals_block_outer_kernel)
        const size_t smem = static_cast<size_t>((k + 3) / 4) * 64 * sizeof(double);
        // resident CTAs per SM the outer pass is compiled for (register budget 65536 / (256 * MINB)); SMK_HALS_OUTER_OCC overrides
        static const int outer_occ = [] { const char* e = getenv("SMK_HALS_OUTER_OCC"); const int v = e ? atoi(e) : 3; return v == 3 ? 3 : (v == 1 ? 1 : 2); }();      // 3: C3 step 25.2 -> 24.3 ms
        auto outer_kernel = outer_occ == 3 ? hals_block_outer_kernel<3> : (outer_occ == 1 ? hals_block_outer_kernel<1> : hals_block_outer_kernel<2>);
        SMK_CUDA(cudaFuncSetAttribute(outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int outer_blocks = std::max(1, std::min(ceil_div(q, 64), 8 * num_sms));
        const int step_blocks = std::max(1, std::min(std::min(ceil_div(q, threads), 8 * num_sms), kSweepBlocks));
        double* Qs = scratch;
        double* Xb[2] = {scratch + static_cast<size_t>(kHalsB) * q, scratch + 2 * static_cast<size_t>(kHalsB) * q};
        double* rowinfo = scratch + 3 * static_cast<size_t>(kHalsB) * q;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(rowinfo + 2 * 256);
        SMK_CUDA(cudaMemsetAsync(ticket, 0, 2 * sizeof(unsigned int), stream));
        // phase B as one cooperative kernel per block when a CTA per SM can hold its share of 4 rows in shared memory (SMK_HALS_SWEEP=0: step kernels)
        const char* sweep_env = getenv("SMK_HALS_SWEEP");          // read per call: the tests run both forms in one process
        const bool sweep_on = !(sweep_env && atoi(sweep_env) == 0);
        const int sweep_grid = std::max(1, std::min(num_sms, ceil_div(q, 256)));
        const int cpc = ceil_div(q, sweep_grid);
        const size_t sweep_smem = static_cast<size_t>(kSubB) * cpc * sizeof(double);
        const bool use_sweep = sweep_on && sweep_smem <= 216 * 1024 + 512 && sweep_grid <= kSweepBlocks && cpc <= kSweepColsMax * kSweepThreads;
        if (use_sweep)
        {
            static bool attr_set = false;
            if (!attr_set) { SMK_CUDA(cudaFuncSetAttribute(hals_block_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); attr_set = true; }
        }
        int cur = 0, c_last = 0;
        for (int c0 = 0; c0 < k; c0 += kHalsB, cur ^= 1)
        {
            outer_kernel<<<outer_blocks, threads, smem, stream>>>(k, q, c0, X, G, R, Qs, Xb[cur], c0 > 0 ? Xb[cur ^ 1] : nullptr,
                                                                            rowinfo);
            SMK_LAUNCH_CHECK();
            const int nb = std::min(kHalsB, k - c0);
            if (use_sweep)
            {
                int k_ = k, q_ = q, c0_ = c0, nb_ = nb, cpc_ = cpc;
                double* xb = Xb[cur];
                const double* g_ = G; const double* qs_ = Qs;
                void* args[] = {&k_, &q_, &c0_, &nb_, &cpc_, &xb, &g_, &qs_, &partial, &rowinfo, &ticket, &norms};
                SMK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(hals_block_sweep_kernel), dim3(sweep_grid), dim3(kSweepThreads),
                                                     args, sweep_smem, stream));
                SMK_LAUNCH_CHECK();
            }
            else
            for (int l = 0; l < nb; ++l)
            {
#define SMK_STEP(L) case L: launch_block_step<L>(stream, step_blocks, k, q, c0, Xb[cur], G, Qs, partial, rowinfo, ticket, norms); break;
                switch (l)
                {
                    SMK_STEP(0) SMK_STEP(1) SMK_STEP(2) SMK_STEP(3) SMK_STEP(4) SMK_STEP(5) SMK_STEP(6) SMK_STEP(7)
                    SMK_STEP(8) SMK_STEP(9) SMK_STEP(10) SMK_STEP(11) SMK_STEP(12) SMK_STEP(13) SMK_STEP(14) SMK_STEP(15)
                }
#undef SMK_STEP
            }
            c_last = c0;
        }
        const int wb_blocks = std::max(1, std::min(ceil_div(static_cast<long long>(q) * kHalsB, threads), 8 * num_sms));
        hals_block_writeback_kernel<<<wb_blocks, threads, 0, stream>>>(k, q, c_last, X, Xb[cur ^ 1], rowinfo);
        SMK_LAUNCH_CHECK();
        return;
    }
    dispatch_kpl(k, [&](auto kpl) {
        constexpr int KPL = decltype(kpl)::value;
        if (!normalize_rows && k >= 8 && !(dbgv & 2))
        {
            const int kpad = ((k + 3) / 4) * 4;
            const size_t smem = (static_cast<size_t>(kpad) * 16 + 256 + kColsWarps * (static_cast<size_t>(kpad) * 8 + 256)) * sizeof(double);
            SMK_CUDA(cudaFuncSetAttribute(hals_cols_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            const int per_sm = smem <= 112 * 1024 ? 2 : 1;
            const int blocks = std::max(1, std::min(ceil_div(q, 8 * kColsWarps), per_sm * num_sms));
            hals_cols_fused_kernel<<<blocks, kColsWarps * 32, smem, stream>>>(k, q, X, G, R);
            SMK_LAUNCH_CHECK();
        }
        else if (!normalize_rows)
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
    // partial must hold blocks*k doubles: the context allocates 4096 + 512 * max(k, 256), so cap the grid at 512
    int blocks = std::max(1, std::min(std::min(ceil_div(m, wpb), 2 * num_sms), 512));
    if (k > kMaxKPL * 32)
    {
        row_sumsq_big_kernel<<<blocks, threads, 0, stream>>>(k, m, Wt, partial);
        SMK_LAUNCH_CHECK();
    }
    else
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
