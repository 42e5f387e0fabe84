// smallk_b200 — sparse (CSC) x dense products for the NMF iteration.
//
// The reference's four sparse Gemm variants (common/include/sparse_gemm.hpp:26-74,
// sparse_gemm_ab_impl.hpp, sparse_gemm_ba_impl.hpp) reduce, for the solvers, to two
// products with a "k x big" dense operand:
//     W'A  (k x n):  out(:,j) = sum_{(r,a) in column j of A}  a * Wt(:,r)       [B'*A]
//     H A' (k x m):  out(:,i) = sum_{(c,a) in row i of A}     a * H(:,c)        [A*B' transposed]
// The first walks the CSC arrays; the second walks a CSR copy of A built once on the
// device (= the reference's stable Transpose, sparse_matrix_ops.hpp:37-126), so BOTH
// are gather-reduce kernels: no atomics, each output column written once, and the
// entries of a compressed column are added in storage order exactly as the reference
// adds them. Dense columns Wt(:,r) / H(:,c) are contiguous k-vectors, so each gathered
// operand is one coalesced k*8-byte read.
#include <cub/cub.cuh>
#include <vector>
#include <cstdlib>
#include "context.h"

namespace smk {

namespace {

// L lanes cooperate on one output column, each lane owning KPL rows (row = lane + L*e).
template <int L, int KPL>
__global__ void spmm_gather_kernel(int ncols, const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ idx,
                                   const double* __restrict__ val, int k, const double* __restrict__ B, long long ldb,
                                   double alpha, double beta, double* __restrict__ out, long long ldo)
{
    const int groups_per_block = blockDim.x / L;
    const int lane = threadIdx.x % L;
    for (long long j = blockIdx.x * static_cast<long long>(groups_per_block) + threadIdx.x / L; j < ncols;
         j += static_cast<long long>(gridDim.x) * groups_per_block)
    {
        double acc[KPL];
#pragma unroll
        for (int e = 0; e < KPL; ++e)
        {
            const int r = lane + L * e;
            double c0 = 0.0;
            if (beta != 0.0 && r < k) c0 = out[j * ldo + r] * beta;
            acc[e] = c0;
        }
        const unsigned int beg = ptr[j], end = ptr[j + 1];
        unsigned int o = beg;
        // 4 entries per trip: issue the index/value loads and the gathers together
        for (; o + 4 <= end; o += 4)
        {
            unsigned int i0 = idx[o], i1 = idx[o + 1], i2 = idx[o + 2], i3 = idx[o + 3];
            double a0 = alpha * val[o], a1 = alpha * val[o + 1], a2 = alpha * val[o + 2], a3 = alpha * val[o + 3];
            double b0[KPL], b1[KPL], b2[KPL], b3[KPL];
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                const int r = lane + L * e;
                const bool ok = r < k;
                b0[e] = ok ? B[i0 * ldb + r] : 0.0;
                b1[e] = ok ? B[i1 * ldb + r] : 0.0;
                b2[e] = ok ? B[i2 * ldb + r] : 0.0;
                b3[e] = ok ? B[i3 * ldb + r] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                acc[e] += a0 * b0[e];
                acc[e] += a1 * b1[e];
                acc[e] += a2 * b2[e];
                acc[e] += a3 * b3[e];
            }
        }
        for (; o < end; ++o)
        {
            const unsigned int i0 = idx[o];
            const double a0 = alpha * val[o];
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                const int r = lane + L * e;
                if (r < k) acc[e] += a0 * B[i0 * ldb + r];
            }
        }
#pragma unroll
        for (int e = 0; e < KPL; ++e)
        {
            const int r = lane + L * e;
            if (r < k) out[j * ldo + r] = acc[e];
        }
    }
}

template <int L, int KPL>
void launch_gather(cudaStream_t stream, int ncols, const unsigned int* ptr, const unsigned int* idx, const double* val,
                   int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo, int num_sms)
{
    const int threads = 256;
    const int gpb = threads / L;
    int blocks = std::max(1, std::min(ceil_div(ncols, gpb), 16 * num_sms));
    spmm_gather_kernel<L, KPL><<<blocks, threads, 0, stream>>>(ncols, ptr, idx, val, k, B, ldb, alpha, beta, out, ldo);
    SMK_LAUNCH_CHECK();
}

// column index of every stored entry (expands colptr)
__global__ void expand_cols_kernel(int n, const unsigned int* __restrict__ colptr, unsigned int* __restrict__ colof)
{
    for (int j = blockIdx.x; j < n; j += gridDim.x)
        for (unsigned int o = colptr[j] + threadIdx.x; o < colptr[j + 1]; o += blockDim.x) colof[o] = j;
}

__global__ void iota_kernel(unsigned int n, unsigned int* __restrict__ v)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

__global__ void gather_perm_kernel(unsigned int n, const unsigned int* __restrict__ perm, const unsigned int* __restrict__ colof,
                                   const double* __restrict__ val, unsigned int* __restrict__ colidx, double* __restrict__ valr)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const unsigned int s = perm[i];
        colidx[i] = colof[s];
        valr[i] = val[s];
    }
}

// rowptr[r] = first position in the sorted row keys with key >= r
__global__ void rowptr_kernel(int m, unsigned int nnz, const unsigned int* __restrict__ sorted_rows, unsigned int* __restrict__ rowptr)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= m; r += gridDim.x * blockDim.x)
    {
        unsigned int lo = 0, hi = nnz;
        while (lo < hi)
        {
            const unsigned int mid = lo + ((hi - lo) >> 1);
            if (sorted_rows[mid] < static_cast<unsigned int>(r)) lo = mid + 1; else hi = mid;
        }
        rowptr[r] = lo;
    }
}

// ---------------------------------------------------------------------------
// segmented variant: one group of L lanes per SEGMENT (<= kSpmmSeg stored entries of one compressed column)
// ---------------------------------------------------------------------------
__global__ void seg_count_kernel(int ncols, const unsigned int* __restrict__ ptr, unsigned int* __restrict__ cnt,
                                 unsigned int* __restrict__ mcnt, unsigned int* __restrict__ flag)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= ncols; j += gridDim.x * blockDim.x)
    {
        unsigned int c = 0;
        if (j < ncols) { const unsigned int len = ptr[j + 1] - ptr[j]; c = len == 0 ? 1u : (len + kSpmmSeg - 1) / kSpmmSeg; }
        cnt[j] = c;
        mcnt[j] = c > 1 ? c : 0u;
        flag[j] = c > 1 ? 1u : 0u;
    }
}

// one warp per column fills the records of its segments
__global__ void seg_fill_kernel(int ncols, const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ first,
                                const unsigned int* __restrict__ first_slot, const unsigned int* __restrict__ mpos,
                                unsigned int* __restrict__ scol, unsigned int* __restrict__ sbeg, unsigned int* __restrict__ send,
                                unsigned int* __restrict__ sslot, unsigned int* __restrict__ multi_col)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < ncols; j += warps)
    {
        const unsigned int f = first[j], cnt = first[j + 1] - f, b0 = ptr[j], e0 = ptr[j + 1];
        for (unsigned int sg = lane; sg < cnt; sg += 32)
        {
            const unsigned int b = b0 + sg * kSpmmSeg;
            scol[f + sg] = j;
            sbeg[f + sg] = b;
            send[f + sg] = min(e0, b + kSpmmSeg);
            sslot[f + sg] = cnt > 1 ? first_slot[j] + sg : 0xFFFFFFFFu;
        }
        if (cnt > 1 && lane == 0) multi_col[mpos[j]] = j;
    }
}

template <int L, int KPL>
__global__ void spmm_seg_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                                const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                                const unsigned int* __restrict__ idx, const double* __restrict__ val, int k,
                                const double* __restrict__ B, long long ldb, double alpha, double beta,
                                double* __restrict__ out, long long ldo, double* __restrict__ partial)
{
    const int groups_per_block = blockDim.x / L;
    const int lane = threadIdx.x % L;
    for (long long it = blockIdx.x * static_cast<long long>(groups_per_block) + threadIdx.x / L; it < nseg;
         it += static_cast<long long>(gridDim.x) * groups_per_block)
    {
        const unsigned int j = scol[it], slot = sslot[it];
        const bool direct = slot == 0xFFFFFFFFu;
        double acc[KPL];
#pragma unroll
        for (int e = 0; e < KPL; ++e)
        {
            const int r = lane + L * e;
            acc[e] = (direct && beta != 0.0 && r < k) ? out[j * ldo + r] * beta : 0.0;
        }
        const unsigned int end = send[it];
        unsigned int o = sbeg[it];
        for (; o + 4 <= end; o += 4)
        {
            const unsigned int i0 = idx[o], i1 = idx[o + 1], i2 = idx[o + 2], i3 = idx[o + 3];
            const double a0 = alpha * val[o], a1 = alpha * val[o + 1], a2 = alpha * val[o + 2], a3 = alpha * val[o + 3];
            double b0[KPL], b1[KPL], b2[KPL], b3[KPL];
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                const int r = lane + L * e;
                const bool ok = r < k;
                b0[e] = ok ? B[i0 * ldb + r] : 0.0;
                b1[e] = ok ? B[i1 * ldb + r] : 0.0;
                b2[e] = ok ? B[i2 * ldb + r] : 0.0;
                b3[e] = ok ? B[i3 * ldb + r] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                acc[e] += a0 * b0[e];
                acc[e] += a1 * b1[e];
                acc[e] += a2 * b2[e];
                acc[e] += a3 * b3[e];
            }
        }
        for (; o < end; ++o)
        {
            const unsigned int i0 = idx[o];
            const double a0 = alpha * val[o];
#pragma unroll
            for (int e = 0; e < KPL; ++e)
            {
                const int r = lane + L * e;
                if (r < k) acc[e] += a0 * B[i0 * ldb + r];
            }
        }
        double* dst = direct ? out + j * ldo : partial + static_cast<long long>(slot) * k;
#pragma unroll
        for (int e = 0; e < KPL; ++e)
        {
            const int r = lane + L * e;
            if (r < k) dst[r] = acc[e];
        }
    }
}

// Wide-k variant of spmm_seg_kernel (k even, 64 <= k <= 256): one warp per segment, lane l owns the double2 pieces
// (2l, 2l+1) + 64v of the k-vector, v < NV, so a gathered operand B(:, i) is NV 512-byte LDG.128 wavefronts.
// The (index, value) pairs of a segment are read 32 at a time, one pair per lane (two coalesced loads instead of 64
// broadcast loads), and handed round by shuffles; U gathered operands (U * NV LDG.128 per lane, U KB per warp) are in
// flight before the first FMA consumes one. The gathers are what the kernel waits for: at 8 KB in flight per warp
// and 16 warps per SM the memory system, not the issue rate, sets the pace. The entries are still added in storage
// order, one fused multiply-add per entry and output element, as in the scalar kernel.
template <int NV, int U>
__global__ void __launch_bounds__(256, 2)
spmm_seg_wide_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                     const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                     const unsigned int* __restrict__ idx, const double* __restrict__ val, int k,
                     const double* __restrict__ B, long long ldb, double alpha, double beta,
                     double* __restrict__ out, long long ldo, double* __restrict__ partial)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    bool live[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) live[v] = 2 * lane + 64 * v < k;
    for (long long it = blockIdx.x * static_cast<long long>(wpb) + (threadIdx.x >> 5); it < nseg;
         it += static_cast<long long>(gridDim.x) * wpb)
    {
        const unsigned int j = scol[it], slot = sslot[it];
        const bool direct = slot == 0xFFFFFFFFu;
        double2 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v)
        {
            acc[v] = make_double2(0.0, 0.0);
            if (direct && beta != 0.0 && live[v])
            {
                const double2 c0 = *reinterpret_cast<const double2*>(out + j * ldo + 2 * lane + 64 * v);
                acc[v] = make_double2(c0.x * beta, c0.y * beta);
            }
        }
        const unsigned int end = send[it];
        unsigned int o = sbeg[it];
        unsigned int nxt_i = 0; double nxt_a = 0.0;
        if (o + lane < end) { nxt_i = idx[o + lane]; nxt_a = alpha * val[o + lane]; }
        while (o < end)
        {
            const int cnt = min(32u, end - o);
            const unsigned int my_i = nxt_i; const double my_a = nxt_a;
            o += 32;
            if (o + lane < end) { nxt_i = idx[o + lane]; nxt_a = alpha * val[o + lane]; }      // next batch, behind the gathers
            int t = 0;
            for (; t + U <= cnt; t += U)
            {
                double2 b[U][NV];
                double a[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const unsigned int iu = __shfl_sync(0xffffffffu, my_i, t + u);
                    a[u] = __shfl_sync(0xffffffffu, my_a, t + u);
                    const double* bcol = B + iu * ldb + 2 * lane;
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        b[u][v] = live[v] ? __ldg(reinterpret_cast<const double2*>(bcol + 64 * v)) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < NV; ++v) { acc[v].x += a[u] * b[u][v].x; acc[v].y += a[u] * b[u][v].y; }
            }
            for (; t < cnt; ++t)
            {
                const unsigned int iu = __shfl_sync(0xffffffffu, my_i, t);
                const double au = __shfl_sync(0xffffffffu, my_a, t);
                const double* bcol = B + iu * ldb + 2 * lane;
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (live[v])
                    {
                        const double2 bv = __ldg(reinterpret_cast<const double2*>(bcol + 64 * v));
                        acc[v].x += au * bv.x; acc[v].y += au * bv.y;
                    }
            }
        }
        double* dst = direct ? out + j * ldo : partial + static_cast<long long>(slot) * k;
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (live[v]) *reinterpret_cast<double2*>(dst + 2 * lane + 64 * v) = acc[v];
    }
}

// ---------------------------------------------------------------------------
// Tiered variant (k % 4 == 0, 64 <= k <= 256, fewer than 2^30 gathered vectors): 256-bit gathers + residency classes.
//
// A gather SpMM moves k*8 bytes of the dense operand per stored entry (C3: 1 KB x 7.7e7 = 79 GB per product), ~60x the
// compulsory HBM bytes, so what bounds it is where those bytes come from. Term frequencies are Zipf-like: a few
// hundred rows of A take a quarter of the entries and a few percent of the rows take three quarters, while the
// operand as a whole (C3: Wt = 1 GB) is far larger than L2 and the long tail of cold rows keeps flushing the
// popular ones out of it. The two top bits of a private copy of the index array therefore carry a residency class
// per entry, chosen once per matrix from the degree of the gathered vector (build_gather_tiers):
//     10 | slot : the vector is one of the `smem_rows` most popular ones and is served from shared memory, which
//                 every CTA fills once (conflict-free LDS.128 pairs; no L2 traffic at all for these entries),
//     01 | id   : popular enough to be kept in L2   -> LDG.E.EL.ELL2.256  (L1 and L2 evict_last),
//     11 | id   : cold tail, touched and dropped    -> LDG.E.NA.EFL2.256  (L1 no_allocate, L2 evict_first),
//     00 | id   : no tiers built (small or unskewed operand): plain LDG.E.256.
// Lane l owns the four consecutive doubles 4l..4l+3 (+128v) of the k-vector, so one gathered operand is NV 256-bit
// loads per lane (sm_100 LDG.256) instead of 2 NV LDG.128. Entries are still added in storage order, one fused
// multiply-add per entry and output element: results are those of the kernels above.
// ---------------------------------------------------------------------------
struct d4 { double x, y, z, w; };

__device__ __forceinline__ d4 ldg256(const double* p)
{
    d4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ d4 ldg256_keep(const double* p)
{
    d4 r;
    asm("ld.global.nc.L1::evict_last.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ d4 ldg256_drop(const double* p)
{
    d4 r;
    asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

constexpr int kTierSmemBytes = 200 * 1024;                    // shared-memory tier per CTA (one 512-thread CTA per SM)
constexpr size_t kTierKeepBytes = size_t(48) << 20;           // L2-resident tier: well under half of the 126 MB L2
constexpr size_t kTierMinOperandBytes = size_t(96) << 20;     // smaller operands live in L2 without help
constexpr size_t kSlabMaxBytes = size_t(64) << 20;            // a 32-row operand slab of at most this size is gathered from L2
constexpr unsigned int kTierSmem = 0x80000000u, kTierKeep = 0x40000000u, kTierDrop = 0xC0000000u, kTierMask = 0xC0000000u;

// one gathered operand piece, by residency class (the class is warp-uniform: no divergence)
__device__ __forceinline__ d4 tier_load(unsigned int code, const double* __restrict__ B, long long ldb, const double* sh, int k, int off, int q)
{
    const unsigned int cls = code & kTierMask;
    if (cls == kTierSmem)
    {
        const double* row = sh + static_cast<size_t>(code & ~kTierMask) * k;
        const double2 lo = *reinterpret_cast<const double2*>(row + 2 * q);
        const double2 hi = *reinterpret_cast<const double2*>(row + (k >> 1) + 2 * q);
        d4 r; r.x = lo.x; r.y = lo.y; r.z = hi.x; r.w = hi.y;
        return r;
    }
    const double* p = B + static_cast<long long>(code & ~kTierMask) * ldb + off;
    if (cls == kTierDrop) return ldg256_drop(p);
    if (cls == kTierKeep) return ldg256_keep(p);
    return ldg256(p);
}

template <int NV, int U>
__global__ void __launch_bounds__(512, 1)
spmm_seg_tier_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                     const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                     const unsigned int* __restrict__ idx, const double* __restrict__ val, int k,
                     const double* __restrict__ B, long long ldb, double alpha, double beta,
                     double* __restrict__ out, long long ldo, double* __restrict__ partial,
                     int smem_rows, const unsigned int* __restrict__ smem_ids)
{
    extern __shared__ __align__(16) double tier_sh[];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    bool live[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) live[v] = 4 * lane + 128 * v < k;
    // shared-memory tier: row s = operand vector smem_ids[s]; the piece (lane, v) is stored as two double2 halves,
    // k/2 doubles apart, at double2 index q = 32v + lane, so both LDS.128 of a warp are contiguous
    for (int s = warp; s < smem_rows; s += wpb)
    {
        const double* src = B + static_cast<long long>(smem_ids[s]) * ldb;
        double* row = tier_sh + static_cast<size_t>(s) * k;
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (live[v])
            {
                const d4 x = ldg256_keep(src + 4 * lane + 128 * v);
                const int q = 32 * v + lane;
                *reinterpret_cast<double2*>(row + 2 * q) = make_double2(x.x, x.y);
                *reinterpret_cast<double2*>(row + (k >> 1) + 2 * q) = make_double2(x.z, x.w);
            }
    }
    __syncthreads();

    for (long long it = blockIdx.x * static_cast<long long>(wpb) + warp; it < nseg; it += static_cast<long long>(gridDim.x) * wpb)
    {
        const unsigned int j = scol[it], slot = sslot[it];
        const bool direct = slot == 0xFFFFFFFFu;
        d4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v)
        {
            acc[v].x = acc[v].y = acc[v].z = acc[v].w = 0.0;
            if (direct && beta != 0.0 && live[v])
            {
                const double* c0 = out + j * ldo + 4 * lane + 128 * v;     // plain loads: out is written by this kernel
                const double2 lo = *reinterpret_cast<const double2*>(c0), hi = *reinterpret_cast<const double2*>(c0 + 2);
                acc[v].x = lo.x * beta; acc[v].y = lo.y * beta; acc[v].z = hi.x * beta; acc[v].w = hi.y * beta;
            }
        }
        const unsigned int end = send[it];
        unsigned int o = sbeg[it];
        unsigned int nxt_i = 0; double nxt_a = 0.0;
        if (o + lane < end) { nxt_i = __ldcs(idx + o + lane); nxt_a = alpha * __ldcs(val + o + lane); }
        while (o < end)
        {
            const int cnt = min(32u, end - o);
            const unsigned int my_i = nxt_i; const double my_a = nxt_a;
            o += 32;
            if (o + lane < end) { nxt_i = __ldcs(idx + o + lane); nxt_a = alpha * __ldcs(val + o + lane); }
            int t = 0;
            for (; t + U <= cnt; t += U)
            {
                d4 b[U][NV];
                double a[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const unsigned int code = __shfl_sync(0xffffffffu, my_i, t + u);
                    a[u] = __shfl_sync(0xffffffffu, my_a, t + u);
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                    {
                        if (live[v]) b[u][v] = tier_load(code, B, ldb, tier_sh, k, 4 * lane + 128 * v, 32 * v + lane);
                        else { b[u][v].x = b[u][v].y = b[u][v].z = b[u][v].w = 0.0; }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                    {
                        acc[v].x += a[u] * b[u][v].x; acc[v].y += a[u] * b[u][v].y;
                        acc[v].z += a[u] * b[u][v].z; acc[v].w += a[u] * b[u][v].w;
                    }
            }
            for (; t < cnt; ++t)
            {
                const unsigned int code = __shfl_sync(0xffffffffu, my_i, t);
                const double au = __shfl_sync(0xffffffffu, my_a, t);
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (live[v])
                    {
                        const d4 bv = tier_load(code, B, ldb, tier_sh, k, 4 * lane + 128 * v, 32 * v + lane);
                        acc[v].x += au * bv.x; acc[v].y += au * bv.y; acc[v].z += au * bv.z; acc[v].w += au * bv.w;
                    }
            }
        }
        double* dst = direct ? out + j * ldo : partial + static_cast<long long>(slot) * k;
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (live[v])
            {
                double* p = dst + 4 * lane + 128 * v;
                *reinterpret_cast<double2*>(p) = make_double2(acc[v].x, acc[v].y);
                *reinterpret_cast<double2*>(p + 2) = make_double2(acc[v].z, acc[v].w);
            }
    }
}

// ---------------------------------------------------------------------------
// k-slab variant: an operand that is larger than L2 but whose 32-row slab fits (C3: H = 205 MB, slab 51 MB) is gathered
// one slab per launch, so every gather after the first touch of a vector is an L2 hit; A's index/value arrays are
// streamed once per slab (12 bytes per entry, against the 256 bytes the entry gathers). 8 lanes own a segment (lane g
// the doubles koff + 4g .. 4g+3 of the k-vector: one LDG.E.256 per lane and entry), four segments per warp; the
// (index, value) pairs are read 8 at a time and handed round inside the group. Entries are added in storage order
// with one fused multiply-add per entry and output element: bit for bit the sums of the kernels above.
// ---------------------------------------------------------------------------
constexpr int kSlab = 32;

__global__ void __launch_bounds__(256, 2)
spmm_seg_slab_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                     const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                     const unsigned int* __restrict__ idx, const double* __restrict__ val, int k, int koff,
                     const double* __restrict__ B, long long ldb, double alpha, double beta,
                     double* __restrict__ out, long long ldo, double* __restrict__ partial)
{
    const int lane = threadIdx.x & 31, g = lane & 7;
    const unsigned int mask = 0xFFu << (lane & ~7);
    const long long ngroups = static_cast<long long>(gridDim.x) * (blockDim.x >> 3);
    const int off = koff + 4 * g;
    for (long long it = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3; it < nseg; it += ngroups)
    {
        const unsigned int j = scol[it], slot = sslot[it];
        const bool direct = slot == 0xFFFFFFFFu;
        d4 acc; acc.x = acc.y = acc.z = acc.w = 0.0;
        if (direct && beta != 0.0)
        {
            const double* c0 = out + j * ldo + off;
            const double2 lo = *reinterpret_cast<const double2*>(c0), hi = *reinterpret_cast<const double2*>(c0 + 2);
            acc.x = lo.x * beta; acc.y = lo.y * beta; acc.z = hi.x * beta; acc.w = hi.y * beta;
        }
        const unsigned int end = send[it];
        unsigned int o = sbeg[it];
        unsigned int nxt_i = 0; double nxt_a = 0.0;
        if (o + g < end) { nxt_i = __ldcs(idx + o + g); nxt_a = alpha * __ldcs(val + o + g); }
        while (o < end)
        {
            const int cnt = min(8u, end - o);
            const unsigned int my_i = nxt_i; const double my_a = nxt_a;
            o += 8;
            nxt_i = 0; nxt_a = 0.0;
            if (o + g < end) { nxt_i = __ldcs(idx + o + g); nxt_a = alpha * __ldcs(val + o + g); }
            d4 b[8];
            double a[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
            {
                const unsigned int iu = __shfl_sync(mask, my_i, t, 8);
                a[t] = __shfl_sync(mask, my_a, t, 8);
                if (t < cnt) b[t] = ldg256(B + iu * ldb + off);
                else { b[t].x = b[t].y = b[t].z = b[t].w = 0.0; }
            }
            if (cnt == 8)
            {
#pragma unroll
                for (int t = 0; t < 8; ++t) { acc.x += a[t] * b[t].x; acc.y += a[t] * b[t].y; acc.z += a[t] * b[t].z; acc.w += a[t] * b[t].w; }
            }
            else
            {
                for (int t = 0; t < cnt; ++t) { acc.x += a[t] * b[t].x; acc.y += a[t] * b[t].y; acc.z += a[t] * b[t].z; acc.w += a[t] * b[t].w; }
            }
        }
        double* p = (direct ? out + j * ldo : partial + static_cast<long long>(slot) * k) + off;
        *reinterpret_cast<double2*>(p) = make_double2(acc.x, acc.y);
        *reinterpret_cast<double2*>(p + 2) = make_double2(acc.z, acc.w);
    }
}

// residency class of every gatherable vector from its rank in decreasing degree order
__global__ void tier_code_kernel(int count, const unsigned int* __restrict__ ids_by_degree, int smem_rows, int keep_rows,
                                 unsigned int* __restrict__ code, unsigned int* __restrict__ smem_ids)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x)
    {
        const unsigned int id = ids_by_degree[r];
        unsigned int c;
        if (r < smem_rows) { c = kTierSmem | static_cast<unsigned int>(r); smem_ids[r] = id; }
        else if (r < smem_rows + keep_rows) c = kTierKeep | id;
        else c = kTierDrop | id;
        code[id] = c;
    }
}

__global__ void tier_degree_kernel(int count, const unsigned int* __restrict__ ptr, unsigned int* __restrict__ deg, unsigned int* __restrict__ ids)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x) { deg[r] = ptr[r + 1] - ptr[r]; ids[r] = r; }
}

__global__ void tier_apply_kernel(unsigned int nnz, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ code, unsigned int* __restrict__ out)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x) out[i] = code[idx[i]];
}

// out(:, j) = beta * out(:, j) + sum of the partials of column j, in segment order
__global__ void spmm_combine_kernel(int nmulti, const unsigned int* __restrict__ multi_col, const unsigned int* __restrict__ first_slot,
                                    int k, double beta, const double* __restrict__ partial, double* __restrict__ out, long long ldo)
{
    for (int i = blockIdx.x; i < nmulti; i += gridDim.x)
    {
        const unsigned int j = multi_col[i];
        const unsigned int s0 = first_slot[j], s1 = first_slot[j + 1];
        for (int r = threadIdx.x; r < k; r += blockDim.x)
        {
            double acc = (beta != 0.0) ? out[j * ldo + r] * beta : 0.0;
            unsigned int sl = s0;
            for (; sl + 4 <= s1; sl += 4)
            {
                const double p0 = partial[static_cast<long long>(sl) * k + r], p1 = partial[static_cast<long long>(sl + 1) * k + r];
                const double p2 = partial[static_cast<long long>(sl + 2) * k + r], p3 = partial[static_cast<long long>(sl + 3) * k + r];
                acc += p0; acc += p1; acc += p2; acc += p3;
            }
            for (; sl < s1; ++sl) acc += partial[static_cast<long long>(sl) * k + r];
            out[j * ldo + r] = acc;
        }
    }
}

template <int L, int KPL>
void launch_seg(cudaStream_t stream, const SegTable& T, const unsigned int* idx, const double* val, int k, const double* B,
                long long ldb, double alpha, double beta, double* out, long long ldo, double* partial, int num_sms)
{
    const int threads = 256;
    const int gpb = threads / L;
    const int blocks = std::max(1, std::min(ceil_div(T.nseg, gpb), 32 * num_sms));
    spmm_seg_kernel<L, KPL><<<blocks, threads, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, B, ldb,
                                                           alpha, beta, out, ldo, partial);
    SMK_LAUNCH_CHECK();
}

} // namespace

void spmm_gather(cudaStream_t stream, int ncols, const unsigned int* ptr, const unsigned int* idx, const double* val,
                 int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo, int num_sms)
{
    if (ncols <= 0) return;
#define SMK_G(L, KPL) launch_gather<L, KPL>(stream, ncols, ptr, idx, val, k, B, ldb, alpha, beta, out, ldo, num_sms)
    if (k <= 2) SMK_G(2, 1);
    else if (k <= 4) SMK_G(4, 1);
    else if (k <= 8) SMK_G(8, 1);
    else if (k <= 16) SMK_G(16, 1);
    else if (k <= 32) SMK_G(32, 1);
    else if (k <= 64) SMK_G(32, 2);
    else if (k <= 128) SMK_G(32, 4);
    else if (k <= 256) SMK_G(32, 8);
    else throw std::string("spmm: k > 256 is not supported");
#undef SMK_G
}

void spmm_gather_seg(cudaStream_t stream, int ncols, const SegTable& T, const unsigned int* idx, const double* val,
                     int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo,
                     double* partial, int num_sms, int ngather)
{
    if (ncols <= 0 || T.nseg <= 0) return;
    const uintptr_t addr_bits = reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(partial);
    const bool aligned16 = (addr_bits & 15) == 0;
    const bool wide256 = k >= 64 && k <= 256 && (ldb & 3) == 0 && (ldo & 3) == 0 && (addr_bits & 31) == 0 && ngather > 0 && ngather < (1 << 30);
    const bool tiers = T.tiers_on && T.tier_k == k;
    // SMK_SPMM_SLAB: 0 = never, 2 = whenever the shape allows (tests), otherwise by operand size
    const char* slab_env = getenv("SMK_SPMM_SLAB");
    const int slab_mode = slab_env ? atoi(slab_env) : 1;
    // operand larger than L2 can keep, 32-row slab small enough to stay: one launch per slab
    const size_t operand_bytes = static_cast<size_t>(ngather > 0 ? ngather : 0) * k * sizeof(double);
    const bool slab_fits = operand_bytes > kTierMinOperandBytes && static_cast<size_t>(ngather) * kSlab * sizeof(double) <= kSlabMaxBytes;
    if (wide256 && !tiers && (k % kSlab) == 0 && slab_mode != 0 && (slab_fits || slab_mode == 2))
    {
        const int groups_per_block = 256 / 8;
        const int blocks = std::max(1, std::min(ceil_div(T.nseg, groups_per_block), 2 * num_sms));
        for (int koff = 0; koff < k; koff += kSlab)
        {
            spmm_seg_slab_kernel<<<blocks, 256, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, koff, B, ldb,
                                                             alpha, beta, out, ldo, partial);
            SMK_LAUNCH_CHECK();
        }
    }
    else if (wide256 && k >= 96 && (k & 3) == 0)
    {
        const int smem_rows = tiers ? T.tier_smem_rows : 0;
        const size_t smem_bytes = static_cast<size_t>(smem_rows) * k * sizeof(double);
        const unsigned int* use_idx = tiers ? T.tier_idx.p : idx;
        const int blocks = std::max(1, std::min(ceil_div(T.nseg, 16), num_sms));
#define SMK_T(NV, U)                                                                                                                  \
        do {                                                                                                                          \
            static bool attr_set = false;                                                                                             \
            if (!attr_set) { SMK_CUDA(cudaFuncSetAttribute(spmm_seg_tier_kernel<NV, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTierSmemBytes)); attr_set = true; } \
            spmm_seg_tier_kernel<NV, U><<<blocks, 512, smem_bytes, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, use_idx, val, k, B, ldb, \
                                                                             alpha, beta, out, ldo, partial, smem_rows, T.tier_smem_ids.p);      \
        } while (0)
        if (k <= 128) SMK_T(1, 8);
        else SMK_T(2, 4);
#undef SMK_T
        SMK_LAUNCH_CHECK();
    }
    else if (k >= 64 && k <= 256 && (k & 1) == 0 && (ldb & 1) == 0 && (ldo & 1) == 0 && aligned16)
    {
        const int wpb = 8;
        const int blocks = std::max(1, std::min(ceil_div(T.nseg, wpb), 2 * num_sms * 8));
#define SMK_W(NV, U) spmm_seg_wide_kernel<NV, U><<<blocks, 32 * wpb, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, \
                                                                             B, ldb, alpha, beta, out, ldo, partial)
        if (k <= 64) SMK_W(1, 8);
        else if (k <= 128) SMK_W(2, 8);
        else if (k <= 192) SMK_W(3, 4);
        else SMK_W(4, 4);
#undef SMK_W
        SMK_LAUNCH_CHECK();
    }
    else
    {
#define SMK_S(L, KPL) launch_seg<L, KPL>(stream, T, idx, val, k, B, ldb, alpha, beta, out, ldo, partial, num_sms)
    if (k <= 2) SMK_S(2, 1);
    else if (k <= 4) SMK_S(4, 1);
    else if (k <= 8) SMK_S(8, 1);
    else if (k <= 16) SMK_S(16, 1);
    else if (k <= 32) SMK_S(32, 1);
    else if (k <= 64) SMK_S(32, 2);
    else if (k <= 128) SMK_S(32, 4);
    else if (k <= 256) SMK_S(32, 8);
    else throw std::string("spmm: k > 256 is not supported");
#undef SMK_S
    }
    if (T.nmulti > 0)
    {
        const int threads = k <= 32 ? 32 : (k <= 64 ? 64 : 128);
        spmm_combine_kernel<<<std::min(T.nmulti, 16 * num_sms), threads, 0, stream>>>(T.nmulti, T.multi_col.p, T.first_slot.p, k, beta,
                                                                                     partial, out, ldo);
        SMK_LAUNCH_CHECK();
    }
}

// Residency classes for the gathers of one orientation (see spmm_seg_tier_kernel). count = number of gatherable
// vectors, degree_ptr = offsets of the OTHER orientation (degree of vector r = degree_ptr[r+1] - degree_ptr[r]).
// Built only when the operand is much larger than L2 and the degrees are skewed enough for a static choice to beat
// L2's own replacement: the two resident tiers must cover at least twice the share of the stored entries that the same
// number of average vectors would. Synchronises the stream.
void build_gather_tiers(cudaStream_t stream, SegTable& T, int count, const unsigned int* degree_ptr, const unsigned int* idx,
                        unsigned int nnz, int k, int num_sms)
{
    T.tiers_on = false;
    if (k < 96 || k > 256 || (k & 3) || count <= 0 || count >= (1 << 30) || nnz == 0) return;
    // tuning / test knobs: SMK_SPMM_TIERS=0 turns the classes off; the two *_KB values shrink the thresholds so that
    // small test matrices exercise all four classes
    if (const char* e = getenv("SMK_SPMM_TIERS")) if (atoi(e) == 0) return;
    size_t min_operand = kTierMinOperandBytes, keep_bytes = kTierKeepBytes;
    if (const char* e = getenv("SMK_SPMM_TIER_MIN_KB")) min_operand = static_cast<size_t>(atoll(e)) << 10;
    if (const char* e = getenv("SMK_SPMM_TIER_KEEP_KB")) keep_bytes = static_cast<size_t>(atoll(e)) << 10;
    const size_t vec_bytes = static_cast<size_t>(k) * sizeof(double);
    if (static_cast<size_t>(count) * vec_bytes <= min_operand) return;
    const int smem_rows = static_cast<int>(std::min<size_t>(count, kTierSmemBytes / vec_bytes));
    const int keep_rows = static_cast<int>(std::min<size_t>(count - smem_rows, keep_bytes / vec_bytes));
    DevBuf<unsigned int> deg, ids, deg_s, ids_s, code;
    DevBuf<unsigned char> tmp;
    deg.reserve(count); ids.reserve(count); deg_s.reserve(count); ids_s.reserve(count);
    const int blocks = std::max(1, std::min(ceil_div(count, 256), 8 * num_sms));
    tier_degree_kernel<<<blocks, 256, 0, stream>>>(count, degree_ptr, deg.p, ids.p);
    SMK_LAUNCH_CHECK();
    size_t bytes = 0;
    SMK_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, deg.p, deg_s.p, ids.p, ids_s.p, count, 0, 32, stream));
    tmp.reserve(bytes);
    SMK_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, deg.p, deg_s.p, ids.p, ids_s.p, count, 0, 32, stream));
    launch_counter() += 4;
    std::vector<unsigned int> top(static_cast<size_t>(smem_rows) + keep_rows);
    SMK_CUDA(cudaMemcpyAsync(top.data(), deg_s.p, top.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    SMK_CUDA(cudaStreamSynchronize(stream));
    double covered = 0.0;
    for (unsigned int d : top) covered += d;
    const double share = covered / nnz, uniform_share = static_cast<double>(top.size()) / count;
    if (share < 2.0 * uniform_share) return;
    code.reserve(count);
    T.tier_idx.reserve(nnz);
    T.tier_smem_ids.reserve(std::max(1, smem_rows));
    tier_code_kernel<<<blocks, 256, 0, stream>>>(count, ids_s.p, smem_rows, keep_rows, code.p, T.tier_smem_ids.p);
    SMK_LAUNCH_CHECK();
    tier_apply_kernel<<<std::min<unsigned int>((nnz + 255) / 256, 64u * num_sms), 256, 0, stream>>>(nnz, idx, code.p, T.tier_idx.p);
    SMK_LAUNCH_CHECK();
    SMK_CUDA(cudaStreamSynchronize(stream));
    T.tiers_on = true; T.tier_k = k; T.tier_smem_rows = smem_rows; T.tier_share = share;
}

void build_segments(cudaStream_t stream, int ncols, const unsigned int* ptr, SegTable& T, int num_sms)
{
    T.tiers_on = false;                      // residency classes belong to the matrix the table was built for
    const size_t n1 = static_cast<size_t>(ncols) + 1;
    T.t_cnt.reserve(3 * n1);                // cnt | mcnt | flag
    T.t_first.reserve(n1); T.first_slot.reserve(n1); T.t_mpos.reserve(n1);
    unsigned int* cnt = T.t_cnt.p; unsigned int* mcnt = cnt + n1; unsigned int* flag = mcnt + n1;
    const int blocks = std::max(1, std::min(ceil_div(static_cast<long long>(n1), 256), 8 * num_sms));
    seg_count_kernel<<<blocks, 256, 0, stream>>>(ncols, ptr, cnt, mcnt, flag);
    SMK_LAUNCH_CHECK();
    size_t bytes = 0;
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, T.t_first.p, static_cast<int>(n1), stream));
    T.t_scan.reserve(bytes);
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(T.t_scan.p, bytes, cnt, T.t_first.p, static_cast<int>(n1), stream));
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(T.t_scan.p, bytes, mcnt, T.first_slot.p, static_cast<int>(n1), stream));
    SMK_CUDA(cub::DeviceScan::ExclusiveSum(T.t_scan.p, bytes, flag, T.t_mpos.p, static_cast<int>(n1), stream));
    launch_counter() += 6;
    unsigned int h[3] = {0, 0, 0};
    SMK_CUDA(cudaMemcpyAsync(&h[0], T.t_first.p + ncols, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    SMK_CUDA(cudaMemcpyAsync(&h[1], T.first_slot.p + ncols, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    SMK_CUDA(cudaMemcpyAsync(&h[2], T.t_mpos.p + ncols, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    SMK_CUDA(cudaStreamSynchronize(stream));
    T.nseg = static_cast<int>(h[0]); T.nslots = static_cast<int>(h[1]); T.nmulti = static_cast<int>(h[2]);
    T.col.reserve(T.nseg); T.beg.reserve(T.nseg); T.end.reserve(T.nseg); T.slot.reserve(T.nseg);
    T.multi_col.reserve(std::max(1, T.nmulti));
    const int fblocks = std::max(1, std::min(ceil_div(ncols, 8), 8 * num_sms));
    seg_fill_kernel<<<fblocks, 256, 0, stream>>>(ncols, ptr, T.t_first.p, T.first_slot.p, T.t_mpos.p, T.col.p, T.beg.p, T.end.p,
                                                 T.slot.p, T.multi_col.p);
    SMK_LAUNCH_CHECK();
}

// Stable transpose on the device: a stable radix sort of the entry ids by row index keeps, inside each
// row, the column-ascending / storage order of the CSC arrays.
void build_csr(cudaStream_t stream, SparseDev& S, bool keep_scratch)
{
    const unsigned int nnz = S.nnz;
    S.rowptr.reserve(static_cast<size_t>(S.m) + 1);
    S.colidx.reserve(nnz);
    S.valr.reserve(nnz);
    if (nnz == 0)
    {
        SMK_CUDA(cudaMemsetAsync(S.rowptr.p, 0, (static_cast<size_t>(S.m) + 1) * sizeof(unsigned int), stream));
        return;
    }
    S.t_colof.reserve(nnz); S.t_ids.reserve(nnz); S.t_ids_sorted.reserve(nnz); S.t_rows_sorted.reserve(nnz);
    expand_cols_kernel<<<std::min(S.n, 65535), 128, 0, stream>>>(S.n, S.colptr.p, S.t_colof.p);
    SMK_LAUNCH_CHECK();
    iota_kernel<<<std::min<unsigned int>((nnz + 255) / 256, 65535u), 256, 0, stream>>>(nnz, S.t_ids.p);
    SMK_LAUNCH_CHECK();
    int end_bit = 1;
    while ((1ull << end_bit) < static_cast<unsigned long long>(S.m) && end_bit < 32) ++end_bit;
    size_t tmp_bytes = 0;
    SMK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, S.rowidx.p, S.t_rows_sorted.p, S.t_ids.p, S.t_ids_sorted.p,
                                             static_cast<int>(nnz), 0, end_bit, stream));
    S.t_sort.reserve(tmp_bytes);
    SMK_CUDA(cub::DeviceRadixSort::SortPairs(S.t_sort.p, tmp_bytes, S.rowidx.p, S.t_rows_sorted.p, S.t_ids.p, S.t_ids_sorted.p,
                                             static_cast<int>(nnz), 0, end_bit, stream));
    launch_counter() += 4;   // cub's own passes (approximate; not on the iteration path)
    gather_perm_kernel<<<std::min<unsigned int>((nnz + 255) / 256, 65535u), 256, 0, stream>>>(nnz, S.t_ids_sorted.p, S.t_colof.p,
                                                                                              S.val.p, S.colidx.p, S.valr.p);
    SMK_LAUNCH_CHECK();
    rowptr_kernel<<<std::min((S.m + 256) / 256, 65535), 256, 0, stream>>>(S.m, nnz, S.t_rows_sorted.p, S.rowptr.p);
    SMK_LAUNCH_CHECK();
    if (!keep_scratch)
    {
        SMK_CUDA(cudaStreamSynchronize(stream));
        S.t_colof.release(); S.t_ids.release(); S.t_ids_sorted.release(); S.t_rows_sorted.release(); S.t_sort.release();
    }
}

} // namespace smk
