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
// 256-bit gather kernels (k % 4 == 0, 64 <= k <= 256, 32-byte aligned operands): lane l owns the four consecutive doubles
// 4l..4l+3 (+128v) of the k-vector, so one gathered operand is NV LDG.E.256 per lane. Two shapes:
//   spmm_seg_wide256_kernel: a warp per segment, the whole k-vector per gather (operands that fit L2, or whose popular vectors
//       do: a gather SpMM moves k*8 bytes of dense operand per stored entry — C3: 1 KB x 1e8 = 102 GB per product, 40x the
//       compulsory HBM bytes — so what bounds it is the rate at which L2 delivers gathered vectors to the SMs,
//       tools/l2_gather_peak.cu: 19-20 TB/s for an L2-resident table on B200);
//   spmm_seg_slab_kernel: an operand larger than L2 whose 32-row slab fits (C3: H = 205 MB, slab 51 MB) is gathered one
//       slab per launch, so every gather after the first touch of a vector is an L2 hit; A's index / value arrays are
//       streamed once per slab (12 bytes per entry against the 256 bytes the entry gathers). 8 lanes own a segment, four
//       segments per warp.
// Entries are added in storage order, one fused multiply-add per entry and output element (the reference's order).
// ---------------------------------------------------------------------------
struct d4 { double x, y, z, w; };

constexpr size_t kSlabMinOperandBytes = size_t(96) << 20;     // smaller operands live in L2 whole
constexpr size_t kSlabMaxBytes = size_t(64) << 20;            // a 32-row operand slab of at most this size is gathered from L2
constexpr int kSlab = 32;

// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// How the loops are written, and why (the round-1 forms of these kernels ran at 60 % of this rate; their ncu source view,
// profiles/ncu_r02_c3_spmm_source.txt, showed 69 % of the stall samples on warps waiting, IN ORDER, for a load whose value a
// select or multiply consumed right behind it — `nxt_a = alpha * val[...]` exposed the DRAM latency of the index stream in
// every batch, `b[t] = (t < cnt) ? load : 0` made each gather wait for the previous one — plus a serial tail loop and 270
// instructions per 8 entries for residency-class branches and 64-bit address arithmetic):
//   * every load is unconditional, from a clamped (always valid) address, and nothing touches a loaded register before the
//     multiply-adds: an entry past the end of a segment is a copy of the segment's last entry with weight 0 (the weight is
//     zeroed when the batch is consumed, two batches after its load), so its FMA adds +-0;
//   * the (index, value) stream runs two batches ahead, in two register pairs used alternately (no loop-carried move that
//     would wait for the load just issued);
//   * one code path for full and partial batches; no residency classes (the shared-memory tier bought 3 % and cost a
//     branch per gather; L2's own replacement keeps the popular vectors); addresses are one 32 x 32 -> 64 bit multiply-add.
// ---------------------------------------------------------------------------
__device__ __forceinline__ d4 ldg256_stream(const char* p)
{
    d4 r;
    // volatile: keeps the shuffle -> address -> load sequence of one entry together (ptxas otherwise hoists all the address
    // arithmetic of a group above its first load and runs out of registers)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
template <int NV, int U>
__global__ void __launch_bounds__(512, 1)
spmm_seg_wide256_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                      const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                      const unsigned int* __restrict__ idx, const double* __restrict__ val, int k,
                      const double* __restrict__ B, unsigned int ldb_bytes, double alpha, double beta,
                      double* __restrict__ out, long long ldo, double* __restrict__ partial)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    bool live[NV];
    const char* Bl[NV];          // this lane's piece of vector 0 (dead lanes of a ragged k point at a valid piece; their sums are not stored)
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
        live[v] = 4 * lane + 128 * v < k;
        Bl[v] = reinterpret_cast<const char*>(B + (live[v] ? 4 * lane + 128 * v : 0));
    }
    const long long stride = static_cast<long long>(gridDim.x) * wpb;
    long long it = blockIdx.x * static_cast<long long>(wpb) + warp;
    // record of the next segment, read while the current one is summed
    unsigned int nj = 0, nslot = 0, nbeg = 0, nend = 0;
    if (it < nseg) { nj = scol[it]; nslot = sslot[it]; nbeg = sbeg[it]; nend = send[it]; }
    for (; it < nseg; it += stride)
    {
        const unsigned int j = nj, slot = nslot, end = nend;
        unsigned int o = nbeg;
        if (it + stride < nseg) { nj = scol[it + stride]; nslot = sslot[it + stride]; nbeg = sbeg[it + stride]; nend = send[it + stride]; }
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
        if (o < end)
        {
            const unsigned int last = end - 1;
            unsigned int i0, i1; double v0, v1;
            { const unsigned int p = min(o + lane, last); i0 = __ldcs(idx + p); v0 = __ldcs(val + p); }
            { const unsigned int p = min(o + 32 + lane, last); i1 = __ldcs(idx + p); v1 = __ldcs(val + p); }
            auto batch = [&](unsigned int& ci, double& cv) {
                const unsigned int left = end - o;                       // >= 1 entries of this segment from o on
                const unsigned int my_i = ci;
                const double my_v = static_cast<unsigned int>(lane) < left ? alpha * cv : 0.0;    // past the end: a copy of the last entry, weight 0
                { const unsigned int p = min(o + 64 + lane, last); ci = __ldcs(idx + p); cv = __ldcs(val + p); }   // two batches ahead
                const int ngr = static_cast<int>((min(left, 32u) + U - 1) / U);
                for (int gq = 0; gq < ngr; ++gq)
                {
                    d4 b[U][NV];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        const unsigned int id = __shfl_sync(0xffffffffu, my_i, gq * U + u);
#pragma unroll
                        for (int v = 0; v < NV; ++v) b[u][v] = ldg256_stream(Bl[v] + static_cast<unsigned long long>(id) * ldb_bytes);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        const double au = __shfl_sync(0xffffffffu, my_v, gq * U + u);
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                        {
                            acc[v].x += au * b[u][v].x; acc[v].y += au * b[u][v].y;
                            acc[v].z += au * b[u][v].z; acc[v].w += au * b[u][v].w;
                        }
                    }
                }
                o += 32;
            };
            for (;;)
            {
                batch(i0, v0); if (o >= end) break;
                batch(i1, v1); if (o >= end) break;
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

template <int U, int MINB>      // U gathers in flight per lane (8 / U sub-batches per batch of 8 entries), MINB CTAs per SM
__global__ void __launch_bounds__(256, MINB)
spmm_seg_slab_kernel(int nseg, const unsigned int* __restrict__ scol, const unsigned int* __restrict__ sbeg,
                      const unsigned int* __restrict__ send, const unsigned int* __restrict__ sslot,
                      const unsigned int* __restrict__ idx, const double* __restrict__ val, int k, int koff,
                      const double* __restrict__ B, unsigned int ldb_bytes, double alpha, double beta,
                      double* __restrict__ out, long long ldo, double* __restrict__ partial)
{
    const int lane = threadIdx.x & 31, g = lane & 7;
    const unsigned int mask = 0xFFu << (lane & ~7);
    const long long ngroups = static_cast<long long>(gridDim.x) * (blockDim.x >> 3);
    const int off = koff + 4 * g;
    const char* Bl = reinterpret_cast<const char*>(B + off);
    long long it = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
    unsigned int nj = 0, nslot = 0, nbeg = 0, nend = 0;
    if (it < nseg) { nj = scol[it]; nslot = sslot[it]; nbeg = sbeg[it]; nend = send[it]; }
    for (; it < nseg; it += ngroups)
    {
        const unsigned int j = nj, slot = nslot, end = nend;
        unsigned int o = nbeg;
        if (it + ngroups < nseg) { nj = scol[it + ngroups]; nslot = sslot[it + ngroups]; nbeg = sbeg[it + ngroups]; nend = send[it + ngroups]; }
        const bool direct = slot == 0xFFFFFFFFu;
        d4 acc; acc.x = acc.y = acc.z = acc.w = 0.0;
        if (direct && beta != 0.0)
        {
            const double* c0 = out + j * ldo + off;
            const double2 lo = *reinterpret_cast<const double2*>(c0), hi = *reinterpret_cast<const double2*>(c0 + 2);
            acc.x = lo.x * beta; acc.y = lo.y * beta; acc.z = hi.x * beta; acc.w = hi.y * beta;
        }
        if (o < end)
        {
            const unsigned int last = end - 1;
            unsigned int i0, i1; double v0, v1;
            { const unsigned int p = min(o + g, last); i0 = __ldcs(idx + p); v0 = __ldcs(val + p); }
            { const unsigned int p = min(o + 8 + g, last); i1 = __ldcs(idx + p); v1 = __ldcs(val + p); }
            auto batch = [&](unsigned int& ci, double& cv) {
                const unsigned int left = end - o;
                const unsigned int my_i = ci;
                const double my_v = static_cast<unsigned int>(g) < left ? alpha * cv : 0.0;      // past the end: a copy of the last entry, weight 0
                { const unsigned int p = min(o + 16 + g, last); ci = __ldcs(idx + p); cv = __ldcs(val + p); }      // two batches ahead
#pragma unroll
                for (int h = 0; h < 8; h += U)
                {
                    if (U < 8 && static_cast<unsigned int>(h) >= left) break;      // group-uniform
                    d4 b[U];
#pragma unroll
                    for (int t = 0; t < U; ++t)
                    {
                        const unsigned int iu = __shfl_sync(mask, my_i, h + t, 8);
                        b[t] = ldg256_stream(Bl + static_cast<unsigned long long>(iu) * ldb_bytes);
                    }
#pragma unroll
                    for (int t = 0; t < U; ++t)
                    {
                        const double at = __shfl_sync(mask, my_v, h + t, 8);
                        acc.x += at * b[t].x; acc.y += at * b[t].y; acc.z += at * b[t].z; acc.w += at * b[t].w;
                    }
                }
                o += 8;
            };
            for (;;)
            {
                batch(i0, v0); if (o >= end) break;
                batch(i1, v1); if (o >= end) break;
            }
        }
        double* p = (direct ? out + j * ldo : partial + static_cast<long long>(slot) * k) + off;
        *reinterpret_cast<double2*>(p) = make_double2(acc.x, acc.y);
        *reinterpret_cast<double2*>(p + 2) = make_double2(acc.z, acc.w);
    }
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
    if (k > kSpmmMaxK)
    {
        // any k: the operand in row blocks of kSpmmMaxK (a block of rows of a k x * column-major matrix is the same matrix with
        // another base pointer); entries are still added in storage order, so nothing changes numerically
        for (int k0 = 0; k0 < k; k0 += kSpmmMaxK)
            spmm_gather(stream, ncols, ptr, idx, val, std::min(kSpmmMaxK, k - k0), B + k0, ldb, alpha, beta, out + k0, ldo, num_sms);
        return;
    }
#define SMK_G(L, KPL) launch_gather<L, KPL>(stream, ncols, ptr, idx, val, k, B, ldb, alpha, beta, out, ldo, num_sms)
    if (k <= 2) SMK_G(2, 1);
    else if (k <= 4) SMK_G(4, 1);
    else if (k <= 8) SMK_G(8, 1);
    else if (k <= 16) SMK_G(16, 1);
    else if (k <= 32) SMK_G(32, 1);
    else if (k <= 64) SMK_G(32, 2);
    else if (k <= 128) SMK_G(32, 4);
    else if (k <= 256) SMK_G(32, 8);
    else throw std::string("spmm: internal error (k > kSpmmMaxK reached a register kernel)");
#undef SMK_G
}

void spmm_gather_seg(cudaStream_t stream, int ncols, const SegTable& T, const unsigned int* idx, const double* val,
                     int k, const double* B, long long ldb, double alpha, double beta, double* out, long long ldo,
                     double* partial, int num_sms, int ngather)
{
    if (ncols <= 0 || T.nseg <= 0) return;
    if (k > kSpmmMaxK)
    {
        // as in spmm_gather; every row block runs to completion (per-segment partials included) before the next re-uses `partial`
        for (int k0 = 0; k0 < k; k0 += kSpmmMaxK)
            spmm_gather_seg(stream, ncols, T, idx, val, std::min(kSpmmMaxK, k - k0), B + k0, ldb, alpha, beta, out + k0, ldo, partial,
                            num_sms, ngather);
        return;
    }
    const uintptr_t addr_bits = reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(partial);
    const bool aligned16 = (addr_bits & 15) == 0;
    const bool wide256 = k >= 64 && k <= 256 && (ldb & 3) == 0 && (ldo & 3) == 0 && (addr_bits & 31) == 0 && ngather > 0 &&
                         ldb > 0 && ldb < (1LL << 28);            // byte offsets of a gathered vector: one 32 x 32 -> 64 bit multiply
    // SMK_SPMM_SLAB: 0 = never, 2 = whenever the shape allows (tests), otherwise by operand size
    const char* slab_env = getenv("SMK_SPMM_SLAB");
    const int slab_mode = slab_env ? atoi(slab_env) : 1;
    // operand larger than L2 can keep, 32-row slab small enough to stay: one launch per slab
    const size_t operand_bytes = static_cast<size_t>(ngather > 0 ? ngather : 0) * k * sizeof(double);
    const bool slab_fits = operand_bytes > kSlabMinOperandBytes && static_cast<size_t>(ngather) * kSlab * sizeof(double) <= kSlabMaxBytes;
    const unsigned int ldb_bytes = static_cast<unsigned int>(ldb * 8);
    if (wide256 && (k % kSlab) == 0 && slab_mode != 0 && (slab_fits || slab_mode == 2))
    {
        // 32 warps per SM with 4 gathers in flight per lane (C3 H*A': 6.27 ms against 6.50 ms for 16 warps x 8, r02)
        const int groups_per_block = 256 / 8;
        const int blocks = std::max(1, std::min(ceil_div(T.nseg, groups_per_block), 4 * num_sms));
        for (int koff = 0; koff < k; koff += kSlab)
        {
            spmm_seg_slab_kernel<4, 4><<<blocks, 256, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, koff, B,
                                                                   ldb_bytes, alpha, beta, out, ldo, partial);
            SMK_LAUNCH_CHECK();
        }
    }
    else if (wide256 && k >= 96 && (k & 3) == 0)
    {
        // one 512-thread CTA per SM, 8 (k <= 128) or 4 x 2 gathered pieces in flight per lane: 128 KB of gathers in flight per SM
        const int blocks = std::max(1, std::min(ceil_div(T.nseg, 16), num_sms));
        if (k <= 128)
            spmm_seg_wide256_kernel<1, 8><<<blocks, 512, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, B, ldb_bytes,
                                                                      alpha, beta, out, ldo, partial);
        else
            spmm_seg_wide256_kernel<2, 4><<<blocks, 512, 0, stream>>>(T.nseg, T.col.p, T.beg.p, T.end.p, T.slot.p, idx, val, k, B, ldb_bytes,
                                                                      alpha, beta, out, ldo, partial);
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
    else throw std::string("spmm: internal error (k > kSpmmMaxK reached a register kernel)");
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

void build_segments(cudaStream_t stream, int ncols, const unsigned int* ptr, SegTable& T, int num_sms)
{
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
