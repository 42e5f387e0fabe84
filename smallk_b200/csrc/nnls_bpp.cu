// smallk_b200 — batched NNLS by block principal pivoting, one warp per column.
//
// Replaces NnlsBlockpivot + BppSolveNormalEqNoGroup + UpdatePassiveSet + the
// BitMatrix passes of the reference
//   common/include/nnls.hpp:144-244, common/include/nmf_solver_bpp.hpp:146-219,
//   common/src/nnls.cpp:18-74, common/src/bit_matrix.cpp:432-472,
//   Elemental cholesky::UVar3Unb + SolveAfter (UVar3.hpp:17-44, SolveAfter.hpp:17-42).
//
// The reference runs all right-hand-side columns in lock step: every pivot round
// gathers the non-optimal columns, solves them, scatters them back and then makes
// serial BitMatrix passes over all q columns. Columns never exchange data, so here
// each warp owns one column for its WHOLE pivoting history: the passive set is a
// 64-bit mask held in a register, the k' x k' normal-equation sub-matrix is
// gathered from the shared-memory copy of the k x k Gram matrix into a packed
// upper triangle in shared memory, factored (right-looking upper Cholesky, same
// recurrence as Elemental's UVar3Unb), solved, and the dual y = LHS*x - rhs is
// formed — all without leaving the kernel. Columns are handed out through an
// atomic counter so that long pivoting histories do not unbalance the SMs.
//
// The two places where the reference couples columns are preserved:
//   * ZeroizeSmallValues(X|Y, 1e-12) is applied to EVERY column iff at least one
//     column was non-optimal after the initial solve (nnls.hpp:192,226-227):
//     status[ST_ANY_NONOPT] is raised here and zeroize_if_flag_kernel finishes
//     the columns that never entered the pivot loop.
//   * MAX_ITER = 5k pivot rounds, and a non-positive Cholesky pivot, fail the whole
//     solve (nnls.hpp:195-196, normal_eq.hpp:35-50): status[ST_FAIL_ITER] records
//     the outer iteration in which that happened.
// BitMatrix::MaxRowIndex's off-by-one-word defect for k > 32
// is reproduced in max_row_index_ref() because the backup pivot rule depends on it.
#include "common.cuh"
#include "kernels.h"

namespace smk {

namespace {

constexpr double kZeroThresh = 1.0e-12;    // nnls.hpp:215,226-227
constexpr int kPbar = 3;                   // nnls.hpp:153

__device__ __forceinline__ int tri(int c) { return (c * (c + 1)) >> 1; }

// BitMatrix::MaxRowIndex as the reference computes it (defect included).
__device__ __forceinline__ int max_row_index_ref(unsigned long long mask, int k)
{
    if (mask == 0ull) return 0;
    const int h = 63 - __clzll(static_cast<long long>(mask));
    const int full = k >> 5, extra = k & 31;
    const int w = h >> 5;
    if (extra > 0 && w == full) return h;
    return (w > 0) ? h - 32 : h;
}

struct WarpScratch
{
    double* U;     // packed upper triangle, k(k+1)/2
    double* vb;    // k   sub-vector rhs / solution
    double* sb;    // k   full rhs column
    double* sx;    // k   full x column
    int* ri;       // k   passive row list
};

// Solve LHS[P,P] x_P = b_P for the passive set `pm` (p = popcount rows listed in ri).
// Returns false on a non-positive pivot. On return vb[0..p) holds x_P.
__device__ __forceinline__ bool warp_spd_solve(const double* __restrict__ sL, int k, int p,
                                               const WarpScratch& w, int lane)
{
    double* U = w.U;
    double* vb = w.vb;
    const int* ri = w.ri;
    // gather
    for (int c = 0; c < p; ++c)
    {
        const int rc = ri[c] * k;
        const int base = tri(c);
        for (int i = lane; i <= c; i += 32) U[base + i] = sL[rc + ri[i]];
    }
    for (int i = lane; i < p; i += 32) vb[i] = w.sb[ri[i]];
    __syncwarp();

    // upper Cholesky, right-looking (UVar3Unb)
    for (int j = 0; j < p; ++j)
    {
        const double ajj = U[tri(j) + j];
        if (!(ajj > 0.0)) return false;
        const double d = sqrt(ajj);
        __syncwarp();
        if (lane == 0) U[tri(j) + j] = d;
        for (int c = j + 1 + lane; c < p; c += 32) U[tri(c) + j] /= d;
        __syncwarp();
        for (int c = j + 1 + lane; c < p; c += 32)
        {
            const int base = tri(c);
            const double ujc = U[base + j];
            for (int i = j + 1; i <= c; ++i) U[base + i] -= U[tri(i) + j] * ujc;
        }
        __syncwarp();
    }
    // U' y = b  (forward, column-update form: same per-entry subtraction order as a dot-form solve)
    for (int i = 0; i < p; ++i)
    {
        const double yi = vb[i] / U[tri(i) + i];
        __syncwarp();
        if (lane == 0) vb[i] = yi;
        for (int c = i + 1 + lane; c < p; c += 32) vb[c] -= U[tri(c) + i] * yi;
        __syncwarp();
    }
    // U x = y  (backward)
    for (int c = p - 1; c >= 0; --c)
    {
        const int base = tri(c);
        const double xc = vb[c] / U[base + c];
        __syncwarp();
        if (lane == 0) vb[c] = xc;
        for (int i = lane; i < c; i += 32) vb[i] -= U[base + i] * xc;
        __syncwarp();
    }
    return true;
}

// k <= 64. LHS is k x k (ld = ldl), RHS/X/Y are k x q.
__global__ void nnls_bpp_warp_kernel(int k, int q,
                                     const double* __restrict__ LHS, long long ldl,
                                     const double* __restrict__ RHS, long long ldr,
                                     double* __restrict__ X, long long ldx,
                                     double* __restrict__ Y, long long ldy,
                                     int* __restrict__ status, unsigned int* __restrict__ counter, int outer_iter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sL = reinterpret_cast<double*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int tri_k = (k * (k + 1)) >> 1;
    const size_t per_warp_doubles = static_cast<size_t>(tri_k) + 3 * k;
    double* wbase = sL + static_cast<size_t>(k) * k + warp * per_warp_doubles;
    WarpScratch w;
    w.U = wbase; w.vb = wbase + tri_k; w.sb = w.vb + k; w.sx = w.sb + k;
    w.ri = reinterpret_cast<int*>(sL + static_cast<size_t>(k) * k + nwarps * per_warp_doubles) + warp * k;

    for (int i = threadIdx.x; i < k * k; i += blockDim.x) sL[i] = LHS[static_cast<long long>(i / k) * ldl + (i % k)];
    __syncthreads();

    const int r0 = lane, r1 = lane + 32;
    const bool v0 = r0 < k, v1 = r1 < k;
    const int max_rounds = 5 * k;

    for (;;)
    {
        unsigned int c = 0;
        if (lane == 0) c = atomicAdd(counter, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= static_cast<unsigned int>(q)) break;

        const double* rhs = RHS + static_cast<long long>(c) * ldr;
        double* xcol = X + static_cast<long long>(c) * ldx;
        double* ycol = Y + static_cast<long long>(c) * ldy;

        const double b0 = v0 ? rhs[r0] : 0.0, b1 = v1 ? rhs[r1] : 0.0;
        if (v0) w.sb[r0] = b0;
        if (v1) w.sb[r1] = b1;
        // warm start: passive = (X > 0)   (nnls.hpp:157)
        unsigned long long pm;
        {
            const bool p0 = v0 && (xcol[r0] > 0.0), p1 = v1 && (xcol[r1] > 0.0);
            pm = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, p0)) |
                 (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, p1)) << 32);
        }
        int P = kPbar, Ninf = k + 1, round = 0;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
        bool failed = false;

        for (;;)
        {
            const int p = __popcll(pm);
            // passive row list, ascending
            if (v0 && ((pm >> r0) & 1ull)) w.ri[__popcll(pm & ((1ull << r0) - 1ull))] = r0;
            if (v1 && ((pm >> r1) & 1ull)) w.ri[__popcll(pm & ((1ull << r1) - 1ull))] = r1;
            if (v0) w.sx[r0] = 0.0;
            if (v1) w.sx[r1] = 0.0;
            __syncwarp();
            if (p > 0)
            {
                if (!warp_spd_solve(sL, k, p, w, lane)) { failed = true; break; }
                // ZeroizeSmallValues(Xsub) applies inside the pivot loop only (nnls.hpp:215)
                for (int i = lane; i < p; i += 32)
                {
                    double xv = w.vb[i];
                    if (round > 0 && fabs(xv) < kZeroThresh) xv = 0.0;
                    w.vb[i] = xv;
                    w.sx[w.ri[i]] = xv;
                }
                __syncwarp();
            }
            x0 = v0 ? w.sx[r0] : 0.0;
            x1 = v1 ? w.sx[r1] : 0.0;
            // y = LHS * x - rhs   (nnls.hpp:168-169, 219-220)
            double s0 = 0.0, s1 = 0.0;
            for (int t = 0; t < p; ++t)
            {
                const double xv = w.vb[t];
                const double* col = sL + w.ri[t] * k;
                if (v0) s0 += col[r0] * xv;
                if (v1) s1 += col[r1] * xv;
            }
            y0 = s0 - b0; y1 = s1 - b1;
            if (round > 0)
            {
                if (fabs(y0) < kZeroThresh) y0 = 0.0;
                if (fabs(y1) < kZeroThresh) y1 = 0.0;
            }
            const bool in0 = (pm >> r0) & 1ull, in1 = v1 && ((pm >> r1) & 1ull);
            const unsigned long long nonopt =
                static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v0 && !in0 && y0 < 0.0)) |
                (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v1 && !in1 && y1 < 0.0)) << 32);
            const unsigned long long infeas =
                static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v0 && in0 && x0 < 0.0)) |
                (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v1 && in1 && x1 < 0.0)) << 32);
            const int not_good = __popcll(nonopt) + __popcll(infeas);
            __syncwarp();
            if (not_good == 0) break;
            if (round == 0 && lane == 0) atomicOr(&status[ST_ANY_NONOPT], 1);
            if (round >= max_rounds) { failed = true; break; }     // nnls.hpp:195-196
            // UpdatePassiveSet (common/src/nnls.cpp:18-74)
            if (not_good < Ninf)
            {
                P = kPbar; Ninf = not_good;
                pm = (pm | nonopt) & ~infeas;
            }
            else if (P >= 1)
            {
                P -= 1;
                pm = (pm | nonopt) & ~infeas;
            }
            else
            {
                const int ra = max_row_index_ref(nonopt, k), rb = max_row_index_ref(infeas, k);
                pm ^= (1ull << (ra > rb ? ra : rb));
            }
            ++round;
        }
        if (failed && lane == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        if (v0) { xcol[r0] = x0; ycol[r0] = y0; }
        if (v1) { xcol[r1] = x1; ycol[r1] = y1; }
        __syncwarp();
    }
}

// Finishes ZeroizeSmallValues(X), ZeroizeSmallValues(Y) for columns that never
// entered the pivot loop; a no-op unless some column was non-optimal.
__global__ void zeroize_if_flag_kernel(const int* __restrict__ status, int k, long long q,
                                       double* __restrict__ X, long long ldx, double* __restrict__ Y, long long ldy)
{
    if (status[ST_ANY_NONOPT] == 0) return;
    const long long total = static_cast<long long>(k) * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const long long c = i / k;
        const int r = static_cast<int>(i % k);
        double* x = X + c * ldx + r;
        double* y = Y + c * ldy + r;
        if (fabs(*x) < kZeroThresh) *x = 0.0;
        if (fabs(*y) < kZeroThresh) *y = 0.0;
    }
}

} // namespace

// status: device int[ST_COUNT]; counter: device unsigned. Both are (re)initialised here.
void nnls_bpp(cudaStream_t stream, int k, int q, const double* LHS, long long ldl,
              const double* RHS, long long ldr, double* X, long long ldx, double* Y, long long ldy,
              int* status, unsigned int* counter, int outer_iter, int num_sms)
{
    if (k > 64) throw std::string("nnls_bpp: k > 64 is not supported by the warp kernel yet");
    if (q <= 0) return;
    SMK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    SMK_CUDA(cudaMemsetAsync(&status[ST_ANY_NONOPT], 0, sizeof(int), stream));

    const size_t tri_k = static_cast<size_t>(k) * (k + 1) / 2;
    const size_t per_warp = (tri_k + 3 * k) * sizeof(double) + k * sizeof(int);
    const size_t fixed = static_cast<size_t>(k) * k * sizeof(double);
    const size_t budget = 220 * 1024;
    int warps = static_cast<int>((budget - fixed) / per_warp);
    warps = std::max(1, std::min(warps, 16));
    const size_t smem = fixed + warps * per_warp + 16;
    int blocks_per_sm = std::max(1, std::min<int>(static_cast<int>(budget / smem), 32 / warps > 0 ? 64 / warps : 1));
    int grid = std::min(num_sms * blocks_per_sm, ceil_div(q, warps));
    SMK_CUDA(cudaFuncSetAttribute(nnls_bpp_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    nnls_bpp_warp_kernel<<<grid, warps * 32, smem, stream>>>(k, q, LHS, ldl, RHS, ldr, X, ldx, Y, ldy, status, counter, outer_iter);
    SMK_LAUNCH_CHECK();
    const long long total = static_cast<long long>(k) * q;
    int zb = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * num_sms));
    zeroize_if_flag_kernel<<<zb, 256, 0, stream>>>(status, k, q, X, ldx, Y, ldy);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
