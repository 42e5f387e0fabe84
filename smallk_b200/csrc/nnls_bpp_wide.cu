// smallk_b200 — batched NNLS by block principal pivoting for 64 < k <= 256: one CTA per right-hand side.
//
// Same contract and reference behaviours as nnls_bpp.cu (NnlsBlockpivot, nnls.hpp:144-244; BppSolveNormalEqNoGroup,
// nmf_solver_bpp.hpp:146-219; UpdatePassiveSet, nnls.cpp:18-74; MaxRowIndex defect, bit_matrix.cpp:432-472), for
// the ranks where a passive-set system no longer fits a warp's registers (flatclust BPP at k = 256).
//
// A CTA of 256 threads owns one column for its whole pivoting history; thread t owns row t. The passive set is a
// 256-bit mask (eight ballot words in shared memory). Each pivot round solves G_PP x_P = b_P by an upper Cholesky
// factorization held as a PACKED triangle in shared memory (<= 128 x 128: 66 KB, three CTAs per SM):
//     |P| <= 128 : direct, G_PP gathered from the L2-resident G;
//     |P| >  128 : on the complement A = ~P (|A| < 128) with the block-inverse identity
//                  u = G^-1 b,  z = (G^-1)_AA^-1 u_A,  x_P = u_P - (G^-1)_PA z,
//                  G^-1 formed once per call by invert_spd_kernel. The true residual (G x - b)_P is the acceptance
//                  test; a column that fails it (ill-conditioned G), or any solve when G^-1 could not be formed,
//                  falls back to the direct method with the packed triangle in a per-CTA global scratch (L2).
// The dual y = G x - b is always formed as the full product, as the reference does (nnls.hpp:168-169, 219-220).
#include <cstdlib>
#include "nnls_common.cuh"

namespace smk {

namespace {

constexpr int kWideThreads = 256;
constexpr int kWideNmax = 128;

__device__ __forceinline__ int tri(int c) { return (c * (c + 1)) >> 1; }

// In-place Gauss-Jordan inversion of the SPD k x k matrix in global memory (one CTA). ok[0] = 1 on success.
__global__ void __launch_bounds__(1024, 1)
invert_spd_kernel(int k, const double* __restrict__ G, long long ldg, double* __restrict__ Ginv, int* __restrict__ ok)
{
    __shared__ double scol[256], srow[256];
    __shared__ int s_ok;
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) Ginv[e] = G[static_cast<long long>(e / k) * ldg + (e % k)];
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (int j = 0; j < k; ++j)
    {
        const double piv = Ginv[j + j * k];
        if (!(piv > 0.0)) { if (threadIdx.x == 0) s_ok = 0; break; }     // uniform
        const double ip = 1.0 / piv;
        for (int i = threadIdx.x; i < k; i += blockDim.x) { scol[i] = Ginv[i + j * k]; srow[i] = Ginv[j + i * k] * ip; }
        __syncthreads();
        for (int e = threadIdx.x; e < k * k; e += blockDim.x)
        {
            const int i = e % k, c = e / k;
            double v;
            if (i == j) v = (c == j) ? ip : srow[c];
            else if (c == j) v = -scol[i] * ip;
            else v = Ginv[e] - scol[i] * srow[c];
            Ginv[e] = v;
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) ok[0] = s_ok;
}

// Solves (U'U) x = vb for the SPD matrix whose upper triangle is packed by columns in M (element (i,c), i <= c, at
// tri(c) + i); x overwrites vb. Whole CTA; M may live in shared or global memory. Right-looking factorization with
// the recurrence of Elemental's UVar3Unb (sqrt, divide), then both triangular solves by warp 0.
// Returns false (uniformly) on a non-positive pivot.
__device__ bool cta_spd_solve_packed(double* M, double* __restrict__ vb, int n, double* __restrict__ rowj)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int j = 0; j < n; ++j)
    {
        const double ajj = M[tri(j) + j];
        if (!(ajj > 0.0)) return false;
        const double d = sqrt(ajj);
        __syncthreads();                                   // everyone has read the pivot
        for (int c = j + tid; c < n; c += blockDim.x)
        {
            const double v = (c == j) ? d : M[tri(c) + j] / d;
            M[tri(c) + j] = v;
            rowj[c] = v;
        }
        __syncthreads();
        for (int c = j + 1 + warp; c < n; c += nwarps)
        {
            const double ujc = rowj[c];
            double* col = M + tri(c);
            for (int i = j + 1 + lane; i <= c; i += 32) col[i] = fma(-rowj[i], ujc, col[i]);
        }
        __syncthreads();
    }
    if (warp == 0)
    {
        // U'y = b, right-looking over rows
        for (int i = 0; i < n; ++i)
        {
            const double yi = vb[i] / M[tri(i) + i];
            __syncwarp();
            if (lane == 0) vb[i] = yi;
            for (int c = i + 1 + lane; c < n; c += 32) vb[c] = fma(-M[tri(c) + i], yi, vb[c]);
            __syncwarp();
        }
        // U x = y, right-looking over columns (column c of the packed triangle is contiguous)
        for (int c = n - 1; c >= 0; --c)
        {
            const double* col = M + tri(c);
            const double xc = vb[c] / col[c];
            __syncwarp();
            if (lane == 0) vb[c] = xc;
            for (int i = lane; i < c; i += 32) vb[i] = fma(-col[i], xc, vb[i]);
            __syncwarp();
        }
    }
    __syncthreads();
    return true;
}

// The same solve, blocked by panels of kPB columns with look-ahead, for a triangle in SHARED memory (n <= 128).
//   factor(p)   ONE warp factors the kPB x kPB diagonal block of panel p in registers (a column per lane, rows handed round
//               by shuffles) and publishes it (U above the diagonal, 1 / U(j,j) ON the diagonal) in sD[p & 1];
//   panel(p)    every trailing column — and the right-hand side, which rides along as one more column, so the forward solve
//               costs nothing extra — gets its kPB panel entries by one thread (a kPB-step substitution against sD);
//   update(p)   the kPB rank-1 updates of the trailing columns in one pass; warp 0 takes the kPB columns of panel p + 1 first
//               and then runs factor(p + 1) while the other warps update the rest: the serial factorization of the next
//               diagonal block hides under the bulk of the update (r02 ncu of the first blocked version: 45 % of the stall
//               samples were the other seven warps waiting for factor() and for the diagonal back-solve).
// The backward solve goes panel by panel as well. 2 barriers per kPB columns instead of 3 per column + two triangular solves
// on one warp. Pivots use rsqrt + multiply (as nnls_bpp.cu: 1-2 ulp from the sqrt / divide form; parity is held to 1e-9, not
// bitwise) and the packed triangle keeps 1 / U(j,j) on its diagonal. sD: scratch of 2 * 72 doubles.
constexpr int kPB = 8;
constexpr int kPBS = kPB * kPB + 8;      // one published diagonal block: kPB x kPB + flag

// warp-collective: factor the diagonal block at j0 (nb columns) of M, write it back and to D; D[kPB * kPB] = 1 ok / 0 not SPD
__device__ __forceinline__ void warp_factor_diag(double* M, int j0, int nb, double* __restrict__ D, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    double a[kPB];
    const int cl = j0 + min(lane, nb - 1);
#pragma unroll
    for (int i = 0; i < kPB; ++i) a[i] = (lane < nb && i <= lane) ? M[tri(cl) + j0 + i] : ((i == lane) ? 1.0 : 0.0);
    bool ok = true;
#pragma unroll
    for (int jj = 0; jj < kPB; ++jj)
    {
        if (jj < nb && ok)
        {
            const double ajj = __shfl_sync(FULL, a[jj], jj);
            if (!(ajj > 0.0)) ok = false;                       // uniform: every lane sees the same pivot
            else
            {
                const double r = rsqrt(ajj);
                const double ujc = (lane == jj) ? r : a[jj] * r;    // U(jj, lane) for lane > jj; the reciprocal pivot on the diagonal
                a[jj] = ujc;
#pragma unroll
                for (int i = jj + 1; i < kPB; ++i)
                {
                    const double uji = __shfl_sync(FULL, ujc, i);   // U(jj, i) lives on lane i
                    a[i] = fma(-uji, ujc, a[i]);                    // meaningful for jj < i <= lane
                }
            }
        }
    }
    if (lane < nb)
    {
#pragma unroll
        for (int i = 0; i < kPB; ++i)
            if (i <= lane) { M[tri(j0 + lane) + j0 + i] = a[i]; D[i * kPB + lane] = a[i]; }
    }
    if (lane == 0) D[kPB * kPB] = ok ? 1.0 : 0.0;
}

// rows (rows i in [ibeg, iend]) of column `col` -= sum_t U(j0 + t, i) * uc[t]   (nb pivots), lanes over rows
__device__ __forceinline__ void warp_update_column(const double* M, double* col, int j0, int nb, const double (&uc)[kPB], int ibeg, int iend, int lane)
{
    if (nb == kPB)
    {
        for (int i = ibeg + lane; i <= iend; i += 32)
        {
            const double* ui = M + tri(i) + j0;                 // U(j0 + t, i): contiguous in t; conflict-free across lanes
            double v = col[i];
#pragma unroll
            for (int t = 0; t < kPB; ++t) v = fma(-ui[t], uc[t], v);
            col[i] = v;
        }
    }
    else
    {
        for (int i = ibeg + lane; i <= iend; i += 32)
        {
            const double* ui = M + tri(i) + j0;
            double v = col[i];
            for (int t = 0; t < nb; ++t) v = fma(-ui[t], uc[t], v);
            col[i] = v;
        }
    }
}

__device__ bool cta_spd_solve_blocked(double* M, double* __restrict__ vb, int n, double* __restrict__ sD)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (warp == 0) warp_factor_diag(M, 0, min(kPB, n), sD, lane);
    __syncthreads();
    int pidx = 0;
    for (int j0 = 0; j0 < n; j0 += kPB, pidx ^= 1)
    {
        const int nb = min(kPB, n - j0);
        const double* D = sD + pidx * kPBS;
        if (D[kPB * kPB] == 0.0) return false;                          // uniform
        // ---- panel entries of the trailing columns and of the right-hand side: U_D' u = a, one thread per column
        const int ntrail = n - j0 - nb;
        if (tid <= ntrail)
        {
            double* colp = (tid == ntrail) ? (vb + j0) : (M + tri(j0 + nb + tid) + j0);
            double u[kPB];
#pragma unroll
            for (int t = 0; t < kPB; ++t)
            {
                if (t < nb)
                {
                    double v = colp[t];
#pragma unroll
                    for (int r = 0; r < kPB; ++r)
                        if (r < t) v = fma(-D[r * kPB + t], u[r], v);
                    u[t] = v * D[t * kPB + t];                          // the diagonal holds 1 / U(t,t)
                    colp[t] = u[t];
                }
                else u[t] = 0.0;
            }
        }
        __syncthreads();
        // ---- trailing update with look-ahead. Columns c in (j0 + nb, n), column n = the right-hand side.
        const int j1 = j0 + nb;                                         // first column of the next panel
        const int nb1 = min(kPB, n - j1);                               // its width (<= 0: no next panel)
        if (warp == 0)
        {
            for (int c = j1; c < j1 + nb1; ++c)
            {
                double uc[kPB];
#pragma unroll
                for (int t = 0; t < kPB; ++t) uc[t] = (t < nb) ? M[tri(c) + j0 + t] : 0.0;
                warp_update_column(M, M + tri(c), j0, nb, uc, j1, c, lane);
            }
            __syncwarp();
            if (nb1 > 0) warp_factor_diag(M, j1, nb1, sD + (pidx ^ 1) * kPBS, lane);
        }
        else
        {
            for (int c = j1 + max(nb1, 0) + (warp - 1); c <= n; c += nwarps - 1)
            {
                const bool is_rhs = (c == n);
                const double* pc = is_rhs ? (vb + j0) : (M + tri(c) + j0);
                double uc[kPB];
#pragma unroll
                for (int t = 0; t < kPB; ++t) uc[t] = (t < nb) ? pc[t] : 0.0;
                warp_update_column(M, is_rhs ? vb : (M + tri(c)), j0, nb, uc, j1, is_rhs ? n - 1 : c, lane);
            }
        }
        __syncthreads();
    }
    // ---- U x = y (y sits in vb), panels from the last to the first; the diagonal of M holds 1 / U(j,j)
    for (int j0 = ((n - 1) / kPB) * kPB; j0 >= 0; j0 -= kPB)
    {
        const int nb = min(kPB, n - j0);
        if (warp == 0)
        {
            double yv = (lane < nb) ? vb[j0 + lane] : 0.0;
#pragma unroll
            for (int qq = kPB - 1; qq >= 0; --qq)
            {
                if (qq < nb)
                {
                    const double* colq = M + tri(j0 + qq) + j0;
                    const double xq = __shfl_sync(FULL, yv * colq[qq], qq);
                    if (lane == qq) yv = xq;
                    else if (lane < qq) yv = fma(-colq[lane], xq, yv);
                }
            }
            if (lane < nb) vb[j0 + lane] = yv;
        }
        __syncthreads();
        if (tid < j0)
        {
            double v = vb[tid];
            for (int t = nb - 1; t >= 0; --t) v = fma(-M[tri(j0 + t) + tid], vb[j0 + t], v);
            vb[tid] = v;
        }
        __syncthreads();
    }
    return true;
}

// BitMatrix::MaxRowIndex as the reference computes it (defect included) on a multi-word column mask.
__device__ __forceinline__ int max_row_index_ref_words(const unsigned int* w, int k)
{
    const int nw = (k + 31) >> 5;
    int h = -1;
    for (int q = nw - 1; q >= 0; --q)
        if (w[q]) { h = q * 32 + 31 - __clz(static_cast<int>(w[q])); break; }
    if (h < 0) return 0;
    const int full = k >> 5, extra = k & 31;
    const int wd = h >> 5;
    if (extra > 0 && wd == full) return h;
    return (wd > 0) ? h - 32 : h;
}

// NMAX = the largest system solved in shared memory: 128 for k <= 256 (a passive set or its complement has at most 128 rows), 64 for
// k <= 128 — there a CTA of 128 threads (one per row) and a 64 x 64 triangle leave room for six columns in flight per SM instead of
// three (r02 ncu of the 128 / 256 form at k = 128, profiles/ncu_r02_c3_sparse_bpp_nnls_wide*.txt: 66 % of the stall samples were block
// barriers, seven warps waiting for the one that factors a diagonal block; more columns per SM fill those slots), and a passive set of
// more than 64 rows goes through its complement, which is the smaller system.
constexpr int kWideRowj = 2 * kPBS;      // two published diagonal blocks of the blocked solve; also the scratch of the CTA-wide max reductions
template <int NMAX>
struct WideSmem
{
    __host__ __device__ static constexpr size_t tri_doubles() { return static_cast<size_t>(NMAX) * (NMAX + 1) / 2; }
    __host__ __device__ static constexpr size_t total_bytes()
    {
        // U | rowj[144] | vb[NMAX] | sb[256] | sx[256] | su[256] | list[256] shorts | words[64]
        // (NMAX = 128: 75,136 bytes, 3 CTAs per SM; NMAX = 64: 25,472 bytes)
        return (tri_doubles() + kWideRowj + NMAX + 3 * 256) * sizeof(double) + 256 * sizeof(unsigned short) + 64 * sizeof(unsigned int);
    }
};

template <int NMAX, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 256 ? 3 : 6))
nnls_bpp_wide_kernel(int k, int q, const double* __restrict__ G, long long ldg, const double* __restrict__ Ginv,
                     const int* __restrict__ ginv_flag, const double* __restrict__ RHS, long long ldr,
                     double* __restrict__ X, long long ldx, double* __restrict__ Y, long long ldy,
                     int* __restrict__ status, unsigned int* __restrict__ counter, int outer_iter,
                     double* __restrict__ gscratch, size_t gscratch_stride)
{
    extern __shared__ __align__(16) double smem[];
    using L = WideSmem<NMAX>;
    double* sU = smem;
    double* s_rowj = sU + L::tri_doubles();
    double* s_vb = s_rowj + kWideRowj;
    double* sb = s_vb + NMAX;
    double* sx = sb + 256;
    double* su = sx + 256;
    unsigned short* list = reinterpret_cast<unsigned short*>(su + 256);
    unsigned int* words = reinterpret_cast<unsigned int*>(list + 256);   // [0..7] passive, [8..15] nonopt, [16..23] infeas, [24] column
    // direct method on more than 128 rows: packed triangle, right-hand side and row buffer in this CTA's global scratch
    double* gU = gscratch + static_cast<size_t>(blockIdx.x) * gscratch_stride;
    double* g_vb = gU + (static_cast<size_t>(k) * (k + 1) / 2);
    double* g_rowj = g_vb + 256;
    double* rowj = s_rowj;      // also the 8-entry scratch of the CTA-wide max reductions

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool valid = t < k;
    const int nw = (k + 31) >> 5;
    const bool ginv_ok = ginv_flag[0] != 0;
    const int max_rounds = 5 * k;

    for (;;)
    {
        __syncthreads();
        if (t == 0) words[24] = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int c = words[24];
        if (c >= static_cast<unsigned int>(q)) break;

        const double* rhs = RHS + static_cast<long long>(c) * ldr;
        double* xcol = X + static_cast<long long>(c) * ldx;
        double* ycol = Y + static_cast<long long>(c) * ldy;
        const double b = valid ? rhs[t] : 0.0;
        sb[t] = b;
        {   // warm start: passive = (X > 0)   (nnls.hpp:157)
            const unsigned int w = __ballot_sync(0xffffffffu, valid && xcol[t] > 0.0);
            if (lane == 0) words[warp] = w;
        }
        double bmax = fabs(b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        if (lane == 0) rowj[warp] = bmax;
        __syncthreads();
        bmax = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) bmax = fmax(bmax, rowj[w]);
        __syncthreads();

        int P = kPbar, Ninf = k + 1, round = 0;
        double x = 0.0, y = 0.0;
        bool failed = false, have_u = false;

        for (;;)
        {
            // ---- position of this row inside the passive list / the complement list
            int p = 0, before = 0;
            for (int w = 0; w < nw; ++w)
            {
                const int pc = __popc(words[w]);
                p += pc;
                if (w < warp) before += pc;
            }
            const unsigned int myword = words[warp];
            const bool in = valid && ((myword >> lane) & 1u);
            const int pos_in = before + __popc(myword & ((1u << lane) - 1u));
            const int pos_out = t - pos_in;                 // rows t' < t that are not passive (t < k)
            bool use_complement = (p > NMAX) && ginv_ok;
            bool solved = false;

            if (p == 0) { x = 0.0; solved = true; }
            while (!solved)
            {
                if (p <= NMAX || !use_complement)
                {
                    // ---- direct: G_PP x_P = b_P
                    const bool small = p <= NMAX;
                    double* M = small ? sU : gU;
                    double* vb = small ? s_vb : g_vb;
                    __syncthreads();
                    if (in) list[pos_in] = t;
                    __syncthreads();
                    for (int cc = warp; cc < p; cc += (THREADS >> 5))
                    {
                        const double* gcol = G + static_cast<long long>(list[cc]) * ldg;
                        double* col = M + tri(cc);
                        for (int i = lane; i <= cc; i += 32) col[i] = gcol[list[i]];
                    }
                    if (t < p) vb[t] = sb[list[t]];
                    __syncthreads();
                    if (!(small ? cta_spd_solve_blocked(M, vb, p, s_rowj) : cta_spd_solve_packed(M, vb, p, g_rowj))) { failed = true; break; }
                    x = in ? vb[pos_in] : 0.0;
                    solved = true;
                }
                else
                {
                    // ---- complement: solve on A = ~P, |A| = k - p < 128
                    const int na = k - p;
                    if (!have_u)
                    {
                        // the L2-resident matrix is read with eight loads in flight per thread (the loop is latency-bound otherwise)
                        double u = 0.0;
                        if (valid)
                        {
                            const double* gp = Ginv + t;
#pragma unroll 8
                            for (int cc = 0; cc < k; ++cc) u = fma(__ldg(gp + static_cast<long long>(cc) * k), sb[cc], u);
                        }
                        su[t] = u;
                        have_u = true;
                    }
                    __syncthreads();
                    if (valid && !in) list[pos_out] = t;
                    __syncthreads();
                    for (int cc = warp; cc < na; cc += (THREADS >> 5))
                    {
                        const double* gcol = Ginv + static_cast<long long>(list[cc]) * k;
                        double* col = sU + tri(cc);
                        for (int i = lane; i <= cc; i += 32) col[i] = gcol[list[i]];
                    }
                    if (t < na) s_vb[t] = su[list[t]];
                    __syncthreads();
                    if (na > 0 && !cta_spd_solve_blocked(sU, s_vb, na, s_rowj)) { use_complement = false; continue; }
                    double a = su[t];
                    if (in)
                    {
#pragma unroll 8
                        for (int e = 0; e < na; ++e) a = fma(-__ldg(Ginv + static_cast<long long>(list[e]) * k + t), s_vb[e], a);
                    }
                    x = in ? a : 0.0;
                    solved = true;
                }
            }
            if (failed) break;
            if (round > 0 && fabs(x) < kZeroThresh) x = 0.0;         // ZeroizeSmallValues(Xsub), nnls.hpp:215
            __syncthreads();
            sx[t] = x;
            if (in) list[pos_in] = t;                                  // passive list for the product (the complement path overwrote it)
            __syncthreads();
            // ---- dual y = G x - b over the passive columns (nnls.hpp:168-169, 219-220)
            double s = 0.0;
            if (valid)
            {
#pragma unroll 8
                for (int e = 0; e < p; ++e)
                {
                    const int cc = list[e];
                    s = fma(__ldg(G + static_cast<long long>(cc) * ldg + t), sx[cc], s);
                }
            }
            y = s - b;
            if (use_complement && p > NMAX)
            {
                // acceptance of the complement path: rounding-level residual on the passive rows
                double res = in ? fabs(y) : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
                __syncthreads();
                if (lane == 0) rowj[warp] = res;
                __syncthreads();
                res = 0.0;
                for (int w = 0; w < THREADS / 32; ++w) res = fmax(res, rowj[w]);
                if (!(res <= 1.0e-11 * bmax))
                {
                    // redo this round with the direct method in global scratch (uniform decision)
                    __syncthreads();
                    double* M = gU;
                    for (int cc = warp; cc < p; cc += (THREADS >> 5))
                    {
                        const double* gcol = G + static_cast<long long>(list[cc]) * ldg;
                        double* col = M + tri(cc);
                        for (int i = lane; i <= cc; i += 32) col[i] = gcol[list[i]];
                    }
                    if (t < p) g_vb[t] = sb[list[t]];
                    __syncthreads();
                    if (!cta_spd_solve_packed(M, g_vb, p, g_rowj)) { failed = true; break; }
                    x = in ? g_vb[pos_in] : 0.0;
                    if (round > 0 && fabs(x) < kZeroThresh) x = 0.0;
                    __syncthreads();
                    sx[t] = x;
                    __syncthreads();
                    s = 0.0;
                    if (valid)
                    {
#pragma unroll 8
                        for (int e = 0; e < p; ++e)
                        {
                            const int cc = list[e];
                            s = fma(__ldg(G + static_cast<long long>(cc) * ldg + t), sx[cc], s);
                        }
                    }
                    y = s - b;
                }
            }
            if (round > 0 && fabs(y) < kZeroThresh) y = 0.0;
            // ---- nonopt = (Y < 0) & ~P, infeas = (X < 0) & P   (nnls.hpp:42-140)
            const unsigned int wn = __ballot_sync(0xffffffffu, valid && !in && y < 0.0);
            const unsigned int wi = __ballot_sync(0xffffffffu, in && x < 0.0);
            __syncthreads();
            if (lane == 0) { words[8 + warp] = wn; words[16 + warp] = wi; }
            __syncthreads();
            int not_good = 0;
            for (int w = 0; w < nw; ++w) not_good += __popc(words[8 + w]) + __popc(words[16 + w]);
            if (not_good == 0) break;
            if (round == 0 && t == 0) atomicOr(&status[ST_ANY_NONOPT], 1);
            if (round >= max_rounds) { failed = true; break; }        // nnls.hpp:195-196
            // ---- UpdatePassiveSet (nnls.cpp:18-74)
            __syncthreads();
            if (not_good < Ninf || P >= 1)
            {
                if (not_good < Ninf) { P = kPbar; Ninf = not_good; } else P -= 1;
                if (t < nw) words[t] = (words[t] | words[8 + t]) & ~words[16 + t];
            }
            else if (t == 0)
            {
                const int ra = max_row_index_ref_words(words + 8, k), rb = max_row_index_ref_words(words + 16, k);
                const int r = ra > rb ? ra : rb;
                words[r >> 5] ^= (1u << (r & 31));
                atomicAdd(&status[ST_BACKUP_COUNT], 1);
            }
            __syncthreads();
            ++round;
        }
        if (failed && t == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        if (valid) { xcol[t] = x; ycol[t] = y; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k > 256: the same pivoting history with nothing sized by k on the chip. A CTA owns a column; its threads take the rows
// round-robin; the packed Cholesky triangle, the vectors (right-hand side of the solve, row buffer, x, y), the passive
// list and the three bit masks sit in a per-CTA scratch in global memory (L2). Every solve is the direct one
// (G_PP x_P = b_P, cta_spd_solve_packed). A fallback: correct for any k, not tuned.
// ---------------------------------------------------------------------------------------------------------------
struct BigScratch
{
    int k;
    __host__ __device__ size_t nw() const { return (static_cast<size_t>(k) + 31) >> 5; }
    __host__ __device__ size_t doubles() const { return static_cast<size_t>(k) * (k + 1) / 2 + 4 * static_cast<size_t>(k); }   // U | vb | rowj | sx | sy
    __host__ __device__ size_t ints() const { return static_cast<size_t>(k) + 4 * nw() + 4; }                                  // list | passive, nonopt, infeas | prefix
    __host__ __device__ size_t bytes() const { return ((doubles() * sizeof(double) + ints() * sizeof(int)) + 255) & ~static_cast<size_t>(255); }
};

__global__ void __launch_bounds__(kWideThreads)
nnls_bpp_big_kernel(int k, int q, const double* __restrict__ G, long long ldg, const double* __restrict__ RHS, long long ldr,
                    double* X, long long ldx, double* Y, long long ldy, int* __restrict__ status, unsigned int* __restrict__ counter,
                    int outer_iter, unsigned char* gscratch)
{
    __shared__ unsigned int s_col;
    __shared__ int s_p, s_ng[kWideThreads / 32];
    const BigScratch L{k};
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nw = static_cast<int>(L.nw());
    const int kr = nw * 32;                       // rows rounded up to whole mask words: the ballot loops run whole warps
    unsigned char* mine = gscratch + static_cast<size_t>(blockIdx.x) * L.bytes();
    double* U = reinterpret_cast<double*>(mine);
    double* vb = U + static_cast<size_t>(k) * (k + 1) / 2;
    double* rowj = vb + k;
    double* sx = rowj + k;
    double* sy = sx + k;
    int* list = reinterpret_cast<int*>(sy + k);
    unsigned int* words = reinterpret_cast<unsigned int*>(list + k);      // [0, nw) passive, [nw, 2nw) nonopt, [2nw, 3nw) infeas
    int* wpre = reinterpret_cast<int*>(words + 3 * nw);                   // passive rows before word w
    const int max_rounds = 5 * k;

    for (;;)
    {
        __syncthreads();
        if (t == 0) s_col = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int c = s_col;
        if (c >= static_cast<unsigned int>(q)) break;
        const double* rhs = RHS + static_cast<long long>(c) * ldr;
        double* xcol = X + static_cast<long long>(c) * ldx;
        double* ycol = Y + static_cast<long long>(c) * ldy;
        // warm start: passive = (X > 0)   (nnls.hpp:157)
        for (int r = t; r < kr; r += kWideThreads)
        {
            const unsigned int w = __ballot_sync(0xffffffffu, r < k && xcol[r] > 0.0);
            if (lane == 0) words[r >> 5] = w;
        }
        int P = kPbar, Ninf = k + 1, round = 0;
        bool failed = false;
        for (;;)
        {
            __syncthreads();
            // ---- passive rows before each mask word (warp 0: a scan over the word counts), then the passive list
            if (warp == 0)
            {
                int run = 0;
                for (int w0 = 0; w0 < nw; w0 += 32)
                {
                    const int w = w0 + lane;
                    const int pc = (w < nw) ? __popc(words[w]) : 0;
                    int incl = pc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    if (w < nw) wpre[w] = run + incl - pc;
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) s_p = run;
            }
            __syncthreads();
            const int p = s_p;
            for (int r = t; r < k; r += kWideThreads)
            {
                const unsigned int w = words[r >> 5];
                if ((w >> (r & 31)) & 1u) list[wpre[r >> 5] + __popc(w & ((1u << (r & 31)) - 1u))] = r;
            }
            __syncthreads();
            if (p > 0)
            {
                // ---- G_PP x_P = b_P
                for (int cc = warp; cc < p; cc += (kWideThreads >> 5))
                {
                    const double* gcol = G + static_cast<long long>(list[cc]) * ldg;
                    double* col = U + (static_cast<size_t>(cc) * (cc + 1) >> 1);
                    for (int i = lane; i <= cc; i += 32) col[i] = gcol[list[i]];
                }
                for (int i = t; i < p; i += kWideThreads) vb[i] = rhs[list[i]];
                __syncthreads();
                if (!cta_spd_solve_packed(U, vb, p, rowj)) { failed = true; break; }
            }
            for (int r = t; r < k; r += kWideThreads)
            {
                const unsigned int w = words[r >> 5];
                double x = 0.0;
                if ((w >> (r & 31)) & 1u) x = vb[wpre[r >> 5] + __popc(w & ((1u << (r & 31)) - 1u))];
                if (round > 0 && fabs(x) < kZeroThresh) x = 0.0;      // ZeroizeSmallValues(Xsub), nnls.hpp:215
                sx[r] = x;
            }
            __syncthreads();
            // ---- dual y = G x - b (nnls.hpp:168-169, 219-220); nonopt = (Y < 0) & ~P, infeas = (X < 0) & P (nnls.hpp:42-140)
            int ng = 0;
            for (int r = t; r < kr; r += kWideThreads)
            {
                bool in = false;
                double x = 0.0, y = 0.0;
                if (r < k)
                {
                    double s = 0.0;
                    for (int e = 0; e < p; ++e)
                    {
                        const int cc = list[e];
                        s = fma(__ldg(G + static_cast<long long>(cc) * ldg + r), sx[cc], s);
                    }
                    y = s - rhs[r];
                    if (round > 0 && fabs(y) < kZeroThresh) y = 0.0;
                    sy[r] = y;
                    in = (words[r >> 5] >> (r & 31)) & 1u;
                    x = sx[r];
                }
                const unsigned int wn = __ballot_sync(0xffffffffu, r < k && !in && y < 0.0);
                const unsigned int wi = __ballot_sync(0xffffffffu, r < k && in && x < 0.0);
                if (lane == 0) { words[nw + (r >> 5)] = wn; words[2 * nw + (r >> 5)] = wi; ng += __popc(wn) + __popc(wi); }
            }
            if (lane == 0) s_ng[warp] = ng;
            __syncthreads();
            int not_good = 0;
#pragma unroll
            for (int w = 0; w < kWideThreads / 32; ++w) not_good += s_ng[w];
            if (not_good == 0) break;
            if (round == 0 && t == 0) atomicOr(&status[ST_ANY_NONOPT], 1);
            if (round >= max_rounds) { failed = true; break; }        // nnls.hpp:195-196
            // ---- UpdatePassiveSet (nnls.cpp:18-74)
            if (not_good < Ninf || P >= 1)
            {
                if (not_good < Ninf) { P = kPbar; Ninf = not_good; } else P -= 1;
                for (int w = t; w < nw; w += kWideThreads) words[w] = (words[w] | words[nw + w]) & ~words[2 * nw + w];
            }
            else if (t == 0)
            {
                const int ra = max_row_index_ref_words(words + nw, k), rb = max_row_index_ref_words(words + 2 * nw, k);
                const int r = ra > rb ? ra : rb;
                words[r >> 5] ^= (1u << (r & 31));
                atomicAdd(&status[ST_BACKUP_COUNT], 1);
            }
            ++round;
        }
        if (failed && t == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        __syncthreads();
        if (!failed)
            for (int r = t; r < k; r += kWideThreads) { xcol[r] = sx[r]; ycol[r] = sy[r]; }
    }
}

} // namespace

size_t nnls_wide_scratch_bytes(int k, int num_sms)
{
    if (k <= 64) return 0;
    const size_t tri_k = static_cast<size_t>(k) * (k + 1) / 2;
    return static_cast<size_t>((k <= 128 ? 6 : 3) * num_sms) * (tri_k + 512) * sizeof(double) + 64;
}

// G^-1 by one CTA working in global memory (k > 160: the matrix does not fit shared memory)
void invert_spd_global(cudaStream_t stream, int k, const double* G, long long ldg, double* Ginv, int* ok)
{
    invert_spd_kernel<<<1, 1024, 0, stream>>>(k, G, ldg, Ginv, ok);
    SMK_LAUNCH_CHECK();
}

// scratch: per-CTA packed triangles of the direct-method fallback. Ginv / ginv_flag: nnls_prepare_inverse's output.
void nnls_bpp_wide(cudaStream_t stream, int k, int q, const double* LHS, long long ldl, const double* RHS, long long ldr,
                   double* X, long long ldx, double* Y, long long ldy, int* status, unsigned int* counter, void* scratch,
                   int outer_iter, int num_sms, const double* Ginv, const int* flag)
{
    if (k > 256) throw std::string("nnls_bpp_wide: internal error (k > 256)");
    const size_t tri_k = static_cast<size_t>(k) * (k + 1) / 2;
    double* gscr = static_cast<double*>(scratch);
    // SMK_NNLS_WIDE128=0: the 256-thread / 128 x 128 form for every k (A/B measurements; the tests run both)
    const char* e = getenv("SMK_NNLS_WIDE128");
    const bool narrow = k <= 128 && !(e && atoi(e) == 0);
    if (narrow)
    {
        const int grid = std::min(6 * num_sms, q);
        const size_t smem = WideSmem<64>::total_bytes();
        auto kern = nnls_bpp_wide_kernel<64, 128>;
        SMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, 128, smem, stream>>>(k, q, LHS, ldl, Ginv, flag, RHS, ldr, X, ldx, Y, ldy, status, counter, outer_iter, gscr, tri_k + 512);
    }
    else
    {
        const int grid = std::min(3 * num_sms, q);
        const size_t smem = WideSmem<128>::total_bytes();
        auto kern = nnls_bpp_wide_kernel<128, 256>;
        SMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kern<<<grid, 256, smem, stream>>>(k, q, LHS, ldl, Ginv, flag, RHS, ldr, X, ldx, Y, ldy, status, counter, outer_iter, gscr, tri_k + 512);
    }
    SMK_LAUNCH_CHECK();
}

// k > 256: see nnls_bpp_big_kernel. The grid is what a quarter of a gigabyte of per-CTA scratch allows (at least one CTA).
constexpr size_t kBigScratchBudget = size_t(256) << 20;
static int nnls_big_grid(int k, int num_sms)
{
    const BigScratch L{k};
    return static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(2 * num_sms), kBigScratchBudget / L.bytes())));
}
size_t nnls_big_scratch_bytes(int k, int num_sms)
{
    if (k <= 256) return 0;
    return static_cast<size_t>(nnls_big_grid(k, num_sms)) * BigScratch{k}.bytes() + 256;
}
void nnls_bpp_big(cudaStream_t stream, int k, int q, const double* LHS, long long ldl, const double* RHS, long long ldr,
                  double* X, long long ldx, double* Y, long long ldy, int* status, unsigned int* counter, void* scratch,
                  int outer_iter, int num_sms)
{
    if (k > 32768) throw std::string("nnls_bpp: k > 32768 is not supported");     // the packed triangle is indexed with 32-bit integers
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~static_cast<uintptr_t>(255));
    const int grid = std::min(nnls_big_grid(k, num_sms), q);
    nnls_bpp_big_kernel<<<grid, kWideThreads, 0, stream>>>(k, q, LHS, ldl, RHS, ldr, X, ldx, Y, ldy, status, counter, outer_iter, base);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
