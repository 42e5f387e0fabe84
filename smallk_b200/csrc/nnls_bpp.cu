// smallk_b200 — batched NNLS by block principal pivoting (k <= 64), one warp per column.
//
// Replaces NnlsBlockpivot + BppSolveNormalEqNoGroup + UpdatePassiveSet + the BitMatrix passes
// of the reference
//   common/include/nnls.hpp:144-244, common/include/nmf_solver_bpp.hpp:146-219,
//   common/src/nnls.cpp:18-74, common/src/bit_matrix.cpp:432-472,
//   Elemental cholesky::UVar3Unb + SolveAfter (UVar3.hpp:17-44, SolveAfter.hpp:17-42).
//
// The reference runs all right-hand-side columns in lock step: every pivot round gathers the
// non-optimal columns, solves them (one heap-allocated k' x k' Cholesky per column), scatters them
// back, then makes serial BitMatrix passes over all q columns. Columns never exchange data, so here
// a warp owns one column for its WHOLE pivoting history; the passive set is a 64-bit mask in a
// register and nothing leaves the kernel between rounds. Columns are handed out through an atomic
// counter so long pivoting histories do not unbalance the SMs.
//
// Fast path (nnls_bpp_fast_kernel). With G = LHS (k x k SPD), P the passive set and A its
// complement, the passive-set normal equations G_PP x_P = b_P can be solved on EITHER index set:
//     direct      : x_P = G_PP^-1 b_P                                        (size |P|)
//     complement  : u = G^-1 b,  z = (G^-1)_AA^-1 u_A,  x_P = u_P - (G^-1)_PA z   (size |A|)
// (block-inverse identity; y_A = -z is the dual). For k <= 64 the smaller of the two never exceeds
// 32, so every solve is a <= 32 x 32 upper Cholesky held ENTIRELY IN REGISTERS, one matrix column
// per lane (right-looking, the recurrence of Elemental's UVar3Unb), row j broadcast by shuffles,
// followed by the two triangular solves (forward in column layout, backward after one
// shared-memory transpose). G^-1 is formed once per CTA in shared memory (in-place Gauss-Jordan).
// The residual y_P = (G x - b)_P, which is computed anyway for the dual, is the accuracy check of
// the complement path: a column whose residual is not at rounding level (ill-conditioned G), or
// any solve the fast path cannot do (G not invertible), is DEFERRED with its pivoting state to the
// slow path, which redoes it with the direct method.
//
// Slow path (nnls_bpp_slow_kernel): the direct method for any |P| <= 64 with the packed upper
// triangle in shared memory. Launched after the fast kernel on the deferred list (normally empty).
//
// Reference behaviours preserved:
//   * ZeroizeSmallValues(X|Y, 1e-12) hits EVERY column iff at least one column was non-optimal after
//     the initial solve (nnls.hpp:192,226-227): status[ST_ANY_NONOPT] + zeroize_if_flag_kernel.
//   * MAX_ITER = 5k pivot rounds, or a non-positive Cholesky pivot of a passive-set system, fail the
//     whole solve (nnls.hpp:195-196, normal_eq.hpp:35-50): status[ST_FAIL_ITER] = outer iteration.
//   * BitMatrix::MaxRowIndex's off-by-one-word defect for k > 32, because the
//     backup pivot rule depends on it.
#include "nnls_common.cuh"

namespace smk {

namespace {

__device__ __forceinline__ int tri(int c) { return (c * (c + 1)) >> 1; }

// ===========================================================================
// register-resident SPD solve, n <= NMAX <= 32, lane c holds column c of the upper triangle
// ===========================================================================
// a[i] = M(i, c) for i <= c (c = lane). On return lane c holds x_c.
// sT: per-warp 32 x 33 doubles (U, row-major with ld 33); sRow: per-warp 2 x 32 doubles (row broadcast).
// Pivots use one rsqrt per index: r = 1/sqrt(a_jj), U(j,c) = a_jc * r, and the triangular solves multiply by
// r instead of dividing by U(j,j) (1-2 ulp from the divide form; parity is held to 1e-9, not bitwise).
template <int NMAX>
__device__ __forceinline__ bool spd_solve_reg(int n, double (&a)[NMAX], double rhs, double* __restrict__ sT,
                                              double* __restrict__ sRow, int lane, double& xout)
{
    constexpr unsigned FULL = 0xffffffffu;
    double rinv = 1.0;      // lane j: 1 / U(j,j)
    // ---- factor: A = U'U, right-looking (the recurrence of UVar3Unb)
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
    {
        if (j < n)
        {
            const double ajj = __shfl_sync(FULL, a[j], j);
            if (!(ajj > 0.0)) return false;
            const double r = rsqrt(ajj);
            double ujc;
            if (lane == j) { ujc = ajj * r; rinv = r; }
            else ujc = a[j] * r;
            a[j] = ujc;
            double* row = sRow + (j & 1) * 32;
            row[lane] = ujc;
            __syncwarp();
#pragma unroll
            for (int h = (j + 1) / 2; h < NMAX / 2; ++h)
            {
                // entries a[i] with i > lane lie below the diagonal: never read, so they are updated
                // unconditionally (garbage in, garbage out) instead of paying a predicate per element
                const double2 u2 = reinterpret_cast<const double2*>(row)[h];
                if (2 * h > j) a[2 * h] = fma(-u2.x, ujc, a[2 * h]);
                a[2 * h + 1] = fma(-u2.y, ujc, a[2 * h + 1]);
            }
        }
    }
    // ---- U'y = b, column layout: lane c owns s_c
    double s = rhs;
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
    {
        if (i < n)
        {
            double yi = s * rinv;                   // meaningful on lane i only
            yi = __shfl_sync(FULL, yi, i);
            if (lane == i) s = yi;
            else if (lane > i) s = fma(-a[i], yi, s);
        }
    }
    // ---- U to shared memory (row-major), then U x = y in row layout: lane r reads U(r, q) as it goes
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
        if (i <= lane && lane < n) sT[i * 33 + lane] = a[i];
    __syncwarp();
    double x = s;
    const double* urow = sT + lane * 33;
#pragma unroll
    for (int qq = NMAX - 1; qq >= 0; --qq)
    {
        if (qq < n)
        {
            double xq = x * rinv;                   // meaningful on lane qq only
            xq = __shfl_sync(FULL, xq, qq);
            if (lane == qq) x = xq;
            else if (lane < qq) x = fma(-urow[qq], xq, x);
        }
    }
    __syncwarp();
    xout = x;
    return true;
}

// Gathers M(list,list) from the k x k matrix sM (column-major, ld k) and solves.
template <int NMAX>
__device__ __forceinline__ bool gather_solve(const double* __restrict__ sM, int k, const int* __restrict__ list, int n,
                                             double rhs, double* __restrict__ sT, double* __restrict__ sRow, int lane,
                                             double& xout)
{
    double a[NMAX];
    const int lc = (lane < n) ? list[lane] * k : 0;
#pragma unroll
    for (int i = 0; i < NMAX; ++i) a[i] = (i <= lane && lane < n) ? sM[lc + list[i]] : ((i == lane) ? 1.0 : 0.0);
    return spd_solve_reg<NMAX>(n, a, rhs, sT, sRow, lane, xout);
}

__device__ __forceinline__ bool gather_solve_any(const double* __restrict__ sM, int k, const int* __restrict__ list, int n,
                                                 double rhs, double* __restrict__ sT, double* __restrict__ sRow,
                                                 int lane, double& xout)
{
    if (n <= 8) return gather_solve<8>(sM, k, list, n, rhs, sT, sRow, lane, xout);
    if (n <= 16) return gather_solve<16>(sM, k, list, n, rhs, sT, sRow, lane, xout);
    if (n <= 24) return gather_solve<24>(sM, k, list, n, rhs, sT, sRow, lane, xout);
    return gather_solve<32>(sM, k, list, n, rhs, sT, sRow, lane, xout);
}

// Gauss-Jordan inversion of the SPD k x k matrix G by ONE CTA, the matrix held in shared memory (k <= 160). ok[0] = 1 on
// success, 0 if a pivot is not positive (Ginv is then garbage and the callers fall back to the direct method).
// Thread t owns the entries e = t, t + 1024, ...: their row / column indices follow by increments (no division in the
// loop), every elimination step is two barriers and one multiply-add per owned entry.
constexpr int kInvThreads = 1024;
constexpr int kInvSmemMaxK = 160;

__global__ void __launch_bounds__(kInvThreads, 1)
spd_inverse_smem_kernel(int k, const double* __restrict__ G, long long ldg, double* __restrict__ Ginv, int* __restrict__ ok)
{
    extern __shared__ __align__(16) double sinv[];
    double* S = sinv;                 // k x k, ld k
    double* scol = S + k * k;         // k
    double* srow = scol + k;          // k
    const int kk = k * k;
    // entry e = threadIdx.x + s * kInvThreads is (row i, column c); stepping s moves (i, c) by (di, dc) with one carry
    const int i0 = threadIdx.x % k, c0 = threadIdx.x / k;
    const int di = kInvThreads % k, dc = kInvThreads / k;
    {
        int i = i0, c = c0;
        for (int e = threadIdx.x; e < kk; e += kInvThreads)
        {
            S[e] = G[static_cast<long long>(c) * ldg + i];
            i += di; c += dc;
            if (i >= k) { i -= k; c += 1; }
        }
    }
    __syncthreads();
    bool good = true;
    for (int j = 0; j < k; ++j)
    {
        const double piv = S[j + j * k];
        if (!(piv > 0.0)) { good = false; break; }           // uniform: everyone reads the same word after the barrier
        const double ip = 1.0 / piv;
        for (int i = threadIdx.x; i < k; i += kInvThreads)
        {
            scol[i] = S[i + j * k];
            srow[i] = S[j + i * k] * ip;
        }
        __syncthreads();
        int i = i0, c = c0;
        for (int e = threadIdx.x; e < kk; e += kInvThreads)
        {
            double v;
            if (i == j) v = (c == j) ? ip : srow[c];
            else if (c == j) v = -scol[i] * ip;
            else v = S[e] - scol[i] * srow[c];
            S[e] = v;
            i += di; c += dc;
            if (i >= k) { i -= k; c += 1; }
        }
        __syncthreads();
    }
    if (good)
        for (int e = threadIdx.x; e < kk; e += kInvThreads) Ginv[e] = S[e];
    if (threadIdx.x == 0) ok[0] = good ? 1 : 0;
}

// ---------------------------------------------------------------------------
// fast kernel
// ---------------------------------------------------------------------------
constexpr int kFastWarps = 14;

struct FastSmem
{
    // layout in doubles, computed identically on host and device
    int k;
    __host__ __device__ size_t g() const { return 0; }
    __host__ __device__ size_t ginv() const { return static_cast<size_t>(k) * k; }
    __host__ __device__ size_t rowcol() const { return ginv() + (k > 32 ? static_cast<size_t>(k) * k : 0); }
    __host__ __device__ size_t warp0() const { return (rowcol() + 2 * static_cast<size_t>(k) + 1) & ~static_cast<size_t>(1); }
    __host__ __device__ size_t per_warp() const { return (32 * 33 + 64 + 3 * static_cast<size_t>(k) + 32 + (k + 1) / 2 + 2) & ~static_cast<size_t>(1); }   // even: keeps double2 alignment
    __host__ __device__ size_t total_bytes(int warps) const { return (warp0() + warps * per_warp()) * sizeof(double) + 16; }
};

__global__ void __launch_bounds__(kFastWarps * 32, 1)
nnls_bpp_fast_kernel(int k, int q, const double* __restrict__ LHS, long long ldl,
                     const double* __restrict__ Ginv_g, const int* __restrict__ ginv_flag,
                     const double* __restrict__ RHS, long long ldr,
                     double* __restrict__ X, long long ldx, double* __restrict__ Y, long long ldy,
                     int* __restrict__ status, unsigned int* __restrict__ counter, int outer_iter,
                     BppColState* __restrict__ deferred)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_ginv_ok;
    const FastSmem L{k};
    double* sG = smem + L.g();
    double* sGinv = smem + L.ginv();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wb = smem + L.warp0() + warp * L.per_warp();
    double* sT = wb;                 // 32 x 33
    double* sRow = sT + 32 * 33;     // 2 x 32 row broadcast buffers (16-byte aligned: 1056 doubles precede)
    double* sb = sRow + 64;          // k   rhs column
    double* su = sb + k;             // k   u = Ginv * b
    double* sx = su + k;             // k   x column
    double* sz = sx + k;             // 32  solution of the small system
    int* list = reinterpret_cast<int*>(sz + 32);   // k ints

    // G and (k > 32) its inverse, formed ONCE per call by spd_inverse_kernel (normally on a side stream under the big product
    // that precedes this solve) instead of by every CTA: the in-CTA Gauss-Jordan was a fixed ~0.19 ms of each call
    for (int i = threadIdx.x; i < k * k; i += blockDim.x)
    {
        sG[i] = LHS[static_cast<long long>(i / k) * ldl + (i % k)];
        if (k > 32) sGinv[i] = Ginv_g[i];
    }
    if (threadIdx.x == 0) s_ginv_ok = (k > 32 && ginv_flag[0] != 0) ? 1 : 0;
    __syncthreads();
    const bool ginv_ok = s_ginv_ok != 0;

    const int r0 = lane, r1 = lane + 32;
    const bool v0 = r0 < k, v1 = r1 < k;
    const int max_rounds = 5 * k;
    const unsigned long long kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);

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
        if (v0) sb[r0] = b0;
        if (v1) sb[r1] = b1;
        double bmax = fmax(fabs(b0), fabs(b1));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        unsigned long long pm;
        {   // warm start: passive = (X > 0)   (nnls.hpp:157)
            const bool p0 = v0 && (xcol[r0] > 0.0), p1 = v1 && (xcol[r1] > 0.0);
            pm = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, p0)) |
                 (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, p1)) << 32);
        }
        int P = kPbar, Ninf = k + 1, round = 0;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
        bool failed = false, defer = false, have_u = false;
        __syncwarp();

        for (;;)
        {
            const int p = __popcll(pm);
            const bool in0 = (pm >> r0) & 1ull, in1 = v1 && ((pm >> r1) & 1ull);
            bool quick = false;
            double zq = 0.0;
            // The passive-set system is solved on the smaller of P (direct) and its complement A (block-inverse identity, see the
            // head of the file). Both forms go through ONE call of the register-resident solver: its four size classes are 2 800
            // SASS instructions each, and a second call site doubled a kernel that already overflowed the instruction cache.
            const bool direct = p <= 32;
            const int na = k - p;
            const int nsel = direct ? p : na;
            if (!direct)
            {
                if (!ginv_ok) { defer = true; break; }
                if (!have_u)
                {
                    double u0 = 0.0, u1 = 0.0;
#pragma unroll 4
                    for (int cc = 0; cc < k; ++cc)
                    {
                        const double bv = sb[cc];
                        const double* col = sGinv + cc * k;
                        if (v0) u0 = fma(col[r0], bv, u0);
                        if (v1) u1 = fma(col[r1], bv, u1);
                    }
                    if (v0) su[r0] = u0;
                    if (v1) su[r1] = u1;
                    have_u = true;
                    __syncwarp();
                }
            }
            {
                // the index list of the system: P in ascending order (direct) or A = ~P (complement)
                const unsigned long long sel = direct ? pm : (~pm & kmask);
                if (v0 && ((sel >> r0) & 1ull)) list[__popcll(sel & ((1ull << r0) - 1ull))] = r0;
                if (v1 && ((sel >> r1) & 1ull)) list[__popcll(sel & ((1ull << r1) - 1ull))] = r1;
                __syncwarp();
            }
            double z = 0.0;
            if (nsel > 0)
            {
                const double* rsel = direct ? sb : su;
                const double rv = (lane < nsel) ? rsel[list[lane]] : 0.0;
                if (!gather_solve_any(direct ? sG : sGinv, k, list, nsel, rv, sT, sRow, lane, z))
                {
                    if (direct) failed = true; else defer = true;      // not positive definite: G itself (failure) or the identity's (G^-1)_AA (redo directly)
                    break;
                }
            }
            if (direct)
            {
                // ---- direct: x_P = z (G_PP z = b_P); p == 0 leaves x = 0
                if (round > 0 && fabs(z) < kZeroThresh) z = 0.0;     // ZeroizeSmallValues(Xsub), nnls.hpp:215
                if (v0) sx[r0] = 0.0;
                if (v1) sx[r1] = 0.0;
                __syncwarp();
                if (lane < p) sx[list[lane]] = z;
                __syncwarp();
                x0 = v0 ? sx[r0] : 0.0;
                x1 = v1 ? sx[r1] : 0.0;
            }
            else
            {
                // ---- complement: z = (G^-1)_AA^-1 u_A, x_P = u_P - (G^-1)_PA z, |A| = k - p < 32
                if (lane < na) sz[lane] = z;
                __syncwarp();
                double a0 = v0 ? su[r0] : 0.0, a1 = v1 ? su[r1] : 0.0;
                for (int t = 0; t < na; ++t)
                {
                    const double zt = sz[t];
                    const double* col = sGinv + list[t] * k;
                    if (in0) a0 = fma(-col[r0], zt, a0);
                    if (in1) a1 = fma(-col[r1], zt, a1);
                }
                x0 = in0 ? a0 : 0.0;
                x1 = in1 ? a1 : 0.0;
                if (round > 0)
                {
                    if (fabs(x0) < kZeroThresh) x0 = 0.0;
                    if (fabs(x1) < kZeroThresh) x1 = 0.0;
                }
                if (v0) sx[r0] = x0;
                if (v1) sx[r1] = x1;
                __syncwarp();
                quick = true;
                zq = z;
            }
            // ---- dual y = LHS * x - rhs (nnls.hpp:168-169, 219-220)
            // full_y: the product over the passive columns, as the reference forms it.
            // sx is zero outside the passive set, so the dense loop adds the same terms in the same (ascending) order as a walk
            // over the set bits of pm — fma(g, 0, s) == s exactly — without the 17 instructions of 64-bit mask arithmetic per
            // passive column that the walk costs (it was 31 % of the kernel's instructions, profiles/ncu_r02_nnls_fast_lines.txt)
            auto full_y = [&](double& o0, double& o1) {
                double s0 = 0.0, s1 = 0.0;
                const double* col0 = sG + (v0 ? r0 : 0);
                const double* col1 = sG + (v1 ? r1 : 0);
                if (4 * __popcll(pm) < k)
                {
                    // a small passive set: the walk over its bits is still the shorter loop
                    unsigned long long mm = pm;
                    while (mm)
                    {
                        const int cc = __ffsll(static_cast<long long>(mm)) - 1;
                        mm &= mm - 1ull;
                        const double xv = sx[cc];
                        s0 = fma(col0[cc * k], xv, s0);
                        s1 = fma(col1[cc * k], xv, s1);
                    }
                }
                else
                {
#pragma unroll 8
                    for (int cc = 0; cc < k; ++cc)
                    {
                        const double xv = sx[cc];
                        s0 = fma(col0[cc * k], xv, s0);
                        s1 = fma(col1[cc * k], xv, s1);
                    }
                }
                o0 = s0 - b0; o1 = s1 - b1;
            };
            if (quick)
            {
                // complement path: on the active rows y_A = -z, on the passive rows y_P = 0; the pivoting
                // decisions of this round use that, the product is formed once when the column is accepted
                if (v0) sT[r0] = 0.0;
                if (v1) sT[r1] = 0.0;
                __syncwarp();
                if (lane < k - p) sT[list[lane]] = -zq;
                __syncwarp();
                y0 = v0 ? sT[r0] : 0.0;
                y1 = v1 ? sT[r1] : 0.0;
                __syncwarp();
            }
            else full_y(y0, y1);
            if (round > 0)
            {
                if (fabs(y0) < kZeroThresh) y0 = 0.0;
                if (fabs(y1) < kZeroThresh) y1 = 0.0;
            }
            unsigned long long nonopt =
                static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v0 && !in0 && y0 < 0.0)) |
                (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v1 && !in1 && y1 < 0.0)) << 32);
            const unsigned long long infeas =
                static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v0 && in0 && x0 < 0.0)) |
                (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v1 && in1 && x1 < 0.0)) << 32);
            int not_good = __popcll(nonopt) + __popcll(infeas);
            __syncwarp();
            if (not_good == 0 && quick)
            {
                // acceptance check of the complement path: the true product must (a) leave a rounding-level
                // residual on the passive rows and (b) agree that the column is optimal; otherwise the
                // column goes to the direct method (ill-conditioned LHS).
                full_y(y0, y1);
                double res = fmax(in0 ? fabs(y0) : 0.0, in1 ? fabs(y1) : 0.0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
                if (round > 0)
                {
                    if (fabs(y0) < kZeroThresh) y0 = 0.0;
                    if (fabs(y1) < kZeroThresh) y1 = 0.0;
                }
                nonopt = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v0 && !in0 && y0 < 0.0)) |
                         (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, v1 && !in1 && y1 < 0.0)) << 32);
                if (!(res <= 1.0e-11 * bmax) || nonopt != 0ull) { defer = true; break; }
            }
            if (not_good == 0) break;
            if (round == 0 && lane == 0) atomicOr(&status[ST_ANY_NONOPT], 1);
            if (round >= max_rounds) { failed = true; break; }     // nnls.hpp:195-196
            update_passive_set(pm, P, Ninf, not_good, nonopt, infeas, k, status, lane == 0);
            ++round;
        }
        if (defer)
        {
            if (lane == 0)
            {
                const int slot = atomicAdd(&status[ST_DEFER_COUNT], 1);
                BppColState st;
                st.pm = pm; st.P = P; st.Ninf = Ninf; st.round = round; st.col = static_cast<int>(c);
                deferred[slot] = st;
            }
            __syncwarp();
            continue;
        }
        if (failed && lane == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        if (v0) { xcol[r0] = x0; ycol[r0] = y0; }
        if (v1) { xcol[r1] = x1; ycol[r1] = y1; }
        __syncwarp();
    }
}

// ===========================================================================
// slow path: direct method, packed upper triangle in shared memory, any |P| <= 64
// ===========================================================================
struct WarpScratch
{
    double* U;     // packed upper triangle, k(k+1)/2
    double* vb;    // k   sub-vector rhs / solution
    double* sb;    // k   full rhs column
    double* sx;    // k   full x column
    int* ri;       // k   passive row list
};

__device__ __forceinline__ bool warp_spd_solve(const double* __restrict__ sL, int k, int p,
                                               const WarpScratch& w, int lane)
{
    double* U = w.U;
    double* vb = w.vb;
    const int* ri = w.ri;
    for (int c = 0; c < p; ++c)
    {
        const int rc = ri[c] * k;
        const int base = tri(c);
        for (int i = lane; i <= c; i += 32) U[base + i] = sL[rc + ri[i]];
    }
    for (int i = lane; i < p; i += 32) vb[i] = w.sb[ri[i]];
    __syncwarp();
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
    for (int i = 0; i < p; ++i)
    {
        const double yi = vb[i] / U[tri(i) + i];
        __syncwarp();
        if (lane == 0) vb[i] = yi;
        for (int c = i + 1 + lane; c < p; c += 32) vb[c] -= U[tri(c) + i] * yi;
        __syncwarp();
    }
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

__global__ void nnls_bpp_slow_kernel(int k, const double* __restrict__ LHS, long long ldl,
                                     const double* __restrict__ RHS, long long ldr,
                                     double* __restrict__ X, long long ldx, double* __restrict__ Y, long long ldy,
                                     int* __restrict__ status, unsigned int* __restrict__ counter, int outer_iter,
                                     const BppColState* __restrict__ deferred)
{
    const int ndef = status[ST_DEFER_COUNT];
    if (ndef == 0) return;
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
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(counter, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= static_cast<unsigned int>(ndef)) break;
        const BppColState st = deferred[slot];
        const int c = st.col;

        const double* rhs = RHS + static_cast<long long>(c) * ldr;
        double* xcol = X + static_cast<long long>(c) * ldx;
        double* ycol = Y + static_cast<long long>(c) * ldy;
        const double b0 = v0 ? rhs[r0] : 0.0, b1 = v1 ? rhs[r1] : 0.0;
        if (v0) w.sb[r0] = b0;
        if (v1) w.sb[r1] = b1;
        unsigned long long pm = st.pm;
        int P = st.P, Ninf = st.Ninf, round = st.round;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
        bool failed = false;

        for (;;)
        {
            const int p = __popcll(pm);
            if (v0 && ((pm >> r0) & 1ull)) w.ri[__popcll(pm & ((1ull << r0) - 1ull))] = r0;
            if (v1 && ((pm >> r1) & 1ull)) w.ri[__popcll(pm & ((1ull << r1) - 1ull))] = r1;
            if (v0) w.sx[r0] = 0.0;
            if (v1) w.sx[r1] = 0.0;
            __syncwarp();
            if (p > 0)
            {
                if (!warp_spd_solve(sL, k, p, w, lane)) { failed = true; break; }
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
            if (round >= max_rounds) { failed = true; break; }
            update_passive_set(pm, P, Ninf, not_good, nonopt, infeas, k, status, lane == 0);
            ++round;
        }
        if (failed && lane == 0) atomicMin(&status[ST_FAIL_ITER], outer_iter);
        if (v0) { xcol[r0] = x0; ycol[r0] = y0; }
        if (v1) { xcol[r1] = x1; ycol[r1] = y1; }
        __syncwarp();
    }
}

// the work counters of the two kernels and the per-call flags
__global__ void nnls_reset_kernel(unsigned int* __restrict__ counter, int* __restrict__ status)
{
    if (threadIdx.x == 0) { counter[0] = 0u; counter[1] = 0u; status[ST_ANY_NONOPT] = 0; status[ST_DEFER_COUNT] = 0; }
}

// Finishes ZeroizeSmallValues(X), ZeroizeSmallValues(Y) for columns that never entered the pivot loop;
// a no-op unless some column was non-optimal.
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

size_t nnls_wide_scratch_bytes(int k, int num_sms);
void nnls_bpp_wide(cudaStream_t stream, int k, int q, const double* LHS, long long ldl, const double* RHS, long long ldr,
                   double* X, long long ldx, double* Y, long long ldy, int* status, unsigned int* counter, void* scratch,
                   int outer_iter, int num_sms, const double* Ginv, const int* ginv_flag);
void invert_spd_global(cudaStream_t stream, int k, const double* G, long long ldg, double* Ginv, int* ok);
size_t nnls_big_scratch_bytes(int k, int num_sms);
void nnls_bpp_big(cudaStream_t stream, int k, int q, const double* LHS, long long ldl, const double* RHS, long long ldr,
                  double* X, long long ldx, double* Y, long long ldy, int* status, unsigned int* counter, void* scratch,
                  int outer_iter, int num_sms);

// G^-1 (k x k, tight) and its success flag for nnls_bpp: one launch, meant to run on a side stream under the big product that
// precedes the NNLS solve (solver.cu). Not needed (and not computed) for k <= 32, where every passive-set system is solved directly.
void nnls_prepare_inverse(cudaStream_t stream, int k, const double* LHS, long long ldl, double* Ginv, int* ok)
{
    if (!nnls_uses_inverse(k)) return;
    if (k <= kInvSmemMaxK)
    {
        const size_t smem = (static_cast<size_t>(k) * k + 2 * static_cast<size_t>(k)) * sizeof(double);
        SMK_CUDA(cudaFuncSetAttribute(spd_inverse_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        spd_inverse_smem_kernel<<<1, kInvThreads, smem, stream>>>(k, LHS, ldl, Ginv, ok);
        SMK_LAUNCH_CHECK();
    }
    else invert_spd_global(stream, k, LHS, ldl, Ginv, ok);
}

// bytes of the `deferred` buffer nnls_bpp needs: the fast->slow hand-over list (k <= 64) or the wide kernel's scratch,
// + room for G^-1 and its flag when the caller does not bring them
size_t nnls_deferred_bytes(int q, int k, int num_sms)
{
    if (k > 256) return nnls_big_scratch_bytes(k, num_sms);
    return std::max(static_cast<size_t>(q) * sizeof(BppColState), nnls_wide_scratch_bytes(k, num_sms)) +
           static_cast<size_t>(k) * k * sizeof(double) + 64;
}

// status: device int[ST_COUNT]; counter: device unsigned[2]; deferred: nnls_deferred_bytes(q) bytes.
void nnls_bpp(cudaStream_t stream, int k, int q, const double* LHS, long long ldl,
              const double* RHS, long long ldr, double* X, long long ldx, double* Y, long long ldy,
              int* status, unsigned int* counter, void* deferred, int outer_iter, int num_sms,
              const double* Ginv, const int* ginv_flag)
{
    nnls_reset_kernel<<<1, 32, 0, stream>>>(counter, status);     // one launch instead of three memsets (the solves of a small shard are latency)
    SMK_LAUNCH_CHECK();
    if (q <= 0) return;          // a rank may own no rows; the flags above are still reset for the reduction that follows
    if (k > 256)
    {
        nnls_bpp_big(stream, k, q, LHS, ldl, RHS, ldr, X, ldx, Y, ldy, status, counter, deferred, outer_iter, num_sms);
        return;
    }
    if (!Ginv && k > 32)
    {
        // the caller did not prepare G^-1: form it here, at the tail of the scratch buffer
        unsigned char* tail = static_cast<unsigned char*>(deferred) + nnls_deferred_bytes(q, k, num_sms) - (static_cast<size_t>(k) * k * sizeof(double) + 64);
        tail = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~static_cast<uintptr_t>(15));
        double* gi = reinterpret_cast<double*>(tail);
        int* fl = reinterpret_cast<int*>(gi + static_cast<size_t>(k) * k);
        nnls_prepare_inverse(stream, k, LHS, ldl, gi, fl);
        Ginv = gi; ginv_flag = fl;
    }
    if (k > 64)
    {
        nnls_bpp_wide(stream, k, q, LHS, ldl, RHS, ldr, X, ldx, Y, ldy, status, counter, deferred, outer_iter, num_sms, Ginv, ginv_flag);
        return;
    }
    BppColState* def = static_cast<BppColState*>(deferred);

    {
        const FastSmem L{k};
        const size_t smem = L.total_bytes(kFastWarps);
        const int grid = std::min(num_sms, ceil_div(q, kFastWarps));
        SMK_CUDA(cudaFuncSetAttribute(nnls_bpp_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        nnls_bpp_fast_kernel<<<grid, kFastWarps * 32, smem, stream>>>(k, q, LHS, ldl, Ginv, ginv_flag, RHS, ldr, X, ldx, Y, ldy, status,
                                                                     counter, outer_iter, def);
        SMK_LAUNCH_CHECK();
    }
    {
        const size_t tri_k = static_cast<size_t>(k) * (k + 1) / 2;
        const size_t per_warp = (tri_k + 3 * k) * sizeof(double) + k * sizeof(int);
        const size_t fixed = static_cast<size_t>(k) * k * sizeof(double);
        const size_t budget = 220 * 1024;
        int warps = static_cast<int>((budget - fixed) / per_warp);
        warps = std::max(1, std::min(warps, 16));
        const size_t smem = fixed + warps * per_warp + 16;
        const int grid = std::min(num_sms, ceil_div(q, warps));
        SMK_CUDA(cudaFuncSetAttribute(nnls_bpp_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        nnls_bpp_slow_kernel<<<grid, warps * 32, smem, stream>>>(k, LHS, ldl, RHS, ldr, X, ldx, Y, ldy, status, counter + 1,
                                                                outer_iter, def);
        SMK_LAUNCH_CHECK();
    }
}

// Second half of NnlsBlockpivot's cross-column coupling (nnls.hpp:192,226-227): X and Y are zeroized everywhere iff
// SOME column was non-optimal after its first solve. Kept separate from nnls_bpp so that a multi-GPU caller can
// OR-reduce status[ST_ANY_NONOPT] over the ranks in between.
void nnls_bpp_finish(cudaStream_t stream, int k, int q, double* X, long long ldx, double* Y, long long ldy, int* status, int num_sms)
{
    if (q <= 0) return;
    const long long total = static_cast<long long>(k) * q;
    const int zb = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * num_sms));
    zeroize_if_flag_kernel<<<zb, 256, 0, stream>>>(status, k, q, X, ldx, Y, ldy);
    SMK_LAUNCH_CHECK();
}

} // namespace smk
