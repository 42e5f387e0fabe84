/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * nmf_oracle.c: a plain-C, single-threaded CPU restatement of the reference's
 * NMF iteration hot path, written from the reference's
 * behaviour, not copied from it. It exists so that tests/ can check the CUDA
 * path on any box (the GPU box has no /root/reference). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * PINNING: this restatement is itself checked against the reference's own
 * code (oracle/_ref/libsmallk_ref.so = reference sources + El.hpp shim) in
 * tests/test_oracle_vs_ref.py and against the committed fixtures under
 * tests/golden/ that were produced by that build (tests/golden/make_golden.py).
 * The reference ships no golden vectors for this path in-tree (they live in
 * the external smallk_data repository), so those reference-run fixtures are
 * the pin.
 *
 * All matrices are column-major doubles; "ld" is the leading dimension.
 * Dense products use plain sequential accumulation (the reference calls BLAS
 * dgemm whose summation order is unspecified; agreement is to rounding).
 * Compile with -ffp-contract=off so that the arithmetic is the one written.
 *
 * Every function names the reference lines it follows.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_OK = 0, ORC_BAD_PARAM = -3, ORC_FAILURE = -4 };            /* nmf.hpp:17-26 */
enum { ALG_MU = 0, ALG_HALS = 1, ALG_RANK2 = 2, ALG_BPP = 3 };          /* nmf.hpp:28-34 */
enum { PROG_PG_RATIO = 0, PROG_DELTA_FNORM = 1 };                        /* nmf.hpp:37-41 */

#define AT(p, ld, i, j) ((p)[(size_t)(i) + (size_t)(j) * (size_t)(ld)])

/* ------------------------------------------------------------------ */
/* small dense kernels                                                 */
/* ------------------------------------------------------------------ */

/* C(m x n) = op(A) * op(B), sequential accumulation over the inner index. */
static void gemm(int tA, int tB, int m, int n, int kk,
                 const double* A, int lda, const double* B, int ldb, double* C, int ldc)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i)
        {
            double s = 0.0;
            for (int p = 0; p < kk; ++p)
            {
                double a = tA ? AT(A, lda, p, i) : AT(A, lda, i, p);
                double b = tB ? AT(B, ldb, j, p) : AT(B, ldb, p, j);
                s += a * b;
            }
            AT(C, ldc, i, j) = s;
        }
}

/* G = C*X - R  (k x k times k x q minus k x q): the reference's Gemm + Axpy(-1)
 * pairs, e.g. nmf_solver_bpp.hpp:373-374, nnls.hpp:168-169. */
static void gemm_minus(int k, int q, const double* L, int ldl, const double* X, int ldx,
                       const double* R, int ldr, double* G, int ldg)
{
    for (int j = 0; j < q; ++j)
        for (int i = 0; i < k; ++i)
        {
            double s = 0.0;
            for (int p = 0; p < k; ++p) s += AT(L, ldl, i, p) * AT(X, ldx, p, j);
            AT(G, ldg, i, j) = s + (-1.0) * AT(R, ldr, i, j);
        }
}

/* Elemental FrobeniusNorm: scaled-square accumulation, column-major order.
 * modules/libelemental/src/lapack_like/props/Norm/Frobenius.hpp:16-29 */
static double fro_norm(const double* A, int ld, int m, int n)
{
    double scale = 0.0, ssq = 1.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i)
        {
            double a = fabs(AT(A, ld, i, j));
            if (a != 0.0)
            {
                if (a <= scale) { double r = a / scale; ssq += r * r; }
                else { double r = scale / a; ssq = ssq * r * r + 1.0; scale = a; }
            }
        }
    return scale * sqrt(ssq);
}

/* dnrm2 (reference BLAS formulation) used by NormalizeColumns, normalize.hpp:46 */
static double nrm2(const double* x, int n)
{
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; ++i)
        if (x[i] != 0.0)
        {
            double a = fabs(x[i]);
            if (scale < a) { double r = scale / a; ssq = 1.0 + ssq * r * r; scale = a; }
            else { double r = a / scale; ssq += r * r; }
        }
    return scale * sqrt(ssq);
}

/* ------------------------------------------------------------------ */
/* sparse (CSC) times dense — sparse_gemm_ab_impl.hpp / sparse_gemm_ba_impl.hpp */
/* ------------------------------------------------------------------ */

/* variant 0: C = alpha*A*B  + beta*C   (sparse_gemm_ab_impl.hpp:61-99)
 * variant 1: C = alpha*A*B' + beta*C   (sparse_gemm_ab_impl.hpp:103-160)
 * variant 2: C = alpha*B*A  + beta*C   (sparse_gemm_ba_impl.hpp:26-78)
 * variant 3: C = alpha*B'*A + beta*C   (sparse_gemm_ba_impl.hpp:82-140)
 * Per output element the products are added in the order the reference adds
 * them: A's columns ascending, then storage order inside a column. */
int orc_sparse_gemm(int variant, double alpha, double beta,
                    unsigned m, unsigned n, const unsigned* colp, const unsigned* rowi, const double* val,
                    const double* B, int Bh, int Bw, double* C, int Ch, int Cw)
{
    const int ldb = Bh, ldc = Ch;
    if (variant == 0 || variant == 1)
    {
        const int ncols = Cw;
        if ((unsigned)Ch != m) return ORC_BAD_PARAM;
        if (variant == 0 && ((unsigned)Bh != n || Bw != Cw)) return ORC_BAD_PARAM;
        if (variant == 1 && ((unsigned)Bw != n || Bh != Cw)) return ORC_BAD_PARAM;
        for (int j = 0; j < ncols; ++j)
        {
            double* cj = C + (size_t)j * ldc;
            for (unsigned r = 0; r < m; ++r) cj[r] = (beta == 0.0) ? 0.0 : cj[r] * beta;
            for (unsigned c = 0; c < n; ++c)
            {
                double ab = alpha * (variant == 0 ? AT(B, ldb, c, j) : AT(B, ldb, j, c));
                for (unsigned o = colp[c]; o != colp[c + 1]; ++o) cj[rowi[o]] += ab * val[o];
            }
        }
        return ORC_OK;
    }
    if (variant == 2)
    {
        if ((unsigned)Bw != m || Ch != Bh || (unsigned)Cw != n) return ORC_BAD_PARAM;
        for (unsigned j = 0; j < n; ++j)
        {
            double* cj = C + (size_t)j * ldc;
            for (int r = 0; r < Bh; ++r) cj[r] = (beta == 0.0) ? 0.0 : cj[r] * beta;
            for (unsigned o = colp[j]; o != colp[j + 1]; ++o)
            {
                double aa = alpha * val[o];
                const double* b = B + (size_t)rowi[o] * ldb;
                for (int r = 0; r < Bh; ++r) cj[r] += aa * b[r];
            }
        }
        return ORC_OK;
    }
    if (variant == 3)
    {
        if ((unsigned)Bh != m || Ch != Bw || (unsigned)Cw != n) return ORC_BAD_PARAM;
        for (unsigned j = 0; j < n; ++j)
        {
            double* cj = C + (size_t)j * ldc;
            for (int r = 0; r < Ch; ++r) cj[r] = (beta == 0.0) ? 0.0 : cj[r] * beta;
            for (unsigned o = colp[j]; o != colp[j + 1]; ++o)
            {
                double aa = alpha * val[o];
                unsigned row = rowi[o];
                for (int c = 0; c < Bw; ++c) cj[c] += aa * AT(B, ldb, row, c);
            }
        }
        return ORC_OK;
    }
    return ORC_BAD_PARAM;
}

/* Stable counting-sort transpose of a CSC matrix (sparse_matrix_ops.hpp:37-126):
 * column r of the result lists A's entries of row r by ascending column,
 * duplicates in storage order. */
void orc_csc_transpose(unsigned m, unsigned n, const unsigned* colp, const unsigned* rowi, const double* val,
                       unsigned* tcolp, unsigned* trowi, double* tval)
{
    unsigned nnz = colp[n];
    for (unsigned r = 0; r <= m; ++r) tcolp[r] = 0;
    for (unsigned o = 0; o < nnz; ++o) tcolp[rowi[o] + 1]++;
    for (unsigned r = 0; r < m; ++r) tcolp[r + 1] += tcolp[r];
    unsigned* next = (unsigned*)malloc(sizeof(unsigned) * (m ? m : 1));
    for (unsigned r = 0; r < m; ++r) next[r] = tcolp[r];
    for (unsigned c = 0; c < n; ++c)
        for (unsigned o = colp[c]; o != colp[c + 1]; ++o)
        {
            unsigned d = next[rowi[o]]++;
            trowi[d] = c;
            tval[d] = val[o];
        }
    free(next);
}

/* ------------------------------------------------------------------ */
/* the input matrix A, dense or sparse                                 */
/* ------------------------------------------------------------------ */
typedef struct
{
    int sparse;
    int m, n;
    const double* a; int lda;                                 /* dense */
    const unsigned* colp; const unsigned* rowi; const double* val;   /* CSC */
    unsigned* tcolp; unsigned* trowi; double* tval;            /* CSC of A' (BPP only) */
} MatA;

/* WtA(k x n) = W' * A */
static void prod_WtA(const MatA* A, int k, const double* W, int ldw, double* WtA, int ld)
{
    if (!A->sparse) gemm(1, 0, k, A->n, A->m, W, ldw, A->a, A->lda, WtA, ld);
    else
    {
        /* B'*A with tight copies is not needed: AT() handles ld */
        for (int j = 0; j < A->n; ++j)
        {
            double* cj = WtA + (size_t)j * ld;
            for (int c = 0; c < k; ++c) cj[c] = 0.0;
            for (unsigned o = A->colp[j]; o != A->colp[j + 1]; ++o)
            {
                double aa = 1.0 * A->val[o];
                unsigned row = A->rowi[o];
                for (int c = 0; c < k; ++c) cj[c] += aa * AT(W, ldw, row, c);
            }
        }
    }
}

/* AHt(m x k) = A * H' */
static void prod_AHt(const MatA* A, int k, const double* H, int ldh, double* AHt, int ld)
{
    if (!A->sparse) gemm(0, 1, A->m, k, A->n, A->a, A->lda, H, ldh, AHt, ld);
    else
    {
        for (int j = 0; j < k; ++j)
        {
            double* cj = AHt + (size_t)j * ld;
            for (int r = 0; r < A->m; ++r) cj[r] = 0.0;
            for (int c = 0; c < A->n; ++c)
            {
                double ab = 1.0 * AT(H, ldh, j, c);
                for (unsigned o = A->colp[c]; o != A->colp[c + 1]; ++o) cj[A->rowi[o]] += ab * A->val[o];
            }
        }
    }
}

/* HAt(k x m) = H * A'  (BPP's W-side right-hand side, nmf_solver_bpp.hpp:356) */
static void prod_HAt(const MatA* A, int k, const double* H, int ldh, double* HAt, int ld)
{
    if (!A->sparse) gemm(0, 1, k, A->m, A->n, H, ldh, A->a, A->lda, HAt, ld);
    else
    {
        /* B*At over the stable transpose: sparse_gemm_ba_impl.hpp:26-78 */
        for (int j = 0; j < A->m; ++j)
        {
            double* cj = HAt + (size_t)j * ld;
            for (int r = 0; r < k; ++r) cj[r] = 0.0;
            for (unsigned o = A->tcolp[j]; o != A->tcolp[j + 1]; ++o)
            {
                double aa = 1.0 * A->tval[o];
                const double* b = H + (size_t)A->trowi[o] * ldh;
                for (int r = 0; r < k; ++r) cj[r] += aa * b[r];
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* NNLS by block principal pivoting                                    */
/* ------------------------------------------------------------------ */

/* Upper Cholesky, right-looking (Elemental UVar3Unb, UVar3.hpp:17-44), then
 * U'y=b, Ux=y (cholesky::SolveAfter). L is p x p with ld p and is overwritten.
 * Returns 0 on a non-positive pivot (normal_eq.hpp:35-50 -> solver failure). */
static int hpd_solve(int p, double* L, double* b)
{
    for (int j = 0; j < p; ++j)
    {
        double alpha = AT(L, p, j, j);
        if (alpha <= 0.0) return 0;
        alpha = sqrt(alpha);
        AT(L, p, j, j) = alpha;
        for (int c = j + 1; c < p; ++c) AT(L, p, j, c) /= alpha;
        for (int c = j + 1; c < p; ++c)
            for (int i = j + 1; i <= c; ++i)
                AT(L, p, i, c) -= AT(L, p, j, i) * AT(L, p, j, c);
    }
    for (int i = 0; i < p; ++i)
    {
        double s = b[i];
        for (int q = 0; q < i; ++q) s -= AT(L, p, q, i) * b[q];
        b[i] = s / AT(L, p, i, i);
    }
    for (int i = p - 1; i >= 0; --i)
    {
        double s = b[i];
        for (int q = i + 1; q < p; ++q) s -= AT(L, p, i, q) * b[q];
        b[i] = s / AT(L, p, i, i);
    }
    return 1;
}

/* BitMatrix::MaxRowIndex INCLUDING its off-by-one-word defect for k > 32
 * (common/src/bit_matrix.cpp:432-472): a highest set bit found in a full
 * 32-row word other than word 0 is reported 32 rows too low; an empty column
 * reports 0. bits[] is one byte per row. */
static unsigned max_row_index(const unsigned char* bits, int k)
{
    int full = k / 32, extra = k - 32 * full;
    int ldw = full + (extra ? 1 : 0);
    int w = ldw - 1;
    if (extra > 0)
    {
        for (int q = extra - 1; q >= 0; --q)
            if (bits[full * 32 + q]) return (unsigned)(full * 32 + q);
        --w;
    }
    for (; w >= 0; --w)
        for (int q = 31; q >= 0; --q)
            if (bits[w * 32 + q]) return (unsigned)(w > 0 ? (w - 1) * 32 + q : q);
    return 0;
}

/* Work counters (diagnostics for sizing the GPU kernel; not part of the algorithm). */
static double g_stat_solves = 0.0, g_stat_p3 = 0.0, g_stat_rounds = 0.0, g_stat_calls = 0.0, g_stat_backup = 0.0;
double orc_backup_count(void) { return g_stat_backup; }     /* firings of the backup rule since orc_stats_reset */
void orc_stats_get(double* out4) { out4[0] = g_stat_solves; out4[1] = g_stat_p3; out4[2] = g_stat_rounds; out4[3] = g_stat_calls; }
void orc_stats_reset(void) { g_stat_solves = g_stat_p3 = g_stat_rounds = g_stat_calls = g_stat_backup = 0.0; }

/* BppSolveNormalEqNoGroup (nmf_solver_bpp.hpp:146-219) for the listed columns.
 * X columns are fully overwritten (zeros off the passive set). */
static int bpp_solve(int k, const double* LHS, const double* RHS, double* X,
                     const unsigned char* passive, const int* cols, int ncols, double* Lsub, double* bsub, int* ri)
{
    int ok = 1;
    for (int t = 0; t < ncols; ++t)
    {
        int c = cols[t];
        const unsigned char* pc = passive + (size_t)c * k;
        double* x = X + (size_t)c * k;
        int p = 0;
        for (int r = 0; r < k; ++r) { x[r] = 0.0; if (pc[r]) ri[p++] = r; }
        if (p == 0) continue;
        g_stat_solves += 1.0; g_stat_p3 += (double)p * p * p;
        for (int b = 0; b < p; ++b)
            for (int a = 0; a < p; ++a) AT(Lsub, p, a, b) = AT(LHS, k, ri[a], ri[b]);
        for (int a = 0; a < p; ++a) bsub[a] = AT(RHS, k, ri[a], c);
        if (!hpd_solve(p, Lsub, bsub)) { ok = 0; continue; }
        for (int a = 0; a < p; ++a) x[ri[a]] = bsub[a];
    }
    return ok;
}

/* NnlsBlockpivot (nnls.hpp:144-244) with UpdatePassiveSet (common/src/nnls.cpp:18-74).
 * LHS k x k, RHS/X/Y k x q, tight leading dimensions. X holds the warm start on entry. */
int orc_nnls_bpp(int k, int q, const double* LHS, const double* RHS, double* X, double* Y)
{
    const int PBAR = 3;
    const unsigned MAX_ITER = (unsigned)k * 5u;
    size_t kq = (size_t)k * q;
    unsigned char* passive = (unsigned char*)malloc(kq ? kq : 1);
    unsigned char* nonopt = (unsigned char*)malloc(kq ? kq : 1);
    unsigned char* infeas = (unsigned char*)malloc(kq ? kq : 1);
    int* P = (int*)malloc(sizeof(int) * (q ? q : 1));
    int* Ninf = (int*)malloc(sizeof(int) * (q ? q : 1));
    int* notgood = (int*)malloc(sizeof(int) * (q ? q : 1));
    int* cols = (int*)malloc(sizeof(int) * (q ? q : 1));
    double* Lsub = (double*)malloc(sizeof(double) * k * k);
    double* bsub = (double*)malloc(sizeof(double) * k);
    int* ri = (int*)malloc(sizeof(int) * k);
    int rc = ORC_OK;

    /* passive_set = (X > 0); X = 0; solve every column (nnls.hpp:157-165) */
    for (size_t i = 0; i < kq; ++i) passive[i] = (X[i] > 0.0);
    for (int c = 0; c < q; ++c) cols[c] = c;
    if (!bpp_solve(k, LHS, RHS, X, passive, cols, q, Lsub, bsub, ri)) { rc = ORC_FAILURE; goto done; }
    gemm_minus(k, q, LHS, k, X, k, RHS, k, Y, k);

    int nno = 0;
    for (int c = 0; c < q; ++c)
    {
        P[c] = PBAR; Ninf[c] = k + 1;
        int cnt = 0;
        for (int r = 0; r < k; ++r)
        {
            size_t i = (size_t)c * k + r;
            nonopt[i] = (Y[i] < 0.0) && !passive[i];
            infeas[i] = (X[i] < 0.0) && passive[i];
            cnt += nonopt[i] + infeas[i];
        }
        notgood[c] = cnt;
        if (cnt > 0) cols[nno++] = c;
    }

    unsigned iter = 0;
    g_stat_calls += 1.0;
    while (nno > 0)
    {
        g_stat_rounds += 1.0;
        if (iter >= MAX_ITER) { rc = ORC_FAILURE; goto done; }

        /* UpdatePassiveSet over the non-optimal columns */
        for (int t = 0; t < nno; ++t)
        {
            int c = cols[t];
            unsigned char* pc = passive + (size_t)c * k;
            const unsigned char* no = nonopt + (size_t)c * k;
            const unsigned char* in = infeas + (size_t)c * k;
            if (notgood[c] < Ninf[c])
            {
                P[c] = PBAR; Ninf[c] = notgood[c];
                for (int r = 0; r < k; ++r) { if (no[r]) pc[r] = 1; if (in[r]) pc[r] = 0; }
            }
            else if (P[c] >= 1)
            {
                P[c] -= 1;
                for (int r = 0; r < k; ++r) { if (no[r]) pc[r] = 1; if (in[r]) pc[r] = 0; }
            }
            else
            {
                unsigned r1 = max_row_index(no, k), r2 = max_row_index(in, k);
                unsigned row = r1 > r2 ? r1 : r2;
                pc[row] = !pc[row];
                g_stat_backup += 1.0;
            }
        }

        if (!bpp_solve(k, LHS, RHS, X, passive, cols, nno, Lsub, bsub, ri)) { rc = ORC_FAILURE; goto done; }

        /* ZeroizeSmallValues(Xsub); Ysub = LHS*Xsub - RHSsub (nnls.hpp:215-220) */
        for (int t = 0; t < nno; ++t)
        {
            int c = cols[t];
            double* x = X + (size_t)c * k;
            for (int r = 0; r < k; ++r) if (fabs(x[r]) < 1.0e-12) x[r] = 0.0;
            gemm_minus(k, 1, LHS, k, x, k, RHS + (size_t)c * k, k, Y + (size_t)c * k, k);
        }
        /* ZeroizeSmallValues on ALL of X and Y (nnls.hpp:226-227) */
        for (size_t i = 0; i < kq; ++i)
        {
            if (fabs(X[i]) < 1.0e-12) X[i] = 0.0;
            if (fabs(Y[i]) < 1.0e-12) Y[i] = 0.0;
        }
        /* BppUpdateSets restricted to the columns that were non-optimal (nnls.hpp:42-140) */
        int nn2 = 0;
        for (int t = 0; t < nno; ++t)
        {
            int c = cols[t];
            int cnt = 0;
            for (int r = 0; r < k; ++r)
            {
                size_t i = (size_t)c * k + r;
                nonopt[i] = (Y[i] < 0.0) && !passive[i];
                infeas[i] = (X[i] < 0.0) && passive[i];
                cnt += nonopt[i] + infeas[i];
            }
            notgood[c] = cnt;
            if (cnt > 0) cols[nn2++] = c;       /* stays ascending */
        }
        nno = nn2;
        ++iter;
    }

done:
    free(passive); free(nonopt); free(infeas); free(P); free(Ninf); free(notgood); free(cols);
    free(Lsub); free(bsub); free(ri);
    return rc;
}

/* ------------------------------------------------------------------ */
/* normalisation, progress metrics                                     */
/* ------------------------------------------------------------------ */

/* NormalizeColumns + ScaleRows (normalize.hpp:25-161). Returns 0 when a column
 * norm is below machine epsilon (the reference throws). */
static int normalize_and_scale(int m, int n, int k, double* W, int ldw, double* H, int ldh, double* norms)
{
    for (int c = 0; c < k; ++c)
    {
        double* w = W + (size_t)c * ldw;
        double nr = nrm2(w, m);
        if (fabs(nr) < DBL_EPSILON) return 0;
        double inv = 1.0 / nr;
        for (int r = 0; r < m; ++r) w[r] *= inv;
        norms[c] = nr;
    }
    for (int r = 0; r < k; ++r)
        for (int c = 0; c < n; ++c) AT(H, ldh, r, c) *= norms[r];
    return 1;
}

/* ProjectedGradientNorm (projected_gradient.hpp:125-171) */
static double pg_norm(int m, int n, int k, const double* gW, int ldgw, const double* gH, int ldgh,
                      const double* W, int ldw, const double* H, int ldh)
{
    double sw = 0.0, sh = 0.0;
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < m; ++r)
        {
            double g = AT(gW, ldgw, r, c);
            if (g < 0.0 || AT(W, ldw, r, c) > 0.0) sw += g * g;
        }
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < k; ++r)
        {
            double g = AT(gH, ldgh, r, c);
            if (g < 0.0 || AT(H, ldh, r, c) > 0.0) sh += g * g;
        }
    return sqrt(sw + sh);
}

/* ------------------------------------------------------------------ */
/* solver state and the four update rules                              */
/* ------------------------------------------------------------------ */
typedef struct
{
    int m, n, k;
    double *WtW, *HHt;          /* k x k */
    double *WtA;                /* k x n */
    double *AHt;                /* m x k */
    double *Wt, *gradWt, *HAt;  /* k x m (BPP) */
    double *tmpn, *tmpm;        /* n, m */
    double *WtWH, *WHHt;        /* MU */
    double scale[2];
} Work;

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

static void work_alloc(Work* s, int m, int n, int k)
{
    memset(s, 0, sizeof(*s));
    s->m = m; s->n = n; s->k = k;
    s->WtW = dalloc((size_t)k * k); s->HHt = dalloc((size_t)k * k);
    s->WtA = dalloc((size_t)k * n); s->AHt = dalloc((size_t)m * k);
    s->Wt = dalloc((size_t)k * m); s->gradWt = dalloc((size_t)k * m); s->HAt = dalloc((size_t)k * m);
    s->tmpn = dalloc(n); s->tmpm = dalloc(m);
    s->WtWH = dalloc((size_t)k * n); s->WHHt = dalloc((size_t)m * k);
}
static void work_free(Work* s)
{
    free(s->WtW); free(s->HHt); free(s->WtA); free(s->AHt); free(s->Wt); free(s->gradWt); free(s->HAt);
    free(s->tmpn); free(s->tmpm); free(s->WtWH); free(s->WHHt);
}

/* --- BPP: nmf_solver_bpp.hpp:310-377 --- */
static void bpp_init(Work* s, const MatA* A, const double* W, int ldw)
{
    int m = s->m, k = s->k;
    gemm(1, 0, k, k, m, W, ldw, W, ldw, s->WtW, k);
    prod_WtA(A, k, W, ldw, s->WtA, k);
    for (int c = 0; c < k; ++c) for (int r = 0; r < m; ++r) AT(s->Wt, k, c, r) = AT(W, ldw, r, c);
}
static int bpp_step(Work* s, const MatA* A, double* W, int ldw, double* H, int ldh,
                    double* gW, int ldgw, double* gH, int ldgh)
{
    int m = s->m, n = s->n, k = s->k;
    /* H and gradH may carry a leading dimension > k: work on tight copies */
    double* Ht = (ldh == k) ? H : dalloc((size_t)k * n);
    double* gHt = (ldgh == k) ? gH : dalloc((size_t)k * n);
    if (Ht != H) for (int c = 0; c < n; ++c) memcpy(Ht + (size_t)c * k, H + (size_t)c * ldh, sizeof(double) * k);
    int rc = orc_nnls_bpp(k, n, s->WtW, s->WtA, Ht, gHt);
    if (Ht != H) { for (int c = 0; c < n; ++c) memcpy(H + (size_t)c * ldh, Ht + (size_t)c * k, sizeof(double) * k); }
    if (rc != ORC_OK) { if (Ht != H) free(Ht); if (gHt != gH) free(gHt); return 0; }
    gemm(0, 1, k, k, n, H, ldh, H, ldh, s->HHt, k);
    prod_HAt(A, k, H, ldh, s->HAt, k);
    if (orc_nnls_bpp(k, m, s->HHt, s->HAt, s->Wt, s->gradWt) != ORC_OK) { if (Ht != H) free(Ht); if (gHt != gH) free(gHt); return 0; }
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < m; ++r) { AT(W, ldw, r, c) = AT(s->Wt, k, c, r); AT(gW, ldgw, r, c) = AT(s->gradWt, k, c, r); }
    gemm(1, 0, k, k, m, W, ldw, W, ldw, s->WtW, k);
    prod_WtA(A, k, W, ldw, s->WtA, k);
    gemm_minus(k, n, s->WtW, k, H, ldh, s->WtA, k, gH, ldgh);
    if (Ht != H) free(Ht);
    if (gHt != gH) free(gHt);
    return 1;
}

/* --- HALS: nmf_solver_hals.hpp:26-207 --- */
static void hals_init(Work* s, const MatA* A, const double* H, int ldh)
{
    gemm(0, 1, s->k, s->k, s->n, H, ldh, H, ldh, s->HHt, s->k);
    prod_AHt(A, s->k, H, ldh, s->AHt, s->m);
}
static void hals_update_W(Work* s, double* W, int ldw)
{
    int m = s->m, k = s->k;
    for (int c = 0; c < k; ++c)
    {
        /* WHHt_c = W * HHt(:,c)  (Gemv NORMAL) */
        for (int r = 0; r < m; ++r)
        {
            double t = 0.0;
            for (int p = 0; p < k; ++p) t += AT(W, ldw, r, p) * AT(s->HHt, k, p, c);
            s->tmpm[r] = t;
        }
        double d = AT(s->HHt, k, c, c);
        int zeros = 0;
        for (int r = 0; r < m; ++r)
        {
            double w = AT(W, ldw, r, c);
            w = w + (AT(s->AHt, m, r, c) - s->tmpm[r]) / d;
            if (isnan(w) || w < 0.0) { w = 0.0; ++zeros; }
            AT(W, ldw, r, c) = w;
        }
        if (zeros == m) for (int r = 0; r < m; ++r) AT(W, ldw, r, c) = DBL_EPSILON;
        double nr = fro_norm(W + (size_t)c * ldw, ldw, m, 1);
        double inv = 1.0 / nr;
        for (int r = 0; r < m; ++r) AT(W, ldw, r, c) *= inv;
    }
}
static void hals_update_H(Work* s, double* H, int ldh)
{
    int n = s->n, k = s->k;
    for (int r = 0; r < k; ++r)
    {
        /* WtWH_r = H' * WtW(r,:)'  (Gemv TRANSPOSE) */
        for (int c = 0; c < n; ++c)
        {
            double t = 0.0;
            for (int p = 0; p < k; ++p) t += AT(H, ldh, p, c) * AT(s->WtW, k, r, p);
            s->tmpn[c] = t;
        }
        double d = AT(s->WtW, k, r, r);
        for (int c = 0; c < n; ++c)
        {
            double h = AT(H, ldh, r, c);
            h = h + (AT(s->WtA, k, r, c) - s->tmpn[c]) / d;
            if (isnan(h) || h < 0.0) h = 0.0;
            AT(H, ldh, r, c) = h;
        }
    }
}
static int hals_step(Work* s, const MatA* A, double* W, int ldw, double* H, int ldh,
                     double* gW, int ldgw, double* gH, int ldgh)
{
    int m = s->m, n = s->n, k = s->k;
    hals_update_W(s, W, ldw);
    gemm(1, 0, k, k, m, W, ldw, W, ldw, s->WtW, k);
    prod_WtA(A, k, W, ldw, s->WtA, k);
    hals_update_H(s, H, ldh);
    gemm_minus(k, n, s->WtW, k, H, ldh, s->WtA, k, gH, ldgh);
    gemm(0, 1, k, k, n, H, ldh, H, ldh, s->HHt, k);
    prod_AHt(A, k, H, ldh, s->AHt, m);
    /* gradW = W*HHt - AHt */
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < m; ++r)
        {
            double t = 0.0;
            for (int p = 0; p < k; ++p) t += AT(W, ldw, r, p) * AT(s->HHt, k, p, c);
            AT(gW, ldgw, r, c) = t + (-1.0) * AT(s->AHt, m, r, c);
        }
    return 1;
}

/* --- MU: nmf_solver_mu.hpp:27-169 --- */
static void mu_init(Work* s, const MatA* A, const double* W, int ldw)
{
    prod_WtA(A, s->k, W, ldw, s->WtA, s->k);
    gemm(1, 0, s->k, s->k, s->m, W, ldw, W, ldw, s->WtW, s->k);
}
static int mu_step(Work* s, const MatA* A, double* W, int ldw, double* H, int ldh,
                   double* gW, int ldgw, double* gH, int ldgh)
{
    int m = s->m, n = s->n, k = s->k;
    const double EPS = 1.0e-13;
    gemm(0, 0, k, n, k, s->WtW, k, H, ldh, s->WtWH, k);
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < k; ++r)
            AT(H, ldh, r, c) *= (AT(s->WtA, k, r, c) / (AT(s->WtWH, k, r, c) + EPS));
    gemm(0, 1, k, k, n, H, ldh, H, ldh, s->HHt, k);
    prod_AHt(A, k, H, ldh, s->AHt, m);
    gemm(0, 0, m, k, k, W, ldw, s->HHt, k, s->WHHt, m);
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < m; ++r)
            AT(W, ldw, r, c) *= (AT(s->AHt, m, r, c) / (AT(s->WHHt, m, r, c) + EPS));
    prod_WtA(A, k, W, ldw, s->WtA, k);
    gemm(1, 0, k, k, m, W, ldw, W, ldw, s->WtW, k);
    for (int c = 0; c < k; ++c)
        for (int r = 0; r < m; ++r)
        {
            double t = 0.0;
            for (int p = 0; p < k; ++p) t += AT(W, ldw, r, p) * AT(s->HHt, k, p, c);
            AT(gW, ldgw, r, c) = t + (-1.0) * AT(s->AHt, m, r, c);
        }
    gemm_minus(k, n, s->WtW, k, H, ldh, s->WtA, k, gH, ldgh);
    return 1;
}

/* --- RANK2: nmf_solver_rank2.hpp:25-461 --- */
/* Solve G*x = b column-wise for H (2 x n); G = WtW. (SystemSolveH :25-135) */
static int rank2_solve_H(int n, const double* G, double* X, int ldx, const double* B, int ldb)
{
    double a00 = AT(G, 2, 0, 0), a01 = AT(G, 2, 0, 1), a10 = AT(G, 2, 1, 0), a11 = AT(G, 2, 1, 1);
    double eps = DBL_EPSILON;
    if (fabs(a00) < eps && fabs(a01) < eps) return 0;
    double a2, b2, d2, ia2, id2;
    if (fabs(a00) >= fabs(a01))
    {
        double t = -a10 / a00;
        a2 = a00 - t * a10; b2 = a01 - t * a11; d2 = a11 + t * a01;
        ia2 = 1.0 / a2; id2 = 1.0 / d2;
        if (fabs(d2 / a2) < eps) return 0;
        for (int i = 0; i < n; ++i)
        {
            double e2 = AT(B, ldb, 0, i) - t * AT(B, ldb, 1, i);
            double f2 = AT(B, ldb, 1, i) + t * AT(B, ldb, 0, i);
            double x1 = f2 * id2;
            AT(X, ldx, 1, i) = x1;
            AT(X, ldx, 0, i) = (e2 - b2 * x1) * ia2;
        }
    }
    else
    {
        double ct = -a00 / a10;
        a2 = -a10 + ct * a00; b2 = -a11 + ct * a01; d2 = a01 + ct * a11;
        ia2 = 1.0 / a2; id2 = 1.0 / d2;
        if (fabs(d2 / a2) < eps) return 0;
        for (int i = 0; i < n; ++i)
        {
            double e2 = -AT(B, ldb, 1, i) + ct * AT(B, ldb, 0, i);
            double f2 = AT(B, ldb, 0, i) + ct * AT(B, ldb, 1, i);
            double x1 = f2 * id2;
            AT(X, ldx, 1, i) = x1;
            AT(X, ldx, 0, i) = (e2 - b2 * x1) * ia2;
        }
    }
    return 1;
}
/* Solve x*G = b row-wise for W (m x 2); G = HHt. (SystemSolveW :139-214) */
static int rank2_solve_W(int m, const double* G, double* X, int ldx, const double* B, int ldb)
{
    double a00 = AT(G, 2, 0, 0), a01 = AT(G, 2, 0, 1), a10 = AT(G, 2, 1, 0), a11 = AT(G, 2, 1, 1);
    double eps = DBL_EPSILON;
    if (fabs(a00) < eps && fabs(a01) < eps) return 0;
    double a2, b2, d2, ia2, id2;
    if (fabs(a00) >= fabs(a01))
    {
        double t = a01 / a00;
        a2 = a00 + t * a01; b2 = a10 + t * a11; d2 = a11 - t * a10;
        ia2 = 1.0 / a2; id2 = 1.0 / d2;
        if (fabs(d2 / a2) < eps) return 0;
        for (int i = 0; i < m; ++i)
        {
            double e2 = AT(B, ldb, i, 0) + t * AT(B, ldb, i, 1);
            double f2 = AT(B, ldb, i, 1) - t * AT(B, ldb, i, 0);
            double x1 = f2 * id2;
            AT(X, ldx, i, 1) = x1;
            AT(X, ldx, i, 0) = (e2 - b2 * x1) * ia2;
        }
    }
    else
    {
        double ct = a00 / a01;
        a2 = -a01 - ct * a00; b2 = -a11 - ct * a10; d2 = a10 - ct * a11;
        ia2 = 1.0 / a2; id2 = 1.0 / d2;
        if (fabs(d2 / a2) < eps) return 0;
        for (int i = 0; i < m; ++i)
        {
            double e2 = -AT(B, ldb, i, 1) - ct * AT(B, ldb, i, 0);
            double f2 = AT(B, ldb, i, 0) - ct * AT(B, ldb, i, 1);
            double x1 = f2 * id2;
            AT(X, ldx, i, 1) = x1;
            AT(X, ldx, i, 0) = (e2 - b2 * x1) * ia2;
        }
    }
    return 1;
}
static void rank2_init(Work* s, const MatA* A, const double* W, int ldw)
{
    gemm(1, 0, 2, 2, s->m, W, ldw, W, ldw, s->WtW, 2);
    prod_WtA(A, 2, W, ldw, s->WtA, 2);
}
static int rank2_step(Work* s, const MatA* A, double* W, int ldw, double* H, int ldh,
                      double* gW, int ldgw, double* gH, int ldgh)
{
    int m = s->m, n = s->n;
    if (!rank2_solve_H(n, s->WtW, H, ldh, s->WtA, 2)) return 0;
    {   /* OptimalActiveSetH :218-268 */
        double g0 = AT(s->WtW, 2, 0, 0), g1 = AT(s->WtW, 2, 1, 1);
        double i0 = 1.0 / g0, i1 = 1.0 / g1, q0 = sqrt(g0), q1 = sqrt(g1);
        for (int i = 0; i < n; ++i)
        {
            double v1 = AT(s->WtA, 2, 0, i) * i0, v2 = AT(s->WtA, 2, 1, i) * i1;
            double vv1 = v1 * q0, vv2 = v2 * q1;
            if (vv1 >= vv2) v2 = 0.0; else v1 = 0.0;
            if (AT(H, ldh, 0, i) <= 0.0 || AT(H, ldh, 1, i) <= 0.0) { AT(H, ldh, 0, i) = v1; AT(H, ldh, 1, i) = v2; }
        }
    }
    gemm(0, 1, 2, 2, n, H, ldh, H, ldh, s->HHt, 2);
    prod_AHt(A, 2, H, ldh, s->AHt, m);
    if (!rank2_solve_W(m, s->HHt, W, ldw, s->AHt, m)) return 0;
    {   /* OptimalActiveSetW :272-318 */
        double g0 = AT(s->HHt, 2, 0, 0), g1 = AT(s->HHt, 2, 1, 1);
        double i0 = 1.0 / g0, i1 = 1.0 / g1, q0 = sqrt(g0), q1 = sqrt(g1);
        for (int i = 0; i < m; ++i)
        {
            double v1 = AT(s->AHt, m, i, 0) * i0, v2 = AT(s->AHt, m, i, 1) * i1;
            double vv1 = v1 * q0, vv2 = v2 * q1;
            if (vv1 >= vv2) v2 = 0.0; else v1 = 0.0;
            if (AT(W, ldw, i, 0) <= 0.0 || AT(W, ldw, i, 1) <= 0.0) { AT(W, ldw, i, 0) = v1; AT(W, ldw, i, 1) = v2; }
        }
    }
    if (!normalize_and_scale(m, n, 2, W, ldw, H, ldh, s->scale)) return -1;
    {   /* :422-441 analytic rescale of HHt and AHt */
        double s0 = s->scale[0], s1 = s->scale[1];
        double e00 = AT(s->HHt, 2, 0, 0), e01 = AT(s->HHt, 2, 0, 1), e11 = AT(s->HHt, 2, 1, 1);
        AT(s->HHt, 2, 0, 0) = e00 * s0 * s0;
        AT(s->HHt, 2, 0, 1) = e01 * s0 * s1;
        AT(s->HHt, 2, 1, 0) = e01 * s0 * s1;
        AT(s->HHt, 2, 1, 1) = e11 * s1 * s1;
        for (int i = 0; i < m; ++i) { AT(s->AHt, m, i, 0) *= s0; AT(s->AHt, m, i, 1) *= s1; }
    }
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < m; ++r)
        {
            double t = 0.0;
            for (int p = 0; p < 2; ++p) t += AT(W, ldw, r, p) * AT(s->HHt, 2, p, c);
            AT(gW, ldgw, r, c) = t + (-1.0) * AT(s->AHt, m, r, c);
        }
    gemm(1, 0, 2, 2, m, W, ldw, W, ldw, s->WtW, 2);
    prod_WtA(A, 2, W, ldw, s->WtA, 2);
    gemm_minus(2, n, s->WtW, 2, H, ldh, s->WtA, 2, gH, ldgh);
    return 1;
}

/* ------------------------------------------------------------------ */
/* NmfSolve (nmf_solve_generic.hpp:30-140)                              */
/* ------------------------------------------------------------------ */
static int nmf_solve(const MatA* A, int alg, int prog, int k, double tol, int min_iter, int max_iter,
                     int tolcount, int normalize, double* W, int ldw, double* H, int ldh,
                     int* iterations, double* metrics, double* Wsnap, double* Hsnap)
{
    int m = A->m, n = A->n;
    if (alg == ALG_RANK2 && k != 2) return ORC_BAD_PARAM;
    if (k > n || !(tol > 0.0 && tol < 1.0)) return ORC_BAD_PARAM;      /* nmf_options.cpp:24-111 */
    Work s; work_alloc(&s, m, n, k);
    double* gW = dalloc((size_t)m * k);
    double* gH = dalloc((size_t)k * n);
    double* Wprev = (prog == PROG_DELTA_FNORM) ? dalloc((size_t)m * k) : NULL;
    double pg0 = 0.0;
    int rc = ORC_OK;

    switch (alg)
    {
    case ALG_BPP: bpp_init(&s, A, W, ldw); break;
    case ALG_HALS: hals_init(&s, A, H, ldh); break;
    case ALG_MU: mu_init(&s, A, W, ldw); break;
    case ALG_RANK2: rank2_init(&s, A, W, ldw); break;
    default: rc = ORC_BAD_PARAM; goto done;
    }
    if (Wprev)
    {   /* ProgEstGenericDeltaW::Init: Wprev = 0; Compute(W) */
        for (int c = 0; c < k; ++c) for (int r = 0; r < m; ++r) AT(Wprev, m, r, c) = AT(W, ldw, r, c);
    }
    if (metrics) for (int i = 0; i < max_iter; ++i) metrics[i] = NAN;

    int iter = 0, success = 0, success_count = 0;
    for (iter = 0; iter < max_iter; ++iter)
    {
        int ok;
        switch (alg)
        {
        case ALG_BPP: ok = bpp_step(&s, A, W, ldw, H, ldh, gW, m, gH, k); break;
        case ALG_HALS: ok = hals_step(&s, A, W, ldw, H, ldh, gW, m, gH, k); break;
        case ALG_MU: ok = mu_step(&s, A, W, ldw, H, ldh, gW, m, gH, k); break;
        default: ok = rank2_step(&s, A, W, ldw, H, ldh, gW, m, gH, k); break;
        }
        if (ok <= 0) { rc = (ok < 0) ? -100 : ORC_FAILURE; *iterations = iter; goto done; }

        int evaluate = (iter >= min_iter) || (iter == 0);
        double metric = 1.0;
        if (evaluate)
        {
            if (prog == PROG_PG_RATIO)
            {
                double pg = pg_norm(m, n, k, gW, m, gH, k, W, ldw, H, ldh);
                if (isnan(pg)) { rc = -100; *iterations = iter; goto done; }
                if (iter == 0) { pg0 = pg; metric = 1.0; } else metric = pg / pg0;
            }
            else
            {   /* progress_estimator_generic.hpp:58-69 */
                for (int c = 0; c < k; ++c) for (int r = 0; r < m; ++r) AT(Wprev, m, r, c) += (-1.0) * AT(W, ldw, r, c);
                double nd = fro_norm(Wprev, m, m, k), nc = fro_norm(W, ldw, m, k);
                metric = nd / nc;
                for (int c = 0; c < k; ++c) for (int r = 0; r < m; ++r) AT(Wprev, m, r, c) = AT(W, ldw, r, c);
            }
            if (metrics) metrics[iter] = metric;
            if (Wsnap) for (int c = 0; c < k; ++c) memcpy(Wsnap + ((size_t)iter * k + c) * m, W + (size_t)c * ldw, sizeof(double) * m);
            if (Hsnap) for (int c = 0; c < n; ++c) memcpy(Hsnap + ((size_t)iter * n + c) * k, H + (size_t)c * ldh, sizeof(double) * k);
        }
        if (iter < min_iter) continue;
        if (metric <= tol)
        {
            if (++success_count >= tolcount) { success = 1; break; }
        }
        else success_count = 0;
    }
    if (normalize)
    {
        double* norms = dalloc(k);
        int okn = normalize_and_scale(m, n, k, W, ldw, H, ldh, norms);
        free(norms);
        if (!okn) { rc = -100; *iterations = iter; goto done; }
    }
    if (!success && iter == max_iter) success = 1;
    *iterations = iter;
    rc = success ? ORC_OK : ORC_FAILURE;
done:
    free(gW); free(gH); free(Wprev); work_free(&s);
    return rc;
}

int orc_nmf_dense(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                  int tolcount, int normalize, const double* A, int ldA, double* W, int ldW, double* H, int ldH,
                  int* iterations, double* metrics, double* Wsnap, double* Hsnap)
{
    MatA M; memset(&M, 0, sizeof(M));
    M.sparse = 0; M.m = m; M.n = n; M.a = A; M.lda = ldA;
    int it = 0;
    int rc = nmf_solve(&M, alg, prog, k, tol, min_iter, max_iter, tolcount, normalize, W, ldW, H, ldH, &it, metrics, Wsnap, Hsnap);
    if (iterations) *iterations = it;
    return rc;
}

int orc_nmf_sparse(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                   int tolcount, int normalize, const unsigned* colp, const unsigned* rowi, const double* val,
                   double* W, int ldW, double* H, int ldH,
                   int* iterations, double* metrics, double* Wsnap, double* Hsnap)
{
    MatA M; memset(&M, 0, sizeof(M));
    M.sparse = 1; M.m = m; M.n = n; M.colp = colp; M.rowi = rowi; M.val = val;
    unsigned nnz = colp[n];
    if (alg == ALG_BPP)
    {
        M.tcolp = (unsigned*)malloc(sizeof(unsigned) * ((size_t)m + 1));
        M.trowi = (unsigned*)malloc(sizeof(unsigned) * (nnz ? nnz : 1));
        M.tval = (double*)malloc(sizeof(double) * (nnz ? nnz : 1));
        orc_csc_transpose((unsigned)m, (unsigned)n, colp, rowi, val, M.tcolp, M.trowi, M.tval);
    }
    int it = 0;
    int rc = nmf_solve(&M, alg, prog, k, tol, min_iter, max_iter, tolcount, normalize, W, ldW, H, ldH, &it, metrics, Wsnap, Hsnap);
    if (iterations) *iterations = it;
    free(M.tcolp); free(M.trowi); free(M.tval);
    return rc;
}
