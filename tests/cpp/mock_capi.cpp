// TEST INFRASTRUCTURE, not product code: the entry points of include/smallk_b200.h that the host layer calls, answered on the CPU
// by the oracle (oracle/libsmallk_oracle.so: the C restatement of the reference's solvers). Built by tests/test_host_driver_cpu.py
// into tests/cpp/build/, together with a copy of the host library linked against it, so that the code ABOVE the C ABI — the
// hierclust tree driver with its worker threads, its node factors kept on their own rows and its splits made ahead of time — runs
// in the CPU suite and is held to the reference's own trees (tests/golden/hier_*.npz). Nothing in smallk_b200/ links this.
//
// smk_select_columns restates SubMatrixColsCompact (common/include/sparse_matrix_impl.hpp:479-591: the listed columns in list
// order, rows left without entries dropped, the rest renumbered in ascending order; dense_matrix_impl.hpp:224-285: all rows kept).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/smallk_b200.h"

extern "C" {
int orc_nmf_dense(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter, int tolcount, int normalize,
                  const double* A, int ldA, double* W, int ldW, double* H, int ldH, int* iterations, double* metrics, double* Wsnap,
                  double* Hsnap);
int orc_nmf_sparse(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter, int tolcount, int normalize,
                   const unsigned* colp, const unsigned* rowi, const double* val, double* W, int ldW, double* H, int ldH,
                   int* iterations, double* metrics, double* Wsnap, double* Hsnap);
}

struct smk_ctx
{
    std::string err;
    int device = 0;
    bool dense = false, sparse = false, sub = false;
    int m = 0, n = 0;                       // the loaded matrix
    std::vector<double> A;                  // dense, tight
    std::vector<unsigned> colp, rowi;
    std::vector<double> val;
    int am = 0, an = 0;                     // the active subset
    std::vector<double> subA;
    std::vector<unsigned> scolp, srowi;
    std::vector<double> sval;
};

static int fail(smk_ctx* c, int code, const char* msg) { c->err = msg; return code; }

extern "C" {

int smk_create(smk_ctx** out, int device)
{
    if (!out) return SMK_BAD_PARAM;
    *out = new smk_ctx();
    (*out)->device = device;
    return SMK_OK;
}
void smk_destroy(smk_ctx* c) { delete c; }
const char* smk_last_error(const smk_ctx* c) { return c ? c->err.c_str() : "null context"; }
int smk_device_index(const smk_ctx* c) { return c ? c->device : -1; }

int smk_load_dense(smk_ctx* c, const double* A, long long ldA, int m, int n)
{
    if (!c || !A || m <= 0 || n <= 0 || ldA < m) return SMK_BAD_PARAM;
    c->dense = true; c->sparse = false; c->sub = false; c->m = m; c->n = n;
    c->A.resize(static_cast<size_t>(m) * n);
    for (int j = 0; j < n; ++j) std::memcpy(&c->A[static_cast<size_t>(j) * m], A + static_cast<size_t>(j) * ldA, sizeof(double) * m);
    return SMK_OK;
}

int smk_load_csc(smk_ctx* c, int m, int n, unsigned int nnz, const unsigned int* colp, const unsigned int* rowi, const double* val)
{
    if (!c || !colp || m <= 0 || n <= 0) return SMK_BAD_PARAM;
    c->sparse = true; c->dense = false; c->sub = false; c->m = m; c->n = n;
    c->colp.assign(colp, colp + n + 1);
    c->rowi.assign(rowi, rowi + nnz);
    c->val.assign(val, val + nnz);
    return SMK_OK;
}

int smk_select_all(smk_ctx* c)
{
    if (!c) return SMK_BAD_PARAM;
    c->sub = false;
    return SMK_OK;
}

int smk_select_columns(smk_ctx* c, const unsigned int* cols, int count, int* new_height, unsigned int* new_to_old)
{
    if (!c || !cols || !new_height || !new_to_old) return SMK_BAD_PARAM;
    if (count <= 0) return fail(c, SMK_BAD_PARAM, "SubMatrixColsCompact: empty column set");
    for (int q = 0; q < count; ++q) if (cols[q] >= static_cast<unsigned>(c->n)) return fail(c, SMK_BAD_PARAM, "SubMatrixColsCompact: column index out of range");
    if (c->dense)
    {
        c->subA.resize(static_cast<size_t>(c->m) * count);
        for (int q = 0; q < count; ++q)
            std::memcpy(&c->subA[static_cast<size_t>(q) * c->m], &c->A[static_cast<size_t>(cols[q]) * c->m], sizeof(double) * c->m);
        for (int r = 0; r < c->m; ++r) new_to_old[r] = static_cast<unsigned>(r);
        *new_height = c->m;
        c->am = c->m; c->an = count; c->sub = true;
        return SMK_OK;
    }
    if (!c->sparse) return fail(c, SMK_BAD_PARAM, "no matrix loaded");
    std::vector<unsigned char> used(c->m, 0);
    size_t total = 0;
    for (int q = 0; q < count; ++q)
        for (unsigned e = c->colp[cols[q]]; e < c->colp[cols[q] + 1]; ++e) { used[c->rowi[e]] = 1; ++total; }
    if (total == 0) return fail(c, SMK_BAD_PARAM, "SubMatrixColsCompact: all-zero submatrix");
    std::vector<unsigned> old_to_new(c->m, 0u);
    int h = 0;
    for (int r = 0; r < c->m; ++r) if (used[r]) { old_to_new[r] = static_cast<unsigned>(h); new_to_old[h] = static_cast<unsigned>(r); ++h; }
    c->scolp.assign(static_cast<size_t>(count) + 1, 0u);
    c->srowi.resize(total); c->sval.resize(total);
    size_t at = 0;
    for (int q = 0; q < count; ++q)
    {
        for (unsigned e = c->colp[cols[q]]; e < c->colp[cols[q] + 1]; ++e) { c->srowi[at] = old_to_new[c->rowi[e]]; c->sval[at] = c->val[e]; ++at; }
        c->scolp[q + 1] = static_cast<unsigned>(at);
    }
    *new_height = h;
    c->am = h; c->an = count; c->sub = true;
    return SMK_OK;
}

int smk_nmf(smk_ctx* c, const smk_nmf_options* o, double* W, int ldW, double* H, int ldH, smk_nmf_stats* stats)
{
    if (!c || !o || !W || !H) return SMK_BAD_PARAM;
    const int m = c->sub ? c->am : c->m, n = c->sub ? c->an : c->n;
    if (o->height != m || o->width != n) return fail(c, SMK_BAD_PARAM, "options do not match the active matrix");
    int it = 0, rc;
    if (c->dense)
        rc = orc_nmf_dense(o->algorithm, o->prog_est_algorithm, m, n, o->k, o->tol, o->min_iter, o->max_iter, o->tolcount, o->normalize,
                           c->sub ? c->subA.data() : c->A.data(), m, W, ldW, H, ldH, &it, nullptr, nullptr, nullptr);
    else
        rc = orc_nmf_sparse(o->algorithm, o->prog_est_algorithm, m, n, o->k, o->tol, o->min_iter, o->max_iter, o->tolcount, o->normalize,
                            c->sub ? c->scolp.data() : c->colp.data(), c->sub ? c->srowi.data() : c->rowi.data(),
                            c->sub ? c->sval.data() : c->val.data(), W, ldW, H, ldH, &it, nullptr, nullptr, nullptr);
    if (stats) { stats->elapsed_us = 0; stats->iteration_count = it; }
    if (rc == 0) return SMK_OK;
    if (rc == -100) return fail(c, SMK_FAILURE, "Normalize: column norm < machine epsilon");
    if (rc == SMK_BAD_PARAM) return fail(c, SMK_BAD_PARAM, "invalid options");
    return fail(c, SMK_FAILURE, "NMF solver failure");
}

int smk_argsort_desc(smk_ctx* c, const double* v, int n, int* order)
{
    if (!c || !v || !order || n < 0) return SMK_BAD_PARAM;
    std::iota(order, order + n, 0);
    std::stable_sort(order, order + n, [v](int a, int b) { return v[a] > v[b]; });
    return SMK_OK;
}

int smk_sort_desc(smk_ctx* c, double* v, int n)
{
    if (!c || !v || n < 0) return SMK_BAD_PARAM;
    std::sort(v, v + n, [](double a, double b) { return a > b; });
    return SMK_OK;
}

// NnlsHals (common/include/nnls.hpp:249-316; the flat-clustering step after the tree, clust_flat_generic.hpp:33-74), restated
// as in oracle/nnls_hals_oracle.py: W'W and W'A once; per iteration one HALS sweep over the rows of H (updated rows feed the later
// ones, NaN / negative -> 0), gradH = W'W H - W'A, the projected-gradient norm; stop when it falls below tol times its first value,
// then NormalizeAndScale(W, H).
int smk_nnls_hals(smk_ctx* c, int k, double* W, int ldW, double* H, int ldH, double tol, int max_iter, int* iterations)
{
    if (!c || !W || !H || k <= 0 || max_iter <= 0) return SMK_BAD_PARAM;
    const int m = c->sub ? c->am : c->m, n = c->sub ? c->an : c->n;
    if (ldW < m || ldH < k) return fail(c, SMK_BAD_PARAM, "NnlsHals: non-conformant W and H");
    std::vector<double> WtW(static_cast<size_t>(k) * k, 0.0), WtA(static_cast<size_t>(k) * n, 0.0);
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b)
        {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += W[static_cast<size_t>(a) * ldW + i] * W[static_cast<size_t>(b) * ldW + i];
            WtW[static_cast<size_t>(b) * k + a] = s;
        }
    for (int j = 0; j < n; ++j)
    {
        double* out = &WtA[static_cast<size_t>(j) * k];
        if (c->dense)
        {
            const double* col = (c->sub ? c->subA.data() : c->A.data()) + static_cast<size_t>(j) * m;
            for (int a = 0; a < k; ++a) { double s = 0.0; for (int i = 0; i < m; ++i) s += W[static_cast<size_t>(a) * ldW + i] * col[i]; out[a] = s; }
        }
        else
        {
            const unsigned* cp = c->sub ? c->scolp.data() : c->colp.data();
            const unsigned* ri = c->sub ? c->srowi.data() : c->rowi.data();
            const double* va = c->sub ? c->sval.data() : c->val.data();
            for (unsigned e = cp[j]; e < cp[j + 1]; ++e)
                for (int a = 0; a < k; ++a) out[a] += va[e] * W[static_cast<size_t>(a) * ldW + ri[e]];
        }
    }
    double pg0 = 0.0;
    for (int it = 0; it < max_iter; ++it)
    {
        for (int r = 0; r < k; ++r)
            for (int j = 0; j < n; ++j)
            {
                double* h = H + static_cast<size_t>(j) * ldH;
                double s = 0.0;
                for (int p = 0; p < k; ++p) s += WtW[static_cast<size_t>(p) * k + r] * h[p];
                double v = h[r] + (WtA[static_cast<size_t>(j) * k + r] - s) / WtW[static_cast<size_t>(r) * k + r];
                if (!(v >= 0.0)) v = 0.0;
                h[r] = v;
            }
        double sum = 0.0;
        for (int j = 0; j < n; ++j)
        {
            const double* h = H + static_cast<size_t>(j) * ldH;
            for (int r = 0; r < k; ++r)
            {
                double g = -WtA[static_cast<size_t>(j) * k + r];
                for (int p = 0; p < k; ++p) g += WtW[static_cast<size_t>(p) * k + r] * h[p];
                if (g < 0.0 || h[r] > 0.0) sum += g * g;
            }
        }
        const double pg = std::sqrt(sum);
        if (pg != pg) { if (iterations) *iterations = it + 1; return fail(c, SMK_FAILURE, "ProjectedGradientNorm: NaN"); }
        if (it == 0) { pg0 = pg; continue; }
        if (pg < tol * pg0)
        {
            if (iterations) *iterations = it + 1;
            for (int a = 0; a < k; ++a)
            {
                double s = 0.0;
                for (int i = 0; i < m; ++i) s += W[static_cast<size_t>(a) * ldW + i] * W[static_cast<size_t>(a) * ldW + i];
                const double nr = std::sqrt(s);
                if (nr < 2.220446049250313e-16) return fail(c, SMK_FAILURE, "Normalize: column norm < machine epsilon");
                for (int i = 0; i < m; ++i) W[static_cast<size_t>(a) * ldW + i] /= nr;
                for (int j = 0; j < n; ++j) H[static_cast<size_t>(j) * ldH + a] *= nr;
            }
            return SMK_OK;
        }
    }
    if (iterations) *iterations = max_iter;
    return fail(c, SMK_FAILURE, "NNLS solver reached iteration limit.");
}

} // extern "C"
