// smallk_b200 — the C ABI (include/smallk_b200.h). No exceptions leave this file.
#include <chrono>
#include <cstring>
#include <vector>
#include <cmath>

#include "context.h"
#include "solver.h"

using namespace smk;

namespace {

template <typename F>
int guarded(smk_ctx* c, F&& f)
{
    try { return f(); }
    catch (const CudaError& e)
    {
        if (c) c->err = std::string("CUDA error: ") + cudaGetErrorString(e.code) + " at " + e.file + ":" + std::to_string(e.line);
        return SMK_CUDA_ERROR;
    }
    catch (const std::string& s) { if (c) c->err = s; return SMK_BAD_PARAM; }
    catch (const std::exception& e) { if (c) c->err = e.what(); return SMK_FAILURE; }
}

int fail(smk_ctx* c, int code, const std::string& msg) { c->err = msg; return code; }

// IsValid(opts): common/src/nmf_options.cpp:24-111
bool options_valid(smk_ctx* c, const smk_nmf_options& o)
{
    if (o.k <= 0) { c->err = "k-value must be a positive integer"; return false; }
    if (o.height <= 0) { c->err = "matrix height must be a positive integer"; return false; }
    if (o.width <= 0) { c->err = "matrix width must be a positive integer"; return false; }
    if (o.k > o.width) { c->err = "k value cannot exceed the number of columns"; return false; }
    if (o.tol <= 0.0 || o.tol >= 1.0) { c->err = "tolerance must be in the interval (0.0, 1.0)"; return false; }
    if (o.min_iter <= 0) { c->err = "miniter must be a positive integer"; return false; }
    if (o.max_iter <= 0) { c->err = "maxiter must be a positive integer"; return false; }
    if (o.tolcount <= 0) { c->err = "tolcount must be a positive integer"; return false; }
    if (o.algorithm != SMK_MU && o.algorithm != SMK_HALS && o.algorithm != SMK_RANK2 && o.algorithm != SMK_BPP)
    { c->err = "unknown NMF algorithm specified"; return false; }
    if (o.algorithm == SMK_RANK2 && o.k != 2) { c->err = "RANK2 algorithm requires k == 2"; return false; }
    if (o.prog_est_algorithm != SMK_PG_RATIO && o.prog_est_algorithm != SMK_DELTA_FNORM)
    { c->err = "unknown stopping criterion specified"; return false; }
    return true;
}

void ensure_scratch(smk_ctx* c, int k = 0)
{
    c->status.reserve(ST_COUNT);
    c->counter.reserve(2);
    if (!c->pinned) SMK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->pinned), 16 * sizeof(double), cudaHostAllocDefault));
    if (!c->ticket.p) { c->ticket.reserve(8); SMK_CUDA(cudaMemsetAsync(c->ticket.p, 0, 8 * sizeof(unsigned int), c->stream)); }
    c->partial.reserve(4096 + static_cast<size_t>(512) * std::max(k, 256));      // normalize_and_scale: up to 512 blocks x k partial sums
    c->acc.reserve(8);
}

// host (rows x cols, ld) -> device tight copy
void upload_tight(smk_ctx* c, const double* host, long long ld, int rows, int cols, double* dev)
{
    SMK_CUDA(cudaMemcpy2DAsync(dev, sizeof(double) * rows, host, sizeof(double) * ld, sizeof(double) * rows, cols,
                               cudaMemcpyHostToDevice, c->stream));
}
void download_tight(smk_ctx* c, const double* dev, int rows, int cols, double* host, long long ld)
{
    SMK_CUDA(cudaMemcpy2DAsync(host, sizeof(double) * ld, dev, sizeof(double) * rows, sizeof(double) * rows, cols,
                               cudaMemcpyDeviceToHost, c->stream));
}

// W (m x k host) -> Wt (k x m device)
void upload_W(smk_ctx* c, const double* W, int ldW)
{
    const int m = c->m, k = c->opts.k;
    c->io.reserve(static_cast<size_t>(m) * k);
    upload_tight(c, W, ldW, m, k, c->io.p);
    transpose_f64(c->stream, m, k, c->io.p, m, c->Wt.p, k);
}
void download_Wt(smk_ctx* c, const double* dWt, double* W, int ldW)
{
    const int m = c->m, k = c->opts.k;
    c->io.reserve(static_cast<size_t>(m) * k);
    transpose_f64(c->stream, k, m, dWt, k, c->io.p, m);
    download_tight(c, c->io.p, m, k, W, ldW);
}

int begin_impl(smk_ctx* c, const smk_nmf_options* opts, const double* W0, int ldW, const double* H0, int ldH)
{
    if (!c->has_dense && !c->has_sparse) return fail(c, SMK_BAD_PARAM, "no matrix loaded");
    if (!options_valid(c, *opts)) return SMK_BAD_PARAM;
    if (opts->height != c->m) return fail(c, SMK_BAD_PARAM, "options.height does not match the loaded matrix");
    if (c->nranks == 1 && opts->width != c->n) return fail(c, SMK_BAD_PARAM, "options.width does not match the loaded matrix");
    // common/src/nmf.cpp:191-219
    if (static_cast<unsigned long long>(c->m) * opts->k > 2147483647ull) return fail(c, SMK_SIZE_TOO_LARGE, "W matrix size too large");
    if (static_cast<unsigned long long>(c->n) * opts->k > 2147483647ull) return fail(c, SMK_SIZE_TOO_LARGE, "H matrix size too large");
    if (ldW < c->m) return fail(c, SMK_BAD_PARAM, "leading dimension of W return buffer too small");
    if (ldH < opts->k) return fail(c, SMK_BAD_PARAM, "leading dimension of H return buffer too small");
    c->opts = *opts;
    c->steps_done = 0;
    c->pg0 = 0.0;
    ensure_scratch(c, opts->k);
    solver_alloc(c);
    upload_W(c, W0, ldW);
    upload_tight(c, H0, ldH, opts->k, c->n, c->H.p);
    solver_init(c);
    c->active = true;
    return SMK_OK;
}

} // namespace

extern "C" {

int smk_create(smk_ctx** out, int device)
{
    if (!out) return SMK_BAD_PARAM;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return SMK_CUDA_ERROR;
    smk_ctx* c = new smk_ctx();
    int rc = guarded(c, [&]() {
        c->device = device;
        SMK_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        SMK_CUDA(cudaGetDeviceProperties(&prop, device));
        c->num_sms = prop.multiProcessorCount;
        if (prop.major < 10) throw std::string("smallk_b200 requires an sm_100a device (found sm_") + std::to_string(prop.major * 10 + prop.minor) + ")";
        SMK_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        SMK_CUDA(cudaEventCreate(&c->ev0));
        SMK_CUDA(cudaEventCreate(&c->ev1));
        ensure_scratch(c);
        return SMK_OK;
    });
    if (rc != SMK_OK) { fprintf(stderr, "smk_create: %s\n", c->err.c_str()); delete c; return rc; }
    *out = c;
    return SMK_OK;
}

void smk_destroy(smk_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    peer_release(c);
    c->Wt.release(); c->HAt.release();      // possibly views of the exchange region just freed
    for (cudaEvent_t e : c->phase_pool) cudaEventDestroy(e);
    for (smk_ctx::InvBuf* b : {&c->invH, &c->invW})
    {
        if (b->fork) cudaEventDestroy(b->fork);
        if (b->join) cudaEventDestroy(b->join);
    }
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->pinned) cudaFreeHost(c->pinned);
    delete c;
}

const char* smk_last_error(const smk_ctx* c) { return c ? c->err.c_str() : "null context"; }
int smk_device_sm_count(const smk_ctx* c) { return c ? c->num_sms : 0; }
int smk_device_index(const smk_ctx* c) { return c ? c->device : -1; }

int smk_set_stream(smk_ctx* c, void* s)
{
    if (!c) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
        return SMK_OK;
    });
}

int smk_synchronize(smk_ctx* c)
{
    if (!c) return SMK_BAD_PARAM;
    return guarded(c, [&]() { SMK_CUDA(cudaStreamSynchronize(c->stream)); return SMK_OK; });
}

int smk_comm_unique_id(void* id128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return SMK_FAILURE;
    std::memcpy(id128, &id, sizeof(id));
    return SMK_OK;
}

int smk_comm_init(smk_ctx* c, int rank, int nranks, const void* id128)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        peer_release(c);
        c->Wt.release(); c->HAt.release();
        if (c->comm) { ncclCommDestroy(c->comm); c->comm = nullptr; }
        c->rank = rank; c->nranks = nranks;
        if (nranks == 1) return (int)SMK_OK;
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) throw std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
        return (int)SMK_OK;
    });
}

int smk_load_dense(smk_ctx* c, const double* A, long long ldA, int m, int n)
{
    if (!c || !A || m <= 0 || n <= 0 || ldA < m) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        select_all(c);
        c->A_store.reserve(static_cast<size_t>(m) * n);
        SMK_CUDA(cudaMemcpy2DAsync(c->A_store.p, sizeof(double) * m, A, sizeof(double) * ldA, sizeof(double) * m, n,
                                   cudaMemcpyHostToDevice, c->stream));
        c->dA = c->A_store.p; c->ldA = m; c->m = m; c->n = n;
        c->has_dense = true; c->has_sparse = false; c->active = false;
        return (int)SMK_OK;
    });
}

int smk_load_dense_device(smk_ctx* c, const double* A, long long ldA, int m, int n)
{
    if (!c || !A || m <= 0 || n <= 0 || ldA < m) return SMK_BAD_PARAM;
    select_all(c);
    c->dA = A; c->ldA = ldA; c->m = m; c->n = n;
    c->has_dense = true; c->has_sparse = false; c->active = false;
    return SMK_OK;
}

int smk_load_csc(smk_ctx* c, int m, int n, unsigned int nnz, const unsigned int* colp, const unsigned int* rowi, const double* val)
{
    if (!c || m <= 0 || n <= 0 || !colp || (nnz && (!rowi || !val))) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        select_all(c);
        SparseDev& S = c->S;
        S.m = m; S.n = n; S.nnz = nnz;
        S.colptr.reserve(static_cast<size_t>(n) + 1);
        S.rowidx.reserve(nnz); S.val.reserve(nnz);
        SMK_CUDA(cudaMemcpyAsync(S.colptr.p, colp, sizeof(unsigned int) * (static_cast<size_t>(n) + 1), cudaMemcpyHostToDevice, c->stream));
        if (nnz)
        {
            SMK_CUDA(cudaMemcpyAsync(S.rowidx.p, rowi, sizeof(unsigned int) * nnz, cudaMemcpyHostToDevice, c->stream));
            SMK_CUDA(cudaMemcpyAsync(S.val.p, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, c->stream));
        }
        build_csr(c->stream, S);
        build_segments(c->stream, n, S.colptr.p, S.seg_cols, c->num_sms);
        build_segments(c->stream, m, S.rowptr.p, S.seg_rows, c->num_sms);
        c->m = m; c->n = n;
        c->has_sparse = true; c->has_dense = false; c->active = false;
        return (int)SMK_OK;
    });
}

int smk_select_columns(smk_ctx* c, const unsigned int* cols, int count, int* new_height, unsigned int* new_to_old_rows)
{
    if (!c || !new_height || !new_to_old_rows) return SMK_BAD_PARAM;
    if (!c->has_dense && !c->has_sparse) return fail(c, SMK_BAD_PARAM, "no matrix loaded");
    if (!cols || count <= 0) return fail(c, SMK_BAD_PARAM, "SubMatrixColsCompact: empty column set");
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        ensure_scratch(c);
        *new_height = select_columns(c, cols, count, new_to_old_rows);
        return (int)SMK_OK;
    });
}

int smk_select_all(smk_ctx* c)
{
    if (!c) return SMK_BAD_PARAM;
    select_all(c);
    return SMK_OK;
}

int smk_argsort_desc(smk_ctx* c, const double* values, int n, int* order)
{
    if (!c || !values || !order || n < 0) return SMK_BAD_PARAM;
    if (n == 0) return SMK_OK;
    return guarded(c, [&]() { SMK_CUDA(cudaSetDevice(c->device)); device_sort_desc(c, values, n, order, nullptr); return (int)SMK_OK; });
}

int smk_sort_desc(smk_ctx* c, double* values, int n)
{
    if (!c || !values || n < 0) return SMK_BAD_PARAM;
    if (n == 0) return SMK_OK;
    return guarded(c, [&]() { SMK_CUDA(cudaSetDevice(c->device)); device_sort_desc(c, values, n, nullptr, values); return (int)SMK_OK; });
}

int smk_nnls_hals(smk_ctx* c, int k, double* W, int ldW, double* H, int ldH, double tol, int max_iter, int* iterations)
{
    if (!c || !W || !H || k <= 0 || max_iter <= 0) return SMK_BAD_PARAM;
    if (!c->has_dense && !c->has_sparse) return fail(c, SMK_BAD_PARAM, "no matrix loaded");
    if (ldW < c->m || ldH < k) return fail(c, SMK_BAD_PARAM, "NnlsHals: non-conformant W and H");
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        smk_nmf_options o;
        std::memset(&o, 0, sizeof(o));
        o.tol = tol; o.algorithm = SMK_HALS; o.prog_est_algorithm = SMK_PG_RATIO;
        o.height = c->m; o.width = c->n; o.k = k; o.min_iter = 1; o.max_iter = max_iter; o.tolcount = 1;
        c->opts = o;
        c->steps_done = 0;
        ensure_scratch(c, k);
        solver_alloc(c);
        upload_W(c, W, ldW);
        upload_tight(c, H, ldH, k, c->n, c->H.p);
        c->active = true;
        int rc = solver_nnls_hals(c, tol, max_iter, iterations);
        if (rc != SMK_OK) return rc;
        download_Wt(c, c->Wt.p, W, ldW);
        download_tight(c, c->H.p, k, c->n, H, ldH);
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        return (int)SMK_OK;
    });
}

int smk_solver_begin(smk_ctx* c, const smk_nmf_options* opts, const double* W0, int ldW, const double* H0, int ldH)
{
    if (!c || !opts || !W0 || !H0) return SMK_BAD_PARAM;
    return guarded(c, [&]() { SMK_CUDA(cudaSetDevice(c->device)); return begin_impl(c, opts, W0, ldW, H0, ldH); });
}

int smk_solver_step(smk_ctx* c, int count)
{
    if (!c || !c->active || count < 0) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const long long l0 = launch_counter();
        SMK_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < count; ++i) solver_step(c);
        SMK_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->last_launches = launch_counter() - l0;
        if (solver_fail_iter(c) != INT_MAX) return fail(c, SMK_FAILURE, "NMF solver failure");
        SMK_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
        return (int)SMK_OK;
    });
}

int smk_solver_run(smk_ctx* c, int count, double* metrics)
{
    if (!c || !c->active || count < 0) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const long long l0 = launch_counter();
        SMK_CUDA(cudaEventRecord(c->ev0, c->stream));
        const int rc = solver_run(c, count, metrics);
        SMK_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->last_launches = launch_counter() - l0;
        SMK_CUDA(cudaEventSynchronize(c->ev1));
        SMK_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
        return rc;
    });
}

int smk_phase_report(smk_ctx* c, char* buf, int len)
{
    if (!c || !buf || len <= 0) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const std::string r = solver_phase_report(c);
        std::strncpy(buf, r.c_str(), static_cast<size_t>(len) - 1);
        buf[len - 1] = 0;
        return (int)SMK_OK;
    });
}

int smk_solver_last_step_ms(smk_ctx* c, float* ms, long long* launches)
{
    if (!c) return SMK_BAD_PARAM;
    if (ms) *ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    return SMK_OK;
}

int smk_solver_time_product(smk_ctx* c, int which, int reps, float* mean_ms)
{
    if (!c || !c->active || reps <= 0 || !mean_ms || which < 0 || which > 1) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        solver_product(c, which);      // warm
        SMK_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; ++i) solver_product(c, which);
        SMK_CUDA(cudaEventRecord(c->ev1, c->stream));
        SMK_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        SMK_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *mean_ms = ms / reps;
        return (int)SMK_OK;
    });
}

int smk_solver_progress(smk_ctx* c, double* metric)
{
    if (!c || !c->active || !metric) return SMK_BAD_PARAM;
    return guarded(c, [&]() { SMK_CUDA(cudaSetDevice(c->device)); return solver_progress(c, metric); });
}

int smk_solver_normalize(smk_ctx* c)
{
    if (!c || !c->active) return SMK_BAD_PARAM;
    return guarded(c, [&]() { SMK_CUDA(cudaSetDevice(c->device)); return solver_normalize(c); });
}

int smk_solver_get(smk_ctx* c, double* W, int ldW, double* H, int ldH, double* gW, int ldgW, double* gH, int ldgH)
{
    if (!c || !c->active) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const int k = c->opts.k;
        if (W) { download_Wt(c, c->Wt.p, W, ldW); SMK_CUDA(cudaStreamSynchronize(c->stream)); }
        if (gW) { download_Wt(c, c->gradWt.p, gW, ldgW); SMK_CUDA(cudaStreamSynchronize(c->stream)); }
        if (H) download_tight(c, c->H.p, k, c->n, H, ldH);
        if (gH) download_tight(c, c->gradH.p, k, c->n, gH, ldgH);
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        return (int)SMK_OK;
    });
}

// NmfSolve, common/include/nmf_solve_generic.hpp:30-140, around the device solver.
int smk_nmf(smk_ctx* c, const smk_nmf_options* opts, double* W, int ldW, double* H, int ldH, smk_nmf_stats* stats)
{
    if (!c || !opts || !W || !H) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        auto t0 = std::chrono::high_resolution_clock::now();
        auto finish = [&](int iters) {
            if (stats)
            {
                auto t1 = std::chrono::high_resolution_clock::now();
                stats->elapsed_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
                stats->iteration_count = iters;
            }
        };
        int rc = begin_impl(c, opts, W, ldW, H, ldH);
        if (rc != SMK_OK) return rc;

        bool success = false;
        int iter = 0, success_count = 0;
        // Failure flags are polled only where the reference would look at the metric (a host sync
        // exists there anyway) and once at the end; a failed iteration index is still reported exactly.
        for (iter = 0; iter < opts->max_iter; ++iter)
        {
            if (iter == opts->min_iter && iter > 0)
            {
                // from here on every iteration is followed by the stop test: hand the rest of the loop to the device
                // (one CUDA graph with a WHILE node, nmf_loop.cu); falls through to the host loop when that is not possible
                int giter = iter, grc = SMK_OK;
                bool gsuccess = false;
                if (nmf_loop_graph(c, iter, &giter, &gsuccess, &grc))
                {
                    iter = giter;
                    if (grc != SMK_OK) { finish(iter); return grc; }
                    success = gsuccess;
                    break;
                }
            }
            solver_step(c);
            if (iter < opts->min_iter)
            {
                if (iter == 0)
                {
                    double m0;
                    rc = solver_progress(c, &m0);
                    if (rc != SMK_OK) { finish(iter); return rc; }
                    int fi = solver_fail_iter(c);
                    if (fi != INT_MAX) { finish(fi); return fail(c, SMK_FAILURE, "NMF solver failure on iteration " + std::to_string(fi + 1)); }
                }
                if (opts->verbose) printf("%d:\tprogress metric: \t(min_iter)\n", iter + 1);
                continue;
            }
            double metric;
            rc = solver_progress(c, &metric);
            if (rc != SMK_OK) { finish(iter); return rc; }
            {
                int fi = solver_fail_iter(c);
                if (fi != INT_MAX) { finish(fi); return fail(c, SMK_FAILURE, "NMF solver failure on iteration " + std::to_string(fi + 1)); }
                // Solver_Generic_Rank2 normalises inside the iteration (nmf_solver_rank2.hpp:420): the reference throws from there
                if (opts->algorithm == SMK_RANK2 && c->status_cached && c->status_host[ST_NORM_EPS])
                { finish(iter); return fail(c, SMK_FAILURE, "Normalize: column norm < machine epsilon"); }
            }
            if (opts->verbose && ((iter + 1) <= 9 || (iter + 1) % 10 == 0)) printf("%d:\tprogress metric: \t%g\n", iter + 1, metric);
            if (metric <= opts->tol)
            {
                if (++success_count >= opts->tolcount)
                {
                    success = true;
                    if (opts->verbose) printf("\nSolution converged after %d iterations.\n\n", iter + 1);
                    break;
                }
            }
            else success_count = 0;
        }
        {
            int fi = solver_fail_iter(c);
            if (fi != INT_MAX) { finish(fi); return fail(c, SMK_FAILURE, "NMF solver failure on iteration " + std::to_string(fi + 1)); }
        }
        if (opts->normalize)
        {
            rc = solver_normalize(c);
            if (rc != SMK_OK) { finish(iter); return rc; }
        }
        download_Wt(c, c->Wt.p, W, ldW);
        download_tight(c, c->H.p, opts->k, c->n, H, ldH);
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        if (!success && iter == opts->max_iter) success = true;
        finish(iter);
        return success ? (int)SMK_OK : (int)SMK_FAILURE;
    });
}

// ---- primitives ------------------------------------------------------------

int smk_gemm(smk_ctx* c, int transA, int transB, int M, int N, int K,
             const double* A, int ldA, const double* B, int ldB, double* C, int ldC)
{
    if (!c || !A || !B || !C || M <= 0 || N <= 0 || K < 0) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        // The device kernel wants op(A) column-major M x K. op(B) is consumed either as K x N column-major
        // (NN) or from an N x K column-major array (NT), so only a transposed A needs an explicit transpose.
        const int a_rows = transA ? K : M, a_cols = transA ? M : K;
        const int b_rows = transB ? N : K, b_cols = transB ? K : N;
        DevBuf<double> dA, dAt, dB, dC, ws;
        dA.reserve(static_cast<size_t>(a_rows) * a_cols);
        dB.reserve(static_cast<size_t>(b_rows) * b_cols);
        dC.reserve(static_cast<size_t>(M) * N);
        upload_tight(c, A, ldA, a_rows, a_cols, dA.p);
        upload_tight(c, B, ldB, b_rows, b_cols, dB.p);
        const double* Aop = dA.p;
        if (transA)
        {
            dAt.reserve(static_cast<size_t>(M) * K);
            transpose_f64(c->stream, K, M, dA.p, K, dAt.p, M);
            Aop = dAt.p;
        }
        ws.reserve(std::min<size_t>(static_cast<size_t>(64) * std::max(M, 64) * std::max(N, 128), (size_t(256) << 20) / sizeof(double)) + 8192);
        gemm_workspace_prepare(c->stream, ws.p, ws.n * sizeof(double));
        gemm_f64(c->stream, transB != 0, M, N, K, Aop, M, dB.p, b_rows, dC.p, M, nullptr, 0, ws.p, ws.n * sizeof(double), c->num_sms);
        download_tight(c, dC.p, M, N, C, ldC);
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        return (int)SMK_OK;
    });
}

int smk_nnls_bpp(smk_ctx* c, int k, int q, const double* LHS, const double* RHS, double* X, double* Y)
{
    if (!c || k <= 0 || q <= 0 || !LHS || !RHS || !X || !Y) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        ensure_scratch(c);
        DevBuf<double> dL, dR, dX, dY;
        dL.reserve(static_cast<size_t>(k) * k); dR.reserve(static_cast<size_t>(k) * q);
        dX.reserve(static_cast<size_t>(k) * q); dY.reserve(static_cast<size_t>(k) * q);
        upload_tight(c, LHS, k, k, k, dL.p);
        upload_tight(c, RHS, k, k, q, dR.p);
        upload_tight(c, X, k, k, q, dX.p);
        static const int init[ST_COUNT] = {0, INT_MAX, 0, 0, 0, 0, 0, 0};
        SMK_CUDA(cudaMemcpyAsync(c->status.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        c->deferred.reserve(nnls_deferred_bytes(q, k, c->num_sms));
        nnls_bpp(c->stream, k, q, dL.p, k, dR.p, k, dX.p, k, dY.p, k, c->status.p, c->counter.p, c->deferred.p, 0, c->num_sms);
        nnls_bpp_finish(c->stream, k, q, dX.p, k, dY.p, k, c->status.p, c->num_sms);
        download_tight(c, dX.p, k, q, X, k);
        download_tight(c, dY.p, k, q, Y, k);
        int st[ST_COUNT];
        SMK_CUDA(cudaMemcpyAsync(st, c->status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        if (st[ST_FAIL_ITER] != INT_MAX) return fail(c, SMK_FAILURE, "NnlsBlockpivot failed (non-HPD sub-problem or iteration limit)");
        return (int)SMK_OK;
    });
}

int smk_preprocess_tf(smk_ctx* c, unsigned int m, unsigned int n, unsigned int nnz, const unsigned int* col_offsets,
                      const unsigned int* row_indices, const double* counts, unsigned int max_iter, unsigned int docs_per_term,
                      unsigned int terms_per_doc, unsigned int* out_m, unsigned int* out_n, unsigned int* out_nnz,
                      unsigned int* out_col_offsets, unsigned int* out_row_indices, unsigned int* out_counts, double* out_scores,
                      unsigned int* term_indices, unsigned int* doc_indices)
{
    if (!c || !col_offsets || (nnz && (!row_indices || !counts)) || !out_m || !out_n || !out_nnz || !out_col_offsets || !out_row_indices ||
        !out_counts || !out_scores || !term_indices || !doc_indices || m == 0 || n == 0)
        return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const int rc = preprocess_tf_device(c, m, n, nnz, col_offsets, row_indices, counts, max_iter, docs_per_term, terms_per_doc, out_m, out_n,
                                            out_nnz, out_col_offsets, out_row_indices, out_counts, out_scores, term_indices, doc_indices);
        if (rc != SMK_OK) c->err = "preprocess_tf: every document was pruned";
        return rc;
    });
}

int smk_nnls_backup_count(smk_ctx* c, int* count)
{
    if (!c || !count || !c->status.p) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        int st[ST_COUNT];
        SMK_CUDA(cudaMemcpyAsync(st, c->status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        *count = st[ST_BACKUP_COUNT];
        return (int)SMK_OK;
    });
}

int smk_sparse_gemm(smk_ctx* c, int variant, double alpha, const double* B, int Bh, int Bw,
                    double beta, double* C, int Ch, int Cw)
{
    if (!c || !c->has_sparse || !B || !C || variant < 0 || variant > 3) return SMK_BAD_PARAM;
    return guarded(c, [&]() {
        SMK_CUDA(cudaSetDevice(c->device));
        const SparseDev& S = *c->Sa;      // the active matrix (a column subset after smk_select_columns)
        const int m = S.m, n = S.n;
        // reference shape checks: sparse_gemm_ab_impl.hpp / sparse_gemm_ba_impl.hpp
        int k;
        switch (variant)
        {
        case 0: if (Bh != n || Ch != m || Bw != Cw) return fail(c, SMK_BAD_PARAM, "Gemm: non-conformant matrices"); k = Bw; break;
        case 1: if (Bw != n || Ch != m || Bh != Cw) return fail(c, SMK_BAD_PARAM, "Gemm: non-conformant matrices"); k = Bh; break;
        case 2: if (Bw != m || Ch != Bh || Cw != n) return fail(c, SMK_BAD_PARAM, "Gemm: non-conformant matrices"); k = Bh; break;
        default: if (Bh != m || Ch != Bw || Cw != n) return fail(c, SMK_BAD_PARAM, "Gemm: non-conformant matrices"); k = Bw; break;
        }
        // Device kernels want the dense operand as k x (m or n) and produce k x (n or m).
        DevBuf<double> dB, dBt, dC, dCt;
        dB.reserve(static_cast<size_t>(Bh) * Bw);
        dC.reserve(static_cast<size_t>(Ch) * Cw);
        upload_tight(c, B, Bh, Bh, Bw, dB.p);
        if (beta != 0.0) upload_tight(c, C, Ch, Ch, Cw, dC.p);
        const bool b_is_kmajor = (variant == 1 || variant == 2);    // B is k x n (1) or k x m (2)
        const double* Bk = dB.p;
        if (!b_is_kmajor)
        {
            dBt.reserve(static_cast<size_t>(Bh) * Bw);
            transpose_f64(c->stream, Bh, Bw, dB.p, Bh, dBt.p, Bw);
            Bk = dBt.p;
        }
        const bool c_is_kmajor = (variant == 2 || variant == 3);    // C is k x n
        double* Ck = dC.p;
        if (!c_is_kmajor)
        {
            dCt.reserve(static_cast<size_t>(Ch) * Cw);
            if (beta != 0.0) transpose_f64(c->stream, Ch, Cw, dC.p, Ch, dCt.p, Cw);
            Ck = dCt.p;
        }
        c->spmm_partial.reserve(static_cast<size_t>(std::max(S.seg_cols.nslots, S.seg_rows.nslots)) * k + 1);
        if (variant <= 1)   // C = A*op(B): rows of A -> CSR walk, output k x m
            spmm_gather_seg(c->stream, m, S.seg_rows, S.colidx.p, S.valr.p, k, Bk, k, alpha, beta, Ck, k, c->spmm_partial.p, c->num_sms, n);
        else                // C = op(B)*A: columns of A -> CSC walk, output k x n
            spmm_gather_seg(c->stream, n, S.seg_cols, S.rowidx.p, S.val.p, k, Bk, k, alpha, beta, Ck, k, c->spmm_partial.p, c->num_sms, m);
        if (!c_is_kmajor) transpose_f64(c->stream, Cw, Ch, dCt.p, Cw, dC.p, Ch);
        download_tight(c, dC.p, Ch, Cw, C, Ch);
        SMK_CUDA(cudaStreamSynchronize(c->stream));
        return (int)SMK_OK;
    });
}

} // extern "C"
