// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// ref_capi.cpp: extern "C" entry points around the REFERENCE's own functions,
// compiled by oracle/Makefile together with the reference's unmodified sources
// (read in place from /root/reference) and the El.hpp shim into
// oracle/_ref/libsmallk_ref.so. Nothing here restates an algorithm: every call
// lands in reference code —
//   Nmf / NmfSparse            common/src/nmf.cpp:173,232
//   NmfSolve                   common/include/nmf_solve_generic.hpp:30-140
//   NnlsBlockpivot             common/include/nnls.hpp:144-244
//   sparse Gemm x4             common/include/sparse_gemm.hpp:26-74
//   Solver_Generic_*           common/include/nmf_solver_{bpp,hals,mu,rank2}.hpp
// Used only by tests/, tests/golden/make_golden.py and bench.py's CPU arm.

#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <iostream>
#include <limits>
#include <chrono>

#include "nmf.hpp"
#include "nnls.hpp"
#include "dense_matrix.hpp"
#include "sparse_matrix.hpp"
#include "sparse_gemm.hpp"
#include "nmf_solve_generic.hpp"
#include "progress_estimator_generic.hpp"
#include "thread_utils.hpp"

typedef double R;

namespace {

// A progress estimator that forwards to the reference's own estimator and
// records what it saw. NmfSolve is a template over the estimator type
// (nmf_solve_generic.hpp:30-40), so this needs no change to reference files.
struct TraceSink
{
    double* metrics;      // [max_iter], NaN where the reference did not evaluate
    double* Wsnap;        // optional [max_iter][m*k] snapshots (column-major, ld=m)
    double* Hsnap;        // optional [max_iter][k*n]
    int max_iter;
    double* stamps;       // optional [max_iter] wall-clock seconds at which iteration i's metric was taken
};

template <typename T, template <typename> class MatrixType>
class TracingEst
{
public:
    TracingEst(ProgEstGeneric<T, MatrixType>* inner, TraceSink* sink) : inner_(inner), sink_(sink) {}

    T Init(const MatrixType<T>& A, const DenseMatrix<T>& W, const DenseMatrix<T>& H)
    { return inner_->Init(A, W, H); }

    T Update(const unsigned int iter, const DenseMatrix<T>& W, const DenseMatrix<T>& H,
             const DenseMatrix<T>& gradW, const DenseMatrix<T>& gradH)
    {
        T v = inner_->Update(iter, W, H, gradW, gradH);
        if (sink_ && static_cast<int>(iter) < sink_->max_iter)
        {
            if (sink_->metrics) sink_->metrics[iter] = v;
            if (sink_->stamps)
                sink_->stamps[iter] = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            if (sink_->Wsnap)
            {
                const size_t m = W.Height(), k = W.Width();
                double* dst = sink_->Wsnap + static_cast<size_t>(iter) * m * k;
                for (size_t c = 0; c < k; ++c)
                    std::memcpy(dst + c * m, W.LockedBuffer() + c * W.LDim(), sizeof(double) * m);
            }
            if (sink_->Hsnap)
            {
                const size_t k = H.Height(), n = H.Width();
                double* dst = sink_->Hsnap + static_cast<size_t>(iter) * k * n;
                for (size_t c = 0; c < n; ++c)
                    std::memcpy(dst + c * k, H.LockedBuffer() + c * H.LDim(), sizeof(double) * k);
            }
        }
        return v;
    }

private:
    ProgEstGeneric<T, MatrixType>* inner_;
    TraceSink* sink_;
};

NmfOptions MakeOpts(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                    int tolcount, int max_threads, int normalize, int verbose)
{
    NmfOptions o;
    o.tol = tol;
    o.algorithm = static_cast<NmfAlgorithm>(alg);          // MU=0 HALS=1 RANK2=2 BPP=3 (nmf.hpp:28-34)
    o.prog_est_algorithm = static_cast<NmfProgressAlgorithm>(prog);  // PG_RATIO=0 DELTA_FNORM=1
    o.height = m; o.width = n; o.k = k;
    o.min_iter = min_iter; o.max_iter = max_iter; o.tolcount = tolcount;
    o.max_threads = max_threads;
    o.verbose = (verbose != 0);
    o.normalize = (normalize != 0);
    return o;
}

template <template <typename> class MatrixType>
int RunTraced(const NmfOptions& opts, const MatrixType<R>& A, DenseMatrix<R>& W, DenseMatrix<R>& H,
              NmfStats& stats, TraceSink* sink)
{
    ProgEstGeneric<R, MatrixType>* inner =
        ProgEstGeneric<R, MatrixType>::Create(opts.algorithm, opts.prog_est_algorithm);
    TracingEst<R, MatrixType> est(inner, sink);
    bool ok = false;
    try
    {
        switch (opts.algorithm)
        {
        case NmfAlgorithm::MU:    { Solver_Generic_MU<R, MatrixType> s;      ok = NmfSolve<R>(A, W, H, s, &est, opts, stats); break; }
        case NmfAlgorithm::HALS:  { Solver_Generic_HALS_Da<R, MatrixType> s; ok = NmfSolve<R>(A, W, H, s, &est, opts, stats); break; }
        case NmfAlgorithm::RANK2: { Solver_Generic_Rank2<R, MatrixType> s;   ok = NmfSolve<R>(A, W, H, s, &est, opts, stats); break; }
        case NmfAlgorithm::BPP:   { Solver_Generic_BPP<R, MatrixType> s;     ok = NmfSolve<R>(A, W, H, s, &est, opts, stats); break; }
        }
    }
    catch (std::exception& e)
    {
        std::cerr << "ref_capi: exception from reference: " << e.what() << std::endl;
        delete inner;
        return -100;
    }
    delete inner;
    return ok ? Result::OK : Result::FAILURE;
}

void EnsureInit()
{
    if (Result::INITIALIZED != NmfIsInitialized())
    {
        static int argc = 0;
        static char** argv = nullptr;
        NmfInitialize(argc, argv);
    }
}

} // namespace

extern "C" {

const char* ref_blas_backend()
{
#ifdef SHIM_USE_OPENBLAS
    return "scipy-openblas (venv)";
#else
    return "plain loops";
#endif
}

void ref_set_blas_threads(int n)
{
#ifdef SHIM_USE_OPENBLAS
    scipy_openblas_set_num_threads(n);
#else
    (void)n;
#endif
}

// The reference's library entry point, verbatim (common/src/nmf.cpp:173).
int ref_nmf_dense(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                  int tolcount, int max_threads, int normalize, int verbose,
                  double* A, int ldA, double* W, int ldW, double* H, int ldH,
                  int* iterations, unsigned long long* elapsed_us)
{
    EnsureInit();
    NmfOptions o = MakeOpts(alg, prog, m, n, k, tol, min_iter, max_iter, tolcount, max_threads, normalize, verbose);
    NmfStats st;
    int rc;
    try { rc = Nmf(o, A, ldA, W, ldW, H, ldH, st); }
    catch (std::exception& e) { std::cerr << "ref_capi: " << e.what() << std::endl; rc = -100; }
    if (iterations) *iterations = st.iteration_count;
    if (elapsed_us) *elapsed_us = st.elapsed_us;
    return rc;
}

// common/src/nmf.cpp:232
int ref_nmf_sparse(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                   int tolcount, int max_threads, int normalize, int verbose,
                   unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                   const double* data, double* W, int ldW, double* H, int ldH,
                   int* iterations, unsigned long long* elapsed_us)
{
    EnsureInit();
    NmfOptions o = MakeOpts(alg, prog, m, n, k, tol, min_iter, max_iter, tolcount, max_threads, normalize, verbose);
    NmfStats st;
    int rc;
    try { rc = NmfSparse(o, m, n, nz, col_offsets, row_indices, data, W, ldW, H, ldH, st); }
    catch (std::exception& e) { std::cerr << "ref_capi: " << e.what() << std::endl; rc = -100; }
    if (iterations) *iterations = st.iteration_count;
    if (elapsed_us) *elapsed_us = st.elapsed_us;
    return rc;
}

// Same solve, driven through NmfSolve directly with the recording estimator.
int ref_nmf_dense_trace(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                        int tolcount, int max_threads, int normalize,
                        double* A, int ldA, double* W, int ldW, double* H, int ldH,
                        int* iterations, double* metrics, double* Wsnap, double* Hsnap)
{
    EnsureInit();
    NmfOptions o = MakeOpts(alg, prog, m, n, k, tol, min_iter, max_iter, tolcount, max_threads, normalize, 0);
    if (!IsValid(o)) return Result::BAD_PARAM;
    SetMaxThreadCount(o.max_threads);
    DenseMatrix<R> Am(m, n, A, ldA), Wm(m, k, W, ldW), Hm(k, n, H, ldH);
    TraceSink sink = {metrics, Wsnap, Hsnap, max_iter, nullptr};
    if (metrics) for (int i = 0; i < max_iter; ++i) metrics[i] = std::numeric_limits<double>::quiet_NaN();
    NmfStats st;
    int rc = RunTraced<DenseMatrix>(o, Am, Wm, Hm, st, &sink);
    if (iterations) *iterations = st.iteration_count;
    return rc;
}

int ref_nmf_sparse_trace(int alg, int prog, int m, int n, int k, double tol, int min_iter, int max_iter,
                         int tolcount, int max_threads, int normalize,
                         unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                         const double* data, double* W, int ldW, double* H, int ldH,
                         int* iterations, double* metrics, double* Wsnap, double* Hsnap)
{
    EnsureInit();
    NmfOptions o = MakeOpts(alg, prog, m, n, k, tol, min_iter, max_iter, tolcount, max_threads, normalize, 0);
    if (!IsValid(o)) return Result::BAD_PARAM;
    SetMaxThreadCount(o.max_threads);
    SparseMatrix<R> Am(m, n, nz, col_offsets, row_indices, data);
    DenseMatrix<R> Wm(m, k, W, ldW), Hm(k, n, H, ldH);
    TraceSink sink = {metrics, Wsnap, Hsnap, max_iter, nullptr};
    if (metrics) for (int i = 0; i < max_iter; ++i) metrics[i] = std::numeric_limits<double>::quiet_NaN();
    NmfStats st;
    int rc = RunTraced<SparseMatrix>(o, Am, Wm, Hm, st, &sink);
    if (iterations) *iterations = st.iteration_count;
    return rc;
}

// Per-iteration wall-clock stamps of a dense solve (bench.py --impl reference): the reference's NmfSolve
// with its own estimator; stamps[i] is taken right after iteration i's progress update (min_iter = 1).
int ref_nmf_dense_stamped(int alg, int prog, int m, int n, int k, int max_iter, int max_threads,
                          double* A, int ldA, double* W, int ldW, double* H, int ldH,
                          int* iterations, double* stamps, double* t_start)
{
    EnsureInit();
    // tol so small that the stop rule never fires: exactly max_iter iterations are run
    NmfOptions o = MakeOpts(alg, prog, m, n, k, 1.0e-15, 1, max_iter, 1, max_threads, 0, 0);
    if (!IsValid(o)) return Result::BAD_PARAM;
    SetMaxThreadCount(o.max_threads);
    DenseMatrix<R> Am(m, n, A, ldA), Wm(m, k, W, ldW), Hm(k, n, H, ldH);
    TraceSink sink = {nullptr, nullptr, nullptr, max_iter, stamps};
    NmfStats st;
    if (t_start) *t_start = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    int rc = RunTraced<DenseMatrix>(o, Am, Wm, Hm, st, &sink);
    if (iterations) *iterations = st.iteration_count;
    return rc;
}

// common/include/nnls.hpp:144 — LHS k x k, RHS/X/Y k x q, all column-major, tight ld.
int ref_nnls_blockpivot(int k, int q, double* LHS, double* RHS, double* X, double* Y, int max_threads)
{
    EnsureInit();
    SetMaxThreadCount(max_threads);
    DenseMatrix<R> L(k, k, LHS, k), B(k, q, RHS, k), Xm(k, q, X, k), Ym(k, q, Y, k);
    try { return NnlsBlockpivot(L, B, Xm, Ym) ? 0 : Result::FAILURE; }
    catch (std::exception& e) { std::cerr << "ref_capi: " << e.what() << std::endl; return -100; }
}

// common/include/sparse_gemm.hpp:26-74. variant: 0 = A*B, 1 = A*B', 2 = B*A, 3 = B'*A.
// A sparse m x n (CSC); B and C dense column-major with tight leading dimension.
int ref_sparse_gemm(int variant, double alpha, double beta,
                    unsigned int m, unsigned int n, unsigned int nz,
                    const unsigned int* col_offsets, const unsigned int* row_indices, const double* data,
                    double* B, int Bh, int Bw, double* C, int Ch, int Cw, int max_threads)
{
    SetMaxThreadCount(max_threads);
    SparseMatrix<R> A(m, n, nz, col_offsets, row_indices, data);
    DenseMatrix<R> Bm(Bh, Bw, B, Bh), Cm(Ch, Cw, C, Ch);
    try
    {
        switch (variant)
        {
        case 0: Gemm(NORMAL, NORMAL, alpha, A, Bm, beta, Cm); break;
        case 1: Gemm(NORMAL, TRANSPOSE, alpha, A, Bm, beta, Cm); break;
        case 2: Gemm(NORMAL, NORMAL, alpha, Bm, A, beta, Cm); break;
        case 3: Gemm(TRANSPOSE, NORMAL, alpha, Bm, A, beta, Cm); break;
        default: return Result::BAD_PARAM;
        }
    }
    catch (std::exception& e) { std::cerr << "ref_capi: " << e.what() << std::endl; return -100; }
    return 0;
}

} // extern "C"
