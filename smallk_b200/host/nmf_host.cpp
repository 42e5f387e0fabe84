// Implementation of host/nmf.hpp over the C ABI. Reference behaviour followed: common/src/nmf.cpp:36-295.
#include "nmf.hpp"

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>

#include "host_internal.hpp"

namespace {
smk_ctx* g_ctx = nullptr;
std::string g_err;

smk_nmf_options ToAbi(const NmfOptions& o)
{
    smk_nmf_options a;
    a.tol = o.tol;
    a.algorithm = static_cast<int>(o.algorithm);                    // same numeric values as the reference enum
    a.prog_est_algorithm = static_cast<int>(o.prog_est_algorithm);
    a.height = o.height; a.width = o.width; a.k = o.k;
    a.min_iter = o.min_iter; a.max_iter = o.max_iter; a.tolcount = o.tolcount;
    a.max_threads = o.max_threads;
    a.verbose = o.verbose ? 1 : 0;
    a.normalize = o.normalize ? 1 : 0;
    return a;
}

Result FromAbi(int rc)
{
    switch (rc)
    {
    case SMK_OK: return Result::OK;
    case SMK_NOTINITIALIZED: return Result::NOTINITIALIZED;
    case SMK_INITIALIZED: return Result::INITIALIZED;
    case SMK_BAD_PARAM: return Result::BAD_PARAM;
    case SMK_SIZE_TOO_LARGE: return Result::SIZE_TOO_LARGE;
    case SMK_FLATCLUST_FAILURE: return Result::FLATCLUST_FAILURE;
    default: return Result::FAILURE;
    }
}

template <typename T>
bool FitsWithin(uint64_t v) { return v <= static_cast<uint64_t>(std::numeric_limits<T>::max()); }

Result CheckCommon(const NmfOptions& options, int ldim_w, int ldim_h)
{
    if (!g_ctx)
    {
        std::cerr << "nmflib error: nmf_initialize() must be called prior to any factorization routine\n" << std::endl;
        return Result::NOTINITIALIZED;
    }
    if (!IsValid(options)) return Result::BAD_PARAM;
    const uint64_t m = options.height, n = options.width, k = options.k;
    if (!FitsWithin<int>(m * k)) { std::cerr << "W matrix size too large" << std::endl; return Result::SIZE_TOO_LARGE; }
    if (!FitsWithin<int>(n * k)) { std::cerr << "H matrix size too large" << std::endl; return Result::SIZE_TOO_LARGE; }
    if (ldim_w < options.height) throw std::logic_error("nmflib error: leading dimension of W return buffer too small");
    if (ldim_h < options.k) throw std::logic_error("nmflib error: leading dimension of H return buffer too small");
    return Result::OK;
}

Result Run(const NmfOptions& options, double* W, int ldw, double* H, int ldh, NmfStats& stats)
{
    smk_nmf_options a = ToAbi(options);
    smk_nmf_stats st = {0, 0};
    int rc = smk_nmf(g_ctx, &a, W, ldw, H, ldh, &st);
    stats.elapsed_us = st.elapsed_us;
    stats.iteration_count = st.iteration_count;
    if (rc != SMK_OK)
    {
        g_err = smk_last_error(g_ctx);
        // the reference lets this exception escape from NormalizeColumns (normalize.hpp:47-48)
        if (g_err.find("Normalize:") == 0 || g_err.find("ProjectedGradientNorm") == 0) throw std::runtime_error(g_err);
        if (rc == SMK_FAILURE) std::cerr << "\t" << g_err << std::endl;
    }
    return FromAbi(rc);
}
} // namespace

void NmfInitialize(int, char*[])
{
    if (g_ctx) return;
    int device = 0;
    if (const char* e = std::getenv("SMALLK_B200_DEVICE")) device = std::atoi(e);
    if (smk_create(&g_ctx, device) != SMK_OK)
    {
        g_ctx = nullptr;
        throw std::runtime_error("smallk_b200: no usable sm_100 CUDA device (there is no CPU fallback)");
    }
}

Result NmfIsInitialized() { return g_ctx ? Result::INITIALIZED : Result::NOTINITIALIZED; }

void NmfFinalize()
{
    HierReleaseWorkers();
    if (g_ctx) smk_destroy(g_ctx);
    g_ctx = nullptr;
}

const char* NmfLastError() { return g_err.c_str(); }
smk_ctx* NmfContext() { return g_ctx; }
smk_nmf_options NmfToAbi(const NmfOptions& o) { return ToAbi(o); }
Result NmfFromAbi(int rc) { return FromAbi(rc); }
void NmfSetLastError(const char* msg) { g_err = msg ? msg : ""; }

bool IsValid(const NmfOptions& opts, bool validate_matrix)
{
    using std::cerr; using std::endl;
    if (opts.k <= 0) { cerr << "nmflib error: k-value must be a positive integer" << endl; return false; }
    if (validate_matrix)
    {
        if (opts.height <= 0) { cerr << "nmflib error: matrix height must be a positive integer" << endl; return false; }
        if (opts.width <= 0) { cerr << "nmflib error: matrix width must be a positive integer" << endl; return false; }
        if (opts.k > opts.width) { cerr << "nmflib error: k value cannot exceed the number of columns" << endl; return false; }
    }
    if ((opts.tol <= 0.0) || (opts.tol >= 1.0)) { cerr << "nmflib error: tolerance must be in the interval (0.0, 1.0)" << endl; return false; }
    if (opts.min_iter <= 0) { cerr << "nmflib error: miniter must be a positive integer" << endl; return false; }
    if (opts.max_iter <= 0) { cerr << "nmflib error: maxiter must be a positive integer" << endl; return false; }
    if (opts.tolcount <= 0) { cerr << "nmflib error: tolcount must be a positive integer" << endl; return false; }
    if ((NmfAlgorithm::MU != opts.algorithm) && (NmfAlgorithm::HALS != opts.algorithm) &&
        (NmfAlgorithm::RANK2 != opts.algorithm) && (NmfAlgorithm::BPP != opts.algorithm))
    { cerr << "nmflib error: unknown NMF algorithm specified" << endl; return false; }
    if (NmfAlgorithm::RANK2 == opts.algorithm && 2 != opts.k) { cerr << "nmflib error: RANK2 algorithm requires k == 2" << endl; return false; }
    if ((NmfProgressAlgorithm::PG_RATIO != opts.prog_est_algorithm) && (NmfProgressAlgorithm::DELTA_FNORM != opts.prog_est_algorithm))
    { cerr << "nmflib error: unknown stopping criterion specified" << endl; return false; }
    return true;
}

Result Nmf(const NmfOptions& options, double* buf_a, int ldim_a, double* buf_w, int ldim_w, double* buf_h, int ldim_h,
           NmfStats& stats)
{
    Result r = CheckCommon(options, ldim_w, ldim_h);
    if (r != Result::OK) return r;
    int rc = smk_load_dense(g_ctx, buf_a, ldim_a, options.height, options.width);
    if (rc != SMK_OK) { g_err = smk_last_error(g_ctx); return FromAbi(rc); }
    return Run(options, buf_w, ldim_w, buf_h, ldim_h, stats);
}

Result NmfSparse(const NmfOptions& options, const unsigned int height, const unsigned int width, const unsigned int nz,
                 const unsigned int* col_offsets, const unsigned int* row_indices, const double* data,
                 double* buf_w, int ldim_w, double* buf_h, int ldim_h, NmfStats& stats)
{
    Result r = CheckCommon(options, ldim_w, ldim_h);
    if (r != Result::OK) return r;
    int rc = smk_load_csc(g_ctx, static_cast<int>(height), static_cast<int>(width), nz, col_offsets, row_indices, data);
    if (rc != SMK_OK) { g_err = smk_last_error(g_ctx); return FromAbi(rc); }
    return Run(options, buf_w, ldim_w, buf_h, ldim_h, stats);
}
