// smallk_b200 host interface — the reference's L3 library functions, same names, same argument meaning,
// same Result codes, implemented over the C ABI (include/smallk_b200.h) instead of Elemental.
//
// Mirrors common/include/nmf.hpp:17-92 of the reference. A program written against the reference's
// nmf.hpp compiles against this header unchanged and links with libsmallk_host.so + libsmallk_b200.so.
// There is no CPU implementation behind these calls: without an sm_100 GPU NmfInitialize() throws.
#pragma once

// Result values are shared with the C ABI (SMK_OK ... SMK_FLATCLUST_FAILURE in include/smallk_b200.h).
enum Result { OK = 0, NOTINITIALIZED = -1, INITIALIZED = -2, BAD_PARAM = -3, FAILURE = -4, SIZE_TOO_LARGE = -5, FLATCLUST_FAILURE = -6 };

// Update rule of the outer iteration; the enumerator values (0..3) are what the C ABI takes as `algorithm`.
//   MU    multiplicative updating (Lee, Seung)          -> csrc/solver.cu SMK_MU
//   HALS  hierarchical alternating least squares        -> SMK_HALS
//   RANK2 the k = 2 specialisation used by hierclust    -> SMK_RANK2 (csrc/rank2_fused.cu for sparse input)
//   BPP   block principal pivoting NNLS                 -> SMK_BPP
enum NmfAlgorithm { MU, HALS, RANK2, BPP };

// Stopping metric: projected-gradient norm relative to iteration 1, or the relative change of W in Frobenius norm.
enum NmfProgressAlgorithm { PG_RATIO, DELTA_FNORM };

struct NmfStats
{
    unsigned long long elapsed_us = 0u;     // wall time of the call, microseconds
    int iteration_count = 0;                // outer iterations run (the failing one on FAILURE)
    NmfStats() {}
};

struct NmfOptions
{
    double tol;                                     // stop when the metric is <= tol for tolcount consecutive checks
    NmfAlgorithm algorithm; NmfProgressAlgorithm prog_est_algorithm;
    int height, width, k;                           // A is height x width, rank k
    int min_iter, max_iter, tolcount;
    int max_threads;                                // kept for source compatibility; the GPU path ignores it
    bool verbose, normalize;
};

// common/src/nmf.cpp:36-52. NmfInitialize creates the GPU context (device from SMALLK_B200_DEVICE, default 0).
void NmfInitialize(int argc, char* argv[]);
Result NmfIsInitialized();
void NmfFinalize();

// common/src/nmf_options.cpp:24-111
bool IsValid(const NmfOptions& opts, bool validate_matrix = true);

// common/src/nmf.cpp:173. Column-major buffers; W (m x k) and H (k x n) carry the initial guess in and the
// factors out. Throws std::logic_error if ldW < m or ldH < k, as the reference does.
Result Nmf(const NmfOptions& options, double* A, int ldA, double* W, int ldW, double* H, int ldH, NmfStats& stats);

// common/src/nmf.cpp:232. CSC input, 32-bit indices (m x n with nz stored entries).
Result NmfSparse(const NmfOptions& options, const unsigned int m, const unsigned int n, const unsigned int nz,
                 const unsigned int* col_offsets, const unsigned int* row_indices, const double* data,
                 double* W, int ldW, double* H, int ldH, NmfStats& stats);

// Text of the last error reported by the GPU library (no counterpart in the reference, which prints to cerr).
const char* NmfLastError();
