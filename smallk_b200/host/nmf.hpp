// smallk_b200 host interface — the reference's L3 library functions, same names, same argument meaning,
// same Result codes, implemented over the C ABI (include/smallk_b200.h) instead of Elemental.
//
// Mirrors common/include/nmf.hpp:17-92 of the reference. A program written against the reference's
// nmf.hpp compiles against this header unchanged and links with libsmallk_host.so + libsmallk_b200.so.
// There is no CPU implementation behind these calls: without an sm_100 GPU NmfInitialize() throws.
#pragma once

enum Result
{
    OK                =  0,
    NOTINITIALIZED    = -1,
    INITIALIZED       = -2,
    BAD_PARAM         = -3,
    FAILURE           = -4,
    SIZE_TOO_LARGE    = -5,
    FLATCLUST_FAILURE = -6
};

enum NmfAlgorithm
{
    MU,       // Lee & Seung, multiplicative updating
    HALS,     // Cichocki & Pan, hierarchical alternating least squares
    RANK2,    // Kuang and Park, rank2 specialization
    BPP       // Kim and Park, block principal pivoting
};

enum NmfProgressAlgorithm
{
    PG_RATIO,    // PG_i / PG_1
    DELTA_FNORM  // relative change in the Frobenius norm of W
};

struct NmfStats
{
    NmfStats() : elapsed_us(0u), iteration_count(0) {}
    unsigned long long elapsed_us;
    int iteration_count;
};

struct NmfOptions
{
    double tol;
    NmfAlgorithm algorithm;
    NmfProgressAlgorithm prog_est_algorithm;
    int height;
    int width;
    int k;
    int min_iter;
    int max_iter;
    int tolcount;
    int max_threads;     // kept for source compatibility; the GPU path ignores it
    bool verbose;
    bool normalize;
};

// common/src/nmf.cpp:36-52. NmfInitialize creates the GPU context (device from SMALLK_B200_DEVICE, default 0).
void NmfInitialize(int argc, char* argv[]);
Result NmfIsInitialized();
void NmfFinalize();

// common/src/nmf_options.cpp:24-111
bool IsValid(const NmfOptions& opts, bool validate_matrix = true);

// common/src/nmf.cpp:173. Column-major buffers; W (m x k) and H (k x n) carry the initial guess in and the
// factors out. Throws std::logic_error if ldim_W < m or ldim_H < k, as the reference does.
Result Nmf(const NmfOptions& options,
           double* buf_A, int ldim_A,
           double* buf_W, int ldim_W,
           double* buf_H, int ldim_H,
           NmfStats& stats);

// common/src/nmf.cpp:232. CSC input, 32-bit indices.
Result NmfSparse(const NmfOptions& options,
                 const unsigned int height,
                 const unsigned int width,
                 const unsigned int nz,
                 const unsigned int* col_offsets,
                 const unsigned int* row_indices,
                 const double* data,
                 double* buf_W, int ldim_W,
                 double* buf_H, int ldim_H,
                 NmfStats& stats);

// Text of the last error reported by the GPU library (no counterpart in the reference, which prints to cerr).
const char* NmfLastError();
