// Implementation of host/flat_clust.hpp over the NMF entry points of host/nmf.hpp (GPU library underneath).
#include "flat_clust.hpp"

#include <iostream>
#include <stdexcept>

#include "tree.hpp"      // TopTerms(single column)

bool IsValid(const FlatClustOptions& opts, bool validate_matrix)
{
    // flatclust/src/flat_clust_options.cpp:24-98: the same checks in the same order. Note what it does NOT look at: the
    // algorithm, tolcount and the RANK2 <-> k == 2 rule (FlatClust rejects an unknown algorithm later, flat_clust.cpp:60-75).
    using std::cerr; using std::endl;
    const NmfOptions& o = opts.nmf_opts;
    if (validate_matrix)
    {
        if (o.height <= 0) { cerr << "clustlib error: matrix height must be a positive integer" << endl; return false; }
        if (o.width <= 0) { cerr << "clustlib error: matrix width must be a positive integer" << endl; return false; }
        if (o.k <= 0) { cerr << "clustlib error: cluster count must be a positive integer" << endl; return false; }
        if (o.k > o.width) { cerr << "clustlib error: k value cannot exceed the matrix width" << endl; return false; }
    }
    if (opts.num_clusters <= 0) { cerr << "clustlib error: value for --clusters must be a positive integer" << endl; return false; }
    if (o.tol <= 0.0 || o.tol >= 1.0) { cerr << "clustlib error: tolerance must be in the interval (0.0, 1.0)" << endl; return false; }
    if (o.min_iter <= 0) { cerr << "clustlib error: miniter must be a positive integer" << endl; return false; }
    if (o.max_iter <= 0) { cerr << "clustlib error: iteration count must be a positive integer" << endl; return false; }
    if (opts.maxterms <= 0) { cerr << "clustlib error: maxterms must be a positive integer" << endl; return false; }
    if (NmfProgressAlgorithm::PG_RATIO != o.prog_est_algorithm && NmfProgressAlgorithm::DELTA_FNORM != o.prog_est_algorithm)
    { cerr << "clustlib error: unknown stopping criterion " << endl; return false; }
    return true;
}

namespace {
void check_algorithm(const NmfOptions& o)
{
    // flat_clust.cpp:52-81: HALS, RANK2 (k == 2) and BPP only
    if (NmfAlgorithm::RANK2 == o.algorithm && 2 != o.k) throw std::runtime_error("rank2 algorithm requires k == 2");
    if (NmfAlgorithm::HALS != o.algorithm && NmfAlgorithm::RANK2 != o.algorithm && NmfAlgorithm::BPP != o.algorithm)
        throw std::runtime_error("unknown NMF algorithm");
}
} // namespace

Result FlatClust(const NmfOptions& options, double* buf_A, int ldim_A, double* buf_W, int ldim_W, double* buf_H, int ldim_H,
                 NmfStats& stats)
{
    if (Result::INITIALIZED != NmfIsInitialized())
    {
        std::cerr << "flatclust error: NmfInitialize() must be called prior to any clustering routine\n" << std::endl;
        return Result::NOTINITIALIZED;
    }
    if (!IsValid(options)) return Result::BAD_PARAM;
    check_algorithm(options);
    return Nmf(options, buf_A, ldim_A, buf_W, ldim_W, buf_H, ldim_H, stats);
}

Result FlatClustSparse(const NmfOptions& options, const unsigned int height, const unsigned int width, const unsigned int nz,
                       const unsigned int* col_offsets, const unsigned int* row_indices, const double* data,
                       double* buf_W, int ldim_W, double* buf_H, int ldim_H, NmfStats& stats)
{
    if (Result::INITIALIZED != NmfIsInitialized())
    {
        std::cerr << "flatclust error: NmfInitialize() must be called prior to any clustering routine\n" << std::endl;
        return Result::NOTINITIALIZED;
    }
    if (!IsValid(options)) return Result::BAD_PARAM;
    check_algorithm(options);
    return NmfSparse(options, height, width, nz, col_offsets, row_indices, data, buf_W, ldim_W, buf_H, ldim_H, stats);
}

void ComputeAssignments(std::vector<unsigned int>& assignments, const double* buf_h, const unsigned int ldim_h,
                        const unsigned int k, const unsigned int n)
{
    if (k > n) throw std::logic_error("ComputeAssignments: dimensions of matrix H are invalid");
    assignments.assign(n, 0u);
    for (unsigned int c = 0; c < n; ++c)
    {
        const double* col = buf_h + static_cast<size_t>(c) * ldim_h;
        unsigned int best = 0;
        for (unsigned int r = 1; r < k; ++r) if (col[r] > col[best]) best = r;
        assignments[c] = best;
    }
}

void ComputeFuzzyAssignments(std::vector<float>& probabilities, const double* buf_h, const unsigned int ldim_h,
                             const unsigned int k, const unsigned int n)
{
    if (probabilities.size() < static_cast<size_t>(k) * n) probabilities.resize(static_cast<size_t>(k) * n);
    for (unsigned int c = 0; c < n; ++c)
    {
        const double* col = buf_h + static_cast<size_t>(c) * ldim_h;
        double sum = 0.0;
        for (unsigned int r = 0; r < k; ++r) sum += col[r];
        const double inv = 1.0 / sum;
        for (unsigned int r = 0; r < k; ++r) probabilities[static_cast<size_t>(c) * ldim_h + r] = static_cast<float>(col[r] * inv);
    }
}

void TopTerms(const int maxterms, const double* buf_w, const unsigned int ldim, const unsigned int height,
              const unsigned int width, std::vector<int>& term_indices)
{
    if (height < width) throw std::logic_error("TopTerms: height of W buffer must be >= width");
    if (term_indices.size() < static_cast<size_t>(maxterms) * width) term_indices.resize(static_cast<size_t>(maxterms) * width);
    std::vector<int> scratch, one(maxterms);
    for (unsigned int c = 0; c < width; ++c)
    {
        TopTerms<double>(maxterms, buf_w + static_cast<size_t>(c) * ldim, static_cast<int>(height), scratch, one);
        const int keep = std::min<int>(maxterms, static_cast<int>(height));
        for (int q = 0; q < keep; ++q) term_indices[static_cast<size_t>(c) * maxterms + q] = one[q];
    }
}
