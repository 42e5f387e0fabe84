// smallk_b200 host — flat clustering: the reference's FlatClust / FlatClustSparse (flatclust/include/flat_clust.hpp:21-46,
// flatclust/src/flat_clust.cpp:144-265) = the NMF solvers minus MU, plus the assignment / top-term post-processing
// of common/include/assignments.hpp:32-113 and common/include/terms.hpp:62-108.
#pragma once

#include <string>
#include <vector>

#include "nmf.hpp"

struct FlatClustOptions
{
    NmfOptions nmf_opts;
    int maxterms;
    int num_clusters;
    bool verbose;
};

bool IsValid(const FlatClustOptions& opts, bool validate_matrix = true);

Result FlatClust(const NmfOptions& options, double* buf_A, int ldim_A, double* buf_W, int ldim_W, double* buf_H, int ldim_H,
                 NmfStats& stats);
Result FlatClustSparse(const NmfOptions& options, const unsigned int height, const unsigned int width, const unsigned int nz,
                       const unsigned int* col_offsets, const unsigned int* row_indices, const double* data,
                       double* buf_W, int ldim_W, double* buf_H, int ldim_H, NmfStats& stats);

// assignments.hpp:80-113: cluster of document c = row of the first maximum of H(:,c)
void ComputeAssignments(std::vector<unsigned int>& assignments, const double* buf_h, const unsigned int ldim_h,
                        const unsigned int k, const unsigned int n);
// assignments.hpp:32-76: column-normalised H as float probabilities
void ComputeFuzzyAssignments(std::vector<float>& probabilities, const double* buf_h, const unsigned int ldim_h,
                             const unsigned int k, const unsigned int n);
// terms.hpp:62-108: the maxterms largest rows of every column of W, packed column after column
void TopTerms(const int maxterms, const double* buf_w, const unsigned int ldim, const unsigned int height,
              const unsigned int width, std::vector<int>& term_indices);
