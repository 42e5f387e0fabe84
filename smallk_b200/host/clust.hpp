// smallk_b200 host — hierarchical clustering by recursive rank-2 NMF (HierNMF2) on the GPU library.
// Mirrors hierclust/include/clust.hpp:25-63 of the reference: same structs, same function names, argument
// meaning and Result codes. The rank-2 factorizations, the column-subset extraction and the flat NNLS step run
// on the device through the C ABI (smk_select_columns, smk_nmf, smk_nnls_hals); the tree, the trial-split policy
// and the priority scores are host control flow (hierclust/include/clust_hier_generic.hpp:77-517,
// clust_hier_util.hpp:105-173).
#pragma once

#include <string>
#include <vector>

#include "nmf.hpp"
#include "random.hpp"
#include "sparse_matrix.hpp"
#include "tree.hpp"

typedef double R;

struct ClustStats
{
    ClustStats() : nmf_count(0), max_count(0), iteration_count(0), t_extract(0), t_init(0), t_factor(0), t_priority(0), t_terms(0), t_priority_worker(0) {}
    int nmf_count;        // factorizations performed
    int max_count;        // factorizations that reached the iteration limit
    // not in the reference: throughput accounting
    long long iteration_count;                                  // rank-2 outer iterations over all factorizations
    double t_extract, t_init, t_factor, t_priority, t_terms;    // seconds on the calling thread: subset extraction, initialisers, smk_nmf, priority scores (evaluating or waiting for one), top terms
    double t_priority_worker;                                   // seconds of priority-score evaluation that ran on the worker thread, under the other child's factorization
};

struct ClustOptions
{
    NmfOptions nmf_opts;            // per-node rank-2 factorization: tol, min_iter, max_iter, ... (k and algorithm are overridden)
    int maxterms;                   // top terms kept per node
    R unbalanced;                   // a split with a child below this share of the parent is retried on the smaller child
    int trial_allowance, num_clusters;
    bool verbose, flat;             // flat: finish with the NnlsHals flat clustering from the leaf topic vectors
    std::string initdir;            // Winit_<i>.csv / Hinit_<i>.csv initialisers instead of the random ones
};

bool IsValid(const ClustOptions& opts, bool validate_matrix = true);

// buf_w (m x num_clusters) and buf_h (num_clusters x n) receive the flat-clustering factors when options.flat.
Result Clust(const ClustOptions& options, R* buf_A, int ldim_A, R* buf_w, R* buf_h, Tree<R>& tree, ClustStats& stats, Random& rng);
Result ClustSparse(const ClustOptions& options, const SparseMatrix<R>& A, R* buf_w, R* buf_h, Tree<R>& tree,
                   ClustStats& stats, Random& rng);

// The priority score of a node from its topic vector and the two topic vectors of its trial split
// (clust_hier_util.hpp:105-173): a product of two modified-NDCG values. W_parent: m, W_child: m x 2 (ld = m).
R compute_priority(const R* W_parent, const R* W_child, int m);
