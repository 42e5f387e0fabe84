// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// ref_hier_capi.cpp: extern "C" entry points around the REFERENCE's own hierarchical / flat clustering
// drivers, compiled by oracle/Makefile with the reference's unmodified sources into
// oracle/_ref/libsmallk_ref.so. Nothing here restates an algorithm; every call lands in reference code:
//   ClustSparse / Clust        hierclust/src/clust.cpp:108,160
//   ClustHier/TrialSplit/...   hierclust/include/clust_hier_generic.hpp:77-517
//   Tree<R>                    hierclust/include/tree.hpp
//   FlatClust/FlatClustSparse  flatclust/src/flat_clust.cpp:144,203
//   ComputeAssignments         common/include/assignments.hpp:80-113
// The tree is read back through the reference's own IHierclustWriter interface (tree.hpp:510-546) with a
// writer that records what it is told instead of formatting it, and through Tree::Assignments().

#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>
#include <map>
#include <set>
#include <limits>
#include <string>
#include <cassert>
#include <algorithm>
#include <sstream>
#include "file_format.hpp"
#include "dense_matrix.hpp"
#include "vector_utils.hpp"
#include "hierclust_writer.hpp"

// node priorities and document lists are private members of Tree<T>; the checker reads them for diagnostics
#define private public
#include "tree.hpp"
#undef private
#include "clust.hpp"
#include "nmf.hpp"
#include "random.hpp"
#include "flat_clust.hpp"
#include "assignments.hpp"
#include "terms.hpp"
#include "thread_utils.hpp"
#include "sparse_matrix.hpp"
#include "clust_hier_util.hpp"

typedef double R;

namespace {

struct CaptureWriter : public IHierclustWriter
{
    std::vector<int> parent, left, right, is_left, doc_count;
    std::vector<std::vector<int>> terms;
    int leaf_docs = 0;
    void WriteHeader(std::ofstream&, const int doc_count_) override { leaf_docs = doc_count_; }
    void WriteNodeBegin(std::ofstream&, const int) override {}
    void WriteParentId(std::ofstream&, const int parent_id) override { parent.push_back(parent_id); }
    void WriteLeftChild(std::ofstream&, const bool is_left_child, const int lc_label) override
    { is_left.push_back(is_left_child ? 1 : 0); left.push_back(lc_label); }
    void WriteRightChild(std::ofstream&, const int rc_label) override { right.push_back(rc_label); }
    void WriteDocCount(std::ofstream&, const int count) override { doc_count.push_back(count); }
    void WriteTopTerms(std::ofstream&, const std::vector<int>& term_indices, const std::vector<std::string>&) override
    { terms.push_back(term_indices); }
    void WriteNodeEnd(std::ofstream&) override {}
    void WriteFooter(std::ofstream&) override {}
};

void EnsureInitH()
{
    if (Result::INITIALIZED != NmfIsInitialized())
    {
        static int argc = 0;
        static char** argv = nullptr;
        NmfInitialize(argc, argv);
    }
}

ClustOptions MakeClustOpts(int m, int n, int num_clusters, double tol, int min_iter, int max_iter, int maxterms,
                           double unbalanced, int trial_allowance, int flat, int max_threads, int normalize)
{
    ClustOptions o;
    o.nmf_opts.tol = tol;
    o.nmf_opts.algorithm = NmfAlgorithm::RANK2;
    o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.nmf_opts.height = m; o.nmf_opts.width = n; o.nmf_opts.k = 2;
    o.nmf_opts.min_iter = min_iter; o.nmf_opts.max_iter = max_iter; o.nmf_opts.tolcount = 1;
    o.nmf_opts.max_threads = max_threads;
    o.nmf_opts.verbose = false;
    o.nmf_opts.normalize = (normalize != 0);
    o.maxterms = maxterms;
    o.unbalanced = unbalanced;
    o.trial_allowance = trial_allowance;
    o.num_clusters = num_clusters;
    o.verbose = false;
    o.flat = (flat != 0);
    return o;
}

// Copies the tree out. node arrays have 2*(num_clusters-1) entries; terms is node_count x maxterms (-1 padded).
void ExportTree(Tree<R>& tree, int num_clusters, int maxterms, int n,
                int* assignments, int* parent, int* left, int* right, int* is_left, int* doc_count,
                int* terms, double* priority, int* is_leaf, int* n_outliers)
{
    CaptureWriter w;
    std::vector<std::string> dict;
    tree.WriteTree(&w, "/dev/null", dict);
    const int node_count = 2 * (num_clusters - 1);
    for (int q = 0; q < node_count; ++q)
    {
        parent[q] = w.parent[q]; left[q] = w.left[q]; right[q] = w.right[q];
        is_left[q] = w.is_left[q]; doc_count[q] = w.doc_count[q];
        for (int t = 0; t < maxterms; ++t)
            terms[q * maxterms + t] = (t < static_cast<int>(w.terms[q].size())) ? w.terms[q][t] : -1;
        if (priority) priority[q] = tree.nodes_[q].is_valid ? tree.nodes_[q].priority : 0.0;
        if (is_leaf) is_leaf[q] = tree.is_leaf_[q] ? 1 : 0;
    }
    const std::vector<unsigned int>& a = tree.Assignments();
    for (int j = 0; j < n; ++j) assignments[j] = (a[j] == 0xFFFFFFFFu) ? -1 : static_cast<int>(a[j]);
    if (n_outliers) *n_outliers = static_cast<int>(tree.Outliers().size());
}

} // namespace

extern "C" {

// HierNMF2 on a sparse (CSC) matrix through the reference's ClustSparse. With max_threads = 1 every random
// initialiser comes from the sequential generator (matrix_generator.hpp:229-248), so the run is a
// deterministic function of `seed`. buf_w (m x num_clusters) and buf_h (num_clusters x n) receive the flat
// clustering factors when flat != 0. stats[0] = nmf_count, stats[1] = max_count.
int ref_hierclust_sparse(int m, int n, unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                         const double* data, int num_clusters, double tol, int min_iter, int max_iter, int maxterms,
                         double unbalanced, int trial_allowance, int flat, int normalize, int seed, int max_threads,
                         int* assignments, int* parent, int* left, int* right, int* is_left, int* doc_count, int* terms,
                         double* priority, int* is_leaf, int* n_outliers, double* buf_w, double* buf_h, int* stats,
                         int* flat_assignments, double* elapsed_s)
{
    EnsureInitH();
    ClustOptions opts = MakeClustOpts(m, n, num_clusters, tol, min_iter, max_iter, maxterms, unbalanced, trial_allowance,
                                      flat, max_threads, normalize);
    SparseMatrix<R> A(m, n, nz, col_offsets, row_indices, data);
    Tree<R> tree;
    ClustStats cs;
    Random rng;
    rng.SeedFromInt(seed);
    std::vector<R> w(static_cast<size_t>(m) * num_clusters), h(static_cast<size_t>(num_clusters) * n);
    Result r;
    const auto t0 = std::chrono::steady_clock::now();
    try { r = ClustSparse(opts, A, &w[0], &h[0], tree, cs, rng); }
    catch (std::exception& e) { std::cerr << "ref_hier_capi: " << e.what() << std::endl; return -100; }
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) { stats[0] = cs.nmf_count; stats[1] = cs.max_count; }
    if (Result::OK != r) return static_cast<int>(r);
    ExportTree(tree, num_clusters, maxterms, n, assignments, parent, left, right, is_left, doc_count, terms, priority,
               is_leaf, n_outliers);
    if (flat)
    {
        if (buf_w) std::memcpy(buf_w, &w[0], sizeof(R) * w.size());
        if (buf_h) std::memcpy(buf_h, &h[0], sizeof(R) * h.size());
        if (flat_assignments)
        {
            std::vector<unsigned int> fa;
            ComputeAssignments(fa, &h[0], num_clusters, num_clusters, n);
            for (int j = 0; j < n; ++j) flat_assignments[j] = static_cast<int>(fa[j]);
        }
    }
    return 0;
}

// Same for a dense column-major A through the reference's Clust.
int ref_hierclust_dense(int m, int n, double* A, int ldA, int num_clusters, double tol, int min_iter, int max_iter,
                        int maxterms, double unbalanced, int trial_allowance, int flat, int normalize, int seed,
                        int max_threads, int* assignments, int* parent, int* left, int* right, int* is_left,
                        int* doc_count, int* terms, double* priority, int* is_leaf, int* n_outliers, double* buf_w,
                        double* buf_h, int* stats, int* flat_assignments)
{
    EnsureInitH();
    ClustOptions opts = MakeClustOpts(m, n, num_clusters, tol, min_iter, max_iter, maxterms, unbalanced, trial_allowance,
                                      flat, max_threads, normalize);
    Tree<R> tree;
    ClustStats cs;
    Random rng;
    rng.SeedFromInt(seed);
    std::vector<R> w(static_cast<size_t>(m) * num_clusters), h(static_cast<size_t>(num_clusters) * n);
    Result r;
    try { r = Clust(opts, A, ldA, &w[0], &h[0], tree, cs, rng); }
    catch (std::exception& e) { std::cerr << "ref_hier_capi: " << e.what() << std::endl; return -100; }
    if (stats) { stats[0] = cs.nmf_count; stats[1] = cs.max_count; }
    if (Result::OK != r) return static_cast<int>(r);
    ExportTree(tree, num_clusters, maxterms, n, assignments, parent, left, right, is_left, doc_count, terms, priority,
               is_leaf, n_outliers);
    if (flat)
    {
        if (buf_w) std::memcpy(buf_w, &w[0], sizeof(R) * w.size());
        if (buf_h) std::memcpy(buf_h, &h[0], sizeof(R) * h.size());
        if (flat_assignments)
        {
            std::vector<unsigned int> fa;
            ComputeAssignments(fa, &h[0], num_clusters, num_clusters, n);
            for (int j = 0; j < n; ++j) flat_assignments[j] = static_cast<int>(fa[j]);
        }
    }
    return 0;
}

// flatclust/src/flat_clust.cpp:144,203 + assignments.hpp + terms.hpp: factors, assignments and top terms.
int ref_flatclust_dense(int alg, int m, int n, int k, double tol, int min_iter, int max_iter, int max_threads,
                        int maxterms, double* A, int ldA, double* W, double* H, int* assignments, int* term_indices,
                        int* iterations)
{
    EnsureInitH();
    FlatClustOptions o;
    o.nmf_opts.tol = tol;
    o.nmf_opts.algorithm = static_cast<NmfAlgorithm>(alg);
    o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.nmf_opts.height = m; o.nmf_opts.width = n; o.nmf_opts.k = k;
    o.nmf_opts.min_iter = min_iter; o.nmf_opts.max_iter = max_iter; o.nmf_opts.tolcount = 1;
    o.nmf_opts.max_threads = max_threads; o.nmf_opts.verbose = false; o.nmf_opts.normalize = true;
    o.maxterms = maxterms; o.num_clusters = k; o.verbose = false;
    NmfStats st;
    Result r;
    try { r = FlatClust(o.nmf_opts, A, ldA, W, m, H, k, st); }
    catch (std::exception& e) { std::cerr << "ref_hier_capi: " << e.what() << std::endl; return -100; }
    if (iterations) *iterations = st.iteration_count;
    if (Result::OK != r) return static_cast<int>(r);
    std::vector<unsigned int> fa;
    ComputeAssignments(fa, H, k, k, n);
    for (int j = 0; j < n; ++j) assignments[j] = static_cast<int>(fa[j]);
    std::vector<int> ti(static_cast<size_t>(maxterms) * k);
    TopTerms(maxterms, W, m, m, k, ti);
    for (size_t i = 0; i < ti.size(); ++i) term_indices[i] = ti[i];
    return 0;
}

int ref_flatclust_sparse(int alg, int m, int n, int k, double tol, int min_iter, int max_iter, int max_threads,
                         int maxterms, unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                         const double* data, double* W, double* H, int* assignments, int* term_indices, int* iterations)
{
    EnsureInitH();
    NmfOptions o;
    o.tol = tol;
    o.algorithm = static_cast<NmfAlgorithm>(alg);
    o.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.height = m; o.width = n; o.k = k;
    o.min_iter = min_iter; o.max_iter = max_iter; o.tolcount = 1;
    o.max_threads = max_threads; o.verbose = false; o.normalize = true;
    NmfStats st;
    Result r;
    try { r = FlatClustSparse(o, m, n, nz, col_offsets, row_indices, data, W, m, H, k, st); }
    catch (std::exception& e) { std::cerr << "ref_hier_capi: " << e.what() << std::endl; return -100; }
    if (iterations) *iterations = st.iteration_count;
    if (Result::OK != r) return static_cast<int>(r);
    std::vector<unsigned int> fa;
    ComputeAssignments(fa, H, k, k, n);
    for (int j = 0; j < n; ++j) assignments[j] = static_cast<int>(fa[j]);
    std::vector<int> ti(static_cast<size_t>(maxterms) * k);
    TopTerms(maxterms, W, m, m, k, ti);
    for (size_t i = 0; i < ti.size(); ++i) term_indices[i] = ti[i];
    return 0;
}

// compute_priority, hierclust/include/clust_hier_util.hpp:105-173
double ref_compute_priority(double* W_parent, double* W_child, int m)
{
    DenseMatrix<R> P(m, 1, W_parent, m), C(m, 2, W_child, m);
    return compute_priority(P, C);
}

// TopTerms for one column, common/include/terms.hpp:24-60
void ref_top_terms(int maxterms, double* v, int m, int* out)
{
    DenseMatrix<R> V(m, 1, v, m);
    std::vector<int> sort_indices(m), terms(maxterms, -1);
    TopTerms(maxterms, V, sort_indices, terms);
    for (int i = 0; i < maxterms; ++i) out[i] = terms[i];
}

} // extern "C"
