// smallk_b200 host — extern "C" veneer over the C++ host interface (Clust/ClustSparse, FlatClust*), for language
// bindings (the reference ships a Cython binding, pysmallk; here ctypes loads these symbols). Tests call the
// C++ host layer through this file so that what is checked against the reference is the shipped driver code.
#include <chrono>
#include <cstring>
#include <iostream>
#include <vector>

#include "clust.hpp"
#include "flat_clust.hpp"
#include "host_internal.hpp"
#include "tree.hpp"
#include "hierclust_writer.hpp"
#include <sstream>
#include "matrix_io.hpp"
#include "flat_clust_output.hpp"
#include "random.hpp"

R compute_priority_on(smk_ctx* ctx, const R* W_parent, const R* W_child, int n);
R compute_priority_plain(smk_ctx* ctx, const R* W_parent, const R* W_child, int n);
R compute_priority_rows(smk_ctx* ctx, const R* W_parent, const R* W_child, int n, const unsigned int* child_rows, int n_child_rows,
                        const unsigned int* parent_rows = nullptr, int n_parent_rows = 0);

namespace {

double g_profile[6] = {0, 0, 0, 0, 0, 0};

void EnsureInit()
{
    if (Result::INITIALIZED != NmfIsInitialized()) { static int argc = 0; NmfInitialize(argc, nullptr); }
}

ClustOptions MakeOpts(int m, int n, int num_clusters, double tol, int min_iter, int max_iter, int maxterms, double unbalanced,
                      int trial_allowance, int flat, int normalize, int verbose)
{
    ClustOptions o;
    o.nmf_opts.tol = tol;
    o.nmf_opts.algorithm = NmfAlgorithm::RANK2;
    o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.nmf_opts.height = m; o.nmf_opts.width = n; o.nmf_opts.k = 2;
    o.nmf_opts.min_iter = min_iter; o.nmf_opts.max_iter = max_iter; o.nmf_opts.tolcount = 1;
    o.nmf_opts.max_threads = 1; o.nmf_opts.verbose = false; o.nmf_opts.normalize = (normalize != 0);
    o.maxterms = maxterms; o.unbalanced = unbalanced; o.trial_allowance = trial_allowance;
    o.num_clusters = num_clusters; o.verbose = (verbose != 0); o.flat = (flat != 0);
    return o;
}

void ExportTree(Tree<R>& tree, int num_clusters, int maxterms, int n, int* assignments, int* parent, int* left, int* right,
                int* is_left, int* doc_count, int* terms, double* priority, int* is_leaf, int* n_outliers)
{
    const int nodes = 2 * (num_clusters - 1);
    for (int q = 0; q < nodes; ++q)
    {
        const Tree<R>::NodeView v = tree.Node(q);
        parent[q] = v.parent; left[q] = v.left; right[q] = v.right; is_left[q] = v.is_left_child ? 1 : 0;
        doc_count[q] = static_cast<int>(v.doc_count);
        for (int t = 0; t < maxterms; ++t) terms[q * maxterms + t] = (t < static_cast<int>(v.terms->size())) ? (*v.terms)[t] : -1;
        if (priority) priority[q] = v.is_valid ? v.priority : 0.0;
        if (is_leaf) is_leaf[q] = v.is_leaf ? 1 : 0;
    }
    const std::vector<unsigned int>& a = tree.Assignments();
    for (int j = 0; j < n; ++j) assignments[j] = (a[j] == Tree<R>::NONE) ? -1 : static_cast<int>(a[j]);
    if (n_outliers) *n_outliers = static_cast<int>(tree.Outliers().size());
}

int Finish(Result r, const ClustOptions& o, Tree<R>& tree, ClustStats& cs, int n, int* assignments, int* parent, int* left,
           int* right, int* is_left, int* doc_count, int* terms, double* priority, int* is_leaf, int* n_outliers,
           const std::vector<R>& w, const std::vector<R>& h, double* buf_w, double* buf_h, long long* stats, int* flat_assignments)
{
    if (stats) { stats[0] = cs.nmf_count; stats[1] = cs.max_count; stats[2] = cs.iteration_count; }
    g_profile[0] = cs.t_extract; g_profile[1] = cs.t_init; g_profile[2] = cs.t_factor; g_profile[3] = cs.t_priority; g_profile[4] = cs.t_terms; g_profile[5] = cs.t_priority_worker;
    if (Result::OK != r) return static_cast<int>(r);
    ExportTree(tree, o.num_clusters, o.maxterms, n, assignments, parent, left, right, is_left, doc_count, terms, priority, is_leaf, n_outliers);
    if (o.flat)
    {
        if (buf_w) std::memcpy(buf_w, w.data(), sizeof(R) * w.size());
        if (buf_h) std::memcpy(buf_h, h.data(), sizeof(R) * h.size());
        if (flat_assignments)
        {
            std::vector<unsigned int> fa;
            ComputeAssignments(fa, h.data(), o.num_clusters, o.num_clusters, n);
            for (int j = 0; j < n; ++j) flat_assignments[j] = static_cast<int>(fa[j]);
        }
    }
    return 0;
}

} // namespace

extern "C" {

const char* smkh_last_error() { return NmfLastError(); }

// seconds spent by the last smkh_hierclust_* call in: subset extraction, initialisers, smk_nmf, priority scores, top terms
void smkh_last_hier_profile(double* out6) { for (int i = 0; i < 6; ++i) out6[i] = g_profile[i]; }

// ClustSparse (hierclust/src/clust.cpp:160). stats[0..2] = nmf_count, max_count, total rank-2 iterations.
int smkh_hierclust_sparse(int m, int n, unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                          const double* data, int num_clusters, double tol, int min_iter, int max_iter, int maxterms,
                          double unbalanced, int trial_allowance, int flat, int normalize, int seed, int verbose,
                          int* assignments, int* parent, int* left, int* right, int* is_left, int* doc_count, int* terms,
                          double* priority, int* is_leaf, int* n_outliers, double* buf_w, double* buf_h, long long* stats,
                          int* flat_assignments, double* elapsed_s)
{
    try
    {
        EnsureInit();
        ClustOptions o = MakeOpts(m, n, num_clusters, tol, min_iter, max_iter, maxterms, unbalanced, trial_allowance, flat, normalize, verbose);
        SparseMatrix<R> A(m, n, nz, col_offsets, row_indices, data);
        Tree<R> tree; ClustStats cs; Random rng; rng.SeedFromInt(seed);
        std::vector<R> w(static_cast<size_t>(m) * num_clusters), h(static_cast<size_t>(num_clusters) * n);
        const auto t0 = std::chrono::steady_clock::now();
        const Result r = ClustSparse(o, A, w.data(), h.data(), tree, cs, rng);
        if (elapsed_s) *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return Finish(r, o, tree, cs, n, assignments, parent, left, right, is_left, doc_count, terms, priority, is_leaf, n_outliers,
                      w, h, buf_w, buf_h, stats, flat_assignments);
    }
    catch (std::exception& e) { std::cerr << "smkh_hierclust_sparse: " << e.what() << std::endl; return -100; }
}

// Clust (hierclust/src/clust.cpp:108)
int smkh_hierclust_dense(int m, int n, double* A, int ldA, int num_clusters, double tol, int min_iter, int max_iter,
                         int maxterms, double unbalanced, int trial_allowance, int flat, int normalize, int seed, int verbose,
                         int* assignments, int* parent, int* left, int* right, int* is_left, int* doc_count, int* terms,
                         double* priority, int* is_leaf, int* n_outliers, double* buf_w, double* buf_h, long long* stats,
                         int* flat_assignments, double* elapsed_s)
{
    try
    {
        EnsureInit();
        ClustOptions o = MakeOpts(m, n, num_clusters, tol, min_iter, max_iter, maxterms, unbalanced, trial_allowance, flat, normalize, verbose);
        Tree<R> tree; ClustStats cs; Random rng; rng.SeedFromInt(seed);
        std::vector<R> w(static_cast<size_t>(m) * num_clusters), h(static_cast<size_t>(num_clusters) * n);
        const auto t0 = std::chrono::steady_clock::now();
        const Result r = Clust(o, A, ldA, w.data(), h.data(), tree, cs, rng);
        if (elapsed_s) *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return Finish(r, o, tree, cs, n, assignments, parent, left, right, is_left, doc_count, terms, priority, is_leaf, n_outliers,
                      w, h, buf_w, buf_h, stats, flat_assignments);
    }
    catch (std::exception& e) { std::cerr << "smkh_hierclust_dense: " << e.what() << std::endl; return -100; }
}

// FlatClust / FlatClustSparse + ComputeAssignments + TopTerms. csc == null pointers -> dense A.
int smkh_flatclust(int alg, int m, int n, int k, double tol, int min_iter, int max_iter, int maxterms,
                   double* A, int ldA, unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                   const double* data, double* W, double* H, int* assignments, int* term_indices, int* iterations)
{
    try
    {
        EnsureInit();
        NmfOptions o;
        o.tol = tol; o.algorithm = static_cast<NmfAlgorithm>(alg); o.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
        o.height = m; o.width = n; o.k = k; o.min_iter = min_iter; o.max_iter = max_iter; o.tolcount = 1;
        o.max_threads = 1; o.verbose = false; o.normalize = true;
        NmfStats st;
        const Result r = A ? FlatClust(o, A, ldA, W, m, H, k, st)
                           : FlatClustSparse(o, m, n, nz, col_offsets, row_indices, data, W, m, H, k, st);
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
    catch (std::exception& e) { std::cerr << "smkh_flatclust: " << e.what() << std::endl; return -100; }
}

// compute_priority (clust_hier_util.hpp:105-173) on host buffers — unit-testable without a GPU.
double smkh_compute_priority(const double* W_parent, const double* W_child, int m) { return compute_priority(W_parent, W_child, m); }
// the full-length evaluation the streaming one is checked against
double smkh_compute_priority_plain(const double* W_parent, const double* W_child, int m) { return compute_priority_plain(nullptr, W_parent, W_child, m); }
// the evaluation the tree driver uses: the caller names the rows where the child factors can be non-zero
double smkh_compute_priority_rows(const double* W_parent, const double* W_child, int m, const unsigned int* child_rows, int n_child_rows)
{
    return compute_priority_rows(nullptr, W_parent, W_child, m, child_rows, n_child_rows);
}
// ... and with the rows named outside of which the parent vector is zero (what the tree driver passes below the root)
double smkh_compute_priority_rows2(const double* W_parent, const double* W_child, int m, const unsigned int* child_rows, int n_child_rows,
                                   const unsigned int* parent_rows, int n_parent_rows)
{
    return compute_priority_rows(nullptr, W_parent, W_child, m, child_rows, n_child_rows, parent_rows, n_parent_rows);
}
// the same with the large sorts on the GPU (what the tree driver uses): must give the identical value
double smkh_compute_priority_gpu(const double* W_parent, const double* W_child, int m)
{
    EnsureInit();
    return compute_priority_on(NmfContext(), W_parent, W_child, m);
}

// ---- file formats (host/matrix_io.hpp), unit-testable without a GPU: same contract as oracle/ref_io_capi.cpp ----
static std::string g_io_error;
const char* smkh_io_last_exception() { return g_io_error.c_str(); }

int smkh_load_matrix_market(const char* path, unsigned int* height, unsigned int* width, unsigned int* nnz,
                            unsigned int* col_offsets, unsigned int cap_cols, unsigned int* row_indices, double* data,
                            unsigned int cap_nz)
{
    smallk_io::CscMatrix A;
    try { if (!smallk_io::LoadMatrixMarketFile(path, A)) return -1; }
    catch (std::exception& e) { g_io_error = e.what(); return -3; }      // the reader throws where the reference's throws
    *height = A.height; *width = A.width; *nnz = A.nnz();
    if (A.width + 1 > cap_cols || A.nnz() > cap_nz) return -2;
    for (unsigned int c = 0; c <= A.width; ++c) col_offsets[c] = A.col_offsets[c];
    for (unsigned int e = 0; e < A.nnz(); ++e) { row_indices[e] = A.row_indices[e]; data[e] = A.data[e]; }
    return 0;
}

int smkh_load_delimited(const char* path, unsigned int* height, unsigned int* width, double* buffer, unsigned int cap)
{
    std::vector<double> buf;
    unsigned int h = 0, w = 0;
    if (!smallk_io::LoadDelimitedFile(buf, h, w, path)) return -1;
    *height = h; *width = w;
    if (static_cast<size_t>(h) * w > cap) return -2;
    for (size_t i = 0; i < static_cast<size_t>(h) * w; ++i) buffer[i] = buf[i];
    return 0;
}

int smkh_write_delimited(const double* buffer, unsigned int ldim, unsigned int height, unsigned int width, const char* path,
                         unsigned int precision)
{
    return smallk_io::WriteDelimitedFile(buffer, ldim, height, width, path, precision) ? 0 : -1;
}

// ---- clustering post-processing (host/flat_clust.cpp), same contract as oracle/ref_io_capi.cpp ----
void smkh_compute_assignments(const double* H, unsigned int ldH, unsigned int k, unsigned int n, unsigned int* out)
{
    std::vector<unsigned int> a;
    ComputeAssignments(a, H, ldH, k, n);
    for (unsigned int j = 0; j < n; ++j) out[j] = a[j];
}

void smkh_compute_fuzzy_assignments(const double* H, unsigned int ldH, unsigned int k, unsigned int n, float* out)
{
    std::vector<float> p;
    ComputeFuzzyAssignments(p, H, ldH, k, n);
    for (size_t i = 0; i < static_cast<size_t>(k) * n; ++i) out[i] = p[i];
}

void smkh_top_terms_matrix(int maxterms, const double* W, unsigned int ldW, unsigned int m, unsigned int k, int* out)
{
    std::vector<int> ti(static_cast<size_t>(maxterms) * k);
    TopTerms(maxterms, W, ldW, m, k, ti);
    for (size_t i = 0; i < ti.size(); ++i) out[i] = ti[i];
}

// ---- random initialisers (host/random.hpp), same contract as oracle/ref_io_capi.cpp ----
void smkh_random_matrices(int seed, unsigned int h1, unsigned int w1, double* buf1, unsigned int h2, unsigned int w2, double* buf2,
                          int* next_int)
{
    Random rng;
    rng.SeedFromInt(seed);
    RandomMatrix(buf1, h1, h1, w1, rng, 0.5, 0.5);
    RandomMatrix(buf2, h2, h2, w2, rng, 0.5, 0.5);
    if (next_int) *next_int = rng.RandomInt();
}

// ---- option validation (host/nmf_host.cpp, clust.cpp, flat_clust.cpp), same contract as oracle/ref_io_capi.cpp ----
// which: 0 = IsValid(NmfOptions), 1 = IsValid(ClustOptions), 2 = IsValid(FlatClustOptions). v = {tol, algorithm, prog, height, width, k,
// min_iter, max_iter, tolcount, max_threads, maxterms, unbalanced, trial_allowance, num_clusters}.
int smkh_is_valid(int which, const double* v, int validate_matrix)
{
    NmfOptions o;
    o.tol = v[0]; o.algorithm = static_cast<NmfAlgorithm>(static_cast<int>(v[1]));
    o.prog_est_algorithm = static_cast<NmfProgressAlgorithm>(static_cast<int>(v[2]));
    o.height = static_cast<int>(v[3]); o.width = static_cast<int>(v[4]); o.k = static_cast<int>(v[5]);
    o.min_iter = static_cast<int>(v[6]); o.max_iter = static_cast<int>(v[7]); o.tolcount = static_cast<int>(v[8]);
    o.max_threads = static_cast<int>(v[9]); o.verbose = false; o.normalize = true;
    if (which == 0) return IsValid(o, validate_matrix != 0) ? 1 : 0;
    if (which == 1)
    {
        ClustOptions c;
        c.nmf_opts = o; c.maxterms = static_cast<int>(v[10]); c.unbalanced = v[11]; c.trial_allowance = static_cast<int>(v[12]);
        c.num_clusters = static_cast<int>(v[13]); c.verbose = false; c.flat = false;
        return IsValid(c, validate_matrix != 0) ? 1 : 0;
    }
    FlatClustOptions f;
    f.nmf_opts = o; f.maxterms = static_cast<int>(v[10]); f.num_clusters = static_cast<int>(v[13]); f.verbose = false;
    return IsValid(f, validate_matrix != 0) ? 1 : 0;
}

// ---- flat clustering result files (host/flat_clust_output.hpp), same contract as oracle/ref_io_capi.cpp ----
int smkh_flatclust_write_results(const char* outdir, const unsigned int* assignments, const float* probabilities, const char* dictionary,
                                 int dict_count, const int* term_indices, int format, unsigned int maxterms, unsigned int num_docs,
                                 unsigned int num_clusters)
{
    std::vector<unsigned int> a(assignments, assignments + num_docs);
    std::vector<float> p(probabilities, probabilities + static_cast<size_t>(num_clusters) * num_docs);
    std::vector<std::string> d;
    const char* s = dictionary;
    for (int i = 0; i < dict_count; ++i) { d.push_back(std::string(s)); s += d.back().size() + 1; }
    std::vector<int> t(term_indices, term_indices + static_cast<size_t>(maxterms) * num_clusters);
    try { FlatClustWriteResults(std::string(outdir), a, p, d, t, static_cast<FileFormat>(format), maxterms, num_docs, num_clusters); }
    catch (std::exception&) { return -1; }
    return 0;
}

// ---- the tree and its writers driven by a script (host/tree.hpp, hierclust_writer.hpp), same contract as oracle/ref_io_capi.cpp ----
// compact != 0: the same script through the operations the hierclust driver uses — every split goes in with W held on its non-zero
// rows only (SplitCompact), and before it another leaf (the next best one, if there is one) is split and taken back again
// (UndoSplit), as the driver does when a split made ahead of time turns out wrong. The files must not change.
static int tree_script(int seed, int m, int n, int num_clusters, int maxterms, int format, const char* assign_path, const char* tree_path,
                       int compact)
{
    // deterministic pseudo-random stream shared by both drivers (values in (0, 1), a quarter of them exactly 0)
    unsigned long long state = 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(seed + 1);
    auto next = [&state]() {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const unsigned int bits = static_cast<unsigned int>(state >> 33);
        if ((bits & 3u) == 0u) return 0.0;
        return (static_cast<double>(bits >> 2) + 0.5) / 536870912.0;
    };

    Tree<double> tree;
    tree.Init(num_clusters, 2 * (num_clusters - 1), m, n);
    std::vector<unsigned int> doc_count(2 * (num_clusters - 1), 0u);
    std::vector<double> W(static_cast<size_t>(m) * 2), H(static_cast<size_t>(n) * 2);
    auto fill = [&](std::vector<double>& M, size_t count) { M.resize(count); for (size_t i = 0; i < count; ++i) M[i] = next(); };   // column-major order
    fill(W, static_cast<size_t>(m) * 2); fill(H, static_cast<size_t>(n) * 2);
    tree.SplitRoot(W.data(), H.data(), n);
    for (int split = 0; ; ++split)
    {
        const unsigned int i0 = tree.LeftChildIndex(), i1 = tree.RightChildIndex();
        doc_count[i0] = tree.LeftChildDocs().size(); doc_count[i1] = tree.RightChildDocs().size();
        tree.SetNodePriority(i0, doc_count[i0] > 3 ? next() + 0.01 : -1.0);
        tree.SetNodePriority(i1, doc_count[i1] > 3 ? next() + 0.01 : -1.0);
        if (split == num_clusters - 2) break;
        double mn = 0, mx = 0; unsigned int idx = 0;
        tree.MinMaxLeafPriorities(mn, mx, idx);
        if (mx < 0.0) break;
        std::vector<double> Hs;
        fill(W, static_cast<size_t>(m) * 2); fill(Hs, static_cast<size_t>(doc_count[idx]) * 2);
        if (!compact) { tree.Split(idx, W.data(), Hs.data(), doc_count[idx]); continue; }
        std::vector<unsigned int> rows;
        std::vector<double> Wc;
        for (int r = 0; r < m; ++r) if (W[r] != 0.0 || W[static_cast<size_t>(m) + r] != 0.0) rows.push_back(static_cast<unsigned int>(r));
        Wc.resize(rows.size() * 2);
        for (size_t r = 0; r < rows.size(); ++r) { Wc[r] = W[rows[r]]; Wc[rows.size() + r] = W[static_cast<size_t>(m) + rows[r]]; }
        double mn2 = 0, mx2 = 0; unsigned int other = 0;
        tree.MinMaxLeafPrioritiesWithout(idx, Tree<double>::NONE, mn2, mx2, other);
        if (mx2 >= 0.0 && doc_count[other] > 0)
        {
            // a decoy: split the runner-up with some factors, look at its children, take it back
            std::vector<double> Hd(static_cast<size_t>(doc_count[other]) * 2);
            for (size_t q = 0; q < Hd.size(); ++q) Hd[q] = static_cast<double>((q * 7 + split) % 5);
            tree.SplitCompact(other, rows.data(), static_cast<unsigned int>(rows.size()), Wc.data(), Hd.data(), doc_count[other]);
            if (tree.LeftChildDocs().size() + tree.RightChildDocs().size() != doc_count[other]) return -3;
            tree.UndoSplit(other);
        }
        tree.SplitCompact(idx, rows.data(), static_cast<unsigned int>(rows.size()), Wc.data(), Hs.data(), doc_count[idx]);
    }
    tree.ComputeTopTerms(maxterms);
    tree.ComputeAssignments();
    if (!tree.WriteAssignments(std::string(assign_path))) return -1;
    std::vector<std::string> dict;
    for (int i = 0; i < m; ++i) { std::ostringstream s; s << "w" << i; dict.push_back(s.str()); }
    IHierclustWriter* writer = CreateHierclustWriter(static_cast<FileFormat>(format));
    const bool ok = tree.WriteTree(writer, std::string(tree_path), dict);
    delete writer;
    return ok ? 0 : -2;
}

int smkh_tree_script(int seed, int m, int n, int num_clusters, int maxterms, int format, const char* assign_path, const char* tree_path)
{
    return tree_script(seed, m, n, num_clusters, maxterms, format, assign_path, tree_path, 0);
}
int smkh_tree_script_compact(int seed, int m, int n, int num_clusters, int maxterms, int format, const char* assign_path, const char* tree_path)
{
    return tree_script(seed, m, n, num_clusters, maxterms, format, assign_path, tree_path, 1);
}

// dictionary files (host/flat_clust_output.hpp LoadStringsFromFile): the strings joined by '\n' into out (capacity cap)
int smkh_load_strings(const char* path, char* out, unsigned int cap, int* count)
{
    std::vector<std::string> v(1, "preexisting");              // the reader appends
    if (!LoadStringsFromFile(path, v)) return -1;
    std::string joined;
    for (const auto& t : v) { joined += t; joined += '\n'; }
    if (joined.size() + 1 > cap) return -2;
    std::memcpy(out, joined.c_str(), joined.size() + 1);
    *count = static_cast<int>(v.size());
    return 0;
}

} // extern "C"
