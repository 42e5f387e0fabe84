// TEST INFRASTRUCTURE — not part of the product.
//
// ref_io_capi.cpp: extern "C" entry points around the REFERENCE's own file readers / writers, compiled by
// oracle/Makefile with the reference's unmodified sources into oracle/_ref/libsmallk_ref.so. Nothing here restates
// an algorithm; every call lands in reference code:
//   LoadMatrixMarketFile   common/include/sparse_matrix_io.hpp:117-260 (+ common/src/matrix_market_file.cpp)
//   LoadDelimitedFile      common/include/delimited_file.hpp:79-135
//   WriteDelimitedFile     common/include/delimited_file.hpp:49-76
#include <string>
#include <vector>
#include "sparse_matrix.hpp"
#include "sparse_matrix_io.hpp"
#include "delimited_file.hpp"

extern "C" {

// Returns 0 and the CSC arrays (up to the capacities given) of the matrix the reference builds from the file; -1 if the
// reference rejects the file; -2 if a capacity is too small (dimensions are still reported).
static std::string g_io_error;
const char* ref_io_last_exception() { return g_io_error.c_str(); }

int ref_load_matrix_market(const char* path, unsigned int* height, unsigned int* width, unsigned int* nnz,
                           unsigned int* col_offsets, unsigned int cap_cols, unsigned int* row_indices, double* data,
                           unsigned int cap_nz)
{
    SparseMatrix<double> A;
    unsigned int h = 0, w = 0, nz = 0;
    try { if (!LoadMatrixMarketFile(std::string(path), A, h, w, nz)) return -1; }
    catch (std::exception& e) { g_io_error = e.what(); return -3; }      // SparseMatrix::Load / Compress throw through the reader
    *height = A.Height(); *width = A.Width(); *nnz = A.Size();
    if (A.Width() + 1 > cap_cols || A.Size() > cap_nz) return -2;
    const unsigned int* cp = A.LockedColBuffer();
    const unsigned int* ri = A.LockedRowBuffer();
    const double* dv = A.LockedDataBuffer();
    for (unsigned int c = 0; c <= A.Width(); ++c) col_offsets[c] = cp[c];
    for (unsigned int e = 0; e < A.Size(); ++e) { row_indices[e] = ri[e]; data[e] = dv[e]; }
    return 0;
}

// Column-major buffer (ld = height) the reference reads from a delimited file.
int ref_load_delimited(const char* path, unsigned int* height, unsigned int* width, double* buffer, unsigned int cap)
{
    std::vector<double> buf;
    unsigned int h = 0, w = 0;
    if (!LoadDelimitedFile(buf, h, w, std::string(path))) return -1;
    *height = h; *width = w;
    if (static_cast<size_t>(h) * w > cap) return -2;
    for (size_t i = 0; i < static_cast<size_t>(h) * w; ++i) buffer[i] = buf[i];
    return 0;
}

int ref_write_delimited(const double* buffer, unsigned int ldim, unsigned int height, unsigned int width, const char* path,
                        unsigned int precision)
{
    return WriteDelimitedFile(buffer, ldim, height, width, std::string(path), precision) ? 0 : -1;
}

} // extern "C"

// ---- clustering post-processing -------------------------------------------------------------
//   ComputeAssignments / ComputeFuzzyAssignments   common/include/assignments.hpp:32-113
//   TopTerms                                       common/include/terms.hpp:62-108
#include "assignments.hpp"
#include "terms.hpp"

extern "C" {

void ref_compute_assignments(const double* H, unsigned int ldH, unsigned int k, unsigned int n, unsigned int* out)
{
    std::vector<unsigned int> a;
    ComputeAssignments(a, H, ldH, k, n);
    for (unsigned int j = 0; j < n; ++j) out[j] = a[j];
}

void ref_compute_fuzzy_assignments(const double* H, unsigned int ldH, unsigned int k, unsigned int n, float* out)
{
    std::vector<float> p;
    ComputeFuzzyAssignments(p, H, ldH, k, n);
    for (size_t i = 0; i < static_cast<size_t>(k) * n; ++i) out[i] = p[i];
}

void ref_top_terms_matrix(int maxterms, const double* W, unsigned int ldW, unsigned int m, unsigned int k, int* out)
{
    std::vector<int> ti(static_cast<size_t>(maxterms) * k);
    TopTerms(maxterms, W, ldW, m, k, ti);
    for (size_t i = 0; i < ti.size(); ++i) out[i] = ti[i];
}

} // extern "C"

// ---- random initialisers --------------------------------------------------------------------------------------
//   Random                 common/include/random.hpp:22-200
//   RandomMatrix           common/include/matrix_generator.hpp:61-82 (sequential generator when max_threads == 1)
#include "random.hpp"
#include "matrix_generator.hpp"
#include "thread_utils.hpp"

extern "C" {

// Two consecutive matrices from one seeded generator (W then H, as the clustering drivers draw them), max_threads threads.
void ref_random_matrices(int seed, int max_threads, unsigned int h1, unsigned int w1, double* buf1, unsigned int h2,
                         unsigned int w2, double* buf2, int* next_int)
{
    SetMaxThreadCount(max_threads);
    Random rng;
    rng.SeedFromInt(seed);
    RandomMatrix(buf1, h1, h1, w1, rng, 0.5, 0.5);
    RandomMatrix(buf2, h2, h2, w2, rng, 0.5, 0.5);
    if (next_int) *next_int = rng.RandomInt();
}

} // extern "C"

// ---- option validation: common/src/nmf_options.cpp:23-110, hierclust/src/clust_options.cpp:24-110, flatclust/src/flat_clust_options.cpp
#include "nmf.hpp"
#include "clust.hpp"
#include "flat_clust.hpp"
extern "C" {

// which: 0 = IsValid(NmfOptions), 1 = IsValid(ClustOptions), 2 = IsValid(FlatClustOptions). v = {tol, algorithm, prog, height, width, k,
// min_iter, max_iter, tolcount, max_threads, maxterms, unbalanced, trial_allowance, num_clusters}.
int ref_is_valid(int which, const double* v, int validate_matrix)
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

} // extern "C"

// ---- flat clustering result files: common/src/flat_clust_output.cpp:52-170 (+ flatclust_{xml,json}_writer.cpp, assignments.cpp)
#include "flat_clust_output.hpp"
extern "C" {

// format: 1 = XML, 2 = JSON (FileFormat, file_format.hpp:17-23). dictionary: `m` NUL-terminated strings back to back.
int ref_flatclust_write_results(const char* outdir, const unsigned int* assignments, const float* probabilities, const char* dictionary,
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

} // extern "C"

// ---- the tree and its writers, driven by a script (no factorization): Tree<T> hierclust/include/tree.hpp + src/tree.cpp,
//      HierclustXmlWriter / HierclustJsonWriter hierclust/src/hierclust_{xml,json}_writer.cpp, CreateHierclustWriter
#include "tree.hpp"
#include "hierclust_writer_factory.hpp"
extern "C" {

// Grows a tree of num_clusters leaves from scripted factors (m x 2 topic matrices, 2 x docs membership matrices) and scripted
// priorities, always splitting the leaf MinMaxLeafPriorities names; writes assignments and the tree. format: 1 XML, 2 JSON.
int ref_tree_script(int seed, int m, int n, int num_clusters, int maxterms, int format, const char* assign_path, const char* tree_path)
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
    DenseMatrix<double> W(m, 2), H(2, n);
    auto fill = [&](DenseMatrix<double>& M) { for (int c = 0; c < M.Width(); ++c) for (int r = 0; r < M.Height(); ++r) M.Set(r, c, next()); };
    fill(W); fill(H);
    tree.SplitRoot(&W, &H);
    for (int split = 0; ; ++split)
    {
        const unsigned int i0 = tree.LeftChildIndex(), i1 = tree.RightChildIndex();
        doc_count[i0] = tree.LeftChildDocs().size(); doc_count[i1] = tree.RightChildDocs().size();
        tree.SetNodePriority(i0, doc_count[i0] > 3 ? next() + 0.01 : -1.0);
        tree.SetNodePriority(i1, doc_count[i1] > 3 ? next() + 0.01 : -1.0);
        if (split == num_clusters - 2) break;
        double mn, mx; unsigned int idx;
        tree.MinMaxLeafPriorities(mn, mx, idx);
        if (mx < 0.0) break;
        DenseMatrix<double> Hs(2, doc_count[idx]);
        fill(W); fill(Hs);
        tree.Split(idx, &W, &Hs);
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

} // extern "C"

// ---- dictionary files: LoadStringsFromFile common/src/utils.cpp:220-239
#include <cstring>
#include "utils.hpp"
extern "C" {

int ref_load_strings(const char* path, char* out, unsigned int cap, int* count)
{
    std::vector<std::string> v(1, "preexisting");              // the reader appends
    if (!LoadStringsFromFile(std::string(path), v)) return -1;
    std::string joined;
    for (const auto& t : v) { joined += t; joined += '\n'; }
    if (joined.size() + 1 > cap) return -2;
    std::memcpy(out, joined.c_str(), joined.size() + 1);
    *count = static_cast<int>(v.size());
    return 0;
}

} // extern "C"

// ---- tf-idf preprocessing: preprocess_tf preprocessor/src/preprocess.cpp:81-250 on a TermFrequencyMatrix
//      (common/src/term_frequency_matrix.cpp) built from CSC arrays of term counts
#include "term_frequency_matrix.hpp"
#include "preprocess.hpp"
extern "C" {

// In: m x n CSC of term counts. Out (capacities = the input sizes): the pruned matrix as CSC with row indices and counts,
// its tf-idf scores aligned with them, and for every surviving row / column its original index.
// Returns 0, or -1 if every column was pruned.
int ref_preprocess_tf(unsigned int m, unsigned int n, unsigned int nz, const unsigned int* col_offsets, const unsigned int* row_indices,
                      const double* counts, unsigned int max_iter, unsigned int docs_per_term, unsigned int terms_per_doc,
                      unsigned int* out_m, unsigned int* out_n, unsigned int* out_nz, unsigned int* out_cols, unsigned int* out_rows,
                      unsigned int* out_counts, double* out_scores, unsigned int* term_indices, unsigned int* doc_indices)
{
    SparseMatrix<double> S(m, n, nz, col_offsets, row_indices, counts);
    TermFrequencyMatrix M(S, false);
    std::vector<unsigned int> ti(m), di(n);
    std::vector<double> scores;
    const bool ok = preprocess_tf(M, ti, di, scores, max_iter, docs_per_term, terms_per_doc);
    if (!ok) return -1;
    *out_m = M.Height(); *out_n = M.Width(); *out_nz = M.Size();
    const unsigned int* cp = M.LockedColBuffer();
    const TFData* tf = M.LockedTFDataBuffer();
    for (unsigned int c = 0; c <= M.Width(); ++c) out_cols[c] = cp[c];
    for (unsigned int e = 0; e < M.Size(); ++e) { out_rows[e] = tf[e].row; out_counts[e] = tf[e].count; out_scores[e] = scores[e]; }
    for (unsigned int r = 0; r < M.Height(); ++r) term_indices[r] = ti[r];
    for (unsigned int c = 0; c < M.Width(); ++c) doc_indices[c] = di[c];
    return 0;
}

} // extern "C"
