// smallk_b200 — namespace smallk over host/nmf.hpp. Follows smallk/src/smallk.cpp of the reference:
// global state (:46-68), setters with clamping (:383-468), LoadMatrix (:205-330), Nmf (:471-650).
#include "smallk.hpp"

#include <algorithm>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>
#include <stdexcept>
#include <thread>

#include "clust.hpp"
#include "flat_clust.hpp"
#include "flat_clust_output.hpp"
#include "matrix_io.hpp"
#include "nmf.hpp"

namespace {
bool matrix_loaded = false, is_sparse = false;
unsigned int m = 0, n = 0, k = 0, ldim_a = 0;
std::vector<double> buf_a, buf_w, buf_h;
smallk_io::CscMatrix A;
unsigned int outprecision = 6, max_iter = 5000, min_iter = 5;
unsigned int max_threads = std::max(2u, std::thread::hardware_concurrency());
double nmf_tolerance = 0.005;
std::string outdir;
Random rng;                 // one stream for Nmf and HierNmf2, as in the reference (smallk.cpp:47)
bool dict_loaded = false;
std::vector<std::string> dictionary;
unsigned int maxterms = 5;
double hier_nmf2_tolerance = 0.0001;
smallk::OutputFormat clustfile_format = smallk::JSON;
const char* DEFAULT_FILENAME_W = "w.csv";
const char* DEFAULT_FILENAME_H = "h.csv";

// RandomMatrixSequential with center 0.5, radius 0.5 (common/include/matrix_generator.hpp:229-248,
// random.hpp:154-158). The reference switches to a thread-count-dependent parallel generator for large
// matrices; this build always uses the sequential stream.
void RandomMatrix(double* buf, unsigned int ldim, unsigned int height, unsigned int width)
{
    ::RandomMatrix(buf, ldim, height, width, rng, 0.5, 0.5);
}

std::string EnsureTrailingSep(const std::string& s)
{
    if (s.empty() || s.back() == '/') return s;
    return s + "/";
}
} // namespace

namespace smallk
{
// smallk.cpp:114-119: defaults, a clock-seeded generator (SeedRNG overrides it), then the library
void Initialize(int& argc, char**& argv) { Reset(); rng.SeedFromTime(); NmfInitialize(argc, argv); }
bool IsInitialized() { return Result::INITIALIZED == NmfIsInitialized(); }
void Finalize() { NmfFinalize(); }

unsigned int GetMajorVersion() { return SMALLK_MAJOR_VERSION; }
unsigned int GetMinorVersion() { return SMALLK_MINOR_VERSION; }
unsigned int GetPatchLevel() { return SMALLK_PATCH_LEVEL; }
std::string GetVersionString()
{
    std::ostringstream v;
    v << SMALLK_MAJOR_VERSION << "." << SMALLK_MINOR_VERSION << "." << SMALLK_PATCH_LEVEL;
    return v.str();
}

unsigned int GetOutputPrecision() { return outprecision; }
void SetOutputPrecision(const unsigned int num_digits)
{
    outprecision = num_digits;
    if (0 == outprecision) outprecision = 1;
    if (outprecision > static_cast<unsigned int>(std::numeric_limits<double>::max_digits10))
        outprecision = std::numeric_limits<double>::max_digits10;
}
double GetNmfTolerance() { return nmf_tolerance; }
void SetNmfTolerance(const double tol)
{
    if (tol <= 0.0 || tol >= 1.0) throw std::logic_error("smallk error (SetNmfTolerance): require 0.0 < tol < 1.0");
    nmf_tolerance = tol;
}
unsigned int GetMaxIter() { return max_iter; }
void SetMaxIter(const unsigned int v) { max_iter = v ? v : 1u; }
unsigned int GetMinIter() { return min_iter; }
void SetMinIter(const unsigned int v) { min_iter = v ? v : 1u; }
unsigned int GetMaxThreads() { return max_threads; }
void SetMaxThreads(const unsigned int v)
{
    // smallk.cpp:427-433: capped at the hardware thread count, at least 1 (the GPU path does not use it)
    max_threads = std::max(1u, std::min(v, std::thread::hardware_concurrency()));
}
void Reset()
{
    matrix_loaded = false; is_sparse = false;
    m = n = k = ldim_a = 0;
    buf_a.clear(); buf_w.clear(); buf_h.clear();
    A = smallk_io::CscMatrix();
    outprecision = 6; max_iter = 5000; min_iter = 5; nmf_tolerance = 0.005;
    max_threads = std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 2u;     // smallk.cpp:88-92
    outdir.clear();
    dict_loaded = false; dictionary.clear();
    maxterms = 5; hier_nmf2_tolerance = 0.0001; clustfile_format = JSON;
    // the random generator is left alone, as in the reference (smallk.cpp:81-111): Reset() does not undo SeedRNG()
}
void SeedRNG(const int seed) { rng.SeedFromInt(seed); }

// smallk.cpp:652-735
void LoadDictionary(const std::string& filepath)
{
    dictionary.clear();
    dict_loaded = false;
    if (!LoadStringsFromFile(filepath, dictionary))
        throw std::runtime_error("smallk error (LoadDictionary): load failed for file " + filepath);
    dict_loaded = true;
}
void LoadDictionary(const std::vector<std::string>& terms)
{
    dictionary.assign(terms.begin(), terms.end());
    dict_loaded = true;
}
unsigned int GetMaxTerms() { return maxterms; }
void SetMaxTerms(const unsigned int max_terms) { maxterms = max_terms ? max_terms : 1; }
OutputFormat GetOutputFormat() { return clustfile_format; }
void SetOutputFormat(const OutputFormat format) { clustfile_format = format; }
double GetHierNmf2Tolerance() { return hier_nmf2_tolerance; }
void SetHierNmf2Tolerance(const double tol)
{
    if (tol <= 0.0 || tol >= 1.0) throw std::logic_error("smallk error (SetHierNmf2Tolerance): require 0.0 < tol < 1.0");
    hier_nmf2_tolerance = tol;
}

void LoadMatrix(const std::string& filepath)
{
    // smallk.cpp:168-206: same checks, exception types and texts
    if (filepath.empty()) throw std::runtime_error("smallk error (LoadMatrix): matrix filename is invalid.");
    matrix_loaded = false;
    bool ok;
    if (smallk_io::IsMatrixMarketFile(filepath))
    {
        ok = smallk_io::LoadMatrixMarketFile(filepath, A);
        if (ok) { m = A.height; n = A.width; is_sparse = true; }
    }
    else
    {
        ok = smallk_io::LoadDelimitedFile(buf_a, m, n, filepath);
        if (ok) { ldim_a = m; is_sparse = false; }
    }
    if (!ok) throw std::runtime_error("smallk error (LoadMatrix): load failed for file " + filepath);
    matrix_loaded = true;
}

void LoadMatrix(const double* buffer, const unsigned int ldim, const unsigned int height, const unsigned int width)
{
    // smallk.cpp:209-263 (its messages say "LoadSparseMatrixFromBuffer" for the dense overload too)
    matrix_loaded = false;
    if (0 == height) throw std::runtime_error("smallk error (LoadSparseMatrixFromBuffer): invalid height input.");
    if (0 == width) throw std::runtime_error("smallk error (LoadSparseMatrixFromBuffer): invalid width input.");
    if (nullptr == buffer) throw std::runtime_error("smallk error (LoadSparseMatrixFromBuffer): empty data pointer.");
    if (ldim < height) throw std::logic_error("smallk error (LoadMatrix): leading dimension too small");      // not checked by the reference
    // The reference swaps its loop bounds here (smallk.cpp:249-255) and is only right for square inputs;
    // this copies every column of the height x width matrix.
    m = height; n = width; ldim_a = height;
    buf_a.resize(static_cast<size_t>(m) * n);
    for (unsigned int c = 0; c < n; ++c)
        for (unsigned int r = 0; r < m; ++r) buf_a[static_cast<size_t>(c) * m + r] = buffer[static_cast<size_t>(c) * ldim + r];
    is_sparse = false;
    matrix_loaded = true;
}

void LoadMatrix(const unsigned int height, const unsigned int width, const unsigned int nz,
                const std::vector<double>& data, const std::vector<unsigned int>& row_indices,
                const std::vector<unsigned int>& col_offsets)
{
    // smallk.cpp:266-333: the reference's checks in the reference's order
    matrix_loaded = false;
    const char* who = "smallk error (LoadSparseMatrixFromBuffer): ";
    if (row_indices.size() != data.size()) throw std::runtime_error(std::string(who) + "invalid input vectors.");
    if (0 == height) throw std::runtime_error(std::string(who) + "invalid height input.");
    if (0 == width) throw std::runtime_error(std::string(who) + "invalid width input.");
    if (data.size() > static_cast<size_t>(height) * width) throw std::runtime_error(std::string(who) + "invalid inputs.");
    if (data.empty()) throw std::runtime_error(std::string(who) + "empty data vector.");
    if (row_indices.empty()) throw std::runtime_error(std::string(who) + "empty row_indices vector.");
    if (col_offsets.empty()) throw std::runtime_error(std::string(who) + "empty col_offsets vector.");
    // not checked by the reference (it would read past the vectors): the arrays must hold what nz and width promise
    if (data.size() < nz || col_offsets.size() < static_cast<size_t>(width) + 1)
        throw std::logic_error("smallk error (LoadMatrix): inconsistent sparse arrays");
    A.height = height; A.width = width;
    A.data.assign(data.begin(), data.begin() + nz);
    A.row_indices.assign(row_indices.begin(), row_indices.begin() + nz);
    A.col_offsets.assign(col_offsets.begin(), col_offsets.begin() + width + 1);
    m = height; n = width;
    is_sparse = true;
    matrix_loaded = true;
}

bool IsMatrixLoaded() { return matrix_loaded; }
std::string GetOutputDir() { return outdir; }
void SetOutputDir(const std::string& output_dir)
{
    // smallk.cpp:346-380: a relative path is taken from the current directory; the directory must exist
    std::string full_path;
    if (!output_dir.empty() && '/' != output_dir[0])
    {
        char cwd[4096];
        if (nullptr == getcwd(cwd, sizeof(cwd))) throw std::runtime_error("smallk error (SetOutputDir): could not determine current directory.");
        full_path = EnsureTrailingSep(cwd);
    }
    full_path += output_dir;
    struct stat st;
    if (0 != stat(full_path.c_str(), &st) || !S_ISDIR(st.st_mode))
        throw std::logic_error("smallk error (SetOutputDir): the directory \"" + full_path + "\" does not exist.");
    outdir = EnsureTrailingSep(full_path);
}

void Nmf(const unsigned int kval, const Algorithm algorithm, const std::string& csv_file_w, const std::string& csv_file_h)
{
    using std::cout; using std::cerr; using std::endl;
    if (!matrix_loaded) throw std::logic_error("smallk error (NMF): no matrix has been loaded.");
    if (max_iter < min_iter) throw std::logic_error("smallk error (NMF): min_iterations exceeds max_iterations.");
    if (0 == kval) throw std::logic_error("smallk error (NMF): k must be greater than 0.");
    const unsigned long long lim = std::numeric_limits<int>::max();
    if (static_cast<unsigned long long>(m) * kval > lim) throw std::logic_error("smallk error (Nmf): mxk matrix W is too large.");
    if (static_cast<unsigned long long>(kval) * n > lim) throw std::logic_error("smallk error (Nmf): kxn matrix H is too large.");

    NmfOptions nmf_opts;
    k = kval;
    switch (algorithm)
    {
    case Algorithm::MU: nmf_opts.algorithm = NmfAlgorithm::MU; break;
    case Algorithm::HALS: nmf_opts.algorithm = NmfAlgorithm::HALS; break;
    case Algorithm::RANK2: nmf_opts.algorithm = NmfAlgorithm::RANK2; break;
    case Algorithm::BPP: nmf_opts.algorithm = NmfAlgorithm::BPP; break;
    default: throw std::logic_error("smallk error (NMF): unknown NMF algorithm.");
    }
    if (NmfAlgorithm::RANK2 == nmf_opts.algorithm) k = 2;
    const unsigned int ldim_w = m, ldim_h = k;
    if (buf_w.size() < static_cast<size_t>(m) * k) buf_w.resize(static_cast<size_t>(m) * k);
    if (buf_h.size() < static_cast<size_t>(k) * n) buf_h.resize(static_cast<size_t>(k) * n);

    bool ok = true;
    unsigned int height_w = m, width_w = k, height_h = k, width_h = n;
    cout << "Initializing matrix W..." << endl;
    if (csv_file_w.empty()) RandomMatrix(&buf_w[0], ldim_w, m, k);
    else ok = smallk_io::LoadDelimitedFile(buf_w, height_w, width_w, csv_file_w);
    if (!ok) throw std::runtime_error("smallk error (Nmf): load failed for file \"" + csv_file_w + "\"");
    if (height_w != m || width_w != k)
    {
        cerr << "\tdimensions of matrix W are " << height_w << " x " << width_w << endl;
        cerr << "\texpected " << m << " x " << k << endl;
        throw std::logic_error("smallk error (Nmf): non-conformant matrix W.");
    }
    cout << "Initializing matrix H..." << endl;
    if (csv_file_h.empty()) RandomMatrix(&buf_h[0], ldim_h, k, n);
    else ok = smallk_io::LoadDelimitedFile(buf_h, height_h, width_h, csv_file_h);
    if (!ok) throw std::runtime_error("smallk error (Nmf): load failed for file \"" + csv_file_h + "\"");
    if (height_h != k || width_h != n)
    {
        cerr << "\tdimensions of matrix H are " << height_h << " x " << width_h << endl;
        cerr << "\texpected " << k << " x " << n << endl;
        throw std::logic_error("smallk error (Nmf): non-conformant matrix H.");
    }

    // smallk.cpp:581-595
    nmf_opts.prog_est_algorithm = (NmfAlgorithm::MU == nmf_opts.algorithm) ? NmfProgressAlgorithm::DELTA_FNORM
                                                                            : NmfProgressAlgorithm::PG_RATIO;
    nmf_opts.tol = nmf_tolerance;
    nmf_opts.height = m; nmf_opts.width = n; nmf_opts.k = k;
    nmf_opts.min_iter = min_iter; nmf_opts.max_iter = max_iter;
    nmf_opts.tolcount = 1;
    nmf_opts.max_threads = max_threads;
    nmf_opts.verbose = true;
    nmf_opts.normalize = true;

    NmfStats stats;
    Result result;
    if (is_sparse)
        result = NmfSparse(nmf_opts, A.height, A.width, A.nnz(), A.col_offsets.data(), A.row_indices.data(), A.data.data(),
                           &buf_w[0], ldim_w, &buf_h[0], ldim_h, stats);
    else
        result = ::Nmf(nmf_opts, &buf_a[0], ldim_a, &buf_w[0], ldim_w, &buf_h[0], ldim_h, stats);
    cout << "Elapsed wall clock time: " << stats.elapsed_us / 1000.0 << " ms." << endl << endl;
    if (Result::OK != result) throw std::runtime_error("smallk error (Nmf): NMF solver failure.");

    const std::string outfile_w = outdir + DEFAULT_FILENAME_W, outfile_h = outdir + DEFAULT_FILENAME_H;
    cout << "Writing output files..." << endl;
    if (!smallk_io::WriteDelimitedFile(&buf_w[0], ldim_w, m, k, outfile_w, outprecision))
        throw std::runtime_error("smallk error (Nmf): could not write W result.");
    if (!smallk_io::WriteDelimitedFile(&buf_h[0], ldim_h, k, n, outfile_h, outprecision))
        throw std::runtime_error("smallk error (Nmf): could not write H result.");
}

const double* LockedBufferW(unsigned int& ldim, unsigned int& height, unsigned int& width)
{
    ldim = m; height = m; width = k;
    return buf_w.empty() ? nullptr : &buf_w[0];
}
const double* LockedBufferH(unsigned int& ldim, unsigned int& height, unsigned int& width)
{
    ldim = k; height = k; width = n;
    return buf_h.empty() ? nullptr : &buf_h[0];
}

// smallk.cpp:737-856: HierNMF2 on the loaded matrix; writes assignments_N.csv and tree_N.{xml,json} (and, with the
// flat step, assignments_flat_N.csv, assignments_fuzzy_N.csv, clusters_N.*) into the output directory.
static void HierNmf2Internal(const bool generate_flat, const unsigned int num_clusters)
{
    using std::cout; using std::cerr; using std::endl;
    if (!matrix_loaded) throw std::logic_error("smallk error (HierNmf2): no matrix has been loaded.");
    if (!dict_loaded) throw std::logic_error("smallk error (HierNmf2): no dictionary has been loaded.");
    if (0 == num_clusters) throw std::logic_error("smallk error (HierNmf2): num_clusters must be greater than 0.");
    const unsigned long long lim = std::numeric_limits<int>::max();
    if (2ull * m > lim) throw std::logic_error("smallk error (HierNmf2): matrix height too large.");
    if (2ull * n > lim) throw std::logic_error("smallk error (HierNmf2): matrix width too large.");
    if (dictionary.size() < m) throw std::logic_error("smallk error (HierNmf2): dictionary has fewer terms than the matrix has rows.");

    ClustOptions o;
    o.nmf_opts.tol = hier_nmf2_tolerance;
    o.nmf_opts.algorithm = NmfAlgorithm::RANK2;
    o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.nmf_opts.height = m; o.nmf_opts.width = n; o.nmf_opts.k = 2;
    o.nmf_opts.min_iter = min_iter; o.nmf_opts.max_iter = max_iter; o.nmf_opts.tolcount = 1;
    o.nmf_opts.max_threads = max_threads; o.nmf_opts.verbose = false;
    o.nmf_opts.normalize = true;                 // smallk.cpp:766 (the hierclust CLI passes false)
    o.maxterms = maxterms; o.unbalanced = 0.1; o.trial_allowance = 3;
    o.num_clusters = num_clusters; o.verbose = true; o.flat = generate_flat;

    const FileFormat format = (XML == clustfile_format) ? FileFormat::XML : FileFormat::JSON;
    std::ostringstream an, tn;
    an << "assignments_" << num_clusters;
    tn << "tree_" << num_clusters;
    const std::string assignfile = outdir + AppendExtension(an.str(), FileFormat::CSV);
    const std::string treefile = outdir + AppendExtension(tn.str(), format);

    Tree<R> tree;
    ClustStats stats;
    std::vector<R> flat_w(static_cast<size_t>(m) * num_clusters), flat_h(static_cast<size_t>(num_clusters) * n);
    Result result;
    if (is_sparse)
    {
        SparseMatrix<R> S(A.height, A.width, A.nnz(), A.col_offsets.data(), A.row_indices.data(), A.data.data());
        result = ClustSparse(o, S, flat_w.data(), flat_h.data(), tree, stats, rng);
    }
    else result = Clust(o, &buf_a[0], ldim_a, flat_w.data(), flat_h.data(), tree, stats, rng);
    if (Result::OK != result) throw std::runtime_error("smallk error (HierNMF2): HierNMF2 fatal error.");
    cout << (stats.nmf_count - stats.max_count) << "/" << stats.nmf_count << " factorizations converged." << endl << endl;
    cout << "Writing output files..." << endl;
    if (!tree.WriteAssignments(assignfile)) cerr << "\terror writing assignments file" << endl;
    IHierclustWriter* writer = CreateHierclustWriter(format);
    if (!tree.WriteTree(writer, treefile, dictionary)) cerr << "\terror writing hierarchical results file" << endl;
    delete writer;
    if (generate_flat)
    {
        std::vector<float> probabilities;
        std::vector<unsigned int> assignments_flat;
        std::vector<int> term_indices(static_cast<size_t>(maxterms) * num_clusters);
        ComputeFuzzyAssignments(probabilities, flat_h.data(), num_clusters, num_clusters, n);
        ComputeAssignments(assignments_flat, flat_h.data(), num_clusters, num_clusters, n);
        TopTerms(static_cast<int>(maxterms), flat_w.data(), m, m, num_clusters, term_indices);
        FlatClustWriteResults(outdir, assignments_flat, probabilities, dictionary, term_indices, format, maxterms, n, num_clusters);
    }
}

void HierNmf2(const unsigned int num_clusters) { HierNmf2Internal(false, num_clusters); }
void HierNmf2WithFlat(const unsigned int num_clusters) { HierNmf2Internal(true, num_clusters); }
} // namespace smallk
