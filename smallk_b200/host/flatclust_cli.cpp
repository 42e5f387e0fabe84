// flatclust — command-line front end with the reference's option set (flatclust/src/command_line.cpp:36-58,
// defaults :166-200; flow flatclust/src/main.cpp:46-294) on the GPU library: FlatClust / FlatClustSparse.
// One addition: --seed <int> fixes the random initialisers (the reference always seeds from the clock).
#include <getopt.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "flat_clust.hpp"
#include "flat_clust_output.hpp"
#include "matrix_io.hpp"
#include "random.hpp"

namespace {
struct CommandLineOptions
{
    FlatClustOptions clust_opts;
    std::string infile_A, infile_W, infile_H, dictfile, outdir, clustfile, assignfile, fuzzyfile;
    bool show_help = false;
    FileFormat format = FileFormat::XML;
    int seed = -1;
};

option longopts[] = {
    {"matrixfile", required_argument, NULL, 'a'}, {"dictfile", required_argument, NULL, 'b'},
    {"clusters", required_argument, NULL, 'c'},   {"tol", required_argument, NULL, 'd'},
    {"outdir", required_argument, NULL, 'e'},     {"miniter", required_argument, NULL, 'f'},
    {"maxiter", required_argument, NULL, 'g'},    {"help", no_argument, NULL, 'h'},
    {"algorithm", required_argument, NULL, 'i'},  {"verbose", required_argument, NULL, 'k'},
    {"maxthreads", required_argument, NULL, 'l'}, {"maxterms", required_argument, NULL, 'm'},
    {"infile_W", required_argument, NULL, 'n'},   {"infile_H", required_argument, NULL, 'o'},
    {"clustfile", required_argument, NULL, 'q'},  {"assignfile", required_argument, NULL, 'r'},
    {"format", required_argument, NULL, 's'},     {"fuzzyfile", required_argument, NULL, 't'},
    {"seed", required_argument, NULL, 'u'},       {0, 0, 0, 0}};

bool DirectoryExists(const std::string& d) { struct stat st; return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

void ShowHelp(const char* prog)
{
    std::cout << "\nUsage: " << prog << "\n"
              << "        --matrixfile <filename>      Filename of the matrix to be factored.\n"
              << "                                     Either CSV format for dense or MatrixMarket format for sparse.\n"
              << "        --dictfile <filename>        The name of the dictionary file.\n"
              << "        --clusters <integer>         The number of clusters to generate.\n"
              << "        [--algorithm  BPP]           The NMF algorithm to use: HALS, RANK2 (clusters == 2), BPP\n"
              << "        [--infile_W  (empty)]        Dense matrix to initialize W, CSV file (m x clusters).\n"
              << "        [--infile_H  (empty)]        Dense matrix to initialize H, CSV file (clusters x n).\n"
              << "        [--tol  0.0001]              Tolerance value for the progress metric.\n"
              << "        [--outdir  (empty)]          Output directory.\n"
              << "        [--miniter  5]               Minimum number of iterations to perform.\n"
              << "        [--maxiter  5000]            Maximum number of iterations to perform.\n"
              << "        [--maxterms  5]              Number of terms per node.\n"
              << "        [--maxthreads  N]            Accepted for compatibility; the GPU path ignores it.\n"
              << "        [--verbose  1]               Whether to print updates to the screen. 1 == yes, 0 == no\n"
              << "        [--format  XML]              Format of the output file containing the clusters: XML or JSON\n"
              << "        [--clustfile clusters_N.ext] Name of the output file containing the clusters (relative to the outdir).\n"
              << "        [--assignfile assignments_N.csv]  Name of the file containing final assignments (relative to the outdir).\n"
              << "        [--fuzzyfile assignments_fuzzy_N.csv] Name of fuzzy assignment file (relative to the outdir).\n"
              << "        [--seed  (clock)]            Seed of the random initialisers.\n\n";
}

bool Parse(int argc, char* argv[], CommandLineOptions& o)
{
    NmfOptions& n = o.clust_opts.nmf_opts;
    n.height = n.width = n.k = 0;
    n.min_iter = 5; n.max_iter = 5000; n.tol = 0.0001; n.tolcount = 1;
    n.verbose = true; n.normalize = true; n.algorithm = NmfAlgorithm::BPP;
    n.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO; n.max_threads = 1;
    o.clust_opts.maxterms = 5; o.clust_opts.num_clusters = 0; o.clust_opts.verbose = true;
    int c, index;
    while (-1 != (c = getopt_long(argc, argv, ":a:b:c:d:e:f:g:hi:k:l:m:n:o:q:r:s:t:u:", longopts, &index)))
    {
        const std::string arg = optarg ? optarg : "";
        std::string up = arg;
        std::transform(up.begin(), up.end(), up.begin(), ::toupper);
        switch (c)
        {
        case 'a': o.infile_A = arg; break;
        case 'b': o.dictfile = arg; break;
        case 'c': o.clust_opts.num_clusters = n.k = std::atoi(optarg); break;
        case 'd': n.tol = std::atof(optarg); break;
        case 'e': o.outdir = arg; break;
        case 'f': n.min_iter = std::atoi(optarg); break;
        case 'g': n.max_iter = std::atoi(optarg); break;
        case 'h': o.show_help = true; break;
        case 'i':
            if (up == "HALS") n.algorithm = NmfAlgorithm::HALS;
            else if (up == "RANK2") n.algorithm = NmfAlgorithm::RANK2;
            else if (up == "BPP") n.algorithm = NmfAlgorithm::BPP;
            else { std::cerr << "Invalid value specified for command-line argument " << up << std::endl; return false; }
            break;
        case 'k': o.clust_opts.verbose = (0 != std::atoi(optarg)); n.verbose = o.clust_opts.verbose; break;
        case 'l': n.max_threads = std::max(1, std::atoi(optarg)); break;
        case 'm': o.clust_opts.maxterms = std::atoi(optarg); break;
        case 'n': o.infile_W = arg; break;
        case 'o': o.infile_H = arg; break;
        case 'q': o.clustfile = arg; break;
        case 'r': o.assignfile = arg; break;
        case 's':
            if (up == "XML") o.format = FileFormat::XML;
            else if (up == "JSON") o.format = FileFormat::JSON;
            else { std::cerr << "Invalid value specified for command-line argument " << up << std::endl; return false; }
            break;
        case 't': o.fuzzyfile = arg; break;
        case 'u': o.seed = std::atoi(optarg); break;
        case ':': std::cerr << "missing argument for option " << argv[optind - 1] << std::endl; return false;
        default: std::cerr << "invalid option: " << argv[optind - 1] << std::endl; return false;
        }
    }
    if (1 == argc) o.show_help = true;
    if (o.show_help) return false;
    if (o.infile_A.empty()) { std::cerr << "required command line argument --matrixfile not found" << std::endl; return false; }
    if (o.dictfile.empty()) { std::cerr << "required command line argument --dictfile not found" << std::endl; return false; }
    if (0 == o.clust_opts.num_clusters) { std::cerr << "required command line argument --clusters not found" << std::endl; return false; }
    const std::string dir = EnsureTrailingPathSep(o.outdir);
    std::ostringstream a, f, r;
    a << "assignments_" << o.clust_opts.num_clusters;
    f << "assignments_fuzzy_" << o.clust_opts.num_clusters;
    r << "clusters_" << o.clust_opts.num_clusters;
    o.assignfile = dir + (o.assignfile.empty() ? AppendExtension(a.str(), FileFormat::CSV) : o.assignfile);
    o.fuzzyfile = dir + (o.fuzzyfile.empty() ? AppendExtension(f.str(), FileFormat::CSV) : o.fuzzyfile);
    o.clustfile = dir + (o.clustfile.empty() ? AppendExtension(r.str(), o.format) : o.clustfile);
    return true;
}
} // namespace

int main(int argc, char* argv[])
{
    CommandLineOptions opts;
    if (!Parse(argc, argv, opts))
    {
        if (opts.show_help) { ShowHelp(argv[0]); return 0; }
        return -1;
    }
    if (!opts.outdir.empty() && !DirectoryExists(opts.outdir))
    { std::cerr << "the specified output directory \"" << opts.outdir << "\" does not exist" << std::endl; return -1; }
    {
        // flatclust/src/command_line.cpp:386-441: the tool's own checks (its own texts, not FlatClustOptions' IsValid)
        const NmfOptions& no = opts.clust_opts.nmf_opts;
        const char* complaint = nullptr;
        if (opts.clust_opts.num_clusters <= 0) complaint = "value for --clusters must be a positive integer";
        else if (no.tol <= 0.0 || no.tol >= 1.0) complaint = "tolerance must be in the interval (0.0, 1.0)";
        else if (no.min_iter <= 0) complaint = "miniter must be a positive integer";
        else if (no.max_iter <= 0) complaint = "maxiter must be a positive integer";
        else if (opts.clust_opts.maxterms <= 0) complaint = "maxterms must be a positive integer";
        else if (NmfProgressAlgorithm::PG_RATIO != no.prog_est_algorithm && NmfProgressAlgorithm::DELTA_FNORM != no.prog_est_algorithm)
            complaint = "clustlib error: unknown stopping criterion ";
        if (complaint) { std::cerr << complaint << std::endl; return -1; }
    }
    Random rng;
    if (opts.seed >= 0) rng.SeedFromInt(opts.seed); else rng.SeedFromTime();
    try { NmfInitialize(argc, argv); }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; return -1; }

    const bool verbose = opts.clust_opts.verbose;
    if (verbose) std::cout << "loading dictionary..." << std::endl;
    std::vector<std::string> dictionary;
    if (!LoadStringsFromFile(opts.dictfile, dictionary))
    { std::cerr << "\ncould not load dictionary file " << opts.dictfile << std::endl; NmfFinalize(); return -1; }
    if (verbose) std::cout << "loading matrix..." << std::endl;
    const bool sparse = smallk_io::IsMatrixMarketFile(opts.infile_A);
    smallk_io::CscMatrix A;
    std::vector<double> buf_a;
    unsigned int m = 0, n = 0;
    bool ok = false;
    try { ok = sparse ? smallk_io::LoadMatrixMarketFile(opts.infile_A, A) : smallk_io::LoadDelimitedFile(buf_a, m, n, opts.infile_A); }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }      // index out of bounds, no entries
    if (!ok) { std::cerr << "\nload failed for file " << opts.infile_A << std::endl; NmfFinalize(); return -1; }
    if (sparse) { m = A.height; n = A.width; }
    if (dictionary.size() < m) { std::cerr << "\ndictionary has fewer terms than the matrix has rows" << std::endl; NmfFinalize(); return -1; }
    NmfOptions& no = opts.clust_opts.nmf_opts;
    no.height = m; no.width = n;
    const unsigned int k = no.k;

    std::vector<double> buf_w(static_cast<size_t>(m) * k), buf_h(static_cast<size_t>(k) * n);
    unsigned int hw = m, ww = k, hh = k, wh = n;
    if (verbose) std::cout << "Initializing matrix W..." << std::endl;
    if (opts.infile_W.empty()) RandomMatrix(buf_w.data(), m, m, k, rng, 0.5, 0.5);
    else if (!smallk_io::LoadDelimitedFile(buf_w, hw, ww, opts.infile_W))
    { std::cerr << "\nload failed for file " << opts.infile_W << std::endl; NmfFinalize(); return -1; }
    if (hw != m || ww != k)
    { std::cerr << "\tdimensions of matrix W are " << hw << " x " << ww << "\n\texpected " << m << " x " << k << std::endl; NmfFinalize(); return -1; }
    if (verbose) std::cout << "Initializing matrix H..." << std::endl;
    if (opts.infile_H.empty()) RandomMatrix(buf_h.data(), k, k, n, rng, 0.5, 0.5);
    else if (!smallk_io::LoadDelimitedFile(buf_h, hh, wh, opts.infile_H))
    { std::cerr << "\nload failed for file " << opts.infile_H << std::endl; NmfFinalize(); return -1; }
    if (hh != k || wh != n)
    { std::cerr << "\tdimensions of matrix H are " << hh << " x " << wh << "\n\texpected " << k << " x " << n << std::endl; NmfFinalize(); return -1; }

    NmfStats stats;
    Result result;
    try
    {
        if (sparse)
            result = FlatClustSparse(no, A.height, A.width, A.nnz(), A.col_offsets.data(), A.row_indices.data(), A.data.data(),
                                     buf_w.data(), m, buf_h.data(), k, stats);
        else
            result = FlatClust(no, buf_a.data(), m, buf_w.data(), m, buf_h.data(), k, stats);
    }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }
    if (Result::OK != result) std::cerr << "\nNMF solver failure." << std::endl;
    else
    {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<float> probabilities;
        std::vector<unsigned int> assignments(n);
        std::vector<int> term_indices(static_cast<size_t>(opts.clust_opts.maxterms) * k);
        ComputeFuzzyAssignments(probabilities, buf_h.data(), k, k, n);
        ComputeAssignments(assignments, buf_h.data(), k, k, n);
        TopTerms(opts.clust_opts.maxterms, buf_w.data(), m, m, k, term_indices);
        FlatClustWriteResults(opts.assignfile, opts.fuzzyfile, opts.clustfile, assignments, probabilities, dictionary, term_indices,
                              opts.format, opts.clust_opts.maxterms, n, opts.clust_opts.num_clusters);
        const double post_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (verbose)
            std::cout << "Iterations: " << stats.iteration_count << "\nElapsed wall clock time: " << (stats.elapsed_us / 1000.0 + post_ms) << " ms.\n" << std::endl;
    }
    NmfFinalize();
    return Result::OK == result ? 0 : -1;
}
