// hierclust — command-line front end with the reference's option set (hierclust/src/command_line.cpp:37-58,
// defaults :139-172; flow hierclust/src/main.cpp:47-264) on the GPU library: Clust / ClustSparse of host/clust.hpp.
// One addition: --seed <int> fixes the random initialisers (the reference always seeds from the clock).
#include <getopt.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "clust.hpp"
#include "flat_clust.hpp"
#include "flat_clust_output.hpp"
#include "matrix_io.hpp"

namespace {
struct CommandLineOptions
{
    ClustOptions clust_opts;
    std::string infile_A, dictfile, outdir, treefile, assignfile;
    bool show_help = false;
    FileFormat format = FileFormat::XML;
    int seed = -1;
};

option longopts[] = {
    {"matrixfile", required_argument, NULL, 'a'}, {"dictfile", required_argument, NULL, 'b'},
    {"clusters", required_argument, NULL, 'c'},   {"tol", required_argument, NULL, 'd'},
    {"outdir", required_argument, NULL, 'e'},     {"miniter", required_argument, NULL, 'f'},
    {"maxiter", required_argument, NULL, 'g'},    {"help", no_argument, NULL, 'h'},
    {"trial_allowance", required_argument, NULL, 'i'}, {"unbalanced", required_argument, NULL, 'j'},
    {"verbose", required_argument, NULL, 'k'},    {"maxthreads", required_argument, NULL, 'l'},
    {"maxterms", required_argument, NULL, 'm'},   {"initdir", required_argument, NULL, 'n'},
    {"treefile", required_argument, NULL, 'q'},   {"assignfile", required_argument, NULL, 'r'},
    {"flat", required_argument, NULL, 's'},       {"format", required_argument, NULL, 't'},
    {"seed", required_argument, NULL, 'u'},       {0, 0, 0, 0}};

bool DirectoryExists(const std::string& d) { struct stat st; return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

void ShowHelp(const char* prog)
{
    std::cout << "\nUsage: " << prog << "\n"
              << "        --matrixfile <filename>     Filename of the matrix to be factored.\n"
              << "                                    Either CSV format for dense or MatrixMarket format for sparse.\n"
              << "        --dictfile <filename>       The name of the dictionary file.\n"
              << "        --clusters <integer>        The number of clusters to generate.\n"
              << "        [--initdir  (empty)]        Directory of initializers for all Rank2 factorizations.\n"
              << "                                    If unspecified, random init will be used.\n"
              << "        [--tol  0.0001]             Tolerance value for each factorization.\n"
              << "        [--outdir  (empty)]         Output directory.  If unspecified, results will be\n"
              << "                                    written to the current directory.\n"
              << "        [--miniter  5]              Minimum number of iterations to perform.\n"
              << "        [--maxiter  5000]           Maximum number of  iterations to perform.\n"
              << "        [--maxterms  5]             Number of terms per node.\n"
              << "        [--maxthreads  N]           Accepted for compatibility; the GPU path ignores it.\n"
              << "        [--unbalanced  0.1]         Threshold for determining leaf node imbalance.\n"
              << "        [--trial_allowance  3]      Number of split attempts.\n"
              << "        [--flat  0]                 Whether to generate a flat clustering result. 1 == yes, 0 == no\n"
              << "        [--verbose  1]              Whether to print updates to the screen. 1 == yes, 0 == no\n"
              << "        [--format  XML]             Format of the output file containing the tree: XML or JSON\n"
              << "        [--treefile  tree_N.ext]    Name of the output file containing the tree (relative to the outdir).\n"
              << "        [--assignfile assignments_N.csv]  Name of the file containing final assignments (relative to the outdir).\n"
              << "        [--seed  (clock)]           Seed of the random initialisers.\n\n";
}

bool Parse(int argc, char* argv[], CommandLineOptions& o)
{
    NmfOptions& n = o.clust_opts.nmf_opts;
    n.height = n.width = n.k = 0;
    n.min_iter = 5; n.max_iter = 5000; n.tol = 0.0001; n.tolcount = 1;
    n.verbose = false; n.normalize = false; n.algorithm = NmfAlgorithm::RANK2;
    n.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO; n.max_threads = 1;
    o.clust_opts.maxterms = 5; o.clust_opts.trial_allowance = 3; o.clust_opts.unbalanced = 0.1;
    o.clust_opts.num_clusters = 0; o.clust_opts.verbose = true; o.clust_opts.flat = false;
    int c, index;
    while (-1 != (c = getopt_long(argc, argv, ":a:b:c:d:e:f:g:hi:j:k:l:m:n:q:r:s:t:u:", longopts, &index)))
    {
        const std::string arg = optarg ? optarg : "";
        switch (c)
        {
        case 'a': o.infile_A = arg; break;
        case 'b': o.dictfile = arg; break;
        case 'c': o.clust_opts.num_clusters = std::atoi(optarg); break;
        case 'd': n.tol = std::atof(optarg); break;
        case 'e': o.outdir = arg; break;
        case 'f': n.min_iter = std::atoi(optarg); break;
        case 'g': n.max_iter = std::atoi(optarg); break;
        case 'h': o.show_help = true; break;
        case 'i': o.clust_opts.trial_allowance = std::atoi(optarg); break;
        case 'j': o.clust_opts.unbalanced = std::atof(optarg); break;
        case 'k': o.clust_opts.verbose = (0 != std::atoi(optarg)); break;
        case 'l': n.max_threads = std::max(1, std::atoi(optarg)); break;
        case 'm': o.clust_opts.maxterms = std::atoi(optarg); break;
        case 'n': o.clust_opts.initdir = arg; break;
        case 'q': o.treefile = arg; break;
        case 'r': o.assignfile = arg; break;
        case 's': o.clust_opts.flat = (0 != std::atoi(optarg)); break;
        case 't':
        {
            std::string up = arg;
            std::transform(up.begin(), up.end(), up.begin(), ::toupper);
            if (up == "XML") o.format = FileFormat::XML;
            else if (up == "JSON") o.format = FileFormat::JSON;
            else { std::cerr << "Invalid value specified for command-line argument " << up << std::endl; return false; }
            break;
        }
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
    if (!o.clust_opts.initdir.empty()) o.clust_opts.initdir = EnsureTrailingPathSep(o.clust_opts.initdir);
    const std::string dir = EnsureTrailingPathSep(o.outdir);
    std::ostringstream a, t;
    a << "assignments_" << o.clust_opts.num_clusters;
    t << "tree_" << o.clust_opts.num_clusters;
    o.assignfile = dir + (o.assignfile.empty() ? AppendExtension(a.str(), FileFormat::CSV) : o.assignfile);
    o.treefile = dir + (o.treefile.empty() ? AppendExtension(t.str(), o.format) : o.treefile);
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
    if (!opts.clust_opts.initdir.empty() && !DirectoryExists(opts.clust_opts.initdir))
    { std::cerr << "the specified init directory \"" << opts.clust_opts.initdir << "\" does not exist" << std::endl; return -1; }
    if (!IsValid(opts.clust_opts, false)) return -1;

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
    smallk_io::CscMatrix csc;
    std::vector<R> buf_a;
    unsigned int m = 0, n = 0;
    bool ok = false;
    try { ok = sparse ? smallk_io::LoadMatrixMarketFile(opts.infile_A, csc) : smallk_io::LoadDelimitedFile(buf_a, m, n, opts.infile_A); }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }      // index out of bounds, no entries
    if (!ok) { std::cerr << "\nload failed for file " << opts.infile_A << std::endl; NmfFinalize(); return -1; }
    if (sparse) { m = csc.height; n = csc.width; }
    if (dictionary.size() < m) { std::cerr << "\ndictionary has fewer terms than the matrix has rows" << std::endl; NmfFinalize(); return -1; }
    opts.clust_opts.nmf_opts.height = m; opts.clust_opts.nmf_opts.width = n; opts.clust_opts.nmf_opts.k = 2;
    const unsigned int num_clusters = opts.clust_opts.num_clusters;
    if (verbose)
        std::cout << "\n     Command line options: \n\n\t            height: " << m << "\n\t             width: " << n
                  << "\n\t        matrixfile: " << opts.infile_A << "\n\t          clusters: " << num_clusters
                  << "\n\t               tol: " << opts.clust_opts.nmf_opts.tol << "\n\t          maxterms: " << opts.clust_opts.maxterms
                  << "\n\t        unbalanced: " << opts.clust_opts.unbalanced << "\n\t   trial_allowance: " << opts.clust_opts.trial_allowance
                  << "\n\t              flat: " << opts.clust_opts.flat << "\n" << std::endl;

    std::vector<R> buf_w(static_cast<size_t>(m) * num_clusters), buf_h(static_cast<size_t>(num_clusters) * n);
    Tree<R> tree;
    ClustStats stats;
    std::vector<float> probabilities;
    std::vector<unsigned int> assignments_flat;
    std::vector<int> term_indices(static_cast<size_t>(opts.clust_opts.maxterms) * num_clusters);
    Result result;
    const auto t0 = std::chrono::steady_clock::now();
    try
    {
        if (sparse)
        {
            SparseMatrix<R> A(csc.height, csc.width, csc.nnz(), csc.col_offsets.data(), csc.row_indices.data(), csc.data.data());
            result = ClustSparse(opts.clust_opts, A, buf_w.data(), buf_h.data(), tree, stats, rng);
        }
        else result = Clust(opts.clust_opts, buf_a.data(), m, buf_w.data(), buf_h.data(), tree, stats, rng);
        if (opts.clust_opts.flat && Result::OK == result)
        {
            ComputeFuzzyAssignments(probabilities, buf_h.data(), num_clusters, num_clusters, n);
            ComputeAssignments(assignments_flat, buf_h.data(), num_clusters, num_clusters, n);
            TopTerms(opts.clust_opts.maxterms, buf_w.data(), m, m, num_clusters, term_indices);
        }
    }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }
    const double elapsed = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "\nElapsed wall clock time: ";
    if (elapsed < 1000.0) std::cout << elapsed << " ms." << std::endl; else std::cout << elapsed * 0.001 << " s." << std::endl;
    std::cout << (stats.nmf_count - stats.max_count) << "/" << stats.nmf_count << " factorizations converged." << std::endl << std::endl;

    if (Result::OK != result && Result::FLATCLUST_FAILURE != result)
        std::cerr << "\nHierarchical clustering fatal error." << std::endl;
    else
    {
        if (verbose) std::cout << "Writing output files..." << std::endl;
        if (!tree.WriteAssignments(opts.assignfile)) std::cerr << "\terror writing assignments file" << std::endl;
        IHierclustWriter* writer = CreateHierclustWriter(opts.format);
        if (!tree.WriteTree(writer, opts.treefile, dictionary)) std::cerr << "\terror writing factorization file" << std::endl;
        if (opts.clust_opts.flat && Result::FLATCLUST_FAILURE != result)
            FlatClustWriteResults(opts.outdir, assignments_flat, probabilities, dictionary, term_indices, opts.format,
                                  opts.clust_opts.maxterms, n, num_clusters);
        delete writer;
    }
    NmfFinalize();
    return (Result::OK == result || Result::FLATCLUST_FAILURE == result) ? 0 : -1;
}
