// nmf — command-line front end with the reference's option set (nmf/src/command_line.cpp:34-54, defaults
// :173-194; flow nmf/src/main.cpp:41-256), running on the GPU library through host/nmf.hpp.
#include <getopt.h>

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "matrix_io.hpp"
#include "nmf.hpp"

namespace {
struct CommandLineOptions
{
    NmfOptions nmf_opts;
    bool show_help = false;
    std::string infile_A, infile_W, infile_H, outfile_W = "w.csv", outfile_H = "h.csv";
    int output_precision = 6;
};

option longopts[] = {
    {"matrixfile", required_argument, NULL, 'a'}, {"k", required_argument, NULL, 'b'},
    {"algorithm", required_argument, NULL, 'c'},  {"stopping", required_argument, NULL, 'd'},
    {"tol", required_argument, NULL, 'e'},        {"tolcount", required_argument, NULL, 'f'},
    {"infile_W", required_argument, NULL, 'g'},   {"infile_H", required_argument, NULL, 'h'},
    {"outfile_W", required_argument, NULL, 'i'},  {"outfile_H", required_argument, NULL, 'j'},
    {"miniter", required_argument, NULL, 'k'},    {"maxiter", required_argument, NULL, 'l'},
    {"outprecision", required_argument, NULL, 'm'}, {"maxthreads", required_argument, NULL, 'n'},
    {"normalize", required_argument, NULL, 'o'},  {"verbose", required_argument, NULL, 'p'},
    {"help", no_argument, NULL, 'q'},             {0, 0, 0, 0}};

void ShowHelp(const char* prog)
{
    std::cout << "\nUsage: " << prog << "\n"
              << "        --matrixfile <filename>  Filename of the matrix to be factored.\n"
              << "                                 Either CSV format for dense or MatrixMarket format for sparse.\n"
              << "        --k <integer value>      Inner dimension for factors W and H.\n"
              << "        [--algorithm  BPP]       NMF algorithm: MU, HALS, RANK2, BPP\n"
              << "        [--stopping  PG_RATIO]   Stopping criterion: PG_RATIO, DELTA\n"
              << "        [--tol  0.005]           Tolerance for the selected stopping criterion.\n"
              << "        [--tolcount  1]          Tolerance count; declare convergence after this many\n"
              << "                                 iterations with metric < tolerance.\n"
              << "        [--infile_W  (empty)]    Dense mxk matrix to initialize W; CSV file.\n"
              << "        [--infile_H  (empty)]    Dense kxn matrix to initialize H; CSV file.\n"
              << "        [--outfile_W  w.csv]     Filename for the W matrix result.\n"
              << "        [--outfile_H  h.csv]     Filename for the H matrix result.\n"
              << "        [--miniter  5]           Minimum number of iterations to perform.\n"
              << "        [--maxiter  5000]        Maximum number of iterations to perform.\n"
              << "        [--outprecision  6]      Write results with this many digits after the decimal point.\n"
              << "        [--maxthreads  N]        Accepted for compatibility; the GPU path ignores it.\n"
              << "        [--normalize  1]         Whether to normalize W and scale H. 1 == yes, 0 == no\n"
              << "        [--verbose  1]           Whether to print updates to the screen. 1 == print updates, 0 == silent\n";
}

bool Parse(int argc, char* argv[], CommandLineOptions& o)
{
    o.nmf_opts.algorithm = NmfAlgorithm::BPP;
    o.nmf_opts.height = o.nmf_opts.width = o.nmf_opts.k = 0;
    o.nmf_opts.min_iter = 5; o.nmf_opts.max_iter = 5000;
    o.nmf_opts.verbose = true; o.nmf_opts.normalize = true;
    o.nmf_opts.tol = 0.005; o.nmf_opts.tolcount = 1;
    o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
    o.nmf_opts.max_threads = 1;
    int c, index;
    while (-1 != (c = getopt_long(argc, argv, ":a:b:c:d:e:f:g:h:i:j:k:l:m:n:o:p:q", longopts, &index)))
    {
        std::string tmp = optarg ? optarg : "";
        std::string up = tmp;
        std::transform(up.begin(), up.end(), up.begin(), ::toupper);
        switch (c)
        {
        case 'a': o.infile_A = tmp; break;
        case 'b': o.nmf_opts.k = std::atoi(optarg); break;
        case 'c':
            if (up == "MU") o.nmf_opts.algorithm = NmfAlgorithm::MU;
            else if (up == "HALS") o.nmf_opts.algorithm = NmfAlgorithm::HALS;
            else if (up == "RANK2") o.nmf_opts.algorithm = NmfAlgorithm::RANK2;
            else if (up == "BPP") o.nmf_opts.algorithm = NmfAlgorithm::BPP;
            else { std::cerr << "Invalid value specified for command-line argument " << up << std::endl; return false; }
            break;
        case 'd':
            if (up == "PG_RATIO") o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
            else if (up == "DELTA") o.nmf_opts.prog_est_algorithm = NmfProgressAlgorithm::DELTA_FNORM;
            else { std::cerr << "Invalid value specified for command-line argument " << up << std::endl; return false; }
            break;
        case 'e': o.nmf_opts.tol = std::atof(optarg); break;
        case 'f': o.nmf_opts.tolcount = std::atoi(optarg); break;
        case 'g': o.infile_W = tmp; break;
        case 'h': o.infile_H = tmp; break;
        case 'i': o.outfile_W = tmp; break;
        case 'j': o.outfile_H = tmp; break;
        case 'k': o.nmf_opts.min_iter = std::atoi(optarg); break;
        case 'l': o.nmf_opts.max_iter = std::atoi(optarg); break;
        case 'm':
        {
            int p = std::atoi(optarg);
            if (p <= 0) p = std::numeric_limits<float>::max_digits10;
            else if (p >= std::numeric_limits<double>::max_digits10) p = std::numeric_limits<double>::max_digits10;
            o.output_precision = p;
            break;
        }
        case 'n': o.nmf_opts.max_threads = std::max(1, std::atoi(optarg)); break;
        case 'o': o.nmf_opts.normalize = (0 != std::atoi(optarg)); break;
        case 'p': o.nmf_opts.verbose = (0 != std::atoi(optarg)); break;
        case 'q': o.show_help = true; break;
        case ':': std::cerr << "missing argument for option " << argv[optind - 1] << std::endl; return false;
        default: std::cerr << "invalid option: " << argv[optind - 1] << std::endl; return false;
        }
    }
    if (1 == argc) o.show_help = true;
    if (o.show_help) return false;
    if (o.infile_A.empty()) { std::cerr << "required command line argument --matrixfile not found" << std::endl; return false; }
    // nmf/src/command_line.cpp:331-350: k is "missing" only when it is 0 and the algorithm is not RANK2 (a negative k goes on to
    // IsValid); RANK2 forces k = 2 with a warning
    if (0 == o.nmf_opts.k && NmfAlgorithm::RANK2 != o.nmf_opts.algorithm)
    { std::cerr << "required command line argument --k not found" << std::endl; return false; }
    if (NmfAlgorithm::RANK2 == o.nmf_opts.algorithm && 2 != o.nmf_opts.k)
    { std::cerr << "warning: forcing k=2 for RANK2 algorithm" << std::endl; o.nmf_opts.k = 2; }
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
    if (!IsValid(opts.nmf_opts, false)) return -1;            // nmf/src/main.cpp:60-61: before the library is initialised
    try { NmfInitialize(argc, argv); }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; return -1; }

    const bool sparse = smallk_io::IsMatrixMarketFile(opts.infile_A);
    {
        // nmf/src/main.cpp:70-103: .mtx is sparse, .csv is dense, anything else is refused
        const std::string& f = opts.infile_A;
        const bool csv = f.size() >= 4 && f.compare(f.size() - 4, 4, ".csv") == 0;
        if (!sparse && !csv) { std::cerr << "\nunsupported file type: " << f << std::endl; NmfFinalize(); return -1; }
    }
    std::vector<double> buf_a;
    smallk_io::CscMatrix A;
    unsigned int m = 0, n = 0;
    if (opts.nmf_opts.verbose) std::cout << "Loading matrix..." << std::endl;
    bool ok = false;
    try { ok = sparse ? smallk_io::LoadMatrixMarketFile(opts.infile_A, A) : smallk_io::LoadDelimitedFile(buf_a, m, n, opts.infile_A); }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }      // index out of bounds, no entries
    if (!ok) { std::cerr << "\nload failed for file " << opts.infile_A << std::endl; NmfFinalize(); return -1; }
    if (sparse) { m = A.height; n = A.width; }
    opts.nmf_opts.height = m; opts.nmf_opts.width = n;
    if (!IsValid(opts.nmf_opts)) { NmfFinalize(); return -1; }
    const unsigned int k = opts.nmf_opts.k;

    std::vector<double> buf_w(static_cast<size_t>(m) * k), buf_h(static_cast<size_t>(k) * n);
    std::mt19937 engine(static_cast<unsigned>(time(0)));       // nmf/src/main.cpp seeds from the clock
    std::uniform_real_distribution<double> dist;
    auto random_fill = [&](std::vector<double>& b) { for (auto& v : b) v = 0.5 + 2.0 * 0.5 * dist(engine) - 0.5; };
    unsigned int hw = m, ww = k, hh = k, wh = n;
    if (opts.infile_W.empty()) random_fill(buf_w);
    else if (!smallk_io::LoadDelimitedFile(buf_w, hw, ww, opts.infile_W))
    { std::cerr << "\nload failed for file " << opts.infile_W << std::endl; NmfFinalize(); return -1; }
    if (hw != m || ww != k)
    { std::cerr << "\tdimensions of matrix W are " << hw << " x " << ww << "\n\texpected " << m << " x " << k << std::endl; NmfFinalize(); return -1; }
    if (opts.infile_H.empty()) random_fill(buf_h);
    else if (!smallk_io::LoadDelimitedFile(buf_h, hh, wh, opts.infile_H))
    { std::cerr << "\nload failed for file " << opts.infile_H << std::endl; NmfFinalize(); return -1; }
    if (hh != k || wh != n)
    { std::cerr << "\tdimensions of matrix H are " << hh << " x " << wh << "\n\texpected " << k << " x " << n << std::endl; NmfFinalize(); return -1; }

    NmfStats stats;
    Result result;
    try
    {
        if (sparse)
            result = NmfSparse(opts.nmf_opts, A.height, A.width, A.nnz(), A.col_offsets.data(), A.row_indices.data(),
                               A.data.data(), buf_w.data(), m, buf_h.data(), k, stats);
        else
            result = Nmf(opts.nmf_opts, buf_a.data(), m, buf_w.data(), m, buf_h.data(), k, stats);
    }
    catch (std::exception& e) { std::cerr << e.what() << std::endl; NmfFinalize(); return -1; }

    std::cout << "Elapsed wall clock time: " << stats.elapsed_us / 1000.0 << " ms." << std::endl;
    std::cout << "Iterations: " << stats.iteration_count << std::endl;
    if (Result::OK == result)
    {
        if (opts.nmf_opts.verbose) std::cout << "Writing output files..." << std::endl;
        if (!smallk_io::WriteDelimitedFile(buf_w.data(), m, m, k, opts.outfile_W, opts.output_precision))
            std::cerr << "\tcould not write W result " << std::endl;
        if (!smallk_io::WriteDelimitedFile(buf_h.data(), k, k, n, opts.outfile_H, opts.output_precision))
            std::cerr << "\tcould not write H result " << std::endl;
    }
    else std::cerr << "NMF solver failure: " << NmfLastError() << std::endl;
    NmfFinalize();
    return Result::OK == result ? 0 : -1;
}
