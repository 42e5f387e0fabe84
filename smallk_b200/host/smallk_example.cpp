// Minimal walk through the smallk:: API on the GPU build (the reference ships examples/smallk_example.cpp
// for the same purpose): load a matrix from a file, factor it from W/H init files, read the factors back.
//   smallk_example <matrixfile> <k> <ALG> <initW.csv> <initH.csv> <outdir> [tol] [miniter] [maxiter]
//   smallk_example --hier <matrixfile> <dictfile> <clusters> <outdir> <seed> <flat 0|1> <format XML|JSON>
#include <cstdlib>
#include <iostream>
#include <string>

#include "smallk.hpp"

int main(int argc, char* argv[])
{
    if (argc >= 9 && std::string(argv[1]) == "--hier")
    {
        try
        {
            smallk::Initialize(argc, argv);
            smallk::SetOutputDir(argv[5]);
            smallk::SeedRNG(std::atoi(argv[6]));
            smallk::SetOutputFormat(std::string(argv[8]) == "XML" ? smallk::XML : smallk::JSON);
            smallk::LoadMatrix(argv[2]);
            bool threw = false;
            try { smallk::HierNmf2(4); } catch (std::logic_error&) { threw = true; }     // no dictionary yet
            if (!threw) { std::cerr << "expected std::logic_error without a dictionary" << std::endl; return 4; }
            smallk::LoadDictionary(argv[3]);
            if (std::atoi(argv[7])) smallk::HierNmf2WithFlat(std::atoi(argv[4])); else smallk::HierNmf2(std::atoi(argv[4]));
            smallk::Finalize();
        }
        catch (std::exception& e) { std::cerr << "exception: " << e.what() << std::endl; return 1; }
        return 0;
    }
    if (argc < 7) { std::cerr << "usage: smallk_example <matrixfile> <k> <ALG> <initW> <initH> <outdir> [tol] [miniter] [maxiter]" << std::endl; return 2; }
    try
    {
        smallk::Initialize(argc, argv);
        if (!smallk::IsInitialized()) return 3;
        std::cout << "smallk version " << smallk::GetVersionString() << std::endl;
        smallk::SetOutputDir(argv[6]);
        smallk::SetOutputPrecision(17);
        if (argc > 7) smallk::SetNmfTolerance(std::atof(argv[7]));
        if (argc > 8) smallk::SetMinIter(std::atoi(argv[8]));
        if (argc > 9) smallk::SetMaxIter(std::atoi(argv[9]));
        smallk::LoadMatrix(argv[1]);
        const std::string alg = argv[3];
        smallk::Algorithm a = smallk::BPP;
        if (alg == "MU") a = smallk::MU; else if (alg == "HALS") a = smallk::HALS; else if (alg == "RANK2") a = smallk::RANK2;
        smallk::Nmf(std::atoi(argv[2]), a, argv[4], argv[5]);
        unsigned int ld, h, w;
        const double* W = smallk::LockedBufferW(ld, h, w);
        std::cout << "W is " << h << " x " << w << ", W(0,0) = " << W[0] << std::endl;
        // error behaviour of the reference API: logic_error on misuse
        bool threw = false;
        try { smallk::Nmf(0); } catch (std::logic_error&) { threw = true; }
        if (!threw) { std::cerr << "expected std::logic_error for k == 0" << std::endl; return 4; }
        smallk::Finalize();
    }
    catch (std::exception& e) { std::cerr << "exception: " << e.what() << std::endl; return 1; }
    return 0;
}
