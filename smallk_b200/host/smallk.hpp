// smallk_b200 — the reference's public C++ API (smallk/include/smallk.hpp:34-332), implemented on
// the GPU library. Same namespace, names, default arguments and exception types, so a program using
// smallk::Initialize / LoadMatrix / Nmf / LockedBufferW/H re-links unchanged.
// HierNmf2 / HierNmf2WithFlat drive the rank-2 solver from the host-side tree code of host/clust.cpp.
#pragma once

#include <string>
#include <vector>

#define SMALLK_MAJOR_VERSION 1
#define SMALLK_MINOR_VERSION 6
#define SMALLK_PATCH_LEVEL   2

namespace smallk
{
    enum Algorithm { MU, BPP, HALS, RANK2 };
    enum OutputFormat { XML, JSON };

    // ---- life cycle: Initialize creates the GPU context (throws std::runtime_error without an sm_100 device)
    void Initialize(int& argc, char**& argv);   bool IsInitialized();   void Finalize();
    unsigned int GetMajorVersion();   unsigned int GetMinorVersion();   unsigned int GetPatchLevel();   std::string GetVersionString();

    // ---- settings (defaults as in the reference: 6 digits, tol 0.005, 5..5000 iterations, 5 terms, JSON, HierNMF2 tol 1e-4)
    unsigned int GetOutputPrecision();      void SetOutputPrecision(const unsigned int num_digits = 6);
    double GetNmfTolerance();               void SetNmfTolerance(const double tol = 0.005);
    unsigned int GetMaxIter();              void SetMaxIter(const unsigned int max_iterations = 5000);
    unsigned int GetMinIter();              void SetMinIter(const unsigned int min_iterations = 5);
    unsigned int GetMaxThreads();           void SetMaxThreads(const unsigned int max_threads);
    unsigned int GetMaxTerms();             void SetMaxTerms(const unsigned int max_terms = 5);
    OutputFormat GetOutputFormat();         void SetOutputFormat(const OutputFormat format = JSON);
    double GetHierNmf2Tolerance();          void SetHierNmf2Tolerance(const double tol = 0.0001);
    std::string GetOutputDir();             void SetOutputDir(const std::string& outdir);
    void Reset();                           void SeedRNG(const int seed);

    // ---- the matrix: a .mtx / .csv file, a dense column-major buffer, or CSC arrays; uploaded to the GPU once
    void LoadMatrix(const std::string& filepath);
    void LoadMatrix(const double* buffer, const unsigned int ldim, const unsigned int height, const unsigned int width);
    void LoadMatrix(const unsigned int height, const unsigned int width, const unsigned int nz, const std::vector<double>& data,
                    const std::vector<unsigned int>& row_indices, const std::vector<unsigned int>& col_offsets);
    bool IsMatrixLoaded();

    // ---- NMF: factors stay in library-owned buffers (column-major, ldim_w = m, ldim_h = k) until the next Nmf / Reset
    void Nmf(const unsigned int k, const Algorithm algorithm = BPP, const std::string& initfile_w = std::string(""),
             const std::string& initfile_h = std::string(""));
    const double* LockedBufferW(unsigned int& ldim, unsigned int& height, unsigned int& width);
    const double* LockedBufferH(unsigned int& ldim, unsigned int& height, unsigned int& width);

    // ---- clustering: needs a dictionary (one term per matrix row); results are written to the output directory
    void LoadDictionary(const std::string& filepath);   void LoadDictionary(const std::vector<std::string>& terms);
    void HierNmf2(const unsigned int num_clusters);      void HierNmf2WithFlat(const unsigned int num_clusters);
} // namespace smallk
