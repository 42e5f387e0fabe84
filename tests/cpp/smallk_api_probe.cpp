// Probe of the smallk:: API surface that needs no device: settings, matrix / dictionary loading, argument checks of Nmf and
// HierNmf2. Compiled twice by tests/test_smallk_api_cpu.py — against this repository's host layer (smallk_b200/host/smallk.hpp,
// libsmallk_host.so) and against the reference's own smallk.hpp / smallk.cpp (oracle/_ref) — and the two transcripts are
// compared line by line: same return values, same exception types, same messages.
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "smallk.hpp"
#include "nmf.hpp"
#include "flat_clust.hpp"

static void probe(const char* name, const std::function<std::string()>& f)
{
    std::string out;
    try { out = "ok " + f(); }
    catch (std::logic_error& e) { out = std::string("logic_error: ") + e.what(); }
    catch (std::runtime_error& e) { out = std::string("runtime_error: ") + e.what(); }
    catch (std::exception& e) { out = std::string("exception: ") + e.what(); }
    std::printf("PROBE|%s|%s\n", name, out.c_str());
    std::fflush(stdout);
}

template <typename T> static std::string str(const T& v) { std::ostringstream s; s << v; return s.str(); }

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : ".";      // holds a.csv (3 x 4), a.mtx, dict.txt, bad.mtx
    using namespace smallk;
    Reset();                                               // the reference sets its defaults in Initialize() / Reset(); Initialize needs a device here
    probe("version", [] { return str(GetMajorVersion()) + "." + str(GetMinorVersion()) + "." + str(GetPatchLevel()) + " " + GetVersionString(); });
    probe("defaults", [] { return str(GetOutputPrecision()) + " " + str(GetNmfTolerance()) + " " + str(GetMaxIter()) + " " + str(GetMinIter()) + " " +
                                  str(GetMaxTerms()) + " " + str(static_cast<int>(GetOutputFormat())) + " " + str(GetHierNmf2Tolerance()) + " [" + GetOutputDir() + "]"; });
    probe("loaded0", [] { return str(IsMatrixLoaded()); });
    probe("tol0", [] { SetNmfTolerance(0.0); return std::string(); });
    probe("tol1", [] { SetNmfTolerance(1.0); return std::string(); });
    probe("tolneg", [] { SetNmfTolerance(-0.1); return std::string(); });
    probe("tolok", [] { SetNmfTolerance(0.25); return str(GetNmfTolerance()); });
    probe("toldefault", [] { SetNmfTolerance(); return str(GetNmfTolerance()); });
    probe("htol0", [] { SetHierNmf2Tolerance(0.0); return std::string(); });
    probe("htol2", [] { SetHierNmf2Tolerance(2.0); return std::string(); });
    probe("htolok", [] { SetHierNmf2Tolerance(0.01); return str(GetHierNmf2Tolerance()); });
    probe("prec0", [] { SetOutputPrecision(0); return str(GetOutputPrecision()); });
    probe("prec100", [] { SetOutputPrecision(100); return str(GetOutputPrecision()); });
    probe("prec9", [] { SetOutputPrecision(9); return str(GetOutputPrecision()); });
    probe("maxiter0", [] { SetMaxIter(0); return str(GetMaxIter()); });
    probe("miniter0", [] { SetMinIter(0); return str(GetMinIter()); });
    probe("maxterms0", [] { SetMaxTerms(0); return str(GetMaxTerms()); });
    probe("maxterms7", [] { SetMaxTerms(7); return str(GetMaxTerms()); });
    probe("format", [] { SetOutputFormat(XML); const int a = GetOutputFormat(); SetOutputFormat(); return str(a) + " " + str(static_cast<int>(GetOutputFormat())); });
    probe("threads", [] { SetMaxThreads(1); const unsigned int a = GetMaxThreads(); SetMaxThreads(0); return str(a) + " " + str(GetMaxThreads()); });
    probe("outdir_missing_abs", [] { SetOutputDir("/nonexistent_dir_for_probe/x"); return GetOutputDir(); });
    probe("outdir_missing_rel", [] { SetOutputDir("nonexistent_dir_for_probe"); return std::string("set"); });
    probe("outdir_ok", [&] { SetOutputDir(dir); return GetOutputDir(); });
    probe("outdir_ok_sep", [&] { SetOutputDir(dir + "/"); return GetOutputDir(); });
    probe("nmf_nomatrix", [] { Nmf(2); return std::string(); });
    probe("hier_nomatrix", [] { HierNmf2(4); return std::string(); });
    probe("load_empty_name", [] { LoadMatrix(std::string()); return std::string(); });
    probe("load_missing_csv", [] { LoadMatrix(std::string("/nonexistent_probe.csv")); return std::string(); });
    probe("load_missing_mtx", [] { LoadMatrix(std::string("/nonexistent_probe.mtx")); return std::string(); });
    probe("loaded1", [] { return str(IsMatrixLoaded()); });
    probe("load_bad_mtx", [&] { LoadMatrix(dir + "/bad.mtx"); return std::string(); });
    probe("load_csv", [&] { LoadMatrix(dir + "/a.csv"); return str(IsMatrixLoaded()); });
    probe("nmf_k0", [] { Nmf(0); return std::string(); });
    probe("nmf_minmax", [] { SetMinIter(10); SetMaxIter(5); Nmf(2); return std::string(); });
    probe("restore_iters", [] { SetMinIter(5); SetMaxIter(5000); return str(GetMinIter()) + " " + str(GetMaxIter()); });
    probe("hier_nodict", [] { HierNmf2(4); return std::string(); });
    probe("dict_missing", [] { LoadDictionary(std::string("/nonexistent_probe_dict.txt")); return std::string(); });
    probe("dict_ok", [&] { LoadDictionary(dir + "/dict.txt"); return std::string(); });
    probe("hier_zero", [] { HierNmf2(0); return std::string(); });
    probe("hierflat_zero", [] { HierNmf2WithFlat(0); return std::string(); });
    probe("load_mtx", [&] { LoadMatrix(dir + "/a.mtx"); return str(IsMatrixLoaded()); });
    {
        std::vector<double> buf(12, 1.0);
        probe("buf_null", [] { LoadMatrix(static_cast<const double*>(nullptr), 3, 3, 4); return std::string(); });
        probe("buf_h0", [&] { LoadMatrix(buf.data(), 3, 0, 4); return std::string(); });
        probe("buf_w0", [&] { LoadMatrix(buf.data(), 3, 3, 0); return std::string(); });
        probe("buf_ok", [&] { LoadMatrix(buf.data(), 3, 3, 3); return str(IsMatrixLoaded()); });
    }
    {
        std::vector<double> d = {1.0, 2.0, 3.0};
        std::vector<unsigned int> r = {0, 1, 2}, c = {0, 1, 2, 3};
        const std::vector<double> none_d; const std::vector<unsigned int> none_u;
        probe("sp_mismatch", [&] { LoadMatrix(3, 3, 3, d, std::vector<unsigned int>(2, 0u), c); return std::string(); });
        probe("sp_h0", [&] { LoadMatrix(0, 3, 3, d, r, c); return std::string(); });
        probe("sp_w0", [&] { LoadMatrix(3, 0, 3, d, r, c); return std::string(); });
        probe("sp_toomany", [&] { LoadMatrix(1, 2, 3, d, r, c); return std::string(); });
        probe("sp_empty", [&] { LoadMatrix(3, 3, 0, none_d, none_u, c); return std::string(); });
        probe("sp_nocols", [&] { LoadMatrix(3, 3, 3, d, r, none_u); return std::string(); });
        probe("sp_ok", [&] { LoadMatrix(3, 3, 3, d, r, c); return str(IsMatrixLoaded()); });
    }
    probe("dict_vector", [] { LoadDictionary(std::vector<std::string>(3, "t")); return std::string(); });
    probe("reset", [] { Reset(); return str(IsMatrixLoaded()) + " " + str(GetOutputPrecision()) + " " + str(GetNmfTolerance()) + " " + str(GetMaxIter()) + " " +
                                        str(GetMinIter()) + " " + str(GetMaxTerms()) + " " + str(static_cast<int>(GetOutputFormat())) + " " + str(GetHierNmf2Tolerance()) +
                                        " [" + GetOutputDir() + "]"; });
    probe("hier_after_reset", [] { HierNmf2(4); return std::string(); });
    // the L3 functions before NmfInitialize (common/src/nmf.cpp:173-186, flatclust/src/flat_clust.cpp:144-160)
    {
        NmfOptions o;
        o.tol = 0.005; o.algorithm = NmfAlgorithm::BPP; o.prog_est_algorithm = NmfProgressAlgorithm::PG_RATIO;
        o.height = 3; o.width = 3; o.k = 2; o.min_iter = 1; o.max_iter = 2; o.tolcount = 1; o.max_threads = 1; o.verbose = false; o.normalize = true;
        std::vector<double> A(9, 1.0), W(6, 0.5), H(6, 0.5), d = {1.0, 2.0, 3.0};
        std::vector<unsigned int> r = {0, 1, 2}, c = {0, 1, 2, 3};
        NmfStats st;
        probe("l3_isinit", [] { return str(static_cast<int>(NmfIsInitialized())); });
        probe("l3_nmf_uninit", [&] { return str(static_cast<int>(Nmf(o, A.data(), 3, W.data(), 3, H.data(), 2, st))); });
        probe("l3_nmfsparse_uninit", [&] { return str(static_cast<int>(NmfSparse(o, 3, 3, 3, c.data(), r.data(), d.data(), W.data(), 3, H.data(), 2, st))); });
        probe("l3_flatclust_uninit", [&] { return str(static_cast<int>(FlatClust(o, A.data(), 3, W.data(), 3, H.data(), 2, st))); });
        probe("l3_flatclustsparse_uninit", [&] { return str(static_cast<int>(FlatClustSparse(o, 3, 3, 3, c.data(), r.data(), d.data(), W.data(), 3, H.data(), 2, st))); });
        o.k = 0;
        probe("l3_isvalid_k0", [&] { return str(IsValid(o)); });
        o.k = 2; o.algorithm = NmfAlgorithm::RANK2; o.k = 3;
        probe("l3_isvalid_rank2_k3", [&] { return str(IsValid(o)); });
    }
    return 0;
}
