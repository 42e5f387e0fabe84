"""Command-line front ends without a device: everything the reference's nmf / hierclust / flatclust tools decide before they
initialise the library (option parsing, required arguments, value checks, their order and their texts) must be decided the
same way by smallk_b200/bin/{nmf,hierclust,flatclust}. The reference tools are built here from their own main.cpp /
command_line.cpp against oracle/_ref (skipped where /root/reference is absent) and both are run over a grid of command lines.

Two differences are deliberate: an invalid enumeration value (--algorithm FOO) makes the reference throw an uncaught
std::runtime_error (abort, SIGABRT) where this tool prints the same text and exits with -1; and past validation this tool stops
at "no usable sm_100 CUDA device" on a machine without one, where the reference goes on to load files and factor on the CPU."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
REF_LIB = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "smallk_b200", "bin")
TOOLS = ("nmf", "hierclust", "flatclust")
NO_DEVICE = "no usable sm_100 CUDA device"
# what the reference may say once it is PAST its pre-initialisation checks
PAST_VALIDATION = ("load failed for file", "could not load dictionary file", "dimensions of matrix", "unsupported file type",
                   "warning: forcing k=2", "solver failure", "Cholesky")


@pytest.fixture(scope="module")
def tools(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir(REF) or not os.path.exists(os.path.join(REF_LIB, "libsmallk_ref.so")):
        pytest.skip("g++, /root/reference or oracle/_ref not present on this machine")
    if not all(os.path.exists(os.path.join(BIN, t)) for t in TOOLS):
        pytest.skip("host tools not built")
    d = str(tmp_path_factory.mktemp("cli"))
    for t in TOOLS:
        others = [f"-I{REF}/{o}/include" for o in ("hierclust", "flatclust") if o != t]
        cmd = ["g++", "-std=c++11", "-O1", "-fopenmp", "-DELEM_VER=85", "-DEL_HAVE_OPENMP", "-DNDEBUG", "-include", "functional", "-w",
               f"-I{REF}/{t}/include", f"-I{ROOT}/oracle/shim", f"-I{REF}/common/include"] + others + \
              ["-o", os.path.join(d, "ref_" + t), f"{REF}/{t}/src/main.cpp", f"{REF}/{t}/src/command_line.cpp",
               "-L" + REF_LIB, "-lsmallk_ref", "-Wl,-rpath," + REF_LIB]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    with open(os.path.join(d, "a.csv"), "w") as f:
        f.write("1,2,3,4\n5,6,7,8\n9,10,11,12\n")
    with open(os.path.join(d, "a.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 3\n1 1 1.0\n2 2 2.0\n3 3 3.0\n")
    with open(os.path.join(d, "dict.txt"), "w") as f:
        f.write("a\nb\nc\n")
    return d


def _grid(d):
    A, M, D = (os.path.join(d, x) for x in ("a.csv", "a.mtx", "dict.txt"))
    base_n = ["--matrixfile", A, "--k", "2"]
    base_h = ["--matrixfile", M, "--dictfile", D, "--clusters", "2"]
    base_f = ["--matrixfile", A, "--dictfile", D, "--clusters", "2"]
    return {
        "nmf": [[], ["--help"], ["--k", "2"], ["--matrixfile", A], ["--matrixfile", A, "--k", "0"], ["--matrixfile", A, "--k", "-3"],
                base_n + ["--algorithm", "FOO"], base_n + ["--algorithm", "rank2"], ["--matrixfile", A, "--k", "3", "--algorithm", "RANK2"],
                ["--matrixfile", A, "--algorithm", "RANK2"], base_n + ["--tol", "0"], base_n + ["--tol", "1.5"], base_n + ["--stopping", "XYZ"],
                base_n + ["--stopping", "delta"], base_n + ["--miniter", "0"], base_n + ["--maxiter", "0"], base_n + ["--tolcount", "0"],
                base_n + ["--infile_W", A], base_n + ["--infile_H", A], base_n + ["--outprecision", "0"], base_n + ["--maxthreads", "0"],
                base_n + ["--normalize", "7"], base_n + ["--verbose", "x"], ["--matrixfile", "/nope.csv", "--k", "2"], ["--matrixfile", D, "--k", "2"],
                ["--bogus", "1"], ["--matrixfile", A, "--k"], base_n + ["extra"]],
        "hierclust": [[], ["--help"], ["--matrixfile", M], ["--matrixfile", M, "--dictfile", D], ["--matrixfile", M, "--dictfile", D, "--clusters", "1"],
                      ["--matrixfile", M, "--dictfile", D, "--clusters", "0"], ["--matrixfile", M, "--dictfile", "/nope.txt", "--clusters", "2"],
                      ["--matrixfile", "/nope.mtx", "--dictfile", D, "--clusters", "2"], base_h + ["--tol", "0"], base_h + ["--unbalanced", "1.0"],
                      base_h + ["--unbalanced", "-0.1"], base_h + ["--trial_allowance", "-1"], base_h + ["--maxterms", "0"], base_h + ["--format", "YAML"],
                      base_h + ["--format", "xml"], base_h + ["--outdir", "/nonexistent_dir_q"], base_h + ["--flat", "5"], base_h + ["--miniter", "0"],
                      base_h + ["--maxiter", "-2"], base_h + ["--initdir", "/nonexistent_dir_q"], base_h + ["--bogus", "1"], base_h + ["--tol"]],
        "flatclust": [[], ["--help"], ["--matrixfile", A], ["--matrixfile", A, "--dictfile", D], ["--matrixfile", A, "--dictfile", D, "--clusters", "0"],
                      base_f + ["--algorithm", "MU"], base_f + ["--algorithm", "FOO"], ["--matrixfile", A, "--dictfile", D, "--clusters", "3", "--algorithm", "RANK2"],
                      base_f + ["--tol", "2"], ["--matrixfile", A, "--dictfile", "/nope.txt", "--clusters", "2"], ["--matrixfile", "/nope.csv", "--dictfile", D, "--clusters", "2"],
                      base_f + ["--maxterms", "0"], base_f + ["--format", "YAML"], base_f + ["--outdir", "/nonexistent_dir_q"], base_f + ["--infile_W", A],
                      base_f + ["--miniter", "0"], base_f + ["--maxiter", "0"], base_f + ["--bogus"], base_f + ["--clusters"]],
    }


def _run(exe, args, d):
    r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=d, timeout=120)
    lines = [ln.strip() for ln in r.stderr.replace(exe, "<exe>").splitlines() if ln.strip()]
    return r.returncode, lines, r.stdout


def test_tools_decide_like_the_reference_before_initialisation(tools):
    d = tools
    checked = same = 0
    for tool, cases in _grid(d).items():
        for args in cases:
            ref_rc, ref_err, ref_out = _run(os.path.join(d, "ref_" + tool), args, d)
            our_rc, our_err, our_out = _run(os.path.join(BIN, tool), args, d)
            checked += 1
            if any(NO_DEVICE in ln for ln in our_err):
                # this tool passed validation: so must the reference have (it then runs on the CPU, or fails on a file)
                before = [ln for ln in our_err if NO_DEVICE not in ln]
                assert ref_rc == 0 or any(any(p in ln for p in PAST_VALIDATION) for ln in ref_err), (tool, args, ref_rc, ref_err)
                assert before == [ln for ln in ref_err if ln in before], (tool, args, before, ref_err)       # e.g. the RANK2 warning
                continue
            if ref_rc == -6:                                   # uncaught std::runtime_error in the reference
                assert our_rc == 255, (tool, args)
                assert ref_err[-1].replace("what():", "").strip() == our_err[-1], (tool, args, ref_err, our_err)
                continue
            assert (ref_rc, ref_err) == (our_rc, our_err), (tool, args, ref_rc, ref_err, our_rc, our_err)
            assert ("Usage" in ref_out) == ("Usage" in our_out), (tool, args)
            same += 1
    assert checked >= 65 and same >= 30
