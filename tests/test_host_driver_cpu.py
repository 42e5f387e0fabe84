"""The hierclust tree driver (smallk_b200/host/clust.cpp: trial splits, node factors kept on their own rows, priority scores on
worker threads, the next split made ahead of time and taken back when the last score overrules it) on the CPU: the host library is
built a second time against tests/cpp/mock_capi.cpp, which answers the C ABI with the oracle's solvers, and must grow the
reference's own trees (tests/golden/hier_*.npz) — with the workers and with everything on the calling thread. Test infrastructure
only: nothing in smallk_b200/ links the mock."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import smallk_b200 as sk

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_hier as mh      # noqa: E402
sys.path.insert(0, HERE)

BUILD = os.path.join(HERE, "cpp", "build")
HOST = os.path.join(ROOT, "smallk_b200", "host")
EXACT = ["assignments", "parent", "left", "right", "is_left", "doc_count", "is_leaf", "terms"]


@pytest.fixture(scope="module")
def mock_host():
    import oracle
    if not os.path.exists(oracle.ORACLE_SO):
        oracle.build(ref=False)
    os.makedirs(BUILD, exist_ok=True)
    odir = os.path.dirname(oracle.ORACLE_SO)
    mock = os.path.join(BUILD, "libsmallk_mock.so")
    host = os.path.join(BUILD, "libsmallk_host_mock.so")
    cxx = ["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-pthread"]
    subprocess.check_call(cxx + ["-o", mock, os.path.join(HERE, "cpp", "mock_capi.cpp"), "-L" + odir, "-l:" + os.path.basename(oracle.ORACLE_SO),
                                 "-Wl,-rpath," + odir])
    srcs = [os.path.join(HOST, f) for f in ("nmf_host.cpp", "smallk.cpp", "clust.cpp", "flat_clust.cpp", "host_capi.cpp")]
    subprocess.check_call(cxx + ["-o", host] + srcs + ["-L" + BUILD, "-lsmallk_mock", "-Wl,-rpath," + BUILD])
    lib = ctypes.CDLL(host)
    lib.smkh_last_error.restype = ctypes.c_char_p
    return lib


def check_tree(got, want, name):
    assert got["rc"] == 0, (name, got["rc"])
    for key in EXACT:
        assert np.array_equal(got[key], want[key]), (name, key, got[key], want[key])
    assert int(got["n_outliers"]) == int(want["n_outliers"])
    assert int(got["nmf_count"]) == int(want["nmf_count"])
    assert np.allclose(got["priority"], want["priority"], rtol=1e-4, atol=0), (name, got["priority"], want["priority"])


CASES = sorted(mh.HIER_CASES)


@pytest.mark.parametrize("workers", ["1", "0"])
@pytest.mark.parametrize("name", CASES)
def test_tree_driver_over_the_cpu_mock_grows_the_reference_tree(mock_host, name, workers, monkeypatch):
    monkeypatch.setenv("SMK_HIER_ASYNC", workers)
    g = mh.hier_inputs(name)
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    got = sk.hierclust(A_dense=g["A"], csc=g["csc"], shape=g["shape"], num_clusters=g["num_clusters"], seed=g["seed"], lib=mock_host, **g["extra"])
    check_tree(got, z, name)
    if g["extra"].get("flat"):
        # HierNmf2WithFlat: the leaves' topic vectors as W, NnlsHals for H (the mock's restatement of it), flat assignments
        assert np.array_equal(got["flat_assignments"], z["flat_assignments"])
        assert np.linalg.norm(got["H"] - z["H"]) <= 1e-7 * np.linalg.norm(z["H"])
        assert np.linalg.norm(got["W"] - z["W"]) <= 1e-9 * np.linalg.norm(z["W"])


def test_tree_driver_over_the_cpu_mock_matches_live_reference_on_fresh_seeds(mock_host):
    """Power-law graphs the fixtures do not hold, against the compiled reference run here (one thread: the sequential initialiser):
    more splits made ahead of time, more of them taken back."""
    from oracle import Ref
    from graphgen import powerlaw_graph
    if not Ref.available():
        pytest.skip("oracle/_ref not built on this machine")
    ref = Ref()
    for n, deg, gseed, clusters, seed in ((2500, 9, 77, 7, 21), (1200, 16, 78, 5, 22), (1800, 10, 80, 12, 24)):
        csc = powerlaw_graph(n, deg, gseed)
        want = ref.hierclust(csc=csc, shape=(n, n), num_clusters=clusters, seed=seed, max_threads=1)
        got = sk.hierclust(csc=csc, shape=(n, n), num_clusters=clusters, seed=seed, lib=mock_host)
        check_tree(got, want, f"graph{n}")
