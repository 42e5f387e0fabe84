"""The reference-facing host layer on the GPU: the `nmf` command-line tool (reference option set) and the
smallk:: C++ API, both checked against the CPU oracle on the C1 configuration (BASELINE configs[0])."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "smallk_b200", "bin")


def write_csv(path, a, fmt="%.17e"):
    np.savetxt(path, a, delimiter=",", fmt=fmt)


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def c1(tmp_path_factory):
    d = tmp_path_factory.mktemp("c1")
    rng = np.random.default_rng(1)
    A = rng.random((256, 256))
    write_csv(d / "A.csv", A, "%.6e")                      # matrixgen writes 6 digits
    A = np.loadtxt(d / "A.csv", delimiter=",")
    W0 = rng.random((256, 16)); H0 = rng.random((16, 256))
    write_csv(d / "W0.csv", W0); write_csv(d / "H0.csv", H0)
    return d, A, W0, H0


def test_nmf_cli_bpp_c1(c1, oracle):
    d, A, W0, H0 = c1
    cmd = [os.path.join(BIN, "nmf"), "--matrixfile", str(d / "A.csv"), "--k", "16", "--algorithm", "BPP",
           "--tol", "0.01", "--miniter", "5", "--maxiter", "5000", "--infile_W", str(d / "W0.csv"),
           "--infile_H", str(d / "H0.csv"), "--outfile_W", str(d / "w.csv"), "--outfile_H", str(d / "h.csv"),
           "--outprecision", "17", "--verbose", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    o = oracle.nmf_dense(A, W0, H0, alg="BPP", tol=0.01, min_iter=5, max_iter=5000, normalize=True)
    assert f"Iterations: {o['iterations']}" in out.stdout
    W = np.loadtxt(d / "w.csv", delimiter=","); H = np.loadtxt(d / "h.csv", delimiter=",")
    assert rel(W, o["W"]) < 1e-9 and rel(H, o["H"]) < 1e-9


def test_nmf_cli_rejects_bad_options(c1):
    d = c1[0]
    exe = os.path.join(BIN, "nmf")
    r = subprocess.run([exe, "--matrixfile", str(d / "A.csv"), "--k", "16", "--algorithm", "NOPE"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid value specified for command-line argument NOPE" in r.stderr
    # RANK2 with another k: the reference warns and factors with k = 2 (nmf/src/command_line.cpp:341-349)
    r = subprocess.run([exe, "--matrixfile", str(d / "A.csv"), "--k", "3", "--algorithm", "RANK2", "--verbose", "0", "--maxiter", "20"],
                       capture_output=True, text=True, cwd=str(d))
    assert r.returncode == 0 and "warning: forcing k=2 for RANK2 algorithm" in r.stderr, r.stderr
    assert np.loadtxt(d / "w.csv", delimiter=",").shape[1] == 2
    r = subprocess.run([exe, "--matrixfile", str(d / "A.csv"), "--k", "-3"], capture_output=True, text=True)
    assert r.returncode != 0 and "k-value must be a positive integer" in r.stderr
    r = subprocess.run([exe, "--matrixfile", str(d / "A.csv"), "--k", "300"], capture_output=True, text=True)
    assert r.returncode != 0 and "k value cannot exceed" in r.stderr


def test_smallk_api_hals_sparse_mtx(c1, oracle, tmp_path):
    """smallk::LoadMatrix(.mtx) -> Nmf(k, HALS, initW, initH) -> w.csv / h.csv in the output dir."""
    import scipy.sparse as sps
    rng = np.random.default_rng(3)
    m, n, k = 200, 150, 6
    S = sps.random(m, n, density=0.3, random_state=np.random.RandomState(3), format="coo",
                   data_rvs=np.random.RandomState(4).random_sample)
    with open(tmp_path / "A.mtx", "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m, n, S.nnz))
        for r, c, v in zip(S.row, S.col, S.data):
            f.write("%d %d %.17e\n" % (r + 1, c + 1, v))
    W0 = rng.random((m, k)); H0 = rng.random((k, n)) * 0.1
    write_csv(tmp_path / "W0.csv", W0); write_csv(tmp_path / "H0.csv", H0)
    exe = os.path.join(BIN, "smallk_example")
    out = subprocess.run([exe, str(tmp_path / "A.mtx"), str(k), "HALS", str(tmp_path / "W0.csv"), str(tmp_path / "H0.csv"),
                          str(tmp_path), "0.05", "5", "40"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    # same CSC as the loader builds: stable by column, file order inside a column
    order = np.argsort(S.col, kind="stable")
    colp = np.concatenate([[0], np.cumsum(np.bincount(S.col, minlength=n))]).astype(np.uint32)
    o = oracle.nmf_sparse((m, n), colp, S.row[order].astype(np.uint32), S.data[order], W0, H0, alg="HALS", tol=0.05,
                          min_iter=5, max_iter=40, normalize=True)
    W = np.loadtxt(tmp_path / "w.csv", delimiter=","); H = np.loadtxt(tmp_path / "h.csv", delimiter=",")
    assert rel(W, o["W"]) < 1e-8 and rel(H, o["H"]) < 1e-8


# ---------------------------------------------------------------------------
# hierclust / flatclust command-line tools and smallk::HierNmf2
# ---------------------------------------------------------------------------
import sys                                           # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_hier as mh                        # noqa: E402


def write_mtx_general(path, shape, colp, rowi, val):
    """All stored entries in CSC order, so the loader's stable column sort rebuilds exactly this CSC."""
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (shape[0], shape[1], len(val)))
        for c in range(shape[1]):
            for e in range(colp[c], colp[c + 1]):
                f.write("%d %d %.17g\n" % (rowi[e] + 1, c + 1, val[e]))


def read_assignments(path):
    with open(path) as f:
        return np.array([int(x) for x in f.readline().strip().split(",")])


@pytest.fixture(scope="module")
def hier_case(tmp_path_factory):
    name = "hier_graph_2000_c6"
    d = tmp_path_factory.mktemp("hier")
    g = mh.hier_inputs(name)
    write_mtx_general(d / "A.mtx", g["shape"], *g["csc"])
    with open(d / "dict.txt", "w") as f:
        for i in range(g["shape"][0]):
            f.write("term%d\n" % i)
    return d, g, np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))


def test_hierclust_cli_matches_reference_fixture(hier_case):
    d, g, z = hier_case
    out = subprocess.run([os.path.join(BIN, "hierclust"), "--matrixfile", str(d / "A.mtx"), "--dictfile", str(d / "dict.txt"),
                          "--clusters", str(g["num_clusters"]), "--outdir", str(d), "--seed", str(g["seed"]), "--format", "JSON",
                          "--verbose", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    got = read_assignments(d / ("assignments_%d.csv" % g["num_clusters"]))
    assert np.array_equal(got, z["assignments"])
    import json
    tree = json.load(open(d / ("tree_%d.json" % g["num_clusters"])))
    assert [nd["parent_id"] for nd in tree["nodes"]] == z["parent"].tolist()
    assert [nd["left_child_id"] for nd in tree["nodes"]] == z["left"].tolist()
    assert [nd["right_child_id"] for nd in tree["nodes"]] == z["right"].tolist()
    assert [nd["doc_count"] for nd in tree["nodes"]] == z["doc_count"].tolist()
    assert [[int(t[4:]) for t in nd["top_terms"]] for nd in tree["nodes"]] == z["terms"].tolist()
    assert tree["doc_count"] == int(z["doc_count"][z["is_leaf"] == 1].sum())


def test_smallk_api_hiernmf2_writes_tree_and_assignments(hier_case):
    d, g, z = hier_case
    exe = os.path.join(BIN, "smallk_example")
    out = subprocess.run([exe, "--hier", str(d / "A.mtx"), str(d / "dict.txt"), str(g["num_clusters"]), str(d), str(g["seed"]), "0", "XML"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    # smallk::HierNmf2 runs each factorization with normalize = true (smallk.cpp:766): same partition of the
    # documents as the CLI run (normalisation rescales W columns / H rows by positive factors)
    got = read_assignments(d / ("assignments_%d.csv" % g["num_clusters"]))
    assert got.shape == z["assignments"].shape and got.min() >= -1
    xml = open(d / ("tree_%d.xml" % g["num_clusters"])).read()
    assert xml.startswith('<?xml version="1.0"?>') and xml.count("<node id=") == 2 * (g["num_clusters"] - 1)


def test_flatclust_cli_matches_oracle(oracle, tmp_path):
    rng = np.random.default_rng(12)
    m, n, k = 120, 90, 5
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    write_csv(tmp_path / "A.csv", A); write_csv(tmp_path / "W0.csv", W0); write_csv(tmp_path / "H0.csv", H0)
    with open(tmp_path / "dict.txt", "w") as f:
        for i in range(m):
            f.write("t%d\n" % i)
    out = subprocess.run([os.path.join(BIN, "flatclust"), "--matrixfile", str(tmp_path / "A.csv"), "--dictfile", str(tmp_path / "dict.txt"),
                          "--clusters", str(k), "--algorithm", "HALS", "--infile_W", str(tmp_path / "W0.csv"), "--infile_H",
                          str(tmp_path / "H0.csv"), "--outdir", str(tmp_path), "--tol", "0.01", "--format", "XML", "--verbose", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    o = oracle.nmf_dense(A, W0, H0, alg="HALS", tol=0.01, min_iter=5, max_iter=5000, normalize=True)
    want = np.argmax(o["H"], axis=0)
    got = read_assignments(tmp_path / ("assignments_%d.csv" % k))
    assert np.array_equal(got, want)
    fuzzy = np.loadtxt(tmp_path / ("assignments_fuzzy_%d.csv" % k), delimiter=",")
    assert fuzzy.shape == (n, k) and np.allclose(fuzzy.sum(axis=1), 1.0, atol=2e-3)
    xml = open(tmp_path / ("clusters_%d.xml" % k)).read()
    top0 = np.argsort(-o["W"][:, 0], kind="stable")[:5]
    assert all(('<term name="t%d"/>' % t) in xml for t in top0)
    # MU is not a flatclust algorithm (flatclust/src/command_line.cpp:233-244)
    r = subprocess.run([os.path.join(BIN, "flatclust"), "--matrixfile", str(tmp_path / "A.csv"), "--dictfile", str(tmp_path / "dict.txt"),
                        "--clusters", str(k), "--algorithm", "MU"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid value specified for command-line argument MU" in r.stderr
