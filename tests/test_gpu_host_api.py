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
    assert r.returncode != 0 and "invalid command line value" in r.stderr
    r = subprocess.run([exe, "--matrixfile", str(d / "A.csv"), "--k", "3", "--algorithm", "RANK2"], capture_output=True, text=True)
    assert r.returncode != 0
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
