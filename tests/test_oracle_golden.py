"""Pins the CPU oracle (oracle/nmf_oracle.c) to the reference: against the committed fixtures that the
reference's own code produced (tests/golden/make_golden.py), and — where the reference build exists
(this container) — live against oracle/_ref on fresh inputs."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg      # noqa: E402
from oracle import Ref        # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9       # north_star tolerance; observed agreement is ~1e-13


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def run_oracle(oracle, g, trace=True):
    kw = dict(alg=g["alg"], tol=g["tol"], min_iter=g["min_iter"], max_iter=g["max_iter"], normalize=g["normalize"], trace=trace)
    if g["kind"] == "dense":
        return oracle.nmf_dense(g["A"], g["W0"], g["H0"], **kw)
    return oracle.nmf_sparse((g["m"], g["n"]), *g["sp"], g["W0"], g["H0"], **kw)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_reference_fixture(oracle, name):
    g = mg.golden_inputs(name)
    z = np.load(os.path.join(GOLD, name + ".npz"))
    r = run_oracle(oracle, g)
    assert r["rc"] == 0
    assert r["iterations"] == int(z["iterations"])
    tol = 1e-6 if "hals" in name else TOL      # HALS: discontinuous clamp, see test_gpu_parity
    ok = ~np.isnan(z["metrics"])
    assert np.array_equal(ok, ~np.isnan(r["metrics"]))
    assert np.all(np.abs(r["metrics"][ok] - z["metrics"][ok]) <= tol * np.abs(z["metrics"][ok]))
    assert rel(r["W"], z["W"]) < tol and rel(r["H"], z["H"]) < tol
    for j, it in enumerate(z["snap_iters"]):
        assert rel(r["W_trace"][it], z["W_snaps"][j]) < tol
        assert rel(r["H_trace"][it], z["H_snaps"][j]) < tol


@pytest.mark.parametrize("name", sorted(mg.NNLS_CASES))
def test_oracle_nnls_reproduces_reference_fixture(oracle, name):
    LHS, RHS, X0 = mg.nnls_inputs(*mg.NNLS_CASES[name])
    k, q = RHS.shape
    z = np.load(os.path.join(GOLD, name + ".npz"))
    rc, X, Y = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    assert np.array_equal(X > 0, z["X"] > 0)
    assert rel(X, z["X"]) < 1e-10 and np.abs(Y - z["Y"]).max() < 1e-9 * max(1.0, np.abs(z["Y"]).max())


@pytest.mark.parametrize("name", sorted(mg.BACKUP_CASES))
def test_oracle_backup_rule_cases_reproduce_reference_fixture(oracle, name):
    """UpdatePassiveSet's backup rule (nnls.cpp:64-72) fires in these cases; where MaxRowIndex's defect (bit_matrix.cpp:459-467)
    picks the wrong row the reference itself fails (rc -4), and so must the restatement."""
    import ctypes
    LHS, RHS, X0 = mg.backup_inputs(*mg.BACKUP_CASES[name])
    z = np.load(os.path.join(GOLD, name + ".npz"))
    oracle.lib.orc_backup_count.restype = ctypes.c_double
    oracle.lib.orc_stats_reset()
    rc, X, Y = oracle.nnls_bpp(LHS, RHS, X0)
    assert oracle.lib.orc_backup_count() > 0
    assert rc == int(z["rc"])
    assert (rc == 0) == ("defect" not in name)
    if rc == 0:
        assert np.array_equal(X > 0, z["X"] > 0)
        assert rel(X, z["X"]) < 1e-10 and np.abs(Y - z["Y"]).max() < 1e-9 * max(1.0, np.abs(z["Y"]).max())


needs_ref = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (no /root/reference on this box)")


@needs_ref
@pytest.mark.parametrize("alg,m,n,k", [("BPP", 120, 150, 12), ("BPP", 90, 200, 35), ("HALS", 100, 140, 9),
                                       ("MU", 80, 70, 6), ("RANK2", 200, 150, 2)])
@pytest.mark.parametrize("prog", ["PG_RATIO", "DELTA_FNORM"])
def test_oracle_matches_reference_live_dense(oracle, alg, m, n, k, prog):
    ref = Ref()
    rng = np.random.default_rng(m + n + k)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    kw = dict(alg=alg, prog=prog, tol=1e-12, min_iter=1, max_iter=20, trace=True)
    a = oracle.nmf_dense(A, W0, H0, **kw)
    b = ref.nmf_dense(A, W0, H0, max_threads=2, **kw)
    assert a["rc"] == b["rc"] == 0 and a["iterations"] == b["iterations"]
    assert rel(a["W"], b["W"]) < TOL and rel(a["H"], b["H"]) < TOL
    tol = 1e-6 if alg == "HALS" else TOL
    assert np.nanmax(np.abs(a["metrics"] - b["metrics"]) / np.abs(b["metrics"])) < tol


@needs_ref
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_oracle_sparse_gemm_matches_reference(oracle, variant):
    """The reference's own invariant (tests/src/test_sparse_gemm.cpp:22): sparse Gemm == dense Gemm to 1e-10;
    here additionally oracle == reference for every orientation, including the 2-column 'rank-2' thread path."""
    import scipy.sparse as sps
    ref = Ref()
    m, n = 120, 90
    S = sps.random(m, n, density=0.08, random_state=np.random.RandomState(4), format="csc")
    rng = np.random.default_rng(variant)
    for k in (2, 7):
        shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
        shapeC = (m, k) if variant < 2 else (k, n)
        B = rng.random(shapeB); C = rng.random(shapeC)
        a = oracle.sparse_gemm(variant, 0.7, (m, n), S.indptr, S.indices, S.data, B, -1.3, C)
        b = ref.sparse_gemm(variant, 0.7, (m, n), S.indptr, S.indices, S.data, B, -1.3, C, max_threads=3)
        D = S.toarray()
        dense = [lambda: D @ B, lambda: D @ B.T, lambda: B @ D, lambda: B.T @ D][variant]()
        assert np.linalg.norm(a - b) < 1e-12
        assert np.linalg.norm(b - (0.7 * dense - 1.3 * C)) < 1e-10


@needs_ref
def test_oracle_nnls_matches_reference_live(oracle):
    ref = Ref()
    for args in [(31, 8, 40), (32, 33, 70), (33, 64, 50), (34, 100, 30), (35, 200, 20, -0.5), (36, 130, 25, -0.1)]:
        LHS, RHS, X0 = mg.nnls_inputs(*args)
        rc1, X1, Y1 = oracle.nnls_bpp(LHS, RHS, X0)
        rc2, X2, Y2 = ref.nnls_bpp(LHS, RHS, X0, max_threads=2)
        assert rc1 == rc2 == 0
        assert np.array_equal(X1 > 0, X2 > 0)
        assert rel(X1, X2) < 1e-10


@needs_ref
def test_all_zero_warm_start_reaches_the_reference_solution(oracle):
    """X0 = 0 (zero initialisers read from a file): every passive set starts empty. The reference then takes its global branch —
    the unconstrained solve for all columns while the BitMatrix stays empty (nmf_solver_bpp.hpp:174-178) — and pivots on the signs
    of a residual that is rounding noise; the restatement (and the CUDA kernels) start such a column from x = 0 instead (DESIGN.md
    section 7). The paths differ; the solution BPP converges to is the NNLS optimum either way: same support, same values."""
    ref = Ref()
    for k, q, seed in [(8, 20, 1), (16, 64, 2), (40, 90, 3), (100, 60, 4), (200, 30, 5)]:
        rng = np.random.default_rng(seed)
        W = rng.random((4 * k, k)); A = rng.random((4 * k, q))
        LHS = W.T @ W
        RHS = W.T @ A - 0.35 * rng.random((k, q)) * np.abs(W.T @ A).mean()
        X0 = np.zeros((k, q))
        rc1, X1, Y1 = oracle.nnls_bpp(LHS, RHS, X0.copy())
        rc2, X2, Y2 = ref.nnls_bpp(LHS, RHS, X0.copy(), max_threads=1)
        assert rc1 == rc2 == 0
        assert np.array_equal(X1 > 0, X2 > 0)
        assert rel(X1, X2) < 1e-10
        assert np.abs(Y1 - Y2).max() <= 1e-9 * max(1.0, np.abs(Y2).max())
