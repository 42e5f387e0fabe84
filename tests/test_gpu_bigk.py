"""k > 256: the any-k fallback kernels against the CPU oracle, through the C ABI.

The reference puts no limit on k for MU, HALS and BPP (nmf.cpp:191-219 only checks that m*k and n*k fit an int). The tuned
kernels of this library hold a k-vector in the registers or shared memory of a warp / CTA (k <= 256); beyond that the same
algorithms run on fallback kernels with runtime loops over k (csrc/nnls_bpp_wide.cu: nnls_bpp_big_kernel; csrc/factors.cu:
hals_sweep_*_big_kernel, row_sumsq_big_kernel; csrc/spmm.cu: the dense operand in row blocks of 256). Same tolerances as
tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

import smallk_b200 as sk

pytestmark = pytest.mark.gpu

REL_FACTOR = 1e-9
REL_PRIM = 1e-12


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _nnls_problem(k, q, seed):
    rng = np.random.default_rng(seed)
    m = 4 * k
    W = rng.random((m, k))
    A = rng.random((m, q))
    LHS = W.T @ W
    RHS = W.T @ A - 0.35 * rng.random((k, q)) * np.abs(W.T @ A).mean()
    X0 = rng.random((k, q)) * (rng.random((k, q)) > 0.3)
    return LHS, RHS, X0


@pytest.mark.parametrize("k,q,seed", [(257, 90, 1), (300, 200, 2), (384, 70, 3), (515, 40, 4)])
def test_nnls_bpp_any_k_matches_oracle(gpu, oracle, k, q, seed):
    LHS, RHS, X0 = _nnls_problem(k, q, seed)
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    assert np.array_equal(X > 0, Xo > 0)            # identical passive sets
    assert rel(X, Xo) < 1e-9
    assert np.abs(Y - Yo).max() <= 1e-8 * max(1.0, np.abs(Yo).max())
    assert X.min() >= 0.0


def test_nnls_bpp_any_k_mostly_passive_matches_oracle(gpu, oracle):
    """Large passive sets (a 270 x 270 Cholesky per solve) from the generator of the reference-made NNLS fixtures."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    k, q = 300, 60
    LHS, RHS, X0 = mg.nnls_inputs(321, k, q, -0.1)
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    assert np.array_equal(X > 0, Xo > 0)
    assert rel(X, Xo) < 1e-9
    assert np.abs(Y - Yo).max() <= 1e-8 * max(1.0, np.abs(Yo).max())


def test_nnls_bpp_any_k_non_hpd_fails(gpu):
    k, q = 300, 6
    with pytest.raises(sk.SmallkError) as e:
        gpu.nnls_bpp(-np.eye(k), np.ones((k, q)), np.ones((k, q)))
    assert e.value.code == sk.FAILURE


def _trace_gpu(ctx, W0, H0, opts, iters):
    ctx.solver_begin(W0, H0, opts)
    metrics, Ws, Hs = [], [], []
    for _ in range(iters):
        ctx.solver_step(1)
        metrics.append(ctx.solver_progress())
        W, H = ctx.solver_get()
        Ws.append(W); Hs.append(H)
    return np.array(metrics), Ws, Hs


@pytest.mark.parametrize("alg,m,n,k,iters,prog", [
    ("BPP", 340, 320, 288, 2, "PG_RATIO"),
    ("BPP", 280, 260, 257, 2, "DELTA_FNORM"),
    ("MU", 350, 300, 300, 8, "PG_RATIO"),
    ("HALS", 400, 360, 300, 5, "DELTA_FNORM"),
    ("HALS", 340, 390, 272, 4, "PG_RATIO"),
])
def test_dense_trace_any_k_matches_oracle(gpu, oracle, alg, m, n, k, iters, prog):
    rng = np.random.default_rng(sum(map(ord, alg)) * 7919 + m * 31 + n * 17 + k)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 2.0 / k                      # W0*H0 ~ A (see test_gpu_parity.py::test_dense_trace_matches_oracle)
    o = oracle.nmf_dense(A, W0, H0, alg=alg, prog=prog, tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm=alg, prog=prog, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    mtol = 1e-5 if (alg == "HALS" and prog == "PG_RATIO") else REL_FACTOR      # clamp discontinuity, as in test_gpu_parity.py
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))
        assert abs(metrics[i] - o["metrics"][i]) <= mtol * abs(o["metrics"][i]), (i, metrics[i], o["metrics"][i])


@pytest.mark.parametrize("alg", ["BPP", "HALS"])
def test_nmf_call_any_k_matches_oracle(gpu, oracle, alg):
    """The one-call interface (the stop-tested loop as one CUDA graph) with the fallback kernels inside the capture, final
    NormalizeAndScale (row_sumsq_big_kernel) included."""
    m, n, k = 290, 270, 260
    rng = np.random.default_rng(12)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 2.0 / k
    kw = dict(tol=1e-9, min_iter=2, max_iter=3)
    o = oracle.nmf_dense(A, W0, H0, alg=alg, normalize=True, **kw)
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm=alg, normalize=True, **kw)
    W, H, st = gpu.nmf(W0, H0, opts)
    assert st.iteration_count == o["iterations"]
    assert rel(W, o["W"]) < REL_FACTOR and rel(H, o["H"]) < REL_FACTOR
    assert np.allclose(np.linalg.norm(W, axis=0), 1.0, atol=1e-12)


def _zipf_csc(m, n, per_col, seed):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows = np.minimum((m ** rng.random((n, per_col))).astype(np.int64) - 1, m - 1).clip(0)
    cols = np.repeat(np.arange(n), per_col)
    S = sp.csc_matrix((rng.random(n * per_col) + 0.1, (rows.ravel(), cols)), shape=(m, n))
    S.sort_indices()
    return S


@pytest.mark.parametrize("k", [257, 300, 600])
def test_sparse_gemm_any_k_matches_oracle(gpu, oracle, k):
    """The dense operand in row blocks of 256 (600 = 256 + 256 + 88): all four variants, with and without beta, hub rows cut
    into segments."""
    m, n = 2500, 400
    S = _zipf_csc(m, n, 50, 23)
    rng = np.random.default_rng(k)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    for variant in (0, 1, 2, 3):
        shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
        shapeC = (m, k) if variant < 2 else (k, n)
        B = rng.random(shapeB); C = rng.random(shapeC)
        for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
            got = gpu.sparse_gemm(variant, alpha, B, beta, C)
            want = oracle.sparse_gemm(variant, alpha, (m, n), S.indptr, S.indices, S.data, B, beta, C)
            assert rel(got, want) < REL_PRIM


@pytest.mark.parametrize("alg,iters", [("HALS", 4), ("BPP", 2)])
def test_sparse_nmf_any_k_matches_oracle(gpu, oracle, alg, iters):
    import scipy.sparse as sp
    m, n, k = 360, 300, 264
    rng = np.random.default_rng(31)
    S = sp.random(m, n, density=0.08, random_state=5, format="csc", data_rvs=rng.random)
    S.sort_indices()
    W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 0.1 / k
    kw = dict(tol=1e-12, min_iter=1, max_iter=iters, normalize=True)
    o = oracle.nmf_sparse((m, n), S.indptr, S.indices, S.data, W0, H0, alg=alg, **kw)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    opts = sk.make_options(m, n, k, algorithm=alg, **kw)
    W, H, st = gpu.nmf(W0, H0, opts)
    assert st.iteration_count == o["iterations"]
    assert rel(W, o["W"]) < REL_FACTOR and rel(H, o["H"]) < REL_FACTOR


def test_nnls_hals_any_k_matches_restatement(gpu):
    """NnlsHals (nnls.hpp:249-316) at k = 270 against the NumPy restatement of the reference's loop."""
    from oracle.nnls_hals_oracle import nnls_hals
    rng = np.random.default_rng(19)
    m, n, k = 900, 120, 270
    Wt = rng.random((m, k)) * (rng.random((m, k)) < 0.15); Ht = rng.random((k, n)) * (rng.random((k, n)) < 0.2)
    A = Wt @ Ht
    gpu.load_dense(A)
    H0 = rng.random((k, n))
    ok, Wo, Ho, ito = nnls_hals(A, Wt, H0, 1e-3, 400)
    assert ok and ito > 5
    rc, W, H, it = gpu.nnls_hals(Wt, H0, 1e-3, 400)
    assert rc == 0 and it == ito, (rc, it, ito)
    assert rel(W, Wo) < REL_FACTOR and rel(H, Ho) < REL_FACTOR
