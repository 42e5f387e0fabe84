"""Edge shapes through the C ABI against the CPU oracle: k = 1, matrices of a handful of rows / columns, a single row, and a
sparse input with empty rows and empty columns (first, middle and last). Same tolerances as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import smallk_b200 as sk

pytestmark = pytest.mark.gpu

REL_FACTOR = 1e-9


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _trace_gpu(ctx, W0, H0, opts, iters):
    ctx.solver_begin(W0, H0, opts)
    metrics, Ws, Hs = [], [], []
    for _ in range(iters):
        ctx.solver_step(1)
        metrics.append(ctx.solver_progress())
        W, H = ctx.solver_get()
        Ws.append(W); Hs.append(H)
    return np.array(metrics), Ws, Hs


@pytest.mark.parametrize("alg,m,n,k,iters", [
    ("BPP", 50, 40, 1, 3), ("MU", 50, 40, 1, 3), ("HALS", 50, 40, 1, 3),      # one factor (three iterations: by the fifth the metric is 1e-7 of its start, rounding noise of a cancelled gradient)
    ("BPP", 7, 6, 3, 5), ("HALS", 6, 5, 2, 6), ("MU", 4, 3, 2, 6), ("RANK2", 6, 5, 2, 6),   # a handful of rows and columns
    ("BPP", 1, 30, 1, 4),                                                      # a single row
])
def test_tiny_dense_trace_matches_oracle(gpu, oracle, alg, m, n, k, iters):
    rng = np.random.default_rng(1000 * m + 10 * n + k + len(alg))
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 2.0 / k
    o = oracle.nmf_dense(A, W0, H0, alg=alg, tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    mtol = 1e-5 if alg == "HALS" else REL_FACTOR          # clamp discontinuity of the HALS metric (test_gpu_parity.py)
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))
        if m == 1:
            continue        # a single row is factored exactly by the first iteration: the projected gradient, and so the metric, is rounding noise
        assert abs(metrics[i] - o["metrics"][i]) <= mtol * abs(o["metrics"][i]), (i, metrics[i], o["metrics"][i])


def _holey_csc(m, n, seed):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    S = sp.random(m, n, density=0.08, random_state=seed, format="lil", data_rvs=rng.random)
    S[10:30, :] = 0; S[m - 1, :] = 0; S[0, :] = 0            # empty rows: first, a block, last
    S[:, 5:15] = 0; S[:, n - 1] = 0; S[:, 0] = 0              # empty columns: first, a block, last
    S = S.tocsc(); S.eliminate_zeros(); S.sort_indices()
    return S


@pytest.mark.parametrize("alg,k,iters", [("BPP", 4, 6), ("MU", 4, 8), ("HALS", 4, 6), ("RANK2", 2, 8)])
def test_sparse_input_with_empty_rows_and_columns_matches_oracle(gpu, oracle, alg, k, iters):
    m, n = 120, 90
    S = _holey_csc(m, n, 2)
    assert (np.diff(S.indptr) == 0).sum() >= 12
    rng = np.random.default_rng(7)
    W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 0.3 / k
    o = oracle.nmf_sparse((m, n), S.indptr, S.indices, S.data, W0, H0, alg=alg, tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))


def test_sparse_gemm_on_input_with_empty_rows_and_columns_matches_oracle(gpu, oracle):
    m, n, k = 120, 90, 7
    S = _holey_csc(m, n, 3)
    rng = np.random.default_rng(8)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    for variant in (0, 1, 2, 3):
        shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
        shapeC = (m, k) if variant < 2 else (k, n)
        B = rng.random(shapeB); C = rng.random(shapeC)
        for alpha, beta in [(1.0, 0.0), (-0.4, 2.0)]:
            got = gpu.sparse_gemm(variant, alpha, B, beta, C)
            want = oracle.sparse_gemm(variant, alpha, (m, n), S.indptr, S.indices, S.data, B, beta, C)
            assert rel(got, want) < 1e-12


@pytest.mark.parametrize("k,q,seed", [(8, 20, 1), (40, 90, 3), (100, 60, 4), (300, 40, 6)])
def test_nnls_bpp_from_an_all_zero_warm_start_matches_oracle(gpu, oracle, k, q, seed):
    """X0 = 0: every passive set starts empty (the case of zero initialisers read from a file; DESIGN.md section 7 — the oracle and
    the reference reach the same solution from it, tests/test_oracle_golden.py). All three NNLS kernels (k <= 64, <= 256, any k)."""
    rng = np.random.default_rng(seed)
    W = rng.random((4 * k, k)); A = rng.random((4 * k, q))
    LHS = W.T @ W
    RHS = W.T @ A - 0.35 * rng.random((k, q)) * np.abs(W.T @ A).mean()
    X0 = np.zeros((k, q))
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0.copy())
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0.copy())
    assert np.array_equal(X > 0, Xo > 0)
    assert rel(X, Xo) < 1e-9
    assert np.abs(Y - Yo).max() <= 1e-8 * max(1.0, np.abs(Yo).max())
