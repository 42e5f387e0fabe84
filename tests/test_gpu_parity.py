"""Parity of the CUDA path (through the C ABI) against the CPU oracle, same seeded inputs.

Tolerances: the reference computes in double; north_star asks for factor entries and the
per-iteration progress metric within 1e-9 relative. Primitive products are held to 1e-12.
"""
import numpy as np
import pytest

import smallk_b200 as sk

pytestmark = pytest.mark.gpu

REL_FACTOR = 1e-9      # north_star tolerance for factors / per-iteration metric
REL_PRIM = 1e-12       # single products


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ---------------------------------------------------------------------------
# dense products (El::Gemm call sites of the solvers)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("tA,tB,M,N,K", [
    (True, False, 16, 256, 256),      # W'A  at C1
    (False, True, 300, 16, 257),      # A H' (as Gemm sees it)
    (True, False, 64, 1000, 3000),    # W'A, k = 64, ragged tiles
    (True, True, 7, 33, 129),         # odd everything -> 8-byte cp.async path
    (False, False, 64, 513, 64),      # WtW * H
    (False, True, 64, 64, 5000),      # H H'
    (True, False, 130, 200, 300),     # k > 64: two M tiles
    (True, False, 64, 1300, 4100),    # the TMA kernel (M * N > 65536): NN, ragged N tile, reduction tail of 4
    (False, True, 64, 1304, 2052),    # the TMA kernel: NT, ragged N tile and reduction tail
    (True, False, 128, 700, 1000),    # the TMA kernel: two M tiles
    (True, False, 72, 1000, 1000),    # the TMA kernel: ragged M (zero-filled boxes)
    (False, True, 40, 2000, 777),     # the TMA kernel: NT, M < one tile, odd reduction length
    (False, True, 64, 1302, 2050),    # the TMA kernel: NT, N even but not a multiple of the 8-wide boxes
    (True, False, 72, 1000, 999),     # odd leading dimension: back to the cp.async kernel
])
def test_gemm_matches_numpy(gpu, tA, tB, M, N, K):
    rng = np.random.default_rng(M * 1000 + N + K)
    A = rng.random((K, M) if tA else (M, K))
    B = rng.random((N, K) if tB else (K, N))
    C = gpu.gemm(A, B, transA=tA, transB=tB)
    ref = (A.T if tA else A) @ (B.T if tB else B)
    assert rel(C, ref) < REL_PRIM


@pytest.mark.parametrize("tB,M,N,K", [(False, 64, 1300, 4100), (True, 64, 1304, 2052), (True, 128, 900, 3000)])
def test_gemm_in_kernel_split_reduction_matches_numpy(gpu, tB, M, N, K, monkeypatch):
    """The FIX instantiations (split-R reduction by the last-arriving CTA of a tile; on several GPUs the scatter epilogue lives
    there) forced on one GPU, TMA and cp.async main loops."""
    monkeypatch.setenv("SMK_GEMM_FIXUP", "1")
    rng = np.random.default_rng(M + N + K)
    A = rng.random((K, M))
    B = rng.random((N, K) if tB else (K, N))
    ref = A.T @ (B.T if tB else B)
    assert rel(gpu.gemm(A, B, transA=True, transB=tB), ref) < REL_PRIM


@pytest.mark.parametrize("tB,M,N,K", [(False, 64, 1300, 4100), (True, 64, 1304, 2052)])
def test_gemm_cp_async_kernel_on_a_sized_products_matches_numpy(gpu, tB, M, N, K, monkeypatch):
    """SMK_GEMM_TMA=0: the cp.async kernel (the fallback for unaligned operands) on shapes the TMA kernel normally takes."""
    monkeypatch.setenv("SMK_GEMM_TMA", "0")
    rng = np.random.default_rng(M + N + K + 1)
    A = rng.random((K, M))
    B = rng.random((N, K) if tB else (K, N))
    ref = A.T @ (B.T if tB else B)
    assert rel(gpu.gemm(A, B, transA=True, transB=tB), ref) < REL_PRIM


# ---------------------------------------------------------------------------
# NNLS-BPP (NnlsBlockpivot)
# ---------------------------------------------------------------------------
def _nnls_problem(k, q, seed, m=None):
    rng = np.random.default_rng(seed)
    m = m or 4 * k
    W = rng.random((m, k))
    A = rng.random((m, q))
    LHS = W.T @ W
    RHS = W.T @ A - 0.35 * rng.random((k, q)) * np.abs(W.T @ A).mean()   # push part of the solution to the boundary
    X0 = rng.random((k, q)) * (rng.random((k, q)) > 0.3)
    return LHS, RHS, X0


@pytest.mark.parametrize("k,q,seed", [(16, 256, 1), (5, 37, 2), (32, 500, 3), (33, 300, 4), (64, 700, 5), (48, 1, 6),
                                      (65, 300, 7), (100, 500, 8), (129, 200, 9), (200, 333, 10), (256, 450, 11)])
def test_nnls_bpp_matches_oracle(gpu, oracle, k, q, seed):
    LHS, RHS, X0 = _nnls_problem(k, q, seed)
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    # identical passive sets
    assert np.array_equal(X > 0, Xo > 0)
    assert rel(X, Xo) < 1e-10
    assert np.abs(Y - Yo).max() <= 1e-9 * max(1.0, np.abs(Yo).max())
    # KKT: x >= 0, y >= 0 on the active set, complementary
    assert X.min() >= 0.0
    assert Y[X == 0].min() >= -1e-9 if (X == 0).any() else True


@pytest.mark.parametrize("k,q,frac,seed", [(70, 200, 0.1, 1), (128, 150, 0.5, 2), (200, 120, 0.15, 3), (256, 300, 0.05, 4),
                                           (256, 100, 0.5, 5), (250, 80, 0.45, 6)])
def test_nnls_bpp_wide_mostly_passive_matches_oracle(gpu, oracle, k, q, frac, seed):
    """k > 64 with large passive sets: |P| > 128 takes the complement path of nnls_bpp_wide_kernel."""
    import sys as _s, os as _o
    _s.path.insert(0, _o.path.join(_o.path.dirname(_o.path.abspath(__file__)), "golden"))
    import make_golden as mg
    LHS, RHS, X0 = mg.nnls_inputs(100 + seed, k, q, -frac)
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    assert np.array_equal(X > 0, Xo > 0)
    assert rel(X, Xo) < 1e-9
    assert np.abs(Y - Yo).max() <= 1e-8 * max(1.0, np.abs(Yo).max())


@pytest.mark.parametrize("k,q,seed", [(100, 500, 8), (128, 150, 12)])
def test_nnls_bpp_wide_256_thread_form_at_k_le_128_matches_oracle(gpu, oracle, k, q, seed, monkeypatch):
    """SMK_NNLS_WIDE128=0: the 256-thread / 128 x 128 form of nnls_bpp_wide_kernel at 64 < k <= 128, where the other tests run the
    128-thread / 64 x 64 form."""
    monkeypatch.setenv("SMK_NNLS_WIDE128", "0")
    LHS, RHS, X0 = _nnls_problem(k, q, seed)
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    assert rc == 0
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    assert np.array_equal(X > 0, Xo > 0)
    assert rel(X, Xo) < 1e-10
    assert np.abs(Y - Yo).max() <= 1e-9 * max(1.0, np.abs(Yo).max())


def test_nnls_bpp_all_optimal_after_first_solve_keeps_tiny_values(gpu, oracle):
    """nnls.hpp:192,226-227: X,Y are zeroized only if some column was non-optimal."""
    k, q = 8, 16
    rng = np.random.default_rng(9)
    LHS = np.eye(k) * 2.0
    Xtrue = rng.random((k, q)) + 0.5
    Xtrue[0, 0] = 1e-13           # tiny but positive: survives only when no pivoting round runs
    RHS = LHS @ Xtrue
    rc, Xo, Yo = oracle.nnls_bpp(LHS, RHS, np.ones((k, q)))
    X, Y = gpu.nnls_bpp(LHS, RHS, np.ones((k, q)))
    assert Xo[0, 0] != 0.0 and abs(X[0, 0] - Xo[0, 0]) <= 1e-12 * abs(Xo[0, 0])
    # now make one column non-optimal -> the zeroize pass hits every column
    RHS2 = RHS.copy()
    RHS2[:, 1] = -1.0
    rc, Xo2, Yo2 = oracle.nnls_bpp(LHS, RHS2, np.ones((k, q)))
    X2, Y2 = gpu.nnls_bpp(LHS, RHS2, np.ones((k, q)))
    assert Xo2[0, 0] == 0.0 and X2[0, 0] == 0.0
    assert rel(X2, Xo2) < 1e-12


@pytest.mark.parametrize("k", [6, 100])
def test_nnls_bpp_non_hpd_fails(gpu, k):
    q = 10
    LHS = -np.eye(k)
    with pytest.raises(sk.SmallkError) as e:
        gpu.nnls_bpp(LHS, np.ones((k, q)), np.ones((k, q)))
    assert e.value.code == sk.FAILURE


# ---------------------------------------------------------------------------
# full solves, per-iteration traces
# ---------------------------------------------------------------------------
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
    ("BPP", 256, 256, 16, 30),        # C1 shape
    ("BPP", 300, 400, 40, 20),
    ("BPP", 500, 333, 64, 12),
    ("BPP", 400, 450, 100, 8),        # k > 64: one CTA per column
    ("BPP", 600, 520, 256, 5),
    ("MU", 150, 120, 8, 30),
    ("HALS", 200, 300, 12, 30),
    ("HALS", 260, 200, 40, 15),       # 2 rows per lane
    ("HALS", 300, 260, 72, 10),       # 4 rows per lane
    ("RANK2", 300, 200, 2, 30),
])
@pytest.mark.parametrize("prog", ["PG_RATIO", "DELTA_FNORM"])
def test_dense_trace_matches_oracle(gpu, oracle, alg, m, n, k, iters, prog):
    rng = np.random.default_rng(sum(map(ord, alg)) * 7919 + m * 31 + n * 17 + k)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        # HALS from an init with W0*H0 >> A clamps whole columns of W to zero in its first sweep; the epsilon
        # rule (nmf_solver_hals.hpp:103-115) then makes them identical, W'W singular, and the iteration
        # ill-posed (the reference built with two BLAS orders disagrees by 0.2 after 2 iterations). Scale H0
        # so that W0*H0 ~ A, as a real initialisation would.
        H0 *= 2.0 / k
    o = oracle.nmf_dense(A, W0, H0, alg=alg, prog=prog, tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm=alg, prog=prog, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    # HALS clamps entries to exactly 0; the projected-gradient norm then includes or drops the (large, positive)
    # gradient of an entry depending on whether it is 0 or 1e-17, so its metric is discontinuous in the last ulp
    # of the factors. Factors are still held to 1e-9; the HALS metric to 1e-5.
    mtol = 1e-5 if (alg == "HALS" and prog == "PG_RATIO") else REL_FACTOR
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))
        assert abs(metrics[i] - o["metrics"][i]) <= mtol * abs(o["metrics"][i]), (i, metrics[i], o["metrics"][i])


@pytest.mark.parametrize("m,n,k,iters", [(260, 200, 40, 15), (3000, 150, 72, 6)])
def test_hals_w_side_step_kernels_match_oracle(gpu, oracle, m, n, k, iters, monkeypatch):
    """SMK_HALS_SWEEP=0: the W-side sweep with one launch per step (the form used when the rows of W do not fit the shared memory
    of the grid) instead of the cooperative block sweep the other HALS tests run."""
    monkeypatch.setenv("SMK_HALS_SWEEP", "0")
    rng = np.random.default_rng(m + 3 * n + k)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n)) * (2.0 / k)
    o = oracle.nmf_dense(A, W0, H0, alg="HALS", prog="DELTA_FNORM", tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm="HALS", prog="DELTA_FNORM", tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))
        assert abs(metrics[i] - o["metrics"][i]) <= REL_FACTOR * abs(o["metrics"][i]), (i, metrics[i], o["metrics"][i])


@pytest.mark.parametrize("alg,k", [("BPP", 16), ("HALS", 16), ("MU", 8)])
def test_nmf_call_matches_oracle(gpu, oracle, alg, k):
    """The one-call interface (Nmf): stopping rule, iteration count, final normalisation."""
    m = n = 256
    rng = np.random.default_rng(77)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    kw = dict(tol=0.02, min_iter=5, max_iter=400)
    o = oracle.nmf_dense(A, W0, H0, alg=alg, normalize=True, **kw)
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm=alg, normalize=True, **kw)
    W, H, st = gpu.nmf(W0, H0, opts)
    assert st.iteration_count == o["iterations"]
    assert rel(W, o["W"]) < REL_FACTOR and rel(H, o["H"]) < REL_FACTOR
    assert np.allclose(np.linalg.norm(W, axis=0), 1.0, atol=1e-12)


def test_nmf_call_with_tma_sized_products_matches_oracle(gpu, oracle):
    """smk_nmf (the stop-tested loop as one CUDA graph) on a problem whose A-sized products are large enough for the TMA kernel:
    the tensor maps are kernel parameters of the captured launches."""
    m, n, k = 1500, 1200, 64
    rng = np.random.default_rng(5)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    kw = dict(tol=1e-9, min_iter=3, max_iter=10)            # ten iterations (the limit is a normal exit)
    o = oracle.nmf_dense(A, W0, H0, alg="BPP", normalize=True, **kw)
    gpu.load_dense(A)
    opts = sk.make_options(m, n, k, algorithm="BPP", normalize=True, **kw)
    W, H, st = gpu.nmf(W0, H0, opts)
    assert st.iteration_count == o["iterations"]
    assert rel(W, o["W"]) < REL_FACTOR and rel(H, o["H"]) < REL_FACTOR


# ---------------------------------------------------------------------------
# sparse
# ---------------------------------------------------------------------------
def _random_csc(m, n, density, seed, duplicates=False):
    import scipy.sparse as sp
    S = sp.random(m, n, density=density, random_state=seed, format="csc", data_rvs=np.random.default_rng(seed).random)
    S.sort_indices()
    return S


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [2, 5, 16, 40, 128])
def test_sparse_gemm_matches_oracle(gpu, oracle, variant, k):
    m, n = 300, 170
    S = _random_csc(m, n, 0.05, 11)
    rng = np.random.default_rng(variant * 10 + k)
    shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
    shapeC = (m, k) if variant < 2 else (k, n)
    B = rng.random(shapeB); C = rng.random(shapeC)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
        got = gpu.sparse_gemm(variant, alpha, B, beta, C)
        want = oracle.sparse_gemm(variant, alpha, (m, n), S.indptr, S.indices, S.data, B, beta, C)
        assert rel(got, want) < REL_PRIM


def _zipf_csc(m, n, per_col, seed):
    """Rows drawn log-uniformly (Zipf, s = 1) as in the C3 generator: a few rows take most of the entries."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows = np.minimum((m ** rng.random((n, per_col))).astype(np.int64) - 1, m - 1).clip(0)
    cols = np.repeat(np.arange(n), per_col)
    S = sp.csc_matrix((rng.random(n * per_col) + 0.1, (rows.ravel(), cols)), shape=(m, n))     # duplicates summed
    S.sort_indices()
    return S


@pytest.mark.parametrize("k", [96, 128, 200, 256])
def test_sparse_gemm_wide_gathers_on_skewed_rows_match_oracle(gpu, oracle, k):
    """The 256-bit gather kernel (a warp per segment, whole k-vector per gather; k >= 96) on a Zipf matrix: ragged k (dead lanes),
    partial batches (copies of the last entry with weight 0), hub rows cut into segments, all four variants, with and without beta."""
    m, n = 4000, 500
    S = _zipf_csc(m, n, 60, 17)
    rng = np.random.default_rng(k)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    for variant in (2, 3, 0, 1):
        shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
        shapeC = (m, k) if variant < 2 else (k, n)
        B = rng.random(shapeB); C = rng.random(shapeC)
        for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
            got = gpu.sparse_gemm(variant, alpha, B, beta, C)
            want = oracle.sparse_gemm(variant, alpha, (m, n), S.indptr, S.indices, S.data, B, beta, C)
            assert rel(got, want) < REL_PRIM


@pytest.mark.parametrize("k", [64, 96, 128, 256])
def test_sparse_gemm_in_k_slabs_matches_oracle(gpu, oracle, k, monkeypatch):
    """The k-slab SpMM (one launch per 32 rows of the dense operand; picked by operand size in production) forced on a small
    matrix: all four product variants, with and without beta, ragged columns and empty rows included."""
    monkeypatch.setenv("SMK_SPMM_SLAB", "2")
    m, n = 900, 700
    S = _zipf_csc(m, n, 40, 29)
    rng = np.random.default_rng(100 + k)
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    for variant in (0, 1, 2, 3):
        shapeB = {0: (n, k), 1: (k, n), 2: (k, m), 3: (m, k)}[variant]
        shapeC = (m, k) if variant < 2 else (k, n)
        B = rng.random(shapeB); C = rng.random(shapeC)
        for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
            got = gpu.sparse_gemm(variant, alpha, B, beta, C)
            want = oracle.sparse_gemm(variant, alpha, (m, n), S.indptr, S.indices, S.data, B, beta, C)
            assert rel(got, want) < REL_PRIM


def test_sparse_rank2_with_hub_rows_matches_oracle(gpu, oracle):
    """Rank-2 on a sparse matrix with short rows, rows of 65..512 entries and hubs of more than 512 (the three walks of the
    fused rank-2 iteration, csrc/rank2_fused.cu), against the oracle's trace."""
    import scipy.sparse as sp
    rng = np.random.default_rng(41)
    m, n, iters = 1500, 1400, 15
    A = sp.random(m, n, density=0.004, random_state=3, format="lil", data_rvs=rng.random)
    for r in (3, 700, 1499):                                  # hub rows
        cols = rng.choice(n, 900, replace=False); A[r, cols] = rng.random(900) + 0.1
    for c in (0, 650):                                        # hub columns
        rows = rng.choice(m, 800, replace=False); A[rows, c] = rng.random((800, 1)) + 0.1
    for r in range(20, 60):                                   # rows of 65..512 entries
        cols = rng.choice(n, 70 + 10 * (r - 20), replace=False); A[r, cols] = rng.random(len(cols)) + 0.1
    S = A.tocsc(); S.sort_indices()
    W0 = rng.random((m, 2)); H0 = rng.random((2, n))
    o = oracle.nmf_sparse((m, n), S.indptr, S.indices, S.data, W0, H0, alg="RANK2", tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    opts = sk.make_options(m, n, 2, algorithm="RANK2", tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))


def test_sparse_hals_k128_on_skewed_rows_matches_oracle(gpu, oracle):
    m, n, k, iters = 3000, 400, 128, 6
    S = _zipf_csc(m, n, 80, 5)
    rng = np.random.default_rng(6)
    W0 = rng.random((m, k)); H0 = rng.random((k, n)) * (S.sum() / m / n / (0.25 * k))
    o = oracle.nmf_sparse((m, n), S.indptr, S.indices, S.data, W0, H0, alg="HALS", tol=1e-12, min_iter=1, max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    opts = sk.make_options(m, n, k, algorithm="HALS", tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < REL_FACTOR, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < REL_FACTOR, (i, rel(Hs[i], o["H_trace"][i]))


@pytest.mark.parametrize("alg,k,iters", [("BPP", 10, 20), ("MU", 10, 30), ("RANK2", 2, 30), ("HALS", 10, 12)])
def test_sparse_trace_matches_oracle(gpu, oracle, alg, k, iters):
    m, n = 300, 200
    S = _random_csc(m, n, 0.3 if alg == "HALS" else 0.1, 3)
    rng = np.random.default_rng(5)
    W0 = rng.random((m, k)); H0 = rng.random((k, n))
    if alg == "HALS":
        H0 *= 0.6 / k          # W0*H0 ~ A (see test_dense_trace_matches_oracle)
    o = oracle.nmf_sparse((m, n), S.indptr, S.indices, S.data, W0, H0, alg=alg, tol=1e-12, min_iter=1,
                          max_iter=iters, trace=True)
    assert o["rc"] == 0
    gpu.load_csc((m, n), S.indptr, S.indices, S.data)
    opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
    metrics, Ws, Hs = _trace_gpu(gpu, W0, H0, opts, iters)
    tol = REL_FACTOR
    for i in range(iters):
        assert rel(Ws[i], o["W_trace"][i]) < tol, (i, rel(Ws[i], o["W_trace"][i]))
        assert rel(Hs[i], o["H_trace"][i]) < tol, (i, rel(Hs[i], o["H_trace"][i]))
