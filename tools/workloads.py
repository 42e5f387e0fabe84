"""Synthetic workloads of BASELINE.json at full size: C3 (tf-idf-like CSC 1e6 x 2e5, 1e8 nonzeros) and
C4 (power-law graph, ~320k nodes / ~2M edges, symmetric CSC). Used by bench.py and tools/measure_*.py only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from graphgen import community_graph, powerlaw_graph   # noqa: E402,F401


def c4_graph(n=320000, edges=2000000, seed=31, communities=100):
    """DBLP-scale co-authorship-like graph: power-law degrees with planted communities (a graph without community
    structure gives HierNMF2 nothing to split: it stops after a handful of factorizations)."""
    return community_graph(n, edges, communities, seed)


def c3_tfidf_csc(m=1000000, n=200000, nnz_per_col=500, seed=21, device="cuda"):
    """CSC with ~nnz_per_col entries per column, row popularity ~ Zipf(1) (inverse-CDF sampling r = m^u), duplicates
    removed (the draws per column are oversampled so that the
    distinct rows per column average nnz_per_col: 1e8 stored entries at full size), rows ascending inside a column, values (1 + ln tf) * ln(n / df) with tf ~ Geometric(0.5), columns scaled to
    unit 2-norm (preprocessor/src/preprocess.cpp:193-230). Built on the GPU with torch, returned as host numpy arrays."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    want = nnz_per_col
    nnz_per_col = draws_per_column(m, want)       # oversampled so that ~want DISTINCT rows per column survive the dedup
    chunks = 8
    keys = []
    per = n // chunks
    for ch in range(chunks):
        c0, c1 = ch * per, (n if ch == chunks - 1 else (ch + 1) * per)
        cnt = (c1 - c0) * nnz_per_col
        u = torch.rand(cnt, generator=g, device=device, dtype=torch.float64)
        rows = torch.clamp((torch.pow(torch.tensor(float(m), dtype=torch.float64, device=device), u) - 1.0).to(torch.int64), 0, m - 1)
        cols = torch.arange(c0, c1, device=device, dtype=torch.int64).repeat_interleave(nnz_per_col)
        key = torch.unique(cols * m + rows)          # sorted: by column, then row; duplicates dropped
        keys.append(key)
        del u, rows, cols
    key = torch.cat(keys)
    del keys
    cols = key // m
    rows = (key - cols * m).to(torch.int32)
    nnz = key.numel()
    del key
    colcount = torch.bincount(cols, minlength=n)
    colptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    colptr[1:] = torch.cumsum(colcount, 0)
    df = torch.bincount(rows.to(torch.int64), minlength=m).clamp(min=1).to(torch.float64)
    tf = torch.floor(torch.log(torch.rand(nnz, generator=g, device=device, dtype=torch.float64)) / np.log(0.5)) + 1.0
    val = (1.0 + torch.log(tf)) * torch.log(float(n) / df[rows.to(torch.int64)]).clamp(min=1e-3)
    del tf, df
    sq = torch.zeros(n, dtype=torch.float64, device=device).index_add_(0, cols, val * val)
    val = val / torch.sqrt(sq)[cols]
    out = (colptr.to(torch.int32).cpu().numpy().astype(np.uint32), rows.cpu().numpy().astype(np.uint32), val.cpu().numpy())
    del cols, rows, val, sq, colptr
    torch.cuda.empty_cache()
    return out


def draws_per_column(m, want, seed=5):
    """How many Zipf(1) row draws (r = m^u) give `want` distinct rows per column on average: bisection on a fixed
    512-column sample (668 for m = 1e6, want = 500: a quarter of the draws land on rows the column already has)."""
    rng = np.random.default_rng(seed)
    u = rng.random((512, 2 * want + 16))

    def distinct(d):
        rows = np.clip((np.power(float(m), u[:, :d]) - 1.0).astype(np.int64), 0, m - 1)
        rows.sort(axis=1)
        return float((np.diff(rows, axis=1) != 0).sum(axis=1).mean() + 1.0)

    lo, hi = want, 2 * want + 16
    while lo < hi:
        mid = (lo + hi) // 2
        if distinct(mid) < want:
            lo = mid + 1
        else:
            hi = mid
    return lo


def tfidf_csc_numpy(m, n, nnz_per_col, seed):
    """The C3 generator in NumPy (bit-reproducible on any host, unlike the torch/philox one above): same distribution,
    used for the reduced-size parity fixtures (tests/golden/make_golden_scale.py) and the CPU baselines."""
    rng = np.random.default_rng(seed)
    d = draws_per_column(m, nnz_per_col)
    u = rng.random((n, d))
    rows = np.clip((np.power(float(m), u) - 1.0).astype(np.int64), 0, m - 1)
    del u
    key = np.unique((np.arange(n, dtype=np.int64)[:, None] * m + rows).ravel())
    cols = key // m
    rows = (key - cols * m).astype(np.uint32)
    nnz = key.size
    colptr = np.zeros(n + 1, dtype=np.uint32)
    colptr[1:] = np.cumsum(np.bincount(cols, minlength=n))
    df = np.maximum(np.bincount(rows, minlength=m), 1).astype(np.float64)
    tf = np.floor(np.log(rng.random(nnz)) / np.log(0.5)) + 1.0
    val = (1.0 + np.log(tf)) * np.maximum(np.log(float(n) / df[rows]), 1e-3)
    sq = np.bincount(cols, weights=val * val, minlength=n)
    val = val / np.sqrt(sq)[cols]
    return colptr, rows, val


def dense_columns(m, c0, c1, seed=11):
    """Columns [c0, c1) of the dense U[0,1) workload matrix A (m x n, column-major): entry (i, j) is draw j * m + i of
    PCG64(seed), so any column block can be generated on its own (PCG64.advance) and every rank / the reference arm / the
    parity fixtures see the same matrix. Returned as a C-ordered (c1 - c0, m) array == column-major m x (c1 - c0)."""
    bg = np.random.PCG64(seed)
    bg.advance(int(c0) * int(m))
    return np.random.Generator(bg).random((int(c1) - int(c0), int(m)))


def hals_h0_scale(val_sum, m, n, k):
    """H0 factor so that mean(W0 * H0) = mean(A) for U[0,1) initial factors (HALS note)."""
    return float(val_sum) / m / n / (0.25 * k)
