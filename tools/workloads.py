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
    removed, rows ascending inside a column, values (1 + ln tf) * ln(n / df) with tf ~ Geometric(0.5), columns scaled to
    unit 2-norm (preprocessor/src/preprocess.cpp:193-230). Built on the GPU with torch, returned as host numpy arrays."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n * nnz_per_col
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
