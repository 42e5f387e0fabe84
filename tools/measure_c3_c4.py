"""One-off measurement of the sparse workloads on the GPU box: C4 (hierclust) end to end with the host-side time
breakdown, C3 (sparse HALS) per-iteration device time. Prints JSON lines; not the bench contract (bench.py is)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import smallk_b200 as sk          # noqa: E402
import workloads                  # noqa: E402


def c4(n=320000, edges=2000000, clusters=64):
    t = time.time()
    colp, rowi, val = workloads.c4_graph(n, edges)
    gen_s = time.time() - t
    out = sk.hierclust(csc=(colp, rowi, val), shape=(n, n), num_clusters=clusters, tol=1e-4, min_iter=5, max_iter=5000, seed=32)
    line = {"workload": "C4 hierclust", "nodes": n, "nnz": int(colp[-1]), "clusters": clusters, "rc": out["rc"], "gen_s": gen_s,
            "elapsed_s": out["elapsed_s"], "nmf_count": out["nmf_count"], "rank2_iterations": out["iterations"],
            "iters_per_s": out["iterations"] / out["elapsed_s"], "profile": out["profile"], "outliers": out["n_outliers"],
            "leaf_sizes": out["doc_count"][out["is_leaf"] == 1].tolist()}
    print(json.dumps(line), flush=True)


def c3(m=1000000, n=200000, per_col=500, k=128, iters=5):
    import torch
    t = time.time()
    colp, rowi, val = workloads.c3_tfidf_csc(m, n, per_col)
    gen_s = time.time() - t
    ctx = sk.Context(0)
    t = time.time()
    ctx.load_csc((m, n), colp, rowi, val)
    ctx.synchronize()
    load_s = time.time() - t
    W0 = np.asfortranarray(np.random.default_rng(22).random((m, k)))
    # H0 scaled so that mean(W0*H0) = mean(A): from W0*H0 >> A, HALS clamps whole factors to zero in its first sweep
    H0 = np.asfortranarray(np.random.default_rng(23).random((k, n))) * (float(val.sum()) / m / n / (0.25 * k))
    opts = sk.make_options(m, n, k, algorithm="HALS", tol=1e-15, min_iter=1, max_iter=100, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    ctx.solver_step(1)
    ctx.solver_progress()
    ctx.solver_step(1)
    times = []
    for _ in range(iters):
        ctx.solver_step(1)
        times.append(ctx.last_step()[0])
    metric = ctx.solver_progress()
    prod_ms = {"WtA": ctx.time_product(0, 5), "HAt": ctx.time_product(1, 5)}
    nnz = int(colp[-1])
    B = 2 * (12 * nnz + 4 * (n + 1)) + 64 * k * (m + n)
    ms = float(np.median(times))
    print(json.dumps({"workload": "C3 sparse HALS", "m": m, "n": n, "nnz": nnz, "k": k, "gen_s": gen_s, "load_s": load_s,
                      "ms_per_iter": ms, "iters_per_s": 1000.0 / ms, "algorithmic_GB": B * 1e-9,
                      "achieved_GBs": B / ms * 1e-6, "launches": ctx.last_step()[1], "metric": metric,
                      "spmm_ms": prod_ms, "spmm_gather_TBs": {kk: nnz * k * 8 / v * 1e-9 for kk, v in prod_ms.items()}}), flush=True)
    ctx.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("c4", "all"):
        c4()
    if which in ("c4small",):
        c4(40000, 250000, 16)
    if which in ("c3", "all"):
        c3()
    if which in ("c3small",):
        c3(100000, 20000, 500, 128)
