"""Device time per Rank2 iteration on the root matrix of the C4 graph (no host synchronisation between iterations):
what the kernels cost with their operands hot in L2, next to the per-iteration wall time of a converging smk_nmf run."""
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

for n, edges in ((40000, 250000), (320000, 2000000)):
    colp, rowi, val = workloads.c4_graph(n, edges)
    ctx = sk.Context(0)
    ctx.load_csc((n, n), colp, rowi, val)
    rng = np.random.default_rng(1)
    W0 = rng.random((n, 2)); H0 = rng.random((2, n))
    opts = sk.make_options(n, n, 2, algorithm="RANK2", tol=1e-15, min_iter=1, max_iter=100000, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    ctx.solver_step(20)
    ctx.solver_step(200)
    ms, launches = ctx.last_step()
    t = time.time()
    for _ in range(200):
        ctx.solver_step(1)
        ctx.solver_progress()
    wall = (time.time() - t) / 200
    nnz = int(colp[-1])
    B = 2 * (12 * nnz + 4 * (n + 1)) + 128 * (n + n)
    print(json.dumps({"workload": "rank-2 iteration, C4 root", "nodes": n, "nnz": nnz, "fused": os.environ.get("SMK_RANK2_FUSED", "1"),
                      "device_us_per_iter": ms / 200 * 1e3, "launches_per_iter": launches / 200, "wall_us_per_iter_with_progress_sync": wall * 1e6,
                      "algorithmic_MB": B * 1e-6, "achieved_GBs": B / (ms / 200) * 1e-6}), flush=True)
    ctx.close()
