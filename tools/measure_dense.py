"""Per-iteration device time of a dense NMF workload on one GPU (C5: 100000 x 50000, k = 256, BPP; also MU / HALS).
usage: measure_dense.py m n k ALG [iters]   — prints one JSON line; not the bench contract (bench.py is)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallk_b200 as sk          # noqa: E402


def main():
    import torch
    m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    alg = sys.argv[4] if len(sys.argv) > 4 else "BPP"
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    dev = torch.device("cuda", 0)
    A = torch.empty((n, m), dtype=torch.float64, device=dev)          # row-major [n][m] == column-major m x n
    g = torch.Generator(device=dev)
    for b0 in range(0, n, 500):
        g.manual_seed(41 * 1000003 + b0)
        A[b0:b0 + 500] = torch.rand((min(500, n - b0), m), dtype=torch.float64, device=dev, generator=g)
    ctx = sk.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.load_dense_device(A.data_ptr(), m, m, n)
    W0 = np.asfortranarray(np.random.default_rng(42).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(43).random((k, n)))
    if alg == "HALS":
        H0 *= 2.0 / k
    opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-15, min_iter=1, max_iter=1000, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    for _ in range(3):
        ctx.solver_step(1)
        ctx.solver_progress()
    if os.environ.get("SMK_PHASES"):
        ctx.phase_report()
    times = []
    for _ in range(iters):
        ctx.solver_step(1)
        times.append(ctx.last_step()[0])
    phases = {nm: v / iters for nm, v in ctx.phase_report().items()} if os.environ.get("SMK_PHASES") else None
    metric = ctx.solver_progress()
    ms = float(np.median(times))
    F = 4.0 * k * m * n + (6.0 * k * k * n + 4.0 * k * k * m if alg == "BPP" else 6.0 * k * k * (m + n))
    print(json.dumps({"workload": f"dense {alg} {m}x{n} k={k}", "ms_per_iter": ms, "iters_per_s": 1000.0 / ms, "flop_per_iter": F,
                      "achieved_TFLOPs": F / ms * 1e-9, "launches": ctx.last_step()[1], "times_ms": times, "metric": metric, "phases_ms": phases}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
