"""A/B of the two forms of the 64 < k <= 128 NNLS kernel (SMK_NNLS_WIDE128=0: 256 threads / 128 x 128 triangle; default: 128 threads /
64 x 64 triangle, six columns in flight per SM) through smk_nnls_bpp: wall time of the call (copies included, identical in both)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smallk_b200 as sk


def problem(k, q, seed, keep):
    rng = np.random.default_rng(seed)
    W = rng.random((4 * k, k))
    A = rng.random((4 * k, q))
    LHS = W.T @ W
    RHS = W.T @ A - 0.35 * rng.random((k, q)) * np.abs(W.T @ A).mean()
    X0 = rng.random((k, q)) * (rng.random((k, q)) < keep)
    return LHS, RHS, X0


def main():
    ctx = sk.Context(0)
    for k, q, keep in ((128, 60000, 0.1), (128, 60000, 0.9), (100, 60000, 0.5)):
        LHS, RHS, X0 = problem(k, q, k + q, keep)
        out = {}
        for mode in ("1", "0", "1", "0"):
            os.environ["SMK_NNLS_WIDE128"] = mode
            t0 = time.perf_counter()
            X, Y = ctx.nnls_bpp(LHS, RHS, X0)
            out.setdefault(mode, []).append(time.perf_counter() - t0)
            out["X" + mode] = X
        same = bool(np.array_equal(out["X1"] > 0, out["X0"] > 0)) and float(np.abs(out["X1"] - out["X0"]).max()) < 1e-9 * float(np.abs(out["X0"]).max())
        print(f"k={k} q={q} warm-start density {keep}: 128-thread form {min(out['1']) * 1e3:.1f} ms, 256-thread form {min(out['0']) * 1e3:.1f} ms, same result {same}", flush=True)
    os.environ.pop("SMK_NNLS_WIDE128", None)
    ctx.close()


if __name__ == "__main__":
    main()
