"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/libsmallk_ref.so, i.e. the
unmodified sources under /root/reference compiled against oracle/shim/El.hpp) on seeded synthetic inputs.

Run in the dev container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures travel with the repo; nothing reads /root/reference at test time. Inputs are regenerated from the
seeds stored in each file by `golden_inputs()` below (shared with the tests), outputs are the reference's.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name -> (kind, alg, m, n, k, density, max_iter, tol, min_iter, normalize, seed)
CASES = {
    # BASELINE configs[0] / SURVEY C1: nmf CLI defaults otherwise (min_iter 5, PG_RATIO), tol 1e-4
    "c1_bpp_256_k16":     ("dense", "BPP", 256, 256, 16, None, 400, 1e-4, 5, True, 1),
    "bpp_300x400_k40":    ("dense", "BPP", 300, 400, 40, None, 25, 1e-12, 1, False, 2),
    "bpp_400x300_k64":    ("dense", "BPP", 400, 300, 64, None, 15, 1e-12, 1, False, 3),
    "hals_256_k16":       ("dense", "HALS", 256, 256, 16, None, 60, 1e-12, 1, False, 4),
    "mu_150x120_k8":      ("dense", "MU", 150, 120, 8, None, 60, 1e-12, 1, True, 5),
    "rank2_300x200":      ("dense", "RANK2", 300, 200, 2, None, 40, 1e-12, 1, False, 6),
    "sp_bpp_300x200_k10": ("sparse", "BPP", 300, 200, 10, 0.1, 25, 1e-12, 1, False, 7),
    "sp_mu_300x200_k10":  ("sparse", "MU", 300, 200, 10, 0.1, 40, 1e-12, 1, False, 8),
    "sp_rank2_400x300":   ("sparse", "RANK2", 400, 300, 2, 0.05, 40, 1e-12, 1, True, 9),
    "sp_hals_300x200_k8": ("sparse", "HALS", 300, 200, 8, 0.3, 3, 1e-12, 1, False, 10),
    # k > 64: the one-CTA-per-column NNLS kernel (SURVEY C5 is k = 256)
    "bpp_500x420_k96":    ("dense", "BPP", 500, 420, 96, None, 10, 1e-12, 1, False, 11),
    "bpp_420x400_k160":   ("dense", "BPP", 420, 400, 160, None, 6, 1e-12, 1, False, 12),
    "sp_bpp_900x800_k72": ("sparse", "BPP", 900, 800, 72, 0.15, 8, 1e-12, 1, False, 13),
}


def round6(a):
    """What a matrix looks like after matrixgen wrote it with %.6e and the nmf CLI read it back
    (common/include/delimited_file.hpp:62-63)."""
    return np.array([float("%.6e" % v) for v in a.ravel()]).reshape(a.shape)


def golden_inputs(name):
    kind, alg, m, n, k, density, max_iter, tol, min_iter, normalize, seed = CASES[name]
    rng = np.random.default_rng(1000 + seed)
    if kind == "dense":
        A = round6(rng.random((m, n)))
        sp = None
    else:
        import scipy.sparse as sps
        S = sps.random(m, n, density=density, random_state=np.random.RandomState(seed), format="csc",
                       data_rvs=np.random.RandomState(seed + 50).random_sample)
        S.sort_indices()
        A = None
        sp = (S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data.astype(np.float64))
    W0 = rng.random((m, k))
    H0 = rng.random((k, n))
    return dict(kind=kind, alg=alg, m=m, n=n, k=k, max_iter=max_iter, tol=tol, min_iter=min_iter,
                normalize=normalize, A=A, sp=sp, W0=W0, H0=H0)


def nnls_inputs(seed, k, q, shift=0.35, cold=False):
    """shift sets how much of the unconstrained solution is negative: 0.35 leaves ~10 % of the entries passive,
    0.01 leaves most of them passive (|P| > k/2: the complement path of the GPU kernels)."""
    rng = np.random.default_rng(seed)
    W = rng.random((4 * k, k))
    A = rng.random((4 * k, q))
    LHS = W.T @ W
    if shift < 0:
        # planted mostly-positive solution: |P| is close to (1 + shift) * k, X0 is a cold start
        Xt = rng.random((k, q)) * (rng.random((k, q)) > -shift)
        RHS = LHS @ Xt - 1e-3 * rng.random((k, q)) * np.abs(LHS @ Xt).mean()
        X0 = rng.random((k, q)) * (rng.random((k, q)) > 0.5)
        return LHS, RHS, X0
    RHS = W.T @ A - shift * rng.random((k, q)) * np.abs(W.T @ A).mean()
    X0 = rng.random((k, q)) * (rng.random((k, q)) > 0.3)
    if cold:
        # all-nonpositive initial guess: the passive set starts with NO bit set. IsEmpty(passive_set) in
        # BppSolveNormalEqNoGroup (nmf_solver_bpp.hpp:173) tests for a zero-SIZED BitMatrix (bit_matrix_ops.cpp:24-27), not
        # for an all-zero one, so the reference takes the per-column path and leaves X = 0 for the first dual evaluation
        X0 = -X0
    return LHS, RHS, X0


NNLS_CASES = {"nnls_k16_q64": (21, 16, 64), "nnls_k40_q90": (22, 40, 90), "nnls_k64_q120": (23, 64, 120),
              "nnls_k100_q150": (24, 100, 150), "nnls_k200_q90": (25, 200, 90), "nnls_k256_q64": (26, 256, 64),
              "nnls_k60_q100_dense": (27, 60, 100, -0.15), "nnls_k200_q80_dense": (28, 200, 80, -0.2),
              "nnls_k256_q60_dense": (29, 256, 60, -0.1), "nnls_k250_q40_half": (30, 250, 40, -0.5),
              "nnls_k48_q80_cold": (31, 48, 80, 0.35, True), "nnls_k130_q50_cold": (32, 130, 50, 0.35, True)}


# ---- NNLS cases in which UpdatePassiveSet's BACKUP rule fires (common/src/nnls.cpp:64-72) --------------------------
# Block pivoting falls back to single-variable toggles only when full exchanges cycle, which random well-conditioned
# problems never do. A small system with strongly correlated columns does (found by search: s = 5, seed 6 -> 18 firings in
# 2000 columns); it is embedded at rows [at, at + s) of an otherwise diagonal k x k problem, so that the row the rule
# toggles is reported by BitMatrix::MaxRowIndex from word 0 / the partial last word (correct) or from a full word > 0
# (common/src/bit_matrix.cpp:459-467: 32 rows too low -> the wrong row is toggled, the column cycles until MAX_ITER = 5k
# and the REFERENCE returns failure). name -> (k, at, s, seed); the fixture records the reference's rc, X, Y.
# k > 256 (the any-k fallback kernel): the same 2000-column problem cut down to 42 columns around the four in which the rule
# fires (481, 500, 750, 958 for s = 5, seed = 6) — a passive-set solve is a 300 x 300 Cholesky there, on the CPU and on the GPU.
_BIGK_COLS = tuple(range(470, 510)) + (750, 958)
BACKUP_CASES = {"nnls_backup_k64_word0": (64, 27, 5, 6), "nnls_backup_k48_partial": (48, 43, 5, 6),
                "nnls_backup_k100_word0": (100, 10, 5, 6), "nnls_backup_k200_word0": (200, 20, 6, 9),
                "nnls_backup_k64_defect": (64, 59, 5, 6), "nnls_backup_k100_defect": (100, 60, 5, 6),
                "nnls_backup_k300_word0": (300, 10, 5, 6, _BIGK_COLS), "nnls_backup_k300_partial": (300, 291, 5, 6, _BIGK_COLS)}
# (no k = 300 defect case: the four columns then cycle through 5k = 1500 solves of a 300 x 300 system each, 40 s on the CPU;
# the defective MaxRowIndex is one function shared by every k > 64 and covered at k = 100)


def backup_inputs(k, at, s, seed, cols=None, q=2000):
    rng = np.random.default_rng(seed * 100 + s)
    rows = s + int(rng.integers(0, 3))
    Ws = rng.standard_normal((rows, s)) + 2.0 * rng.standard_normal((rows, 1))
    S = Ws.T @ Ws + 1e-6 * np.eye(s)
    Rs = rng.standard_normal((s, q)) * 3
    Xs = (rng.random((s, q)) > 0.5) * rng.random((s, q))
    rng = np.random.default_rng(seed + 7)
    LHS = np.diag(1.0 + rng.random(k))
    LHS[at:at + s, at:at + s] = S
    RHS = 0.5 + rng.random((k, q))
    RHS[at:at + s] = Rs
    X0 = rng.random((k, q))
    X0[at:at + s] = Xs
    if cols is not None:
        cols = np.asarray(cols)
        RHS, X0 = np.ascontiguousarray(RHS[:, cols]), np.ascontiguousarray(X0[:, cols])
    return LHS, RHS, X0


def main():
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import Ref
    ref = Ref()
    print("reference build BLAS:", ref.blas_backend())
    only = set(sys.argv[1:])
    for name in CASES:
        if only and name not in only:
            continue
        g = golden_inputs(name)
        kw = dict(alg=g["alg"], tol=g["tol"], min_iter=g["min_iter"], max_iter=g["max_iter"], normalize=g["normalize"],
                  trace=True, max_threads=2)
        if g["kind"] == "dense":
            r = ref.nmf_dense(g["A"], g["W0"], g["H0"], **kw)
        else:
            r = ref.nmf_sparse((g["m"], g["n"]), *g["sp"], g["W0"], g["H0"], **kw)
        assert r["rc"] == 0, (name, r["rc"])
        it = r["iterations"]
        # traces hold the un-normalised iterates at the snapshots the estimator saw; keep a few of them
        seen = np.flatnonzero(~np.isnan(r["metrics"]))       # iterations at which the estimator was called
        keep = sorted(set(seen[:3].tolist() + [int(seen[-1])]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), iterations=it, metrics=r["metrics"],
                            W=r["W"], H=r["H"], snap_iters=np.array(keep),
                            W_snaps=r["W_trace"][keep], H_snaps=r["H_trace"][keep])
        print(f"{name}: iterations={it} last metric={r['metrics'][min(it, g['max_iter']) - 1]:.6g}")
    for name, args in NNLS_CASES.items():
        if only and name not in only:
            continue
        LHS, RHS, X0 = nnls_inputs(*args)
        rc, X, Y = ref.nnls_bpp(LHS, RHS, X0)
        assert rc == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, Y=Y)
        print(f"{name}: passive density {np.mean(X > 0):.3f}")
    for name, args in BACKUP_CASES.items():
        if only and name not in only:
            continue
        LHS, RHS, X0 = backup_inputs(*args)
        rc, X, Y = ref.nnls_bpp(LHS, RHS, X0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rc=rc, X=X, Y=Y)
        print(f"{name}: reference rc {rc}")


if __name__ == "__main__":
    main()
