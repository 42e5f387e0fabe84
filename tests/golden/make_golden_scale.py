"""At-scale parity fixtures: the REFERENCE's own code (oracle/_ref/libsmallk_ref.so) run free on the
BASELINE configurations at the sizes SURVEY.md section 8(d) prescribes for parity:

  scale_c2_bpp   dense BPP, the FULL C2 matrix 20000 x 20000, k = 64, min_iter 1, tol 1e-12, 10 iterations
  scale_c5r_bpp  dense BPP, C5 reduced to 20000 x 10000, k = 256, 5 iterations
  scale_c3r_hals sparse HALS, C3 reduced to 100000 x 20000 at the same density (50 entries per column), k = 128, 20 iterations

Run in the dev container:  python tests/golden/make_golden_scale.py [name ...]
Per iterate the files keep the progress metric, ||W||_F, ||H||_F and 10^4 sampled entries of each factor (not the factors:
100 MB per iterate at C2). Inputs are regenerated from seeds by `scale_inputs()` (shared with tests/test_gpu_scale.py and
bench.py, whose C2 workload is the scale_c2_bpp matrix).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

NSAMPLE = 10000

# name -> (kind, alg, m, n, k, iterations, seeds (A, W0, H0))
SCALE_CASES = {
    "scale_c2_bpp":   ("dense", "BPP", 20000, 20000, 64, 10, (11, 12, 13)),
    "scale_c5r_bpp":  ("dense", "BPP", 20000, 10000, 256, 5, (41, 42, 43)),
    "scale_c3r_hals": ("sparse", "HALS", 100000, 20000, 128, 20, (21, 22, 23)),
}


def scale_inputs(name):
    """The inputs of a case. Dense A comes back as a C-ordered (n, m) array (== column-major m x n)."""
    import workloads
    kind, alg, m, n, k, iters, (sa, sw, sh) = SCALE_CASES[name]
    W0 = np.asfortranarray(np.random.default_rng(sw).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(sh).random((k, n)))
    out = dict(kind=kind, alg=alg, m=m, n=n, k=k, iters=iters, W0=W0, H0=H0, A_t=None, sp=None)
    if kind == "dense":
        out["A_t"] = workloads.dense_columns(m, 0, n, seed=sa)
    else:
        colp, rowi, val = workloads.tfidf_csc_numpy(m, n, 50, sa)
        out["sp"] = (colp, rowi, val)
        out["H0"] = np.asfortranarray(H0 * workloads.hals_h0_scale(val.sum(), m, n, k))
    return out


def sample_indices(name, m, n, k):
    rng = np.random.default_rng(sum(map(ord, name)))
    wi = rng.integers(0, m, NSAMPLE), rng.integers(0, k, NSAMPLE)
    hi = rng.integers(0, k, NSAMPLE), rng.integers(0, n, NSAMPLE)
    return wi, hi


def main(names):
    from oracle import Ref
    ref = Ref(blas_threads=os.cpu_count())
    for name in names:
        c = scale_inputs(name)
        m, n, k, iters = c["m"], c["n"], c["k"], c["iters"]
        t0 = time.time()
        if c["kind"] == "dense":
            o = ref.nmf_dense(c["A_t"].T, c["W0"], c["H0"], alg=c["alg"], tol=1e-12, min_iter=1, max_iter=iters,
                              trace=True, max_threads=os.cpu_count())
        else:
            colp, rowi, val = c["sp"]
            o = ref.nmf_sparse((m, n), colp, rowi, val, c["W0"], c["H0"], alg=c["alg"], tol=1e-12, min_iter=1,
                               max_iter=iters, trace=True, max_threads=os.cpu_count())
        assert o["rc"] == 0, o["rc"]
        (wr, wc), (hr, hc) = sample_indices(name, m, n, k)
        Wt, Ht = o["W_trace"], o["H_trace"]            # [iter][m][k], [iter][k][n]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            metrics=o["metrics"][:iters], iterations=o["iterations"],
            normW=np.array([np.linalg.norm(Wt[i]) for i in range(iters)]),
            normH=np.array([np.linalg.norm(Ht[i]) for i in range(iters)]),
            W_samples=np.stack([Wt[i][wr, wc] for i in range(iters)]),
            H_samples=np.stack([Ht[i][hr, hc] for i in range(iters)]),
            nnzW=np.array([np.count_nonzero(Wt[i]) for i in range(iters)]),
            nnzH=np.array([np.count_nonzero(Ht[i]) for i in range(iters)]))
        print(f"{name}: {o['iterations']} iterations in {time.time() - t0:.1f} s, metrics {o['metrics'][:iters]}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(SCALE_CASES))
