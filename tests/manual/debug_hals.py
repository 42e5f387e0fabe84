import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import smallk_b200 as sk
import workloads
from oracle import Oracle

def run(m, n, per_col, k, steps=2):
    colp, rowi, val = workloads.c3_tfidf_csc(m, n, per_col)
    W0 = np.asfortranarray(np.random.default_rng(22).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(23).random((k, n))) * (float(val.sum()) / m / n / (0.25 * k))
    res = {}
    for mode in (3, 2, 1, 0):
        os.environ["SMK_HALS_DEBUG"] = str(mode)
        ctx = sk.Context(0)
        ctx.load_csc((m, n), colp, rowi, val)
        opts = sk.make_options(m, n, k, algorithm="HALS", tol=1e-15, min_iter=1, max_iter=100, normalize=False)
        ctx.solver_begin(W0, H0, opts)
        for s in range(steps):
            ctx.solver_step(1)
            W, H = ctx.solver_get()
            if mode == 3:
                res[s] = (W.copy(), H.copy())
            else:
                Wr, Hr = res[s]
                dW = np.abs(W - Wr); dH = np.abs(H - Hr)
                dW[np.isnan(dW)] = np.inf; dH[np.isnan(dH)] = np.inf
                cw = np.argmax(dW.max(axis=0)); 
                print("mode", mode, "step", s, "nanW", np.isnan(W).sum(), "nanH", np.isnan(H).sum(), "max dW", dW.max(), "at col", cw, "row", np.argmax(dW[:, cw]),
                      "max dH", dH.max(), "ref W colmax", np.abs(Wr).max(axis=0)[8:16], "refH rowmax", np.abs(Hr).max(axis=1)[8:16], flush=True)
        ctx.close()

args = [int(a) for a in sys.argv[1:5]]
run(*args, steps=int(sys.argv[5]))
