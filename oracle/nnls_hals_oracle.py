"""TEST INFRASTRUCTURE — not part of the product.

NumPy restatement of NnlsHals (common/include/nnls.hpp:249-316), the H-only HALS solve against a fixed W that
HierNmf2WithFlat's flat step uses (clust_flat_generic.hpp:33-74):

    W'W, W'A once                                         nnls.hpp:272-273
    per iteration: UpdateH_Hals (row by row, updated rows feed the later ones; NaN / negative -> 0)
                                                          nmf_solver_hals.hpp:33-62
                   gradH = W'W H - W'A                    nnls.hpp:283-284
                   pg = ProjectedGradientNorm(gradH, H)   projected_gradient.hpp:94-121
                   iteration 1 stores pg0; later: stop when pg < tol * pg0, then NormalizeAndScale(W, H)
                                                          nnls.hpp:287-309, normalize.hpp:118-138
Parity unpinned on its own (no reference wrapper is built for this function); the reference's NnlsHals is exercised
end to end through the flat-clustering fixtures of tests/golden (made by oracle/_ref).
"""
import numpy as np


def nnls_hals(A, W, H0, tol, max_iter):
    """Returns (success, W, H, iterations)."""
    W = np.array(W, dtype=np.float64, order="F")
    H = np.array(H0, dtype=np.float64, order="F")
    WtW = W.T @ W
    WtA = W.T @ A
    k = W.shape[1]
    pg0 = 0.0
    for i in range(max_iter):
        for r in range(k):
            h = H[r, :] + (WtA[r, :] - WtW[r, :] @ H) / WtW[r, r]
            h[np.isnan(h) | (h < 0)] = 0.0
            H[r, :] = h
        grad = WtW @ H - WtA
        pg = np.sqrt(np.sum(grad[(grad < 0) | (H > 0)] ** 2))
        if i == 0:
            pg0 = pg
            continue
        if pg < tol * pg0:
            norms = np.linalg.norm(W, axis=0)
            return True, W / norms, H * norms[:, None], i + 1
    return False, W, H, max_iter
