"""The CUDA path against the fixtures produced by the reference's own code (tests/golden/)."""
import os
import sys

import numpy as np
import pytest

import smallk_b200 as sk

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg      # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_gpu_reproduces_reference_fixture(gpu, name):
    g = mg.golden_inputs(name)
    z = np.load(os.path.join(GOLD, name + ".npz"))
    if g["kind"] == "dense":
        gpu.load_dense(g["A"])
    else:
        gpu.load_csc((g["m"], g["n"]), *g["sp"])
    tol = 1e-5 if "hals" in name else TOL
    # per-iteration metrics and snapshots through the solver seam
    opts = sk.make_options(g["m"], g["n"], g["k"], algorithm=g["alg"], tol=g["tol"], min_iter=g["min_iter"],
                           max_iter=g["max_iter"], normalize=False)
    gpu.solver_begin(g["W0"], g["H0"], opts)
    snaps = {int(it): j for j, it in enumerate(z["snap_iters"])}
    for it in range(int(z["iterations"])):
        gpu.solver_step(1)
        if not np.isnan(z["metrics"][it]):
            metric = gpu.solver_progress()
            assert abs(metric - z["metrics"][it]) <= tol * abs(z["metrics"][it]), (it, metric, z["metrics"][it])
        if it in snaps:
            W, H = gpu.solver_get()
            assert rel(W, z["W_snaps"][snaps[it]]) < tol, it
            assert rel(H, z["H_snaps"][snaps[it]]) < tol, it
    # the one-call interface: iteration count, final (possibly normalised) factors
    opts = sk.make_options(g["m"], g["n"], g["k"], algorithm=g["alg"], tol=g["tol"], min_iter=g["min_iter"],
                           max_iter=g["max_iter"], normalize=g["normalize"])
    W, H, st = gpu.nmf(g["W0"], g["H0"], opts)
    assert st.iteration_count == int(z["iterations"])
    assert rel(W, z["W"]) < tol and rel(H, z["H"]) < tol


@pytest.mark.parametrize("name", sorted(mg.NNLS_CASES))
def test_gpu_nnls_reproduces_reference_fixture(gpu, name):
    LHS, RHS, X0 = mg.nnls_inputs(*mg.NNLS_CASES[name])
    k, q = RHS.shape
    z = np.load(os.path.join(GOLD, name + ".npz"))
    X, Y = gpu.nnls_bpp(LHS, RHS, X0)
    assert np.array_equal(X > 0, z["X"] > 0)
    assert rel(X, z["X"]) < 1e-10


@pytest.mark.parametrize("name", sorted(mg.BACKUP_CASES))
def test_gpu_backup_rule_fires_as_in_the_reference(gpu, oracle, name):
    """Crafted NNLS problems in which UpdatePassiveSet's backup rule fires (common/src/nnls.cpp:64-72). The GPU kernels must
    fire it exactly as often as the CPU restatement (itself pinned to the reference's rc / X / Y by the fixture), give the
    reference's solution where the reference succeeds, and FAIL where BitMatrix::MaxRowIndex's off-by-one-word defect
    (common/src/bit_matrix.cpp:459-467) makes the reference cycle to MAX_ITER (k > 32, row in a full word > 0)."""
    import ctypes
    LHS, RHS, X0 = mg.backup_inputs(*mg.BACKUP_CASES[name])
    z = np.load(os.path.join(GOLD, name + ".npz"))
    oracle.lib.orc_backup_count.restype = ctypes.c_double
    oracle.lib.orc_stats_reset()
    rc_o, Xo, Yo = oracle.nnls_bpp(LHS, RHS, X0)
    fired = int(oracle.lib.orc_backup_count())
    assert fired > 0 and rc_o == int(z["rc"])
    if int(z["rc"]) == 0:
        X, Y = gpu.nnls_bpp(LHS, RHS, X0)
        assert np.array_equal(X > 0, z["X"] > 0)
        assert rel(X, z["X"]) < 1e-10
        assert np.abs(Y - z["Y"]).max() < 1e-9 * max(1.0, np.abs(z["Y"]).max())
    else:
        with pytest.raises(sk.SmallkError) as e:
            gpu.nnls_bpp(LHS, RHS, X0)
        assert e.value.code == sk.FAILURE
    assert gpu.nnls_backup_count() == fired
