"""compute_priority (hierclust/include/clust_hier_util.hpp:105-173) of the host driver, no GPU needed: the streaming evaluation
the tree driver uses against the plain full-length one, and both against the reference's own function where oracle/_ref
is built. All three must agree to the last bit (the score is a sequence of sequential sums; nothing is reordered)."""
import ctypes
import os

import numpy as np
import pytest

import smallk_b200 as sk

dp = ctypes.POINTER(ctypes.c_double)


def _host():
    if not os.path.exists(sk.HOST_LIB_PATH):
        pytest.skip("host library not built")
    lib = ctypes.CDLL(sk.HOST_LIB_PATH)
    lib.smkh_compute_priority.restype = ctypes.c_double
    lib.smkh_compute_priority_plain.restype = ctypes.c_double
    lib.smkh_compute_priority_rows.restype = ctypes.c_double
    lib.smkh_compute_priority_rows2.restype = ctypes.c_double
    return lib


def _cases():
    rng = np.random.default_rng(7)
    for n, npart, nchild, mode in [(50, 50, 50, "full"), (400, 120, 60, "sub"), (3000, 900, 300, "sub"), (3000, 900, 300, "ties"),
                                   (2500, 2, 2, "sub"), (2500, 1, 1, "sub"), (6000, 2500, 1200, "stray"), (9000, 4000, 3000, "sub"),
                                   (700, 700, 0, "sub"), (700, 300, 150, "negative")]:
        P = np.zeros(n); C = np.zeros((n, 2), order="F")
        rows = np.sort(rng.choice(n, npart, replace=False))
        P[rows] = rng.random(npart)
        sub = rng.choice(rows, nchild, replace=False) if nchild else rows[:0]
        C[sub, 0] = rng.random(nchild); C[sub, 1] = rng.random(nchild)
        C[sub[: nchild // 3], 0] = 0.0                       # rows present in one child only
        C[sub[nchild // 3: nchild // 2], 1] = 0.0
        if mode == "ties":
            P[rows[::3]] = 0.25; C[sub[::2], 0] = 0.5; C[sub[::4], 1] = 0.125
        if mode == "stray":                                    # child entries on rows where the parent is exactly zero
            stray = rng.choice(np.setdiff1d(np.arange(n), rows), 40, replace=False)
            C[stray[:25], 0] = rng.random(25); C[stray[15:], 1] = rng.random(25)
        if mode == "negative":
            P[rows[5]] = -0.3; C[sub[2], 1] = -1e-3
        yield n, mode, P, C


def test_streaming_priority_equals_plain_evaluation_bit_for_bit():
    lib = _host()
    for n, mode, P, C in _cases():
        a = lib.smkh_compute_priority(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n)
        b = lib.smkh_compute_priority_plain(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n)
        assert a == b, (n, mode, a, b)


def test_row_list_priority_equals_plain_evaluation_bit_for_bit():
    """The evaluation the tree driver uses (compute_priority_rows: the caller names the rows where the child factors can be
    non-zero; everything else in O(rows) + one lean pass for the ideal sum) against the plain one, also with a row list that
    holds more rows than the factors touch and with an m far larger than the node."""
    lib = _host()
    up = ctypes.POINTER(ctypes.c_uint)
    rng = np.random.default_rng(11)
    cases = list(_cases())
    # big, mostly empty vectors: the shape of a deep hierclust node
    for n, npart, nchild in [(120000, 900, 400), (120000, 30000, 9000), (50000, 3, 2)]:
        P = np.zeros(n); C = np.zeros((n, 2), order="F")
        rows = np.sort(rng.choice(n, npart, replace=False))
        P[rows] = rng.random(npart) * (rng.random(npart) > 0.2)
        sub = rng.choice(rows, nchild, replace=False)
        C[sub, 0] = rng.random(nchild) * (rng.random(nchild) > 0.3); C[sub, 1] = rng.random(nchild) * (rng.random(nchild) > 0.3)
        cases.append((n, "deep", P, C))
    for n, mode, P, C in cases:
        touched = np.flatnonzero((C[:, 0] != 0) | (C[:, 1] != 0))
        extra = rng.choice(n, min(n, 17), replace=False)                       # rows of the node that ended up zero in both factors
        child_rows = np.unique(np.concatenate([touched, extra])).astype(np.uint32)
        a = lib.smkh_compute_priority_rows(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n, child_rows.ctypes.data_as(up), len(child_rows))
        b = lib.smkh_compute_priority_plain(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n)
        assert a == b, (n, mode, a, b)
        # the tree driver also names the rows outside of which the parent vector is zero (a superset of its non-zero rows)
        parent_rows = np.unique(np.concatenate([np.flatnonzero(P != 0), rng.choice(n, min(n, 29), replace=False)])).astype(np.uint32)
        c = lib.smkh_compute_priority_rows2(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n, child_rows.ctypes.data_as(up), len(child_rows),
                                            parent_rows.ctypes.data_as(up), len(parent_rows))
        assert c == b, (n, mode, c, b)


def test_priority_equals_reference_function_bit_for_bit():
    from oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref not built on this machine")
    lib = _host()
    ref = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsmallk_ref.so"))
    ref.ref_compute_priority.restype = ctypes.c_double
    for n, mode, P, C in _cases():
        a = lib.smkh_compute_priority(P.ctypes.data_as(dp), C.ctypes.data_as(dp), n)
        Pc, Cc = P.copy(), C.copy(order="F")
        want = ref.ref_compute_priority(Pc.ctypes.data_as(dp), Cc.ctypes.data_as(dp), n)
        assert a == want, (n, mode, a, want)
