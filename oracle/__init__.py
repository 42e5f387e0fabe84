"""TEST INFRASTRUCTURE — ctypes loaders for the two CPU checkers.

* ``Oracle``  -> oracle/libsmallk_oracle.so : our plain-C restatement (nmf_oracle.c)
* ``Ref``     -> oracle/_ref/libsmallk_ref.so : the reference's own sources + El.hpp shim

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this
package. The product (smallk_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsmallk_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsmallk_ref.so")

ALG = {"MU": 0, "HALS": 1, "RANK2": 2, "BPP": 3}       # common/include/nmf.hpp:28-34
PROG = {"PG_RATIO": 0, "DELTA_FNORM": 1}               # common/include/nmf.hpp:37-41

_dp = ctypes.POINTER(ctypes.c_double)
_up = ctypes.POINTER(ctypes.c_uint)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _u(a):
    return a.ctypes.data_as(_up)


def build(ref=True):
    """Compile the checkers (the restatement always; _ref only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/common"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


class _NmfMixin:
    """Shared shape handling: W (m,k) and H (k,n) are returned as new F-ordered arrays."""

    def _prep(self, W0, H0, max_iter, trace):
        W = _f(W0).copy(order="F")
        H = _f(H0).copy(order="F")
        m, k = W.shape
        n = H.shape[1]
        metrics = np.full(max_iter, np.nan)
        Ws = np.zeros((max_iter, k, m)) if trace else None      # [iter][col][row]
        Hs = np.zeros((max_iter, n, k)) if trace else None
        return W, H, m, n, k, metrics, Ws, Hs

    @staticmethod
    def _result(rc, it, W, H, metrics, Ws, Hs):
        out = {"rc": rc, "iterations": it.value, "W": W, "H": H, "metrics": metrics}
        if Ws is not None:
            out["W_trace"] = np.transpose(Ws, (0, 2, 1))
            out["H_trace"] = np.transpose(Hs, (0, 2, 1))
        return out


class Oracle(_NmfMixin):
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = ctypes.CDLL(ORACLE_SO)

    def nmf_dense(self, A, W0, H0, alg="BPP", prog="PG_RATIO", tol=1e-4, min_iter=5, max_iter=5000,
                  tolcount=1, normalize=False, trace=False):
        A = _f(A)
        W, H, m, n, k, metrics, Ws, Hs = self._prep(W0, H0, max_iter, trace)
        it = ctypes.c_int(0)
        rc = self.lib.orc_nmf_dense(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                    tolcount, int(normalize), _d(A), A.shape[0], _d(W), m, _d(H), k,
                                    ctypes.byref(it), _d(metrics), _d(Ws), _d(Hs))
        return self._result(rc, it, W, H, metrics, Ws, Hs)

    def nmf_sparse(self, shape, colp, rowi, val, W0, H0, alg="HALS", prog="PG_RATIO", tol=1e-4, min_iter=5,
                   max_iter=5000, tolcount=1, normalize=False, trace=False):
        colp = np.ascontiguousarray(colp, dtype=np.uint32)
        rowi = np.ascontiguousarray(rowi, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        W, H, m, n, k, metrics, Ws, Hs = self._prep(W0, H0, max_iter, trace)
        assert (m, n) == tuple(shape)
        it = ctypes.c_int(0)
        rc = self.lib.orc_nmf_sparse(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                     tolcount, int(normalize), _u(colp), _u(rowi), _d(val), _d(W), m, _d(H), k,
                                     ctypes.byref(it), _d(metrics), _d(Ws), _d(Hs))
        return self._result(rc, it, W, H, metrics, Ws, Hs)

    def nnls_bpp(self, LHS, RHS, X0):
        LHS = _f(LHS); RHS = _f(RHS)
        X = _f(X0).copy(order="F")
        Y = np.zeros_like(X, order="F")
        k, q = RHS.shape
        rc = self.lib.orc_nnls_bpp(k, q, _d(LHS), _d(RHS), _d(X), _d(Y))
        return rc, X, Y

    def sparse_gemm(self, variant, alpha, shape, colp, rowi, val, B, beta, C):
        colp = np.ascontiguousarray(colp, dtype=np.uint32)
        rowi = np.ascontiguousarray(rowi, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        B = _f(B)
        C = _f(C).copy(order="F")
        rc = self.lib.orc_sparse_gemm(variant, ctypes.c_double(alpha), ctypes.c_double(beta), shape[0], shape[1],
                                      _u(colp), _u(rowi), _d(val), _d(B), B.shape[0], B.shape[1],
                                      _d(C), C.shape[0], C.shape[1])
        assert rc == 0, rc
        return C


class Ref(_NmfMixin):
    """The reference's own code (common/src/nmf.cpp etc.) behind oracle/ref_capi.cpp."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, blas_threads=None):
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.ref_blas_backend.restype = ctypes.c_char_p
        if blas_threads is not None:
            self.lib.ref_set_blas_threads(int(blas_threads))

    def blas_backend(self):
        return self.lib.ref_blas_backend().decode()

    def set_blas_threads(self, n):
        self.lib.ref_set_blas_threads(int(n))

    def nmf_dense(self, A, W0, H0, alg="BPP", prog="PG_RATIO", tol=1e-4, min_iter=5, max_iter=5000,
                  tolcount=1, normalize=False, trace=False, max_threads=1, timed=False):
        A = _f(A)
        W, H, m, n, k, metrics, Ws, Hs = self._prep(W0, H0, max_iter, trace)
        it = ctypes.c_int(0)
        if timed:
            us = ctypes.c_ulonglong(0)
            rc = self.lib.ref_nmf_dense(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                        tolcount, max_threads, int(normalize), 0, _d(A), A.shape[0], _d(W), m,
                                        _d(H), k, ctypes.byref(it), ctypes.byref(us))
            out = self._result(rc, it, W, H, metrics, None, None)
            out["elapsed_us"] = us.value
            return out
        rc = self.lib.ref_nmf_dense_trace(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                          tolcount, max_threads, int(normalize), _d(A), A.shape[0], _d(W), m,
                                          _d(H), k, ctypes.byref(it), _d(metrics), _d(Ws), _d(Hs))
        return self._result(rc, it, W, H, metrics, Ws, Hs)

    def nmf_sparse(self, shape, colp, rowi, val, W0, H0, alg="HALS", prog="PG_RATIO", tol=1e-4, min_iter=5,
                   max_iter=5000, tolcount=1, normalize=False, trace=False, max_threads=1, timed=False):
        colp = np.ascontiguousarray(colp, dtype=np.uint32)
        rowi = np.ascontiguousarray(rowi, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        W, H, m, n, k, metrics, Ws, Hs = self._prep(W0, H0, max_iter, trace)
        it = ctypes.c_int(0)
        nz = int(colp[-1])
        if timed:
            us = ctypes.c_ulonglong(0)
            rc = self.lib.ref_nmf_sparse(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                         tolcount, max_threads, int(normalize), 0, nz, _u(colp), _u(rowi), _d(val),
                                         _d(W), m, _d(H), k, ctypes.byref(it), ctypes.byref(us))
            out = self._result(rc, it, W, H, metrics, None, None)
            out["elapsed_us"] = us.value
            return out
        rc = self.lib.ref_nmf_sparse_trace(ALG[alg], PROG[prog], m, n, k, ctypes.c_double(tol), min_iter, max_iter,
                                           tolcount, max_threads, int(normalize), nz, _u(colp), _u(rowi), _d(val),
                                           _d(W), m, _d(H), k, ctypes.byref(it), _d(metrics), _d(Ws), _d(Hs))
        return self._result(rc, it, W, H, metrics, Ws, Hs)

    def nnls_bpp(self, LHS, RHS, X0, max_threads=1):
        LHS = _f(LHS).copy(order="F"); RHS = _f(RHS).copy(order="F")
        X = _f(X0).copy(order="F")
        Y = np.zeros_like(X, order="F")
        k, q = RHS.shape
        rc = self.lib.ref_nnls_blockpivot(k, q, _d(LHS), _d(RHS), _d(X), _d(Y), max_threads)
        return rc, X, Y

    def sparse_gemm(self, variant, alpha, shape, colp, rowi, val, B, beta, C, max_threads=2):
        colp = np.ascontiguousarray(colp, dtype=np.uint32)
        rowi = np.ascontiguousarray(rowi, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        B = _f(B).copy(order="F")
        C = _f(C).copy(order="F")
        rc = self.lib.ref_sparse_gemm(variant, ctypes.c_double(alpha), ctypes.c_double(beta), shape[0], shape[1],
                                      int(colp[-1]), _u(colp), _u(rowi), _d(val), _d(B), B.shape[0], B.shape[1],
                                      _d(C), C.shape[0], C.shape[1], max_threads)
        assert rc == 0, rc
        return C


_ip = ctypes.POINTER(ctypes.c_int)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _hier_outputs(n, num_clusters, maxterms):
    nodes = 2 * (num_clusters - 1)
    z = lambda *s: np.zeros(s, dtype=np.int32)
    return {"assignments": z(n), "parent": z(nodes), "left": z(nodes), "right": z(nodes), "is_left": z(nodes),
            "doc_count": z(nodes), "terms": z(nodes, maxterms), "priority": np.zeros(nodes), "is_leaf": z(nodes)}


def _ref_hierclust(self, A_dense=None, csc=None, shape=None, num_clusters=4, tol=1e-4, min_iter=5, max_iter=5000,
                   maxterms=5, unbalanced=0.1, trial_allowance=3, flat=False, normalize=False, seed=1, max_threads=1):
    """Reference HierNMF2 (hierclust/src/clust.cpp) on a dense array or a (colp, rowi, val) CSC triple."""
    if csc is not None:
        m, n = shape
    else:
        A_dense = _f(A_dense)
        m, n = A_dense.shape
    out = _hier_outputs(n, num_clusters, maxterms)
    n_out = ctypes.c_int(0)
    stats = np.zeros(2, dtype=np.int32)
    W = np.zeros((m, num_clusters), order="F")
    H = np.zeros((num_clusters, n), order="F")
    fa = np.zeros(n, dtype=np.int32)
    el = ctypes.c_double(0.0)
    if csc is not None:
        colp, rowi, val = csc
        colp = np.ascontiguousarray(colp, dtype=np.uint32)
        rowi = np.ascontiguousarray(rowi, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        rc = self.lib.ref_hierclust_sparse(
            m, n, int(colp[-1]), _u(colp), _u(rowi), _d(val), num_clusters, ctypes.c_double(tol), min_iter, max_iter,
            maxterms, ctypes.c_double(unbalanced), trial_allowance, int(flat), int(normalize), seed, max_threads,
            _i(out["assignments"]), _i(out["parent"]), _i(out["left"]), _i(out["right"]), _i(out["is_left"]),
            _i(out["doc_count"]), _i(out["terms"]), _d(out["priority"]), _i(out["is_leaf"]), ctypes.byref(n_out),
            _d(W), _d(H), _i(stats), _i(fa), ctypes.byref(el))
    else:
        rc = self.lib.ref_hierclust_dense(
            m, n, _d(A_dense), m, num_clusters, ctypes.c_double(tol), min_iter, max_iter,
            maxterms, ctypes.c_double(unbalanced), trial_allowance, int(flat), int(normalize), seed, max_threads,
            _i(out["assignments"]), _i(out["parent"]), _i(out["left"]), _i(out["right"]), _i(out["is_left"]),
            _i(out["doc_count"]), _i(out["terms"]), _d(out["priority"]), _i(out["is_leaf"]), ctypes.byref(n_out),
            _d(W), _d(H), _i(stats), _i(fa))
    out.update(rc=rc, n_outliers=n_out.value, nmf_count=int(stats[0]), max_count=int(stats[1]), elapsed_s=el.value)
    if flat:
        out.update(W=W, H=H, flat_assignments=fa)
    return out


Ref.hierclust = _ref_hierclust
