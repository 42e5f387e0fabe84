"""smallk_b200 — B200-native implementation of SmallK's NMF iteration hot path.

This Python module is only a thin ctypes binding over the C-ABI library
``smallk_b200/lib/libsmallk_b200.so`` (``include/smallk_b200.h``), used by the
tests and by ``bench.py``. The product is the CUDA library and the C++ host
interface in ``smallk_b200/host/``; there is no CPU fallback here — if the
library is missing or no sm_100 device is present, calls raise.

Names mirror the reference's library interface (``common/include/nmf.hpp``):
``NmfOptions`` fields, ``Result`` codes, ``Nmf`` / ``NmfSparse``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsmallk_b200.so")

# common/include/nmf.hpp:17-41
OK, NOTINITIALIZED, INITIALIZED, BAD_PARAM, FAILURE, SIZE_TOO_LARGE, FLATCLUST_FAILURE = 0, -1, -2, -3, -4, -5, -6
CUDA_ERROR = -100
ALGORITHMS = {"MU": 0, "HALS": 1, "RANK2": 2, "BPP": 3}
PROGRESS = {"PG_RATIO": 0, "DELTA_FNORM": 1}

EXPORTS = [
    "smk_create", "smk_destroy", "smk_last_error", "smk_device_sm_count", "smk_device_index", "smk_set_stream", "smk_synchronize",
    "smk_comm_unique_id", "smk_comm_init", "smk_load_dense", "smk_load_dense_device", "smk_load_csc", "smk_nmf",
    "smk_solver_begin", "smk_solver_step", "smk_solver_progress", "smk_solver_get", "smk_solver_normalize",
    "smk_solver_last_step_ms", "smk_solver_time_product", "smk_gemm", "smk_nnls_bpp", "smk_sparse_gemm",
    "smk_select_columns", "smk_select_all", "smk_nnls_hals", "smk_argsort_desc", "smk_sort_desc",
    "smk_solver_run", "smk_phase_report", "smk_nnls_backup_count", "smk_preprocess_tf",
]
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libsmallk_host.so")
HOST_EXPORTS = ["smkh_last_error", "smkh_hierclust_sparse", "smkh_hierclust_dense", "smkh_flatclust", "smkh_compute_priority",
                "smkh_last_hier_profile", "smkh_compute_priority_plain", "smkh_compute_priority_gpu", "smkh_compute_priority_rows", "smkh_compute_priority_rows2", "smkh_load_matrix_market",
                "smkh_load_delimited", "smkh_write_delimited", "smkh_compute_assignments", "smkh_compute_fuzzy_assignments",
                "smkh_top_terms_matrix", "smkh_random_matrices", "smkh_is_valid", "smkh_flatclust_write_results", "smkh_tree_script", "smkh_tree_script_compact"]


class SmallkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"smallk_b200 error {code}: {message}")
        self.code = code
        self.message = message


class NmfOptions(ctypes.Structure):
    """NmfOptions, common/include/nmf.hpp:56-70."""
    _fields_ = [("tol", ctypes.c_double), ("algorithm", ctypes.c_int), ("prog_est_algorithm", ctypes.c_int),
                ("height", ctypes.c_int), ("width", ctypes.c_int), ("k", ctypes.c_int),
                ("min_iter", ctypes.c_int), ("max_iter", ctypes.c_int), ("tolcount", ctypes.c_int),
                ("max_threads", ctypes.c_int), ("verbose", ctypes.c_int), ("normalize", ctypes.c_int)]


class NmfStats(ctypes.Structure):
    """NmfStats, common/include/nmf.hpp:43-53."""
    _fields_ = [("elapsed_us", ctypes.c_ulonglong), ("iteration_count", ctypes.c_int)]


_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_up = ctypes.POINTER(ctypes.c_uint)


def load_library():
    """Load the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmallkError(NOTINITIALIZED, f"{LIB_PATH} not built — run `make -C smallk_b200/csrc` "
                                              "(or __graft_entry__.build())")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.smk_last_error.restype = ctypes.c_char_p
        _lib.smk_last_error.argtypes = [ctypes.c_void_p]
        _lib.smk_destroy.argtypes = [ctypes.c_void_p]
        _lib.smk_destroy.restype = None
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def make_options(m, n, k, algorithm="BPP", prog="PG_RATIO", tol=0.005, min_iter=5, max_iter=5000, tolcount=1,
                 normalize=True, verbose=False, max_threads=1):
    """Defaults follow nmf/src/command_line.cpp:173-194."""
    return NmfOptions(tol, ALGORITHMS[algorithm], PROGRESS[prog], m, n, k, min_iter, max_iter, tolcount,
                      max_threads, int(verbose), int(normalize))


class Context:
    """One GPU context (one host thread, one device)."""

    def __init__(self, device=0):
        lib = load_library()
        h = ctypes.c_void_p()
        rc = lib.smk_create(ctypes.byref(h), int(device))
        if rc != OK:
            raise SmallkError(rc, "smk_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self._h = h
        self._lib = lib
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.smk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise SmallkError(rc, self._lib.smk_last_error(self._h).decode())

    @property
    def sm_count(self):
        return self._lib.smk_device_sm_count(self._h)

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.smk_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(self._lib.smk_synchronize(self._h))

    # ---- multi-GPU -------------------------------------------------------
    def comm_unique_id(self):
        buf = (ctypes.c_ubyte * 128)()
        self._check(self._lib.smk_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.smk_comm_init(self._h, rank, nranks, buf))

    # ---- matrix ----------------------------------------------------------
    def load_dense(self, A):
        A = _f(A)
        m, n = A.shape
        self._check(self._lib.smk_load_dense(self._h, _d(A), ctypes.c_longlong(m), m, n))
        self.synchronize()
        self.shape = self.full_shape = (m, n)

    def load_dense_device(self, ptr, ld, m, n):
        self._check(self._lib.smk_load_dense_device(self._h, ctypes.c_void_p(ptr), ctypes.c_longlong(ld), m, n))
        self.shape = self.full_shape = (m, n)

    def load_csc(self, shape, col_offsets, row_indices, data):
        colp = np.ascontiguousarray(col_offsets, dtype=np.uint32)
        rowi = np.ascontiguousarray(row_indices, dtype=np.uint32)
        val = np.ascontiguousarray(data, dtype=np.float64)
        m, n = shape
        self._check(self._lib.smk_load_csc(self._h, m, n, ctypes.c_uint(int(colp[-1])), colp.ctypes.data_as(_up),
                                           rowi.ctypes.data_as(_up), _d(val)))
        self.shape = self.full_shape = (m, n)

    # ---- column subsets (hierclust) -----------------------------------------
    def select_columns(self, cols):
        """SubMatrixColsCompact on the device: returns new_to_old_rows (length = new height)."""
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        n2o = np.zeros(self.full_shape[0], dtype=np.uint32)
        nh = ctypes.c_int(0)
        self._check(self._lib.smk_select_columns(self._h, cols.ctypes.data_as(_up), len(cols), ctypes.byref(nh),
                                                 n2o.ctypes.data_as(_up)))
        self.shape = (nh.value, len(cols))
        return n2o[:nh.value].copy()

    def select_all(self):
        self._check(self._lib.smk_select_all(self._h))
        self.shape = self.full_shape

    def nnls_hals(self, W, H0, tol, max_iter):
        """NnlsHals (nnls.hpp:249-316): returns (rc, W, H, iterations)."""
        W = _f(W).copy(order="F")
        H = _f(H0).copy(order="F")
        it = ctypes.c_int(0)
        rc = self._lib.smk_nnls_hals(self._h, H.shape[0], _d(W), W.shape[0], _d(H), H.shape[0], ctypes.c_double(tol),
                                     int(max_iter), ctypes.byref(it))
        return rc, W, H, it.value

    # ---- Nmf / NmfSparse ---------------------------------------------------
    def nmf(self, W0, H0, options):
        """Result Nmf(opts, A, W, H, stats): returns (W, H, stats). Raises SmallkError on a non-OK Result."""
        W = _f(W0).copy(order="F")
        H = _f(H0).copy(order="F")
        st = NmfStats()
        rc = self._lib.smk_nmf(self._h, ctypes.byref(options), _d(W), W.shape[0], _d(H), H.shape[0], ctypes.byref(st))
        self._check(rc)
        return W, H, st

    # ---- solver-functor seam -----------------------------------------------
    def solver_begin(self, W0, H0, options):
        W = _f(W0)
        H = _f(H0)
        self._opts = options
        self._check(self._lib.smk_solver_begin(self._h, ctypes.byref(options), _d(W), W.shape[0], _d(H), H.shape[0]))

    def solver_step(self, count=1):
        self._check(self._lib.smk_solver_step(self._h, int(count)))

    def solver_progress(self):
        v = ctypes.c_double(0.0)
        self._check(self._lib.smk_solver_progress(self._h, ctypes.byref(v)))
        return v.value

    def solver_run(self, count):
        """`count` x { solver(); progress update } without host synchronisation in between; returns the metrics."""
        metrics = np.zeros(max(int(count), 1))
        self._check(self._lib.smk_solver_run(self._h, int(count), _d(metrics)))
        return metrics[:int(count)]

    def phase_report(self):
        """{phase: ms} since the last call (SMK_PHASES=1)."""
        buf = ctypes.create_string_buffer(4096)
        self._check(self._lib.smk_phase_report(self._h, buf, 4096))
        out = {}
        for item in buf.value.decode().split(";"):
            if "=" in item:
                name, v = item.split("=")
                out[name] = float(v)
        return out

    def solver_normalize(self):
        self._check(self._lib.smk_solver_normalize(self._h))

    def solver_get(self, grads=False):
        m, n = self.shape
        k = self._opts.k
        W = np.zeros((m, k), order="F")
        H = np.zeros((k, n), order="F")
        gW = np.zeros((m, k), order="F") if grads else None
        gH = np.zeros((k, n), order="F") if grads else None
        self._check(self._lib.smk_solver_get(self._h, _d(W), m, _d(H), k, _d(gW), m, _d(gH), k))
        return (W, H, gW, gH) if grads else (W, H)

    def last_step(self):
        ms = ctypes.c_float(0)
        n = ctypes.c_longlong(0)
        self._lib.smk_solver_last_step_ms(self._h, ctypes.byref(ms), ctypes.byref(n))
        return ms.value, n.value

    def time_product(self, which, reps=5):
        """Mean device ms per launch of W'A (which=0) or H A' (which=1) on the current state."""
        ms = ctypes.c_float(0)
        self._check(self._lib.smk_solver_time_product(self._h, int(which), int(reps), ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.last_step()[1]

    # ---- primitives ---------------------------------------------------------
    def gemm(self, A, B, transA=False, transB=False):
        A = _f(A); B = _f(B)
        M = A.shape[1] if transA else A.shape[0]
        K = A.shape[0] if transA else A.shape[1]
        N = B.shape[0] if transB else B.shape[1]
        C = np.zeros((M, N), order="F")
        self._check(self._lib.smk_gemm(self._h, int(transA), int(transB), M, N, K, _d(A), A.shape[0], _d(B), B.shape[0],
                                       _d(C), M))
        return C

    def nnls_bpp(self, LHS, RHS, X0):
        LHS = _f(LHS); RHS = _f(RHS)
        X = _f(X0).copy(order="F")
        Y = np.zeros_like(X, order="F")
        k, q = RHS.shape
        self._check(self._lib.smk_nnls_bpp(self._h, k, q, _d(LHS), _d(RHS), _d(X), _d(Y)))
        return X, Y

    def preprocess_tf(self, m, n, colptr, rows, counts, max_iter=1000, docs_per_term=3, terms_per_doc=5):
        """preprocess_tf (preprocessor/src/preprocess.cpp:81-250) on the device. Returns None if every document was pruned,
        else a dict like oracle.preprocess_oracle.preprocess_tf's."""
        colptr = np.ascontiguousarray(colptr, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.float64)
        nz = len(rows)
        om, on, onz = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
        oc = np.zeros(n + 1, dtype=np.uint32); orow = np.zeros(max(nz, 1), dtype=np.uint32); ocnt = np.zeros(max(nz, 1), dtype=np.uint32)
        osc = np.zeros(max(nz, 1)); ti = np.zeros(m, dtype=np.uint32); di = np.zeros(n, dtype=np.uint32)
        rc = self._lib.smk_preprocess_tf(self._h, ctypes.c_uint(m), ctypes.c_uint(n), ctypes.c_uint(nz), colptr.ctypes.data_as(_up),
                                         rows.ctypes.data_as(_up), _d(counts), ctypes.c_uint(max_iter), ctypes.c_uint(docs_per_term),
                                         ctypes.c_uint(terms_per_doc), ctypes.byref(om), ctypes.byref(on), ctypes.byref(onz),
                                         oc.ctypes.data_as(_up), orow.ctypes.data_as(_up), ocnt.ctypes.data_as(_up), _d(osc),
                                         ti.ctypes.data_as(_up), di.ctypes.data_as(_up))
        if rc == FAILURE:
            return None
        self._check(rc)
        h, w, z = om.value, on.value, onz.value
        return {"m": h, "n": w, "colptr": oc[: w + 1].astype(np.int64), "rows": orow[:z].astype(np.int64), "counts": ocnt[:z].astype(np.int64),
                "scores": osc[:z], "term_indices": ti[:h].astype(np.int64), "doc_indices": di[:w].astype(np.int64)}

    def nnls_backup_count(self):
        """Firings of UpdatePassiveSet's backup rule since the last nnls_bpp / solver_begin (diagnostic)."""
        n = ctypes.c_int(0)
        self._check(self._lib.smk_nnls_backup_count(self._h, ctypes.byref(n)))
        return n.value

    def sparse_gemm(self, variant, alpha, B, beta, C):
        B = _f(B)
        C = _f(C).copy(order="F")
        self._check(self._lib.smk_sparse_gemm(self._h, variant, ctypes.c_double(alpha), _d(B), B.shape[0], B.shape[1],
                                              ctypes.c_double(beta), _d(C), C.shape[0], C.shape[1]))
        return C


# ---- the C++ host layer (smallk_b200/host/: Clust / ClustSparse / FlatClust*) through its extern "C" veneer ----
_host = None
_ip = ctypes.POINTER(ctypes.c_int)


def load_host_library():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise SmallkError(NOTINITIALIZED, f"{HOST_LIB_PATH} not built — run `make -C smallk_b200/host`")
        load_library()
        _host = ctypes.CDLL(HOST_LIB_PATH)
        _host.smkh_last_error.restype = ctypes.c_char_p
        _host.smkh_compute_priority.restype = ctypes.c_double
    return _host


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def hierclust(A_dense=None, csc=None, shape=None, num_clusters=4, tol=1e-4, min_iter=5, max_iter=5000, maxterms=5,
              unbalanced=0.1, trial_allowance=3, flat=False, normalize=False, seed=1, verbose=False, lib=None):
    """HierNMF2 through the C++ host driver (host/clust.cpp): Clust for a dense array, ClustSparse for a
    (col_offsets, row_indices, data) triple. Returns the tree as arrays, like oracle.Ref.hierclust.
    lib: another build of the host library (tests/test_host_driver_cpu.py runs the driver over a CPU mock of the C ABI)."""
    lib = lib or load_host_library()
    if csc is not None:
        m, n = shape
    else:
        A_dense = _f(A_dense)
        m, n = A_dense.shape
    nodes = 2 * (num_clusters - 1)
    z = lambda *s: np.zeros(s, dtype=np.int32)
    out = {"assignments": z(n), "parent": z(nodes), "left": z(nodes), "right": z(nodes), "is_left": z(nodes),
           "doc_count": z(nodes), "terms": z(nodes, maxterms), "priority": np.zeros(nodes), "is_leaf": z(nodes)}
    n_out = ctypes.c_int(0)
    stats = np.zeros(3, dtype=np.int64)
    W = np.zeros((m, num_clusters), order="F")
    H = np.zeros((num_clusters, n), order="F")
    fa = z(n)
    el = ctypes.c_double(0.0)
    tail = (num_clusters, ctypes.c_double(tol), min_iter, max_iter, maxterms, ctypes.c_double(unbalanced), trial_allowance,
            int(flat), int(normalize), seed, int(verbose), _i(out["assignments"]), _i(out["parent"]), _i(out["left"]),
            _i(out["right"]), _i(out["is_left"]), _i(out["doc_count"]), _i(out["terms"]), _d(out["priority"]),
            _i(out["is_leaf"]), ctypes.byref(n_out), _d(W), _d(H), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
            _i(fa), ctypes.byref(el))
    if csc is not None:
        colp = np.ascontiguousarray(csc[0], dtype=np.uint32)
        rowi = np.ascontiguousarray(csc[1], dtype=np.uint32)
        val = np.ascontiguousarray(csc[2], dtype=np.float64)
        rc = lib.smkh_hierclust_sparse(m, n, ctypes.c_uint(int(colp[-1])), colp.ctypes.data_as(_up),
                                       rowi.ctypes.data_as(_up), _d(val), *tail)
    else:
        rc = lib.smkh_hierclust_dense(m, n, _d(A_dense), m, *tail)
    prof = np.zeros(6)
    lib.smkh_last_hier_profile(_d(prof))
    out.update(rc=rc, n_outliers=n_out.value, nmf_count=int(stats[0]), max_count=int(stats[1]),
               iterations=int(stats[2]), elapsed_s=el.value,
               profile=dict(zip(("extract_s", "init_s", "factor_s", "priority_s", "terms_s", "priority_worker_s"), prof.tolist())))
    if flat:
        out.update(W=W, H=H, flat_assignments=fa)
    return out


def flatclust(W0, H0, A_dense=None, csc=None, shape=None, algorithm="BPP", tol=0.005, min_iter=5, max_iter=5000, maxterms=5):
    """FlatClust / FlatClustSparse + ComputeAssignments + TopTerms through the C++ host layer."""
    lib = load_host_library()
    W = _f(W0).copy(order="F")
    H = _f(H0).copy(order="F")
    m, k = W.shape
    n = H.shape[1]
    assign = np.zeros(n, dtype=np.int32)
    terms = np.zeros((k, maxterms), dtype=np.int32)
    it = ctypes.c_int(0)
    if csc is not None:
        colp = np.ascontiguousarray(csc[0], dtype=np.uint32)
        rowi = np.ascontiguousarray(csc[1], dtype=np.uint32)
        val = np.ascontiguousarray(csc[2], dtype=np.float64)
        rc = lib.smkh_flatclust(ALGORITHMS[algorithm], m, n, k, ctypes.c_double(tol), min_iter, max_iter, maxterms, None, 0,
                                ctypes.c_uint(int(colp[-1])), colp.ctypes.data_as(_up), rowi.ctypes.data_as(_up), _d(val),
                                _d(W), _d(H), _i(assign), _i(terms), ctypes.byref(it))
    else:
        A = _f(A_dense)
        rc = lib.smkh_flatclust(ALGORITHMS[algorithm], m, n, k, ctypes.c_double(tol), min_iter, max_iter, maxterms, _d(A), m,
                                ctypes.c_uint(0), None, None, None, _d(W), _d(H), _i(assign), _i(terms), ctypes.byref(it))
    return {"rc": rc, "W": W, "H": H, "assignments": assign, "terms": terms, "iterations": it.value}
