"""The NumPy restatement of the reference's tf-idf preprocessing (oracle/preprocess_oracle.py) against the reference's own
preprocess_tf compiled into oracle/_ref (preprocessor/src/preprocess.cpp:81-250): identical pruned matrices, index maps and —
bit for bit — scores, on term-count matrices with rare terms, ubiquitous terms, short documents and duplicated documents."""
import ctypes
import os

import numpy as np
import pytest

from oracle.preprocess_oracle import preprocess_tf

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libsmallk_ref.so")
up = ctypes.POINTER(ctypes.c_uint)
dp = ctypes.POINTER(ctypes.c_double)


def _ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built on this machine")
    lib = ctypes.CDLL(REF_SO)
    if not hasattr(lib, "ref_preprocess_tf"):
        pytest.skip("oracle/_ref predates the preprocessing entry point")
    return lib


def _term_counts(m, n, per_doc, seed, duplicates=0, ubiquitous=0, short_docs=0):
    """Zipf-distributed term draws per document (so rare and frequent terms exist), counts = multiplicities."""
    rng = np.random.default_rng(seed)
    cols = []
    for c in range(n):
        k = per_doc if c >= short_docs else int(rng.integers(1, 4))
        t = np.minimum((m ** rng.random(k)).astype(np.int64), m - 1)
        t = np.concatenate([t, np.arange(ubiquitous)])        # terms present in every document
        u, cnt = np.unique(t, return_counts=True)
        perm = rng.permutation(len(u))                         # unsorted rows: the reference sorts them first
        cols.append((u[perm], cnt[perm]))
    for d in range(duplicates):                                # exact copies of earlier documents, later in the matrix
        src = int(rng.integers(short_docs, n // 2))
        cols[n - 1 - d] = (cols[src][0].copy(), cols[src][1].copy())
    colptr = np.concatenate([[0], np.cumsum([len(c[0]) for c in cols])]).astype(np.uint32)
    rows = np.concatenate([c[0] for c in cols]).astype(np.uint32)
    counts = np.concatenate([c[1] for c in cols]).astype(np.float64)
    return colptr, rows, counts


def _run_ref(lib, m, n, colptr, rows, counts, max_iter, dpt, tpd):
    nz = len(rows)
    om, on, onz = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
    oc = np.zeros(n + 1, dtype=np.uint32); orow = np.zeros(nz, dtype=np.uint32); ocnt = np.zeros(nz, dtype=np.uint32)
    osc = np.zeros(nz); ti = np.zeros(m, dtype=np.uint32); di = np.zeros(n, dtype=np.uint32)
    rc = lib.ref_preprocess_tf(m, n, nz, colptr.ctypes.data_as(up), rows.ctypes.data_as(up), counts.ctypes.data_as(dp), max_iter, dpt, tpd,
                               ctypes.byref(om), ctypes.byref(on), ctypes.byref(onz), oc.ctypes.data_as(up), orow.ctypes.data_as(up),
                               ocnt.ctypes.data_as(up), osc.ctypes.data_as(dp), ti.ctypes.data_as(up), di.ctypes.data_as(up))
    if rc != 0:
        return None
    h, w, z = om.value, on.value, onz.value
    return {"m": h, "n": w, "colptr": oc[: w + 1].astype(np.int64), "rows": orow[:z].astype(np.int64), "counts": ocnt[:z].astype(np.int64),
            "scores": osc[:z], "term_indices": ti[:h].astype(np.int64), "doc_indices": di[:w].astype(np.int64)}


@pytest.mark.parametrize("m,n,per_doc,seed,dup,ubi,short,dpt,tpd,max_iter", [
    (300, 200, 25, 1, 0, 0, 0, 3, 5, 1000),
    (300, 200, 25, 2, 12, 2, 9, 3, 5, 1000),          # duplicates, ubiquitous terms, short documents
    (1000, 400, 40, 3, 30, 1, 20, 5, 8, 1000),
    (150, 120, 10, 4, 6, 0, 5, 2, 3, 1),              # one round only
    (500, 400, 14, 5, 10, 3, 0, 4, 6, 1000),
    (80, 300, 30, 6, 40, 0, 0, 1, 1, 1000),           # nothing prunable by counts: duplicates only
])
def test_restatement_matches_reference(m, n, per_doc, seed, dup, ubi, short, dpt, tpd, max_iter):
    lib = _ref()
    colptr, rows, counts = _term_counts(m, n, per_doc, seed, dup, ubi, short)
    want = _run_ref(lib, m, n, colptr, rows, counts, max_iter, dpt, tpd)
    got = preprocess_tf(m, n, colptr, rows, counts, max_iter, dpt, tpd)
    assert (got is None) == (want is None)
    if want is None:
        return
    assert (got["m"], got["n"]) == (want["m"], want["n"])
    assert want["n"] < n or dup == 0                        # the duplicated documents are gone
    for key in ("colptr", "rows", "counts", "term_indices", "doc_indices"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(got["scores"], want["scores"])   # same operations in the same order: bit-identical
    norms = np.sqrt(np.add.reduceat(got["scores"] ** 2, got["colptr"][:-1]))
    assert np.allclose(norms, 1.0, rtol=1e-12)


def test_all_documents_pruned_is_reported():
    lib = _ref()
    colptr, rows, counts = _term_counts(50, 20, 3, 9)
    assert _run_ref(lib, 50, 20, colptr, rows, counts, 1000, 1, 40) is None
    assert preprocess_tf(50, 20, colptr, rows, counts, 1000, 1, 40) is None


def test_restatement_reproduces_reference_fixtures():
    """The committed outputs of the reference (tests/golden/preprocess_*.npz, made by tests/golden/make_golden_preprocess.py):
    needs neither /root/reference nor oracle/_ref."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_preprocess as mg
    for name, c in mg.PREPROCESS_CASES.items():
        z = np.load(os.path.join(HERE, "golden", name + ".npz"))
        got = preprocess_tf(c["m"], c["n"], z["in_colptr"], z["in_rows"], z["in_counts"], c["max_iter"], c["dpt"], c["tpd"])
        assert (got["m"], got["n"]) == (int(z["out_m"]), int(z["out_n"]))
        for key in ("colptr", "rows", "counts", "term_indices", "doc_indices", "scores"):
            assert np.array_equal(got[key], z["out_" + key]), (name, key)
