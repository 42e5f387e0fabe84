"""preprocess_tf on the device (csrc/preprocess.cu, smk_preprocess_tf) against the CPU restatement of the reference's tf-idf
pipeline (oracle/preprocess_oracle.py, itself bit-identical to preprocessor/src/preprocess.cpp:81-250 built into oracle/_ref) and
against the reference's committed outputs (tests/golden/preprocess_*.npz): pruned matrix and index maps bit-exact, scores to
1e-13 (device log() and a warp-ordered sum of squares instead of glibc's log and a sequential sum)."""
import os
import sys

import numpy as np
import pytest

from oracle.preprocess_oracle import preprocess_tf

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_oracle_preprocess import _term_counts      # noqa: E402  the same generator as the CPU tests

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-13


def _same(got, want):
    assert (got is None) == (want is None)
    if want is None:
        return
    assert (got["m"], got["n"]) == (want["m"], want["n"])
    for key in ("colptr", "rows", "counts", "term_indices", "doc_indices"):
        assert np.array_equal(got[key], want[key]), key
    assert np.allclose(got["scores"], want["scores"], rtol=SCORE_TOL, atol=0.0)


@pytest.mark.parametrize("m,n,per_doc,seed,dup,ubi,short,dpt,tpd,max_iter", [
    (300, 200, 25, 1, 0, 0, 0, 3, 5, 1000),
    (300, 200, 25, 2, 12, 2, 9, 3, 5, 1000),          # duplicates, ubiquitous terms, short documents
    (1000, 400, 40, 3, 30, 1, 20, 5, 8, 1000),
    (150, 120, 10, 4, 6, 0, 5, 2, 3, 1),              # one round only
    (500, 400, 14, 5, 10, 3, 0, 4, 6, 1000),
    (80, 300, 30, 6, 40, 0, 0, 1, 1, 1000),           # nothing prunable by counts: duplicates only
    (20000, 6000, 60, 7, 300, 2, 150, 3, 5, 1000),    # larger: several rounds, long runs of equal hashes
])
def test_device_pipeline_matches_oracle(gpu, m, n, per_doc, seed, dup, ubi, short, dpt, tpd, max_iter):
    colptr, rows, counts = _term_counts(m, n, per_doc, seed, dup, ubi, short)
    want = preprocess_tf(m, n, colptr, rows, counts, max_iter, dpt, tpd)
    got = gpu.preprocess_tf(m, n, colptr, rows, counts, max_iter, dpt, tpd)
    _same(got, want)
    if want is not None and dup:
        assert want["n"] < n


def test_all_documents_pruned_is_reported(gpu):
    colptr, rows, counts = _term_counts(50, 20, 3, 9)
    assert preprocess_tf(50, 20, colptr, rows, counts, 1000, 1, 40) is None
    assert gpu.preprocess_tf(50, 20, colptr, rows, counts, 1000, 1, 40) is None


def test_device_pipeline_reproduces_reference_fixtures(gpu):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_preprocess as mg
    for name, c in mg.PREPROCESS_CASES.items():
        z = np.load(os.path.join(HERE, "golden", name + ".npz"))
        got = gpu.preprocess_tf(c["m"], c["n"], z["in_colptr"], z["in_rows"], z["in_counts"], c["max_iter"], c["dpt"], c["tpd"])
        assert (got["m"], got["n"]) == (int(z["out_m"]), int(z["out_n"]))
        for key in ("colptr", "rows", "counts", "term_indices", "doc_indices"):
            assert np.array_equal(got[key], z["out_" + key]), (name, key)
        assert np.allclose(got["scores"], z["out_scores"], rtol=SCORE_TOL, atol=0.0)
