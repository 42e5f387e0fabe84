"""Golden vectors of the reference's tf-idf preprocessing: inputs and the outputs of preprocess_tf (preprocessor/src/preprocess.cpp)
run through oracle/_ref in the development container. Usage: python tests/golden/make_golden_preprocess.py"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PREPROCESS_CASES = {
    "preprocess_300x200": dict(m=300, n=200, per_doc=25, seed=2, dup=12, ubi=2, short=9, dpt=3, tpd=5, max_iter=1000),
    "preprocess_1000x400": dict(m=1000, n=400, per_doc=40, seed=3, dup=30, ubi=1, short=20, dpt=5, tpd=8, max_iter=1000),
}


def main():
    import test_oracle_preprocess as t
    lib = ctypes.CDLL(t.REF_SO)
    for name, c in PREPROCESS_CASES.items():
        colptr, rows, counts = t._term_counts(c["m"], c["n"], c["per_doc"], c["seed"], c["dup"], c["ubi"], c["short"])
        out = t._run_ref(lib, c["m"], c["n"], colptr, rows, counts, c["max_iter"], c["dpt"], c["tpd"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), in_colptr=colptr, in_rows=rows, in_counts=counts,
                            **{"out_" + k: np.asarray(v) for k, v in out.items()})
        print(name, (c["m"], c["n"]), "->", (out["m"], out["n"]), "nnz", len(rows), "->", len(out["rows"]))


if __name__ == "__main__":
    main()
