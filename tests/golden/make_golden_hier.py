"""Generates tests/golden/hier_*.npz: HierNMF2 trees and assignments produced by the REFERENCE's own hierclust code
(oracle/_ref/libsmallk_ref.so = unmodified hierclust/src/clust.cpp + clust_hier_generic.hpp + tree.hpp compiled
against the El.hpp shim) on seeded synthetic inputs, with max_threads = 1 so that every random initialiser comes
from the reference's sequential generator (a deterministic function of the seed).

Run in the dev container:  python tests/golden/make_golden_hier.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from graphgen import powerlaw_graph, topic_matrix   # noqa: E402

# name -> (kind, params, num_clusters, run-seed, extra options)
HIER_CASES = {
    # symmetric power-law graphs (the C4 workload in miniature); outliers appear in the second one
    "hier_graph_2000_c6":  ("graph", dict(n=2000, avg_deg=10, seed=31), 6, 7, dict()),
    "hier_graph_3000_c8":  ("graph", dict(n=3000, avg_deg=12, seed=31), 8, 5, dict()),
    "hier_graph_1500_c5_flat": ("graph", dict(n=1500, avg_deg=14, seed=33), 5, 11, dict(flat=True, tol=1e-3)),
    # term-document style rectangular sparse matrix with planted topics
    "hier_topics_800x600_c7": ("topics", dict(m=800, n=600, topics=9, seed=41), 7, 3, dict()),
    # dense input through Clust()
    "hier_dense_120x90_c4": ("dense", dict(m=120, n=90, topics=5, seed=51), 4, 2, dict()),
}
KEYS = ["assignments", "parent", "left", "right", "is_left", "doc_count", "terms", "priority", "is_leaf"]


def hier_inputs(name):
    kind, p, num_clusters, seed, extra = HIER_CASES[name]
    if kind == "graph":
        colp, rowi, val = powerlaw_graph(p["n"], p["avg_deg"], p["seed"])
        return dict(csc=(colp, rowi, val), shape=(p["n"], p["n"]), A=None, num_clusters=num_clusters, seed=seed, extra=extra)
    S = topic_matrix(p["m"], p["n"], p["topics"], p["seed"])
    if kind == "dense":
        return dict(csc=None, shape=(p["m"], p["n"]), A=np.asfortranarray(S.toarray()), num_clusters=num_clusters, seed=seed, extra=extra)
    return dict(csc=(S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data.astype(np.float64)),
                shape=(p["m"], p["n"]), A=None, num_clusters=num_clusters, seed=seed, extra=extra)


def run_reference(ref, g):
    return ref.hierclust(A_dense=g["A"], csc=g["csc"], shape=g["shape"], num_clusters=g["num_clusters"], seed=g["seed"],
                         max_threads=1, **g["extra"])


def main():
    from oracle import Ref
    ref = Ref()
    for name in sorted(HIER_CASES):
        g = hier_inputs(name)
        o = run_reference(ref, g)
        assert o["rc"] == 0, (name, o["rc"])
        out = {k: o[k] for k in KEYS}
        out.update(n_outliers=o["n_outliers"], nmf_count=o["nmf_count"], max_count=o["max_count"])
        if g["extra"].get("flat"):
            out.update(flat_assignments=o["flat_assignments"], W=o["W"], H=o["H"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "nmf_count", o["nmf_count"], "outliers", o["n_outliers"], "leaf sizes",
              o["doc_count"][o["is_leaf"] == 1].tolist())


if __name__ == "__main__":
    main()
