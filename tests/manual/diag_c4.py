"""Manual check (lives under tests/ because it runs the reference checker). Where does the GPU hierclust tree leave the reference's at C4 scale? Runs both for a few cluster counts on the same graph
and prints the first tree node at which document counts / priorities differ. Needs /root/repo/oracle/_ref (prebuilt)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import smallk_b200 as sk          # noqa: E402
import workloads                  # noqa: E402
from oracle import Ref            # noqa: E402

n, edges, clusters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ref_threads = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # 1: the reference then uses its sequential random initialiser, as this library does
colp, rowi, val = workloads.c4_graph(n, edges)
kw = dict(csc=(colp, rowi, val), shape=(n, n), num_clusters=clusters, tol=1e-4, min_iter=5, max_iter=5000, seed=32)
g = sk.hierclust(**kw)
t = time.time()
r = Ref().hierclust(max_threads=ref_threads, **kw)
ref_s = time.time() - t
nodes = len(g["doc_count"])
first = next((i for i in range(nodes) if g["doc_count"][i] != r["doc_count"][i] or g["parent"][i] != r["parent"][i]), None)
relp = np.abs(g["priority"] - r["priority"]) / np.maximum(np.abs(r["priority"]), 1e-300)
print(json.dumps({"fused": os.environ.get("SMK_RANK2_FUSED", "1"), "ref_threads": ref_threads, "nodes": n, "clusters": clusters, "gpu_s": g["elapsed_s"], "ref_s": ref_s, "gpu_nmf": g["nmf_count"], "ref_nmf": int(r["nmf_count"]),
                  "gpu_iters": g["iterations"], "ref_iters": int(r.get("iterations", -1)),
                  "same_assignments": bool(np.array_equal(g["assignments"], r["assignments"])),
                  "assignments_differing": int(np.sum(g["assignments"] != r["assignments"])),
                  "first_tree_difference_at_node": first, "max_rel_priority_diff": float(relp.max()),
                  "gpu_doc_count": g["doc_count"][:12].tolist(), "ref_doc_count": r["doc_count"][:12].tolist(),
                  "gpu_priority": g["priority"][:8].tolist(), "ref_priority": r["priority"][:8].tolist()}), flush=True)
