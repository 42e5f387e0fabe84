"""Split-count sweep of the two big products on column shards of C2 / C5 (one GPU): python tools/sweep_splits.py [c2|c5]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallk_b200 as sk

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
m, n, k = (20000, 20000, 64) if which == "c2" else (100000, 50000, 256)
dev = torch.device("cuda", 0)
ctx = sk.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
out = {}
for world in (1, 2, 4, 8):
    n_loc = n // world
    A = torch.rand((n_loc, m), dtype=torch.float64, device=dev)
    W0 = np.asfortranarray(np.random.default_rng(1).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(2).random((k, n_loc)))
    ctx.load_dense_device(A.data_ptr(), m, m, n_loc)
    opts = sk.make_options(m, n_loc, k, algorithm="BPP", tol=1e-15, min_iter=1, max_iter=10, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    for name, idx, envn in (("WtA", 0, "SMK_GEMM_SPLITS_NN"), ("HAt", 1, "SMK_GEMM_SPLITS_NT")):
        for fixup in ((0,) if idx == 0 else (0, 1)):
            os.environ["SMK_GEMM_FIXUP"] = str(fixup)
            res = {}
            for s in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32):
                if s > 1 and which == "c5" and s > 8:
                    continue
                os.environ[envn] = str(s)
                try:
                    res[s] = round(ctx.time_product(idx, reps=5), 4)
                except Exception as ex:
                    res[s] = str(ex)[:200]
            os.environ[envn] = "0"
            ok = [(v, s) for s, v in res.items() if isinstance(v, float) and s > 0]
            print(f"{which} world={world} {name} fixup={fixup}: default {res[0]} best {min(ok) if ok else None} all {res}", flush=True)
            out[f"{world}/{name}/{fixup}"] = res
    os.environ["SMK_GEMM_FIXUP"] = "0"
    del A
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_splits_{which}.json"), "w"))
