"""Launched by tests/test_gpu_multi.py under torch.distributed.run, one rank per GPU: column-sharded NMF through the CUDA
library against the single-process CPU oracle, with both exchange back ends (the library's own NVLink peer-memory kernels,
csrc/peer.cu, and NCCL: SMK_PEER=1/0) and both ways of driving the loop (step + progress with a host read-back per
iteration, and smk_solver_run: all iterations enqueued, metrics computed on the device). Every rank checks the full W and
its own block of H."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallk_b200 as sk                                     # noqa: E402
from smallk_b200.sharding import column_block              # noqa: E402
from oracle import Oracle                                     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sk.Context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    orc = Oracle()
    failures = []
    cases = [("BPP", 301, 257, 24, 8, False), ("BPP", 420, 390, 100, 5, False), ("MU", 203, 190, 9, 12, False),
             ("HALS", 230, 200, 12, 10, False), ("RANK2", 250, 240, 2, 12, False), ("BPP", 300, 280, 16, 8, True),
             # k > 256: the any-k fallback kernels; the exchange goes through NCCL whatever SMK_PEER says (solver_alloc)
             ("HALS", 330, 300, 260, 3, False), ("MU", 300, 280, 264, 4, False)]
    for peer in ("1", "0"):
        os.environ["SMK_PEER"] = peer          # read by the library at every solver_begin
        for alg, m, n, k, iters, sparse in cases:
            rng = np.random.default_rng(m * 7 + n)
            A = rng.random((m, n))
            if sparse:
                A *= rng.random((m, n)) < 0.2
            W0 = rng.random((m, k)); H0 = rng.random((k, n))
            if alg == "HALS":
                H0 *= 2.0 / k
            c0, c1 = column_block(n, rank, world)
            if sparse:
                import scipy.sparse as sps
                S = sps.csc_matrix(A[:, c0:c1]); S.sort_indices()
                ctx.load_csc((m, c1 - c0), S.indptr, S.indices, S.data)
                Sf = sps.csc_matrix(A); Sf.sort_indices()
                o = orc.nmf_sparse((m, n), Sf.indptr.astype(np.uint32), Sf.indices.astype(np.uint32), Sf.data, W0, H0, alg=alg, tol=1e-12,
                                   min_iter=1, max_iter=iters, trace=True)
            else:
                ctx.load_dense(np.asfortranarray(A[:, c0:c1]))
                o = orc.nmf_dense(A, W0, H0, alg=alg, tol=1e-12, min_iter=1, max_iter=iters, trace=True)
            opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-12, min_iter=1, max_iter=iters, normalize=False)
            tol = 1e-6 if alg == "HALS" else 1e-9

            def compare(it, metric, tag):
                W, H = ctx.solver_get()
                rw = np.linalg.norm(W - o["W_trace"][it]) / np.linalg.norm(o["W_trace"][it])
                rh = np.linalg.norm(H - o["H_trace"][it][:, c0:c1]) / np.linalg.norm(o["H_trace"][it][:, c0:c1])
                rm = abs(metric - o["metrics"][it]) / abs(o["metrics"][it])
                if not (rw < tol and rh < tol and rm < max(tol, 1e-8)):
                    failures.append((tag, peer, alg, m, n, k, sparse, it, rw, rh, rm))
                    return False
                return True

            ctx.solver_begin(W0, np.asfortranarray(H0[:, c0:c1]), opts)
            for it in range(iters):
                ctx.solver_step(1)
                if not compare(it, ctx.solver_progress(), "step"):
                    break
            # the same iterations enqueued back to back, metrics on the device
            ctx.solver_begin(W0, np.asfortranarray(H0[:, c0:c1]), opts)
            metrics = ctx.solver_run(iters)
            rm = np.abs(metrics - o["metrics"][:iters]) / np.abs(o["metrics"][:iters])
            if not np.all(rm < max(tol, 1e-8)):
                failures.append(("run-metrics", peer, alg, m, n, k, sparse, rm.tolist()))
            compare(iters - 1, metrics[-1], "run")
    ctx.close()
    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MULTI_GPU_RESULT", "OK" if int(t.item()) == 0 else "FAIL", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if not failures else 1)


if __name__ == "__main__":
    main()
