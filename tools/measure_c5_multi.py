"""C5 (BASELINE configs[4]): dense 100000 x 50000, k = 256, column-sharded over the ranks torchrun started (one per GPU),
BPP (default) or MU; NCCL all-reduce of H*H', reduce-scatter of H*A', row-sharded W update, all-gather of W inside the
library. Prints one JSON line from rank 0: outer iterations / s (device time, max over ranks) and the fraction of the FP64
tensor peak. Not the bench contract (bench.py is): a measurement of the largest configuration.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/measure_c5_multi.py [ALG] [iters]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smallk_b200 as sk                              # noqa: E402
from smallk_b200.sharding import column_block       # noqa: E402


def main():
    alg = sys.argv[1] if len(sys.argv) > 1 else "BPP"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    m, n, k = 100000, 50000, 256
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c0, c1 = column_block(n, rank, world)
    n_loc = c1 - c0
    ctx = sk.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    A = torch.empty((n_loc, m), dtype=torch.float64, device=dev)          # row-major [n][m] == column-major m x n
    g = torch.Generator(device=dev)
    BLK = 500
    for b0 in range((c0 // BLK) * BLK, c1, BLK):
        g.manual_seed(41 * 1000003 + b0)
        blk = torch.rand((BLK, m), dtype=torch.float64, device=dev, generator=g)
        lo, hi = max(b0, c0), min(b0 + BLK, c1)
        A[lo - c0:hi - c0] = blk[lo - b0:hi - b0]
        del blk
    W0 = np.asfortranarray(np.random.default_rng(42).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(43).random((k, n))[:, c0:c1])
    ctx.load_dense_device(A.data_ptr(), m, m, n_loc)
    opts = sk.make_options(m, n_loc if world > 1 else n, k, algorithm=alg, tol=1e-15, min_iter=1, max_iter=1000, normalize=False)
    ctx.solver_begin(W0, H0, opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        ctx.solver_step(1)
        ctx.solver_progress()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(iters):
        ctx.solver_step(1)
        metric = ctx.solver_progress()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        F = 4.0 * k * m * n + (6.0 * k * k * n + 4.0 * k * k * m if alg == "BPP" else 6.0 * k * k * (m + n))
        per = ms / iters
        print(json.dumps({"workload": f"C5 dense {alg} {m}x{n} k={k}, column-sharded", "n_gpus": world, "ms_per_iter": per,
                          "iters_per_s": 1000.0 / per, "flop_per_iter": F, "achieved_TFLOPs_total": F / per * 1e-9,
                          "achieved_TFLOPs_per_gpu": F / per * 1e-9 / world, "metric": metric}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
