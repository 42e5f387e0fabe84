"""world_size-2 runs of the column-sharded NMF exchange pattern on CPU (gloo): the partition helpers of
smallk_b200/sharding.py plus the collective sequence the CUDA library issues over NCCL (csrc/solver.cu) reproduce the
single-process CPU oracle. The per-shard arithmetic is the oracle's (tests may use it); what is under test is the
sharding: column blocks of A/H, all-reduce of H*H', row-sliced reduction of H*A', all-gather of W, summed progress metric."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smallk_b200.sharding import column_block, column_block_by_nnz, row_slice      # noqa: E402

WORLD = 2


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    dist.all_reduce(t)
    return t.numpy()


def _allgather_rows(W_slice_padded):
    t = torch.from_numpy(np.ascontiguousarray(W_slice_padded))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.numpy() for o in out], axis=0)


def _pg_sq(G, X):
    mask = (G < 0) | (X > 0)
    return float((G[mask] ** 2).sum())


def _worker(rank, port, alg, m, n, k, iters, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from oracle import Oracle
        orc = Oracle()
        rng = np.random.default_rng(77)
        A = rng.random((m, n)); W = rng.random((m, k)); Hfull = rng.random((k, n))
        c0, c1 = column_block(n, rank, WORLD)
        r0, rows, m_loc = row_slice(m, rank, WORLD)
        Al, H = A[:, c0:c1], Hfull[:, c0:c1].copy()
        WtW = W.T @ W; WtA = W.T @ Al
        metrics = []
        for it in range(iters):
            if alg == "MU":
                H *= WtA / (WtW @ H + 1e-13)
            else:
                rc, H, _ = orc.nnls_bpp(WtW, WtA, H)
                assert rc == 0
            HHt = _allreduce(H @ H.T)                                   # k x k all-reduce
            AHt = _allreduce(Al @ H.T)[r0:r0 + rows]                    # reduce-scatter == all-reduce + own row slice
            Wl = W[r0:r0 + rows]
            if alg == "MU":
                Wl = Wl * (AHt / (Wl @ HHt + 1e-13))
            else:
                rc, Wlt, _ = orc.nnls_bpp(HHt, AHt.T.copy(), Wl.T.copy())
                assert rc == 0
                Wl = Wlt.T
            gradWl = Wl @ HHt - AHt
            pad = np.zeros((m_loc, k)); pad[:rows] = Wl
            W = _allgather_rows(pad)[:m]                                # all-gather of the padded slices
            WtW = W.T @ W; WtA = W.T @ Al
            gradH = WtW @ H - WtA
            pg2 = _allreduce(np.array([_pg_sq(gradWl, Wl), _pg_sq(gradH, H)]))   # two partial sums, one all-reduce
            metrics.append(float(np.sqrt(pg2.sum())))
        Hall = [torch.empty((k, column_block(n, r, WORLD)[1] - column_block(n, r, WORLD)[0]), dtype=torch.float64) for r in range(WORLD)]
        dist.all_gather(Hall, torch.from_numpy(np.ascontiguousarray(H))) if n % WORLD == 0 else None
        if rank == 0:
            q.put((W, np.concatenate([h.numpy() for h in Hall], axis=1) if n % WORLD == 0 else None, metrics))
    finally:
        dist.destroy_process_group()


def _rank2_nnls(G, B):
    """min ||.|| over x >= 0 of the 2 x 2 normal equations G x = b, for every column b of B (2 x q): the unconstrained solution
    where it is positive, else the better single-variable solution (nmf_solver_rank2.hpp:218-318)."""
    X = np.linalg.solve(G, B)
    bad = (X[0] <= 0) | (X[1] <= 0)
    v0, v1 = B[0] / G[0, 0], B[1] / G[1, 1]
    first = v0 * np.sqrt(G[0, 0]) >= v1 * np.sqrt(G[1, 1])
    X[0, bad] = np.where(first, v0, 0.0)[bad]
    X[1, bad] = np.where(first, 0.0, v1)[bad]
    return X


def _worker_rank2(rank, port, m, n, iters, q):
    """RANK2: the in-loop normalisation couples all rows of W, so the library all-reduces H*A' and every rank updates the whole
    W (csrc/solver.cu, generic sequence); the H part of the projected-gradient sum is the only other exchange."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        rng = np.random.default_rng(78)
        A = rng.random((m, n)); W = rng.random((m, 2)); Hfull = rng.random((2, n))
        c0, c1 = column_block(n, rank, WORLD)
        Al, H = A[:, c0:c1], Hfull[:, c0:c1].copy()
        WtW = W.T @ W; WtA = W.T @ Al
        metrics = []
        for it in range(iters):
            H = _rank2_nnls(WtW, WtA)                                   # local columns
            HHt = _allreduce(H @ H.T)
            AHt = _allreduce(Al @ H.T)                                  # m x 2, summed over the column shards
            W = _rank2_nnls(HHt, AHt.T.copy()).T                        # replicated
            s = np.linalg.norm(W, axis=0)
            W = W / s; H = H * s[:, None]
            HHt = HHt * np.outer(s, s); AHt = AHt * s
            gradW = W @ HHt - AHt
            WtW = W.T @ W; WtA = W.T @ Al
            gradH = WtW @ H - WtA
            pg_h = _allreduce(np.array([_pg_sq(gradH, H)]))[0]          # only the H part is a partial sum
            metrics.append(float(np.sqrt(_pg_sq(gradW, W) + pg_h)))
        Hall = [torch.empty((2, column_block(n, r, WORLD)[1] - column_block(n, r, WORLD)[0]), dtype=torch.float64) for r in range(WORLD)]
        dist.all_gather(Hall, torch.from_numpy(np.ascontiguousarray(H)))
        if rank == 0:
            q.put((W, np.concatenate([h.numpy() for h in Hall], axis=1), metrics))
    finally:
        dist.destroy_process_group()


def _worker_hals(rank, port, m, n, k, iters, q):
    """HALS: the in-sweep normalisation of W's columns couples all rows, so H*H' and A*H' are all-reduced (at Init and at the end
    of every iteration, nmf_solver_hals.hpp:141-199) and every rank sweeps the whole W; the H sweep is local to the column block."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        rng = np.random.default_rng(79)
        A = rng.random((m, n)); W = rng.random((m, k)); Hfull = rng.random((k, n)) * (2.0 / k)
        c0, c1 = column_block(n, rank, WORLD)
        Al, H = A[:, c0:c1], Hfull[:, c0:c1].copy()
        HHt = _allreduce(H @ H.T); AHt = _allreduce(Al @ H.T)
        metrics = []
        for it in range(iters):
            for c in range(k):                                           # UpdateW_Hals, replicated
                w = W[:, c] + (AHt[:, c] - W @ HHt[:, c]) / HHt[c, c]
                w[np.isnan(w) | (w < 0)] = 0.0
                if not w.any():
                    w[:] = np.finfo(np.float64).eps
                W[:, c] = w / np.linalg.norm(w)
            WtW = W.T @ W; WtA = W.T @ Al
            for r in range(k):                                           # UpdateH_Hals, local columns
                h = H[r] + (WtA[r] - WtW[r] @ H) / WtW[r, r]
                h[np.isnan(h) | (h < 0)] = 0.0
                H[r] = h
            gradH = WtW @ H - WtA
            HHt = _allreduce(H @ H.T); AHt = _allreduce(Al @ H.T)
            gradW = W @ HHt - AHt
            pg_h = _allreduce(np.array([_pg_sq(gradH, H)]))[0]
            metrics.append(float(np.sqrt(_pg_sq(gradW, W) + pg_h)))
        Hall = [torch.empty((k, column_block(n, r, WORLD)[1] - column_block(n, r, WORLD)[0]), dtype=torch.float64) for r in range(WORLD)]
        dist.all_gather(Hall, torch.from_numpy(np.ascontiguousarray(H)))
        if rank == 0:
            q.put((W, np.concatenate([h.numpy() for h in Hall], axis=1), metrics))
    finally:
        dist.destroy_process_group()


def test_hals_all_reduce_exchange_matches_single_process_oracle():
    from oracle import Oracle
    m, n, k, iters = 64, 50, 6, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_hals, args=(r, port, m, n, k, iters, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    W, H, metrics = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(79)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n)) * (2.0 / k)
    o = Oracle().nmf_dense(A, W0, H0, alg="HALS", tol=1e-12, min_iter=1, max_iter=iters, normalize=False, trace=True)
    assert np.linalg.norm(W - o["W"]) <= 1e-9 * np.linalg.norm(o["W"])
    assert np.linalg.norm(H - o["H"]) <= 1e-9 * np.linalg.norm(o["H"])
    ratios = np.array(metrics[1:]) / metrics[0]
    assert np.allclose(ratios, o["metrics"][1:iters], rtol=1e-5)       # the HALS metric is discontinuous in the last ulp


def test_rank2_all_reduce_exchange_matches_single_process_oracle():
    from oracle import Oracle
    m, n, iters = 70, 48, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_rank2, args=(r, port, m, n, iters, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    W, H, metrics = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(78)
    A = rng.random((m, n)); W0 = rng.random((m, 2)); H0 = rng.random((2, n))
    o = Oracle().nmf_dense(A, W0, H0, alg="RANK2", tol=1e-12, min_iter=1, max_iter=iters, normalize=False, trace=True)
    assert np.linalg.norm(W - o["W"]) <= 1e-9 * np.linalg.norm(o["W"])
    assert np.linalg.norm(H - o["H"]) <= 1e-9 * np.linalg.norm(o["H"])
    ratios = np.array(metrics[1:]) / metrics[0]
    assert np.allclose(ratios, o["metrics"][1:iters], rtol=1e-8)


@pytest.mark.parametrize("alg,m,n,k", [("MU", 61, 40, 5), ("BPP", 50, 36, 6)])
def test_column_sharded_exchange_matches_single_process_oracle(alg, m, n, k):
    from oracle import Oracle
    iters = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, port, alg, m, n, k, iters, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    W, H, metrics = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(77)
    A = rng.random((m, n)); W0 = rng.random((m, k)); H0 = rng.random((k, n))
    o = Oracle().nmf_dense(A, W0, H0, alg=alg, tol=1e-12, min_iter=1, max_iter=iters, normalize=False, trace=True)
    assert np.linalg.norm(W - o["W"]) <= 1e-9 * np.linalg.norm(o["W"])
    assert np.linalg.norm(H - o["H"]) <= 1e-9 * np.linalg.norm(o["H"])
    ratios = np.array(metrics[1:]) / metrics[0]
    assert np.allclose(ratios, o["metrics"][1:iters], rtol=1e-8)


def test_partition_helpers_cover_everything_once():
    for n, world in ((20000, 8), (7, 4), (5, 8), (33, 2)):
        blocks = [column_block(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        slices = [row_slice(n, r, world) for r in range(world)]
        assert sum(s[1] for s in slices) == n
        assert len({s[2] for s in slices}) == 1 and slices[0][2] * world >= n
        assert all(s[0] == r * s[2] for r, s in enumerate(slices))


def test_nnz_balanced_column_blocks():
    rng = np.random.default_rng(5)
    for n, world in ((1000, 8), (37, 4), (5, 8), (200000, 8)):
        lens = np.minimum((1000 ** rng.random(n)).astype(np.int64), 900)            # skewed column lengths
        colp = np.concatenate([[0], np.cumsum(lens)])
        blocks = [column_block_by_nnz(colp, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert all(b[0] <= b[1] for b in blocks)
        share = np.array([colp[b[1]] - colp[b[0]] for b in blocks], dtype=np.float64)
        assert share.sum() == colp[-1]
        if n >= 1000:                                                                 # no rank is off its share by more than one column
            assert np.all(np.abs(share - colp[-1] / world) <= lens.max())
            even = np.array([colp[column_block(n, r, world)[1]] - colp[column_block(n, r, world)[0]] for r in range(world)])
            assert share.max() <= even.max() + lens.max()
    empty = np.zeros(11, dtype=np.int64)
    assert [column_block_by_nnz(empty, r, 2) for r in range(2)] == [column_block(10, r, 2) for r in range(2)]
