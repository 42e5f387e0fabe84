"""bench.py --workload c3 | c4: the sparse configurations of BASELINE.json measured with the same JSON contract as the
headline line (bench.py's default, C2). One GPU; not part of the driver's default run.

  c3  sparse HALS NMF on the synthetic tf-idf-like CSC 1,000,000 x 200,000, k = 128
      step      = one outer iteration: Solver_Generic_HALS_Da::operator() + the progress update
      roofline  = the two sparse products (W'A over the CSC, H A' over the CSR): algorithmic bytes per launch
                  (12 nnz + 4 (cols + 1) index/value bytes + the dense operand read once + the output written once)
                  / measured launch time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  c4  hierclust (HierNMF2) on the synthetic ~320k-node graph, 64 leaves
      step      = one rank-2 outer iteration; value = iterations / time inside the factorizations,
                  e2e = iterations / wall time of the whole Clust call on host CSC arrays (upload, extraction,
                  random inits, factorizations, priority scores, tree)
      roofline  = one rank-2 iteration on the root matrix: 2 (12 nnz + 4 (n + 1)) + 128 (m + n) bytes / device time
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "nmf_outer_iterations_per_second"
UNIT = "iter/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
        except Exception:
            pass
    return 6550.0, "fallback: B200_PROFILING.md measured copy bandwidth (MEASURED_PEAKS.json absent)"


def gather_peak():
    """What random gathers of 1 KB vectors sustain on this GPU, measured now by tools/l2_gather_peak (no index stream, no
    arithmetic): (TB/s from an L2-resident 48 MB table, TB/s from a 1 GB table, source). The first is the ceiling of a gather
    SpMM whose operand (or operand slab) stays in L2, the second that of one whose operand does not, uniform popularity."""
    import subprocess
    exe = os.path.join(ROOT, "tools", "l2_gather_peak")
    l2 = big = None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        for line in out.splitlines():
            if line.startswith("GATHER table=") and "1KB-vectors" in line:
                mb = int(line.split("table=")[1].split("MB")[0])
                rates = [float(tok.split("GB/s")[0].split()[-1]) for tok in line.split("|")]
                if mb == 48:
                    l2 = max(rates) * 1e-3
                if mb == 1024:
                    big = max(rates) * 1e-3
    except Exception:
        pass
    if l2 is None:
        return 19.7, 7.7, "fallback: profiles/gather_peak_r02.txt (tools/l2_gather_peak could not run)"
    return l2, big, "measured now: tools/l2_gather_peak (random 1 KB vectors, LDG.256, 32-64 warps/SM)"


def _clock_sampler(index):
    import bench
    s = bench.ClockSampler(index)
    s.start()
    return s


def run_c3(args, env=None, steps=None, warmup=None):
    """Returns the JSON line (a dict). With env (bench.Env) and more than one rank: A and H are split into contiguous column
    blocks holding equal shares of the stored entries (smallk_b200.sharding.column_block_by_nnz), W is replicated (the in-sweep
    column norms of HALS couple its rows), H A' is summed over the ranks by the library's exchange kernels."""
    import torch
    import smallk_b200 as sk
    import workloads
    from smallk_b200.sharding import column_block_by_nnz
    steps = steps or args.steps
    warmup = warmup or args.warmup
    m, n, per_col, k = 1000000, 200000, 500, 128
    world = env.world if env is not None else 1
    rank = env.rank if env is not None else 0
    local = env.local if env is not None else 0
    torch.cuda.set_device(local)
    colp, rowi, val = workloads.c3_tfidf_csc(m, n, per_col, device=f"cuda:{local}")
    nnz = int(colp[-1])
    val_sum = float(val.sum())
    c0, c1 = column_block_by_nnz(colp, rank, world)
    if world > 1:
        e0, e1 = int(colp[c0]), int(colp[c1])
        colp_l = (colp[c0:c1 + 1].astype(np.int64) - e0).astype(np.uint32)
        rowi_l, val_l = rowi[e0:e1], val[e0:e1]
    else:
        colp_l, rowi_l, val_l = colp, rowi, val
    n_loc, nnz_loc = c1 - c0, int(colp_l[-1])
    if env is not None:
        ctx, stream = env.ctx, env.stream
    else:
        ctx = sk.Context(0)
        stream = torch.cuda.current_stream()
        ctx.set_stream(stream.cuda_stream)
    ctx.load_csc((m, n_loc), colp_l, rowi_l, val_l)
    W0 = np.asfortranarray(np.random.default_rng(22).random((m, k)))
    # H0 scaled so that mean(W0*H0) = mean(A): from W0*H0 >> A, HALS clamps whole factors to zero in its first sweep
    H0 = np.asfortranarray(np.random.default_rng(23).random((k, n))[:, c0:c1]) * workloads.hals_h0_scale(val_sum, m, n, k)
    opts = sk.make_options(m, n, k, algorithm="HALS", tol=1e-15, min_iter=1, max_iter=warmup + steps + 8, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    if env is not None:
        ms_per_step, trace, launches, clocks, phases = env.timed_run(ctx, warmup, steps, sample_clocks=(world == 1))
    else:
        import bench
        trace = list(ctx.solver_run(warmup))
        sampler = _clock_sampler(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        trace += list(ctx.solver_run(steps))
        e1.record(stream)
        torch.cuda.synchronize()
        ms_per_step = e0.elapsed_time(e1) / steps
        launches = ctx.last_step()[1]
        clocks = sampler.stop()
        phases = {name: v / (steps + warmup) for name, v in ctx.phase_report().items()} if os.environ.get("SMK_PHASES") else None
    metric = trace[-1]
    t_wta = ctx.time_product(0, reps=5)
    t_hat = ctx.time_product(1, reps=5)
    peak, peak_src = hbm_peak()
    bytes_wta = 12.0 * nnz_loc + 4.0 * (n_loc + 1) + 8.0 * k * (m + n_loc)       # A once, Wt read once, W'A written once
    bytes_hat = 12.0 * nnz_loc + 4.0 * (m + 1) + 8.0 * k * (n_loc + m)
    achieved = (bytes_wta + bytes_hat) / ((t_wta + t_hat) * 1e-3) * 1e-9
    B_iter = 2 * (12.0 * nnz + 4 * (n + 1)) + 64.0 * k * (m + n)      # SURVEY section 8(d) compulsory bytes per iteration (whole job)
    # every stored entry gathers one k-vector of the dense operand through L2: 8 k nnz bytes per product; the rate at which the
    # memory system delivers gathered vectors to the SMs is what bounds a gather SpMM at this k, not the compulsory HBM bytes
    l2_cap, big_cap, cap_src = gather_peak() if rank == 0 else (None, None, None)
    traffic = None
    for tname in ("ncu_r02_c3_spmm_traffic.json", "ncu_r01_c3_spmm_traffic.json"):
        tp = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tp) and world == 1:
            try:
                tj = json.load(open(tp))
                import bench
                sha_now = bench.kernel_sha16(*bench.SPMM_KERNEL_SPAN)
                # the capture describes the kernel source it was taken from: refused once spmm.cu has changed
                traffic = float(tj["dram_bytes_both_products"]) if tj.get("kernel_source_sha16") in (None, sha_now) else None
            except Exception:
                traffic = None
            break
    # ---- the other sparse solver of the metric ("BPP/HALS, dense+sparse"): BPP on the same matrix, one GPU ----------------
    variants = None
    if world == 1:
        variants = {}
        try:
            ob = sk.make_options(m, n, k, algorithm="BPP", tol=1e-15, min_iter=1, max_iter=1 + 4, normalize=False)
            ctx.solver_begin(W0, H0, ob)
            tr_b = list(ctx.solver_run(1))
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            b0.record(stream)
            tr_b += list(ctx.solver_run(4))
            b1.record(stream)
            torch.cuda.synchronize()
            ms_b = b0.elapsed_time(b1) / 4
            variants["sparse_bpp"] = {"metric": METRIC, "value": 1000.0 / ms_b, "unit": UNIT, "ms_per_step": ms_b, "steps": 4, "warmup": 1,
                                      "gpu_launches": ctx.last_step()[1],
                                      "config": {"workload": f"sparse BPP NMF {m}x{n} nnz={nnz} k={k} (C3's matrix, Solver_Generic_BPP)",
                                                 "algorithm": "BPP", "k": k},
                                      "progress_metric_last": tr_b[-1]}
        except Exception as ex:                       # a variant must never cost the C3 line
            variants["sparse_bpp"] = {"error": f"{type(ex).__name__}: {ex}"}
    if env is None:
        ctx.close()
        del ctx

    e2e = None
    if not args.no_e2e and world == 1:
        ctx2 = sk.Context(local)
        W = W0.copy(order="F"); H = H0.copy(order="F")
        o2 = sk.make_options(m, n, k, algorithm="HALS", tol=1e-15, min_iter=steps, max_iter=steps, normalize=False)
        t0 = time.perf_counter()
        ctx2.load_csc((m, n), colp, rowi, val)
        ctx2.nmf(W, H, o2)
        t_e2e = time.perf_counter() - t0
        e2e = {"value": steps / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": (12.0 * nnz + 4.0 * (n + 1) + 8.0 * k * (m + n)) / steps,
               "d2h_bytes_per_step": 8.0 * k * (m + n) / steps, "seconds": t_e2e,
               "note": f"smk_load_csc (CSR and segment tables built on the device) + smk_nmf ({steps} iterations) on host arrays; wall clock"}
        ctx2.close()

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import Ref
            ms_, ns_ = 100000, 20000
            cp, ri, va = workloads.tfidf_csc_numpy(ms_, ns_, 50, 21)
            Ws = np.asfortranarray(np.random.default_rng(22).random((ms_, k)))
            Hs = np.asfortranarray(np.random.default_rng(23).random((k, ns_))) * workloads.hals_h0_scale(va.sum(), ms_, ns_, k)
            cores = os.cpu_count() or 1
            o = Ref(blas_threads=cores).nmf_sparse((ms_, ns_), cp, ri, va, Ws, Hs, alg="HALS", tol=1e-15, min_iter=3, max_iter=3, max_threads=cores, timed=True)
            sec = o["elapsed_us"] * 1e-6 / 3
            f_full = 4.0 * k * nnz + 6.0 * k * k * (m + n)
            f_s = 4.0 * k * int(cp[-1]) + 6.0 * k * k * (ms_ + ns_)
            cpu = {"value": 1.0 / (sec * f_full / f_s), "unit": UNIT, "cores": cores, "kind": "reference",
                   "measured_iter_per_s_on_sample": 1.0 / sec,
                   "sample": f"reference sparse HALS (oracle/_ref, {cores} threads) on the same generator at {ms_}x{ns_} (nnz {int(cp[-1])}: same density; the "
                             f"matrix of the scale_c3r_hals parity fixture), 3 iterations incl. Init; `value` scales its seconds to the full size by the "
                             f"flop ratio {f_full / f_s:.1f} (the full matrix does not fit a few-minute CPU run: 128 Gemv passes over a 1 GB W per iteration)"}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}

    line = {"metric": METRIC, "value": 1000.0 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"sparse HALS NMF {m}x{n} nnz={nnz} k={k} (BASELINE configs[2], SURVEY C3)", "algorithm": "HALS", "k": k,
                       "sharding": f"columns [{c0}, {c1}) of {n} on rank {rank} of {world}: {nnz_loc} stored entries (balanced by nnz)",
                       "l2": "inputs larger than L2 (CSC + CSR 2.8 GB, W 1 GB)", "step": "solver() + PG_RATIO progress update, enqueued by one smk_solver_run call"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "spmm_seg_wide256_kernel (W'A) + spmm_seg_slab_kernel x4 (H A'): algorithmic bytes of both products / both launch times",
                         "algorithmic_bytes": bytes_wta + bytes_hat, "launch_ms": {"WtA": t_wta, "HAt": t_hat},
                         "gathered_TBs": {"WtA": nnz_loc * k * 8 / t_wta * 1e-9, "HAt": nnz_loc * k * 8 / t_hat * 1e-9},
                         "l2_gather_bound": None if not l2_cap else {
                             "cap_TBs": l2_cap, "cap_1GB_table_TBs": big_cap, "cap_source": cap_src,
                             "frac_WtA": nnz_loc * k * 8 / t_wta * 1e-9 / l2_cap, "frac_HAt": nnz_loc * k * 8 / t_hat * 1e-9 / l2_cap,
                             "note": "a gather SpMM moves 8*k*nnz bytes of dense operand from L2 to the SMs per product (no on-chip reuse exists at 0.05 % "
                                     "density): the ceiling is the gather rate of the memory system (cap_TBs: table resident in L2, as the 51 MB slabs of H "
                                     "are; cap_1GB_table_TBs: uniform gathers from a 1 GB table, W's size, whose Zipf popularity puts W'A between the two)"},
                         "peak_source": peak_src,
                         "step_compulsory_GB": B_iter * 1e-9, "step_frac_of_peak": B_iter / world / (ms_per_step * 1e-3) * 1e-9 / peak},
            "cpu_baseline": cpu, "progress_metric_last": metric}
    if phases:
        line["phases_ms_per_step"] = phases
    if variants:
        line["variants"] = variants
    return line


def run_c4(args):
    import torch
    import smallk_b200 as sk
    import workloads
    n, edges, clusters = 320000, 2000000, 64
    torch.cuda.set_device(0)
    colp, rowi, val = workloads.c4_graph(n, edges)
    nnz = int(colp[-1])
    kw = dict(csc=(colp, rowi, val), shape=(n, n), tol=1e-4, min_iter=5, max_iter=5000, seed=32)
    for _ in range(max(1, args.warmup // 3)):
        sk.hierclust(num_clusters=4, **kw)                       # warm-up: allocations, lazy module load
    sampler = _clock_sampler(0)
    out = sk.hierclust(num_clusters=clusters, **kw)
    clocks = sampler.stop()
    iters = int(out["iterations"])
    prof = out["profile"]

    # one rank-2 iteration on the root matrix, operands hot in L2, no host synchronisation in between
    ctx = sk.Context(0)
    ctx.load_csc((n, n), colp, rowi, val)
    rng = np.random.default_rng(1)
    opts = sk.make_options(n, n, 2, algorithm="RANK2", tol=1e-15, min_iter=1, max_iter=100000, normalize=False)
    ctx.solver_begin(rng.random((n, 2)), rng.random((2, n)), opts)
    ctx.solver_step(20)
    ctx.solver_step(200)
    ms200, l200 = ctx.last_step()
    ctx.close()
    root_us = ms200 / 200 * 1e3
    B_root = 2 * (12.0 * nnz + 4 * (n + 1)) + 128.0 * (n + n)
    peak, peak_src = hbm_peak()
    achieved = B_root / (root_us * 1e-6) * 1e-9

    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import Ref
            ns, es, cs = 40000, 250000, 16
            cp, ri, va = workloads.c4_graph(ns, es)
            kws = dict(csc=(cp, ri, va), shape=(ns, ns), num_clusters=cs, tol=1e-4, min_iter=5, max_iter=5000, seed=32)
            g = sk.hierclust(**kws)
            t0 = time.perf_counter()
            r = Ref().hierclust(max_threads=1, **kws)
            ref_s = time.perf_counter() - t0
            same = bool(np.array_equal(g["assignments"], r["assignments"]))
            cpu = {"value": int(g["iterations"]) / ref_s, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"reference ClustSparse (oracle/_ref) on the {ns}-node graph of the same generator, {cs} leaves, one thread (its "
                             f"fastest setting and the one with the sequential initialiser): {ref_s:.2f} s for the {int(g['iterations'])} rank-2 "
                             f"iterations this library ran in {g['elapsed_s']:.2f} s; identical assignments: {same}"}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"unavailable: {ex}"}

    line = {"metric": METRIC, "value": iters / prof["factor_s"], "unit": UNIT, "n_gpus": 1, "steps": iters, "warmup": args.warmup,
            "ms_per_step": prof["factor_s"] / iters * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"hierclust rank-2 NMF, {n} nodes / {nnz} stored entries, {clusters} leaves (BASELINE configs[3], SURVEY C4)",
                       "algorithm": "RANK2", "k": 2, "nmf_count": int(out["nmf_count"]),
                       "l2": "node matrices fit L2 below the root; the run is latency-bound, not cache-bound",
                       "step": "one rank-2 outer iteration inside smk_nmf (solver + progress update + stop test), summed over all node factorizations"},
            "clocks": clocks,
            "e2e": {"value": iters / out["elapsed_s"], "unit": UNIT, "h2d_bytes_per_step": (12.0 * nnz + 4 * (n + 1)) / iters,
                    "d2h_bytes_per_step": 0.0, "seconds": out["elapsed_s"], "profile_s": prof,
                    "note": "ClustSparse on host CSC arrays: upload, per-node extraction, random inits, factorizations (factors up / down per node), "
                            "priority scores, tree; wall clock of the whole call"},
            "gpu_launches": 3 * iters,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "rank2_h_kernel + rank2_w_kernel + rank2_grad_kernel: one iteration on the root matrix",
                         "algorithmic_bytes": B_root, "launch_us": root_us, "launches_per_iteration": l200 / 200, "peak_source": peak_src},
            "cpu_baseline": cpu}
    return line
