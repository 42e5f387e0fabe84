#!/usr/bin/env python
"""bench.py — NMF outer iterations / second on B200 (BASELINE.json metric). Headline: dense BPP at C2.

Workload of the main line (BASELINE.json configs[1]): dense uniform-random A 20000 x 20000 FP64,
k = 64, BPP, W0/H0 injected. A is the matrix of the at-scale parity fixture (tools/workloads.dense_columns: PCG64(11), any
column block can be generated on its own), so every rank count, the reference arm and tests/golden/scale_c2_bpp.npz (the
reference's own code run on it) see the same numbers. One "step" = one outer iteration = one solver(A, W, H, gradW, gradH)
call plus its progress-metric update (common/include/nmf_solve_generic.hpp:70-98).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras] [--workload c2|c3|c4]

* value     : outer iterations / s with A, W, H resident in HBM: the K steps are enqueued by ONE smk_solver_run call (no
              host synchronisation between iterations; the metric of every iteration is computed on the device), CUDA events on
              the launching stream, barrier + synchronize on both sides, max over ranks.
* parity    : the progress metric of the first iterations against the reference's own trace on this matrix
              (tests/golden/scale_c2_bpp.npz, 1e-9), at every rank count; later iterations against the committed single-GPU trace.
* e2e       : the same metric through the host-buffer call a user of the reference makes (Nmf(opts, A, W, H): smk_load_dense +
              smk_nmf on pinned host buffers): upload of A, W0, H0, K iterations, download of W, H, all inside the timed region.
* roofline  : the dominant kernel (gemm_tma_kernel, the two A-sized contractions) against the FP64 tensor-pipe peak
              measured on this GPU by tools/dmma_peak (MEASURED_PEAKS.json has no FP64 entry).
* cpu_baseline / --impl reference : the reference's own sources (oracle/_ref: reference code + El.hpp shim + the venv's OpenBLAS)
              on the box's host cores, on the FULL workload, iteration count bounded.
* extra     : the other BASELINE configurations, each with its own value / roofline / e2e / cpu_baseline where they apply:
              c1 (nmf CLI, 256 x 256, k = 16), c5 (dense BPP 100000 x 50000, k = 256; any N), c3 (sparse HALS 1e6 x 2e5,
              k = 128; any N, column blocks balanced by nnz), c4 (hierclust, 320k nodes, 64 leaves; N = 1), preprocess (the
              tf-idf pipeline upstream of the sparse path; N = 1).
N > 1: A and H are sharded by column block, one process per GPU (torchrun); the exchanges (k x k Grams, the k x m product
H*A', the row blocks of W) are the library's own kernels over NVLink peer memory (csrc/peer.cu). Total work is fixed ("strong").
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "nmf_outer_iterations_per_second"
UNIT = "iter/s"
K_RANK = 64
SEED_A, SEED_W, SEED_H = 11, 12, 13
GOLD = os.path.join(ROOT, "tests", "golden")


def flops_per_iter(m, n, k):
    # F = 4kmn + 6k^2 n + 4k^2 m (NNLS work excluded)
    return 4.0 * k * m * n + 6.0 * k * k * n + 4.0 * k * k * m


class ClockSampler:
    """SM clock and throttle reasons during the timed region, sampled in-process through NVML every 5 ms
    (the nvidia-smi CLI takes longer per call than a short timed region lasts)."""

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        phys = int(ids[index]) if index < len(ids) else index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.sm, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in names:
                    if bits & int(bit):
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=3)
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "how": "NVML in-process, 5 ms period"}


def measured_fp64_peak():
    """FP64 tensor-pipe peak, measured now with tools/dmma_peak (register-resident DMMA.8x8x4 loop)."""
    exe = os.path.join(ROOT, "tools", "dmma_peak")
    best = None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        for line in out.splitlines():
            if line.startswith("DMMA") and "TFLOP/s" in line:
                v = float(line.split(":")[1].split("TFLOP/s")[0])
                if v < 100.0:       # a failed launch prints nonsense
                    best = v if best is None else max(best, v)
    except Exception:
        pass
    if best is None:
        return 36.9, "fallback: profiles/dmma_peak_r01.txt (tools/dmma_peak could not run)"
    return best, "measured now: tools/dmma_peak (DMMA.8x8x4 register loop); MEASURED_PEAKS.json has no FP64 entry"


def file_sha16(path):
    try:
        return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
    except OSError:
        return None


def kernel_sha16(path, begin, end):
    """sha of the DEVICE code a traffic capture describes: the text of `path` from the line starting with `begin` up to the line
    starting with `end` (host-side dispatch edits do not invalidate an ncu capture; a change to the kernel does)."""
    try:
        text = open(path).read()
        a = text.index("\n" + begin)
        b = text.index("\n" + end, a + 1)
        return hashlib.sha256(text[a:b].encode()).hexdigest()[:16]
    except (OSError, ValueError):
        return None


# the kernels the committed traffic files describe
GEMM_KERNEL_SPAN = (os.path.join(ROOT, "smallk_b200", "csrc", "gemm_f64.cu"), "// What every GEMM kernel of this file does", "// C = sum_s partial[s] (- D)")
SPMM_KERNEL_SPAN = (os.path.join(ROOT, "smallk_b200", "csrc", "spmm.cu"), "// How the loops are written, and why", "// out(:, j) = beta * out(:, j) + sum of the partials")


# ---------------------------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(m, n, k, warm, timed, threads, layouts, seeds=(SEED_A, SEED_W, SEED_H), alg="BPP", h0_scale=1.0):
    """The reference's own NmfSolve (Solver_Generic_BPP, or the solver `alg` names) on the FULL m x n workload: `warm` + `timed` iterations per thread
    layout, timed per iteration by stamps taken inside its loop. Returns (mean seconds per timed iteration of the faster
    layout, its BLAS thread count, BLAS backend, per-layout seconds)."""
    import ctypes
    import workloads
    from oracle import Ref, REF_SO, ALG
    if not os.path.exists(REF_SO):
        raise RuntimeError("oracle/_ref/libsmallk_ref.so is missing (built only where /root/reference exists)")
    ref = Ref()
    A_t = workloads.dense_columns(m, 0, n, seed=seeds[0])          # C-ordered (n, m) == column-major m x n
    W0 = np.asfortranarray(np.random.default_rng(seeds[1]).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(seeds[2]).random((k, n))) * h0_scale
    dp = ctypes.POINTER(ctypes.c_double)
    iters = warm + timed
    out = {}
    # two thread layouts: threaded BLAS (the big products) + OpenMP for the per-column solves, or OpenMP only
    for blas_threads in layouts:
        ref.set_blas_threads(blas_threads)
        W = W0.copy(order="F"); H = H0.copy(order="F")
        stamps = np.zeros(iters)
        t0 = ctypes.c_double(0.0)
        it = ctypes.c_int(0)
        rc = ref.lib.ref_nmf_dense_stamped(ALG[alg], 0, m, n, k, iters, threads, A_t.ctypes.data_as(dp), m,
                                           W.ctypes.data_as(dp), m, H.ctypes.data_as(dp), k,
                                           ctypes.byref(it), stamps.ctypes.data_as(dp), ctypes.byref(t0))
        if rc != 0:
            raise RuntimeError(f"reference solver returned {rc}")
        per_iter = np.diff(np.concatenate([[t0.value], stamps]))      # [0] includes Solver::Init (the A' copy, W'W, W'A)
        out[blas_threads] = float(np.mean(per_iter[warm:]))
    best = min(out, key=out.get)
    return out[best], best, ref.blas_backend(), out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = n = args.size
    k = K_RANK
    cores = os.cpu_count() or 1
    # the FULL workload; what is bounded is the iteration count (a reference iteration at C2 takes seconds; Init alone ~10 s)
    warm = min(args.warmup, args.ref_max_warmup)
    timed = max(1, min(args.steps, args.ref_max_steps))
    layouts = (cores, 1) if args.ref_both_layouts else (cores,)
    t_begin = time.perf_counter()
    sec, blas_threads, backend, per_layout = cpu_reference_run(m, n, k, warm, timed, cores, layouts)
    value = 1.0 / sec
    sample = (f"reference Solver_Generic_BPP via NmfSolve (oracle/_ref: the reference's sources + El.hpp shim) on the FULL {m}x{n} "
              f"workload, {timed} timed iterations after {warm} warm-up (requested {args.steps}/{args.warmup}; a reference iteration "
              f"takes seconds), per-iteration stamps inside its loop, Init excluded; BLAS={backend}, OpenMP threads={cores}; "
              f"seconds per iteration by BLAS thread count: {per_layout}; reported: {blas_threads} BLAS threads; "
              f"whole arm {time.perf_counter() - t_begin:.0f} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"dense BPP NMF {m}x{n} FP64 k={k} (BASELINE configs[1], SURVEY C2)", "algorithm": "BPP", "k": k,
                       "requested_steps": args.steps, "requested_warmup": args.warmup},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """Rank bookkeeping + the one library context every workload of this run shares (one NCCL bootstrap)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import smallk_b200 as sk
        self.torch, self.dist, self.sk = torch, dist, sk
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.ctx = self.new_context()

    def new_context(self):
        ctx = self.sk.Context(self.local)
        ctx.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            uid = [ctx.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(self.rank, self.world, uid[0])
        return ctx

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_run(self, ctx, warm, steps, sample_clocks=False):
        """`warm` untimed iterations, then exactly `steps` timed ones (one smk_solver_run call). Returns
        (ms per step: max over ranks, all metrics, launches in the timed region, clocks, {phase: ms per step} with SMK_PHASES=1)."""
        torch = self.torch
        trace = list(ctx.solver_run(warm)) if warm > 0 else []
        phases_on = bool(os.environ.get("SMK_PHASES"))
        if phases_on:
            ctx.phase_report()                      # reset: only the timed steps are reported
        sampler = None
        if sample_clocks and self.rank == 0:
            sampler = ClockSampler(self.local)
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        trace += list(ctx.solver_run(steps))
        e1.record(self.stream)
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sampler is not None else None
        phases = {name: v / steps for name, v in ctx.phase_report().items()} if phases_on else None
        return ms / steps, trace, ctx.last_step()[1], clocks, phases


def parity_against_fixture(trace, fixture, n1_trace_file, tol=1e-9):
    """The metric of iteration i must be the reference's (first iterations: fixture made by oracle/_ref) and the
    single-GPU run's (all iterations: committed trace), to 1e-9, whatever the number of ranks."""
    out = {"tol": tol}
    ok = True
    fpath = os.path.join(GOLD, fixture)
    if os.path.exists(fpath):
        ref = np.load(fpath)["metrics"]
        nchk = min(len(ref), len(trace))
        err = float(np.max(np.abs(np.array(trace[:nchk]) - ref[:nchk]) / np.abs(ref[:nchk]))) if nchk else None
        out.update({"reference_fixture": f"tests/golden/{fixture} (oracle/_ref on this matrix)", "iterations_checked": nchk,
                    "max_rel_err_vs_reference": err})
        ok = ok and (err is not None and err <= tol)
    tpath = os.path.join(GOLD, n1_trace_file)
    if os.path.exists(tpath):
        t1 = np.array(json.load(open(tpath))["metrics"])
        nchk = min(len(t1), len(trace))
        err = float(np.max(np.abs(np.array(trace[:nchk]) - t1[:nchk]) / np.abs(t1[:nchk]))) if nchk else None
        out.update({"n1_trace": f"tests/golden/{n1_trace_file}", "n1_iterations_checked": nchk, "max_rel_err_vs_n1": err})
        ok = ok and (err is not None and err <= tol)
    out["ok"] = bool(ok)
    return out


def upload_dense_block(env, m, c0, c1, seed):
    """This rank's column block of the workload matrix: generated on the host (pinned), copied to the device.
    Returns (device tensor [n_loc, m], pinned host tensor)."""
    import workloads
    torch = env.torch
    n_loc = c1 - c0
    hA = torch.empty((n_loc, m), dtype=torch.float64, pin_memory=True)
    h = hA.numpy()
    for b0 in range(0, n_loc, 1000):
        b1 = min(b0 + 1000, n_loc)
        h[b0:b1] = workloads.dense_columns(m, c0 + b0, c0 + b1, seed=seed)
    A = torch.empty((n_loc, m), dtype=torch.float64, device=env.dev)      # row-major [n][m] == column-major m x n
    A.copy_(hA, non_blocking=True)
    torch.cuda.synchronize()
    return A, hA


def run_c2(env, args):
    import smallk_b200 as sk
    from smallk_b200.sharding import column_block
    torch = env.torch
    world, rank = env.world, env.rank
    m = n = args.size
    k = K_RANK
    c0, c1 = column_block(n, rank, world)
    n_loc = c1 - c0
    ctx = env.ctx
    A, hA = upload_dense_block(env, m, c0, c1, SEED_A)
    W0 = np.asfortranarray(np.random.default_rng(SEED_W).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(SEED_H).random((k, n))[:, c0:c1])
    ctx.load_dense_device(A.data_ptr(), m, m, n_loc)
    opts = sk.make_options(m, n, k, algorithm="BPP", prog="PG_RATIO", tol=1e-15, min_iter=1,
                           max_iter=args.warmup + args.steps, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    ms_per_step, trace, launches, clocks, phases = env.timed_run(ctx, args.warmup, args.steps, sample_clocks=True)
    value = 1000.0 / ms_per_step

    parity = None
    if m == 20000 and k == 64:
        parity = parity_against_fixture(trace, "scale_c2_bpp.npz", "bench_c2_trace_n1.json")
        if args.write_trace and rank == 0 and world == 1:
            json.dump({"what": "progress metric per iteration of bench.py's C2 workload on ONE GPU (python bench.py --write-trace)",
                       "metrics": trace}, open(os.path.join(GOLD, "bench_c2_trace_n1.json"), "w"))

    # ---- roofline of the dominant kernel: the two A-sized DMMA contractions -------------------------
    t_wta = ctx.time_product(0, reps=5)
    t_hat = ctx.time_product(1, reps=5)
    peak, peak_src = measured_fp64_peak() if rank == 0 else (None, None)
    flops_launch = 2.0 * k * m * n_loc
    achieved = flops_launch / (0.5 * (t_wta + t_hat)) * 1e-9
    # dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed ncu --set full capture;
    # it describes the full-size single-GPU product only, and only the kernel source it was taken from
    traffic, traffic_note = None, None
    if world == 1 and m == 20000 and n == 20000 and k == 64:
        for tname in ("ncu_r02_c2_gemm_traffic.json", "ncu_r01_c2_gemm_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if not os.path.exists(tpath):
                continue
            try:
                tj = json.load(open(tpath))
                sha_now = kernel_sha16(*GEMM_KERNEL_SPAN)
                if tj.get("kernel_source_sha16") in (None, sha_now):
                    traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
                    traffic_note = f"profiles/{tname}" + ("" if tj.get("kernel_source_sha16") else " (capture predates the sha check)")
                else:
                    traffic_note = f"profiles/{tname} is stale: the GEMM kernels in gemm_f64.cu changed since the capture"
            except Exception:
                pass
            break

    # ---- the other dense solver of the metric ("BPP/HALS, dense+sparse"): HALS on the same matrix, one GPU ---------------
    variants = None
    if world == 1 and not args.no_extras and m == 20000:
        variants = {}
        try:
            hs = 2.0 / k                        # mean(W0 * H0) = mean(A) for U[0,1) factors and entries (HALS note)
            oh = sk.make_options(m, n, k, algorithm="HALS", prog="PG_RATIO", tol=1e-15, min_iter=1, max_iter=3 + 20, normalize=False)
            ctx.solver_begin(W0, np.asfortranarray(H0 * hs), oh)
            ms_h, tr_h, l_h, _, ph_h = env.timed_run(ctx, 3, 20)
            fl_h = 4.0 * k * m * n + 6.0 * k * k * (m + n)           # two A-sized products; Gram, sweep and gradient per factor
            v = {"metric": METRIC, "value": 1000.0 / ms_h, "unit": UNIT, "ms_per_step": ms_h, "steps": 20, "warmup": 3, "gpu_launches": l_h,
                 "config": {"workload": f"dense HALS NMF {m}x{n} FP64 k={k} (C2's matrix, Solver_Generic_HALS_Da)", "algorithm": "HALS", "k": k},
                 "step_frac_of_fp64_tensor_peak": (fl_h / (ms_h * 1e-3) * 1e-12 / peak) if peak else None,
                 "progress_metric_last": tr_h[-1]}
            if ph_h:
                v["phases_ms_per_step"] = ph_h
            if not args.no_cpu_baseline:
                try:
                    cores = os.cpu_count() or 1
                    sec, bt, backend, _ = cpu_reference_run(m, n, k, 1, 2, cores, (cores,), alg="HALS", h0_scale=hs)
                    v["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "reference",
                                         "sample": f"reference HALS (oracle/_ref, BLAS={backend}, {bt} BLAS threads) on the FULL {m}x{n} workload: "
                                                   "2 timed iterations after 1 warm-up, Init excluded"}
                except Exception as ex:
                    v["cpu_baseline"] = {"value": None, "sample": f"unavailable: {ex}"}
            variants["dense_hals"] = v
        except Exception as ex:                       # a variant must never cost the headline line
            variants["dense_hals"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- end to end through the host-buffer API -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del A
        torch.cuda.empty_cache()
        hW = torch.from_numpy(np.ascontiguousarray(W0.T)).pin_memory()     # (k, m) C-order == (m, k) F-order
        hH = torch.from_numpy(np.ascontiguousarray(H0.T)).pin_memory()
        import ctypes
        dp = ctypes.POINTER(ctypes.c_double)
        lib = sk.load_library()
        o2 = sk.make_options(m, n, k, algorithm="BPP", prog="PG_RATIO", tol=1e-15,
                             min_iter=args.steps, max_iter=args.steps, normalize=False)
        st = sk.NmfStats()
        e2e_ms = []
        for rep in range(2):          # first call warms allocations
            hW.copy_(torch.from_numpy(np.ascontiguousarray(W0.T))); hH.copy_(torch.from_numpy(np.ascontiguousarray(H0.T)))
            env.barrier()
            t0 = time.perf_counter()
            rc = lib.smk_load_dense(ctx._h, ctypes.cast(hA.data_ptr(), dp), ctypes.c_longlong(m), m, n_loc)
            assert rc == 0, rc
            rc = lib.smk_nmf(ctx._h, ctypes.byref(o2), ctypes.cast(hW.data_ptr(), dp), m,
                             ctypes.cast(hH.data_ptr(), dp), k, ctypes.byref(st))
            assert rc == 0, (rc, lib.smk_last_error(ctx._h))
            env.barrier()
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        t_e2e = env.max_over_ranks(e2e_ms[-1])
        h2d = 8.0 * (m * n_loc + m * k + k * n_loc) * world / args.steps
        d2h = 8.0 * (m * k + k * n_loc) * world / args.steps
        e2e = {"value": args.steps / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "seconds": t_e2e * 1e-3,
               "note": f"smk_load_dense + smk_nmf ({args.steps} iterations, PG at iteration 1 only as in NmfSolve with "
                       f"min_iter = max_iter) on pinned host A/W/H; wall clock around the call. The {8e-9 * m * n_loc:.1f} GB upload of A "
                       f"per rank cannot overlap the iterations (every iteration reads all of A): e2e = upload + K steps"}
    else:
        del A
    del hA
    torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1): the reference on the full workload, iteration count bounded -------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            sec, blas_threads, backend, per_layout = cpu_reference_run(m, n, k, 1, 2, cores, (cores,))
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"reference BPP (oracle/_ref: reference sources + El.hpp shim, BLAS={backend}, {blas_threads} BLAS threads, "
                             f"{cores} OpenMP threads) on the FULL {m}x{n} workload: 2 timed iterations after 1 warm-up, Init excluded "
                             f"({time.perf_counter() - t0:.0f} s in all)"}
        except Exception as ex:      # the checker is optional on the box; say so rather than fail the bench
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense BPP NMF {m}x{n} FP64 k={k} (BASELINE configs[1], SURVEY C2)",
                   "algorithm": "BPP", "k": k, "sharding": f"A,H by column block over {world} GPU(s)",
                   "exchange": ("none (one GPU)" if world == 1 else
                                ("NCCL collectives (SMK_PEER=0)" if os.environ.get("SMK_PEER") == "0" else
                                 "own kernels over NVLink peer memory (csrc/peer.cu)")),
                   "l2": "inputs larger than L2 (A is %.1f GB per rank)" % (8e-9 * m * n_loc),
                   "step": "solver() + PG_RATIO progress update, all K steps enqueued by one smk_solver_run call (metric per iteration "
                           "computed on the device, no host synchronisation between iterations)"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_note,
                     "algorithmic_bytes_per_launch": 8.0 * (m * n_loc + k * (m + n_loc)),
                     "kernel": "gemm_tma_kernel (W'A and H A', 2*k*m*n flop per launch; tiles by TMA)",
                     "launch_ms": {"WtA": t_wta, "HAt": t_hat}, "peak_source": peak_src,
                     "step_frac_of_peak": flops_per_iter(m, n, k) / world / (ms_per_step * 1e-3) * 1e-12 / peak if peak else None},
        "cpu_baseline": cpu,
        "parity": parity,
        "progress_metric_last": trace[-1],
    }
    if phases:
        line["phases_ms_per_step"] = phases
    if variants:
        line["variants"] = variants
    return line


def run_c5(env, args):
    """BASELINE configs[4] / SURVEY C5: dense BPP 100000 x 50000, k = 256, column-sharded over the ranks of this run (one GPU
    holds the 40 GB matrix too). A is generated on the device, block-seeded so its bits do not depend on the rank count."""
    import smallk_b200 as sk
    from smallk_b200.sharding import column_block
    torch = env.torch
    m, n, k = 100000, 50000, 256
    warm, steps = 2, 5
    c0, c1 = column_block(n, env.rank, env.world)
    n_loc = c1 - c0
    A = torch.empty((n_loc, m), dtype=torch.float64, device=env.dev)
    BLK = 500
    g = torch.Generator(device=env.dev)
    for b0 in range((c0 // BLK) * BLK, c1, BLK):
        g.manual_seed(41 * 1000003 + b0)
        blk = torch.rand((BLK, m), dtype=torch.float64, device=env.dev, generator=g)
        lo, hi = max(b0, c0), min(b0 + BLK, c1)
        A[lo - c0:hi - c0] = blk[lo - b0:hi - b0]
        del blk
    W0 = np.asfortranarray(np.random.default_rng(42).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(43).random((k, n))[:, c0:c1])
    ctx = env.ctx
    ctx.load_dense_device(A.data_ptr(), m, m, n_loc)
    opts = sk.make_options(m, n, k, algorithm="BPP", prog="PG_RATIO", tol=1e-15, min_iter=1, max_iter=warm + steps, normalize=False)
    ctx.solver_begin(W0, H0, opts)
    ms_per_step, trace, launches, _, phases = env.timed_run(ctx, warm, steps)
    t_wta = ctx.time_product(0, reps=2)
    t_hat = ctx.time_product(1, reps=2)
    peak, peak_src = measured_fp64_peak() if env.rank == 0 else (None, None)
    achieved = 2.0 * k * m * n_loc / (0.5 * (t_wta + t_hat)) * 1e-9
    del A
    torch.cuda.empty_cache()
    out = {"metric": METRIC, "value": 1000.0 / ms_per_step, "unit": UNIT, "n_gpus": env.world, "steps": steps, "warmup": warm,
           "ms_per_step": ms_per_step, "scaling": "strong", "dtype": "f64", "data": "synthetic (device-generated, block-seeded)",
           "config": {"workload": f"dense BPP NMF {m}x{n} FP64 k={k} (BASELINE configs[4], SURVEY C5)", "algorithm": "BPP", "k": k,
                      "sharding": f"A,H by column block over {env.world} GPU(s)", "l2": "inputs larger than L2 (A is %.1f GB per rank)" % (8e-9 * m * n_loc)},
           "gpu_launches": launches,
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                        "traffic": None, "kernel": "gemm_tma_kernel (W'A and H A')", "launch_ms": {"WtA": t_wta, "HAt": t_hat},
                        "peak_source": peak_src,
                        "step_frac_of_peak": flops_per_iter(m, n, k) / env.world / (ms_per_step * 1e-3) * 1e-12 / peak if peak else None},
           "e2e": None, "e2e_note": "not measured: the reference cannot hold this matrix (32-bit offsets, SURVEY section 0) and a 40 GB "
                                    "pinned host copy is outside a few-minute run; C2's e2e covers the host-buffer path",
           "cpu_baseline": None, "cpu_baseline_note": "the reference overflows its 32-bit element offsets at this size; its timing at the "
                                                      "largest size it can index is in tests/golden/make_golden_scale.py (scale_c5r_bpp: 16 s per iteration on 8 cores)",
           "progress_metric_trace": trace}
    if phases:
        out["phases_ms_per_step"] = phases
    return out


def run_c1(env, args):
    """BASELINE configs[0] / SURVEY C1: the nmf command-line tool on a 256 x 256 uniform-random matrix, k = 16, BPP, tol 1e-4,
    initial factors from files. Reports iterations to converge and iterations / s of the tool's own timer, next to the
    reference's Nmf() on the same inputs (same iteration count = same trajectory)."""
    import ctypes
    import tempfile
    import smallk_b200 as sk
    host = sk.load_host_library()
    m = n = 256
    k = 16
    dp = ctypes.POINTER(ctypes.c_double)

    def rand2(seed, h1, w1, h2, w2):
        a = np.zeros((h1, w1), order="F"); b = np.zeros((h2, w2), order="F")
        host.smkh_random_matrices(seed, h1, w1, a.ctypes.data_as(dp), h2, w2, b.ctypes.data_as(dp), None)
        return a, b
    A, _ = rand2(1, m, n, 1, 1)                       # matrixgen --type UNIFORM: mt19937 U[0,1), column-major fill
    W0, H0 = rand2(2, m, k, k, n)                     # RandomMatrix for W, the same engine continuing for H
    with tempfile.TemporaryDirectory() as d:
        fa, fw, fh = (os.path.join(d, f) for f in ("a.csv", "w0.csv", "h0.csv"))
        np.savetxt(fa, A, delimiter=",", fmt="%.6e")  # what matrixgen writes: 6 significant digits (delimited_file.hpp:62-63)
        np.savetxt(fw, W0, delimiter=",", fmt="%.15e")
        np.savetxt(fh, H0, delimiter=",", fmt="%.15e")
        exe = os.path.join(ROOT, "smallk_b200", "bin", "nmf")
        cmd = [exe, "--matrixfile", fa, "--k", str(k), "--algorithm", "BPP", "--tol", "1e-4", "--miniter", "5", "--maxiter", "5000",
               "--infile_W", fw, "--infile_H", fh, "--outfile_W", os.path.join(d, "w.csv"), "--outfile_H", os.path.join(d, "h.csv"), "--verbose", "0"]
        runs = []
        for _ in range(3):                            # best of three (the first start of the tool pages the library in)
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError(r.stdout[-500:] + r.stderr[-500:])
            ms = float(r.stdout.split("Elapsed wall clock time:")[1].split("ms")[0])
            its = int(r.stdout.split("Iterations:")[1].split()[0])
            runs.append((ms, its, wall))
        A6 = np.loadtxt(fa, delimiter=",")
    ms, its, wall = min(runs)
    out = {"metric": METRIC, "value": its / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "iterations_to_converge": its, "ms_per_step": ms / its,
           "config": {"workload": "nmf CLI, dense 256x256 uniform-random (matrixgen recipe, %.6e), k=16, BPP, tol 1e-4, miniter 5, file-initialised "
                                  "W/H (BASELINE configs[0], SURVEY C1)", "tool": "smallk_b200/bin/nmf"},
           "e2e": {"value": its / wall, "unit": UNIT, "seconds": wall,
                   "note": "whole process: start-up, CUDA context, CSV parsing, upload, iterations, CSV output"},
           "roofline": None, "roofline_note": "latency-bound: 4.8e6 flop per iteration (SURVEY section 8(d): report iterations/s only)"}
    try:
        from oracle import Ref
        ref = Ref(blas_threads=1)
        # one thread: at 256 x 256 the reference is fastest single-threaded (16 OpenMP x 16 BLAS threads: 9.5 it/s, 240 s for this case)
        o = ref.nmf_dense(A6, W0, H0, alg="BPP", tol=1e-4, min_iter=5, max_iter=5000, normalize=True, timed=True, max_threads=1)
        out["cpu_baseline"] = {"value": o["iterations"] / (o["elapsed_us"] * 1e-6), "unit": UNIT, "cores": 1, "kind": "reference",
                               "iterations_to_converge": o["iterations"],
                               "sample": "the reference's Nmf() (oracle/_ref) on the same A, W0, H0, one thread (its fastest setting at this size): "
                                         "its own NmfStats timer"}
        out["parity"] = {"same_iteration_count_as_reference": bool(o["iterations"] == its)}
    except Exception as ex:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}
    return out


def run_preprocess(env, args):
    """the tf-idf preprocessing (preprocess_tf, preprocessor/src/preprocess.cpp:81-250) of a synthetic
    term-count matrix, device pipeline (smk_preprocess_tf, host arrays in / out) next to the reference's own function on one core
    (it is single-threaded). Index outputs must be identical."""
    import ctypes
    m, n, per_doc = 200000, 100000, 120
    rng = np.random.default_rng(51)
    t = np.minimum((float(m) ** rng.random((n, per_doc))).astype(np.int64), m - 1)          # Zipf(1) term draws per document
    t[n - 2000:] = t[rng.integers(0, n // 2, 2000)]                                         # 2 % exact copies of earlier documents
    key, cnt = np.unique((np.arange(n, dtype=np.int64)[:, None] * m + t).ravel(), return_counts=True)
    cols = key // m
    rows = (key - cols * m).astype(np.uint32)
    colptr = np.zeros(n + 1, dtype=np.uint32)
    colptr[1:] = np.cumsum(np.bincount(cols, minlength=n))
    counts = cnt.astype(np.float64)
    nnz = len(rows)
    ctx = env.ctx
    ctx.preprocess_tf(m, n, colptr, rows, counts)                                          # warm-up (allocations, CUB temp sizes)
    t0 = time.perf_counter()
    got = ctx.preprocess_tf(m, n, colptr, rows, counts, 1000, 3, 5)
    gpu_s = time.perf_counter() - t0
    out = {"metric": "term_count_entries_per_second", "value": nnz / gpu_s, "unit": "entries/s", "seconds": gpu_s, "n_gpus": 1,
           "config": {"workload": f"preprocess_tf on a synthetic {m} x {n} term-count matrix, {nnz} entries, docs_per_term 3, terms_per_doc 5",
                      "result": f"{got['m']} x {got['n']}, {len(got['rows'])} entries"},
           "e2e": {"value": nnz / gpu_s, "unit": "entries/s", "h2d_bytes_per_step": 16.0 * nnz + 4.0 * (n + 1),
                   "d2h_bytes_per_step": 16.0 * len(got["rows"]) + 4.0 * (got["m"] + 2 * got["n"]),
                   "note": "smk_preprocess_tf on host arrays: upload, sort, pruning rounds, scores, download; wall clock (value is this same number: "
                           "the pipeline has no device-resident entry point)"},
           "roofline": None, "roofline_note": "a handful of HBM passes over 8-byte (row, count) pairs per round; launch- and readback-bound at this size"}
    try:
        from oracle import REF_SO
        lib = ctypes.CDLL(REF_SO)
        up, dp = ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_double)
        om, on, onz = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
        oc = np.zeros(n + 1, dtype=np.uint32); orow = np.zeros(nnz, dtype=np.uint32); ocnt = np.zeros(nnz, dtype=np.uint32)
        osc = np.zeros(nnz); ti = np.zeros(m, dtype=np.uint32); di = np.zeros(n, dtype=np.uint32)
        sys.stdout.flush()
        saved = os.dup(1)                        # the reference narrates on stdout; this process's stdout carries the JSON line
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            t0 = time.perf_counter()
            rc = lib.ref_preprocess_tf(m, n, nnz, colptr.ctypes.data_as(up), rows.ctypes.data_as(up), counts.ctypes.data_as(dp), 1000, 3, 5,
                                       ctypes.byref(om), ctypes.byref(on), ctypes.byref(onz), oc.ctypes.data_as(up), orow.ctypes.data_as(up),
                                       ocnt.ctypes.data_as(up), osc.ctypes.data_as(dp), ti.ctypes.data_as(up), di.ctypes.data_as(up))
            ref_s = time.perf_counter() - t0
        finally:
            os.dup2(saved, 1); os.close(saved); os.close(devnull)
        same = bool(rc == 0 and om.value == got["m"] and on.value == got["n"] and np.array_equal(orow[:onz.value], got["rows"]) and
                    np.array_equal(oc[:on.value + 1], got["colptr"]) and np.array_equal(di[:on.value], got["doc_indices"]) and
                    np.array_equal(ti[:om.value], got["term_indices"]))
        out["cpu_baseline"] = {"value": nnz / ref_s, "unit": "entries/s", "seconds": ref_s, "cores": 1, "kind": "reference",
                               "sample": "the reference's preprocess_tf (oracle/_ref) on the same matrix, full size, one thread (the reference's is serial)"}
        out["parity"] = {"identical_pruned_matrix_and_index_maps": same,
                         "max_rel_score_diff": float(np.max(np.abs(osc[:onz.value] - got["scores"]) / np.abs(osc[:onz.value]))) if same else None}
    except Exception as ex:
        out["cpu_baseline"] = {"value": None, "unit": "entries/s", "cores": 1, "kind": "reference", "sample": f"unavailable: {ex}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smallk_b200")
    ap.add_argument("--size", type=int, default=20000, help="m = n of the dense workload (20000 = BASELINE C2)")
    ap.add_argument("--ref-max-steps", type=int, default=5, help="reference arm: timed iterations are capped at this (full-size workload)")
    ap.add_argument("--ref-max-warmup", type=int, default=2)
    ap.add_argument("--ref-both-layouts", type=int, default=1, help="reference arm: also time OpenMP-only (1 BLAS thread) and report the faster")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default): the headline dense BPP workload + the other configurations nested under `extra`; "
                         "c3 / c4: only the sparse HALS / hierclust configuration, one GPU (tools/bench_sparse.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="c1,c5,c3,c4,preprocess", help="which of the other configurations to nest under `extra`")
    ap.add_argument("--write-trace", action="store_true", help="N = 1: store the metric trace as tests/golden/bench_c2_trace_n1.json")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.workload != "c2":
        if int(os.environ.get("RANK", "0")) != 0:
            return                                   # one GPU: the other ranks of a torchrun launch have nothing to do
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is wired for the headline workload (c2); "
                              f"--workload {args.workload} reports the reference in its cpu_baseline"}), flush=True)
            return
        import bench_sparse
        line = (bench_sparse.run_c3 if args.workload == "c3" else bench_sparse.run_c4)(args)
        print(json.dumps(line), flush=True)
        return

    if args.impl == "reference":
        run_reference_arm(args)
        return

    env = Env()
    line = run_c2(env, args)
    if not args.no_extras and args.size == 20000:
        import bench_sparse
        extra = {}
        want = [w for w in args.extras.split(",") if w]
        for name in want:
            t0 = time.perf_counter()
            try:
                if name == "c5":
                    res = run_c5(env, args)
                elif name == "c3":
                    res = bench_sparse.run_c3(args, env=env, steps=10, warmup=3)
                elif name == "c1" and env.world == 1:
                    res = run_c1(env, args)
                elif name == "preprocess" and env.world == 1:
                    res = run_preprocess(env, args)
                elif name == "c4" and env.world == 1:
                    # its own process: hierclust's host driver is timed by wall clock, keep it clear of this process's leftovers
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "c4", "--steps", str(args.steps), "--warmup",
                                        str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []),
                                       capture_output=True, text=True, timeout=900)
                    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                    res = json.loads(lines[-1]) if lines else {"error": (r.stdout[-300:] + r.stderr[-600:])}
                else:
                    continue
            except Exception as ex:                  # an extra must never cost the headline line
                res = {"error": f"{type(ex).__name__}: {ex}"}
            if isinstance(res, dict):
                res["seconds_spent"] = time.perf_counter() - t0
            extra[name] = res
            env.torch.cuda.empty_cache()
        line["extra"] = extra
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    env.ctx.close()
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
