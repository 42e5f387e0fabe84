#!/usr/bin/env python
"""bench.py — NMF outer iterations / second on B200 (BASELINE.json metric), dense BPP at C2.

Workload (BASELINE.json configs[1]): dense uniform-random A 20000 x 20000 FP64,
k = 64, BPP, W0/H0 injected. One "step" = one outer iteration = one solver(A, W, H, gradW, gradH) call
plus its progress-metric update (common/include/nmf_solve_generic.hpp:70-98).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size M] [--workload c2|c3|c4]

* value     : outer iterations / s with A, W, H resident in HBM (CUDA events on the launching stream,
              barrier + synchronize on both sides, max over ranks).
* e2e       : the same metric through the host-buffer call a user of the reference makes
              (Nmf(opts, A, W, H): smk_load_dense + smk_nmf on pinned host buffers): upload of A, W0, H0,
              K iterations, download of W, H, all inside the timed region.
* roofline  : the dominant kernel (gemm_skinny_kernel, the two A-sized contractions) against the FP64
              tensor-pipe peak measured on this GPU by tools/dmma_peak (MEASURED_PEAKS.json has no FP64 entry).
* cpu_baseline / --impl reference : the reference's own sources (oracle/_ref, El.hpp shim + the venv's
              OpenBLAS) on the box's host cores, bounded sample.
N > 1: A and H are sharded by column block, one process per GPU (torchrun); H*H' and H*A' are all-reduced
over NCCL inside the library. Total work is fixed ("strong" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nmf_outer_iterations_per_second"
UNIT = "iter/s"
K_RANK = 64
SEED_A, SEED_W, SEED_H = 11, 12, 13


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


def make_inputs(m, n_local, col0, k):
    """U[0,1) synthetic inputs. A is generated per column block so every rank sees its slice of ONE matrix."""
    rng_w = np.random.default_rng(SEED_W)
    W0 = np.asfortranarray(rng_w.random((m, k)))
    return W0


def cpu_reference_run(m, n, k, iters, threads):
    """The reference's own NmfSolve (BPP) on the host: returns (seconds per iteration list, backend, cores)."""
    import ctypes
    from oracle import Ref, REF_SO
    if not os.path.exists(REF_SO):
        raise RuntimeError("oracle/_ref/libsmallk_ref.so is missing (built only where /root/reference exists)")
    ref = Ref()
    rng = np.random.default_rng(SEED_A)
    A = np.asfortranarray(rng.random((m, n)))
    W0 = np.asfortranarray(np.random.default_rng(SEED_W).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(SEED_H).random((k, n)))
    best = None
    # two thread layouts: OpenMP threads for the per-column solves + single-threaded BLAS, or threaded BLAS too
    for blas_threads in (threads, 1):
        ref.set_blas_threads(blas_threads)
        W = W0.copy(order="F"); H = H0.copy(order="F")
        stamps = np.zeros(iters)
        t0 = ctypes.c_double(0.0)
        it = ctypes.c_int(0)
        dp = ctypes.POINTER(ctypes.c_double)
        rc = ref.lib.ref_nmf_dense_stamped(3, 0, m, n, k, iters, threads, A.ctypes.data_as(dp), m,
                                           W.ctypes.data_as(dp), m, H.ctypes.data_as(dp), k,
                                           ctypes.byref(it), stamps.ctypes.data_as(dp), ctypes.byref(t0))
        if rc != 0:
            raise RuntimeError(f"reference solver returned {rc}")
        per_iter = np.diff(np.concatenate([[t0.value], stamps]))
        if best is None or per_iter[1:].mean() < best[0][1:].mean():
            best = (per_iter, blas_threads)
    return best[0], ref.blas_backend(), best[1]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = n = args.size
    k = K_RANK
    cores = os.cpu_count() or 1
    # bounded sample: the reference needs ~1e11 flop per iteration at C2; run W+K iterations of a column block
    # of the same matrix sized so that the run ends within minutes, and scale by the exact flop ratio.
    n_sample = min(n, args.ref_cols)
    iters = args.warmup + args.steps
    per_iter, backend, blas_threads = cpu_reference_run(m, n_sample, k, iters, cores)
    timed = per_iter[args.warmup:]
    sec_sample = float(np.mean(timed))
    scale = flops_per_iter(m, n, k) / flops_per_iter(m, n_sample, k)
    sec_full = sec_sample * scale
    value = 1.0 / sec_full
    sample = (f"reference Solver_Generic_BPP via NmfSolve on a {m}x{n_sample} column block of the workload, "
              f"{args.steps} timed iterations after {args.warmup} warm-up, seconds scaled by the flop ratio "
              f"{scale:.3f} to {m}x{n}; BLAS={backend} ({blas_threads} BLAS threads), OpenMP threads={cores}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"dense BPP NMF {m}x{n} FP64 k={k} (BASELINE configs[1])", "algorithm": "BPP", "k": k},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smallk_b200")
    ap.add_argument("--size", type=int, default=20000, help="m = n of the dense workload (20000 = BASELINE C2)")
    ap.add_argument("--ref-cols", type=int, default=1000, help="columns in the CPU sample of the reference arm")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default): the headline dense BPP workload; c3 / c4: the sparse HALS and hierclust configurations, one GPU "
                         "(tools/bench_sparse.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.workload != "c2":
        if int(os.environ.get("RANK", "0")) != 0:
            return                                   # one GPU: the other ranks of a torchrun launch have nothing to do
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is wired for the headline workload (c2); "
                              f"--workload {args.workload} reports the reference in its cpu_baseline"}), flush=True)
            return
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_sparse
        (bench_sparse.run_c3 if args.workload == "c3" else bench_sparse.run_c4)(args)
        return

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import smallk_b200 as sk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = n = args.size
    k = K_RANK
    # column block of this rank (contiguous, balanced)
    from smallk_b200.sharding import column_block
    c0, c1 = column_block(n, rank, world)
    n_loc = c1 - c0

    ctx = sk.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])

    # synthetic inputs: one global matrix, generated blockwise on the device (philox, seeded per column block
    # of 500 columns so the bits do not depend on the number of ranks)
    A = torch.empty((n_loc, m), dtype=torch.float64, device=dev)     # row-major [n][m] == column-major m x n
    BLK = 500
    g = torch.Generator(device=dev)
    for b0 in range((c0 // BLK) * BLK, c1, BLK):
        g.manual_seed(SEED_A * 1000003 + b0)
        blk = torch.rand((BLK, m), dtype=torch.float64, device=dev, generator=g)
        lo, hi = max(b0, c0), min(b0 + BLK, c1)
        A[lo - c0:hi - c0] = blk[lo - b0:hi - b0]
        del blk
    W0 = np.asfortranarray(np.random.default_rng(SEED_W).random((m, k)))
    H0 = np.asfortranarray(np.random.default_rng(SEED_H).random((k, n))[:, c0:c1])
    ctx.load_dense_device(A.data_ptr(), m, m, n_loc)
    opts = sk.make_options(m, n_loc if world > 1 else n, k, algorithm="BPP", prog="PG_RATIO", tol=1e-15, min_iter=1,
                           max_iter=args.warmup + args.steps, normalize=False)
    ctx.solver_begin(W0, H0, opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    metric_trace = []
    for _ in range(args.warmup):
        ctx.solver_step(1)
        metric_trace.append(ctx.solver_progress())

    sampler = None
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.solver_step(1)
        launches += ctx.last_step()[1] + 4          # + the progress reductions (2 x (partial + final))
        metric_trace.append(ctx.solver_progress())
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- roofline of the dominant kernel: the two A-sized DMMA contractions -------------------------
    t_wta = ctx.time_product(0, reps=5)
    t_hat = ctx.time_product(1, reps=5)
    peak, peak_src = measured_fp64_peak() if rank == 0 else (None, None)
    flops_launch = 2.0 * k * m * n_loc
    achieved = flops_launch / (0.5 * (t_wta + t_hat)) * 1e-9
    # dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed ncu --set full capture
    # (profiles/ncu_r01_c2_gemm_skinny.txt); it describes the full-size single-GPU product only
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_r01_c2_gemm_traffic.json")
    if os.path.exists(tpath) and world == 1 and m == 20000 and n == 20000 and k == 64:
        try:
            tj = json.load(open(tpath))
            traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
        except Exception:
            traffic = None

    # ---- end to end through the host-buffer API -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del A
        torch.cuda.empty_cache()
        hA = torch.empty((n_loc, m), dtype=torch.float64, pin_memory=True)
        rng = np.random.default_rng(SEED_A + 7 + rank)
        hA_np = hA.numpy()
        for b0 in range(0, n_loc, 1000):
            hA_np[b0:b0 + 1000] = rng.random((min(1000, n_loc - b0), m))
        hW = torch.from_numpy(np.ascontiguousarray(W0.T)).pin_memory()     # (k, m) C-order == (m, k) F-order
        hH = torch.from_numpy(np.ascontiguousarray(H0.T)).pin_memory()
        import ctypes
        dp = ctypes.POINTER(ctypes.c_double)
        lib = sk.load_library()
        o2 = sk.make_options(m, n_loc if world > 1 else n, k, algorithm="BPP", prog="PG_RATIO", tol=1e-15,
                             min_iter=args.steps, max_iter=args.steps, normalize=False)
        st = sk.NmfStats()
        e2e_ms = []
        for rep in range(2):          # first call warms allocations
            barrier()
            t0 = time.perf_counter()
            rc = lib.smk_load_dense(ctx._h, ctypes.cast(hA.data_ptr(), dp), ctypes.c_longlong(m), m, n_loc)
            assert rc == 0, rc
            rc = lib.smk_nmf(ctx._h, ctypes.byref(o2), ctypes.cast(hW.data_ptr(), dp), m,
                             ctypes.cast(hH.data_ptr(), dp), k, ctypes.byref(st))
            assert rc == 0, (rc, lib.smk_last_error(ctx._h))
            barrier()
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        t_e2e = e2e_ms[-1]
        if world > 1:
            t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        h2d = 8.0 * (m * n_loc + m * k + k * n_loc) * world / args.steps
        d2h = 8.0 * (m * k + k * n_loc) * world / args.steps
        e2e = {"value": args.steps / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "note": f"smk_load_dense + smk_nmf ({args.steps} iterations, PG at iteration 1 only as in NmfSolve with "
                       f"min_iter = max_iter) on pinned host A/W/H; wall clock around the call"}

    # ---- CPU baseline (rank 0, N = 1) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            n_sample = min(n, 600)
            per_iter, backend, blas_threads = cpu_reference_run(m, n_sample, k, 4, cores)
            sec = float(np.mean(per_iter[1:])) * flops_per_iter(m, n, k) / flops_per_iter(m, n_sample, k)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"reference BPP (oracle/_ref: reference sources + El.hpp shim, BLAS={backend}, "
                             f"{blas_threads} BLAS threads) on a {m}x{n_sample} column block, 3 timed iterations, "
                             f"scaled to {m}x{n} by the flop ratio"}
        except Exception as ex:      # the checker is optional on the box; say so rather than fail the bench
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"dense BPP NMF {m}x{n} FP64 k={k} (BASELINE configs[1], SURVEY C2)",
                       "algorithm": "BPP", "k": k, "sharding": f"A,H by column block over {world} GPU(s)",
                       "l2": "inputs larger than L2 (A is %.1f GB per rank)" % (8e-9 * m * n_loc),
                       "step": "solver() + PG_RATIO progress update (device reduction + 16-byte readback)"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": 8.0 * (m * n_loc + k * (m + n_loc)),
                         "kernel": "gemm_skinny_kernel (W'A and H A', 2*k*m*n flop per launch)",
                         "launch_ms": {"WtA": t_wta, "HAt": t_hat}, "peak_source": peak_src,
                         "step_frac_of_peak": flops_per_iter(m, n, k) / world / (ms_per_step * 1e-3) * 1e-12 / peak if peak else None},
            "cpu_baseline": cpu,
            "progress_metric_last": metric_trace[-1],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
