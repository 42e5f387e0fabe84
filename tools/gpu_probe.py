"""One-off GPU probe: FP64 cuBLAS rate, and the first timings of the dense BPP step at C2 scale."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smallk_b200 as sk

dev = torch.device("cuda:0")
out = {}
# cuBLAS DGEMM (the practical FP64 peak)
for N in (4096, 8192):
    a = torch.rand(N, N, dtype=torch.float64, device=dev); b = torch.rand(N, N, dtype=torch.float64, device=dev)
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(5):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"cublas_dgemm_{N}_tflops"] = 2 * N**3 / best * 1e-9
    del a, b
# skinny cuBLAS: W'A at C2
m = n = int(os.environ.get("PROBE_N", 20000)); k = 64
A = torch.rand(n, m, dtype=torch.float64, device=dev)     # row-major [n][m] == column-major m x n
Wt = torch.rand(m, k, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
best = 1e9
for _ in range(4):
    e0.record(); C = torch.matmul(A, Wt); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["cublas_skinny_WtA_ms"] = best
out["cublas_skinny_WtA_tflops"] = 2 * m * n * k / best * 1e-9
del Wt, C

ctx = sk.Context(0)
ctx.load_dense_device(A.data_ptr(), m, m, n)
rng = np.random.default_rng(12)
W0 = rng.random((m, k)); H0 = rng.random((k, n))
for alg in os.environ.get("PROBE_ALGS", "BPP,MU,HALS").split(","):
    opts = sk.make_options(m, n, k, algorithm=alg, tol=1e-12, min_iter=1, max_iter=100, normalize=False)
    t = time.time(); ctx.solver_begin(W0, H0, opts); ctx.synchronize(); out[f"{alg}_begin_s"] = time.time() - t
    times = []
    for i in range(6):
        ctx.solver_step(1)
        ms, launches = ctx.last_step()
        times.append(ms)
    out[f"{alg}_step_ms"] = times
    out[f"{alg}_launches"] = launches
    out[f"{alg}_metric"] = ctx.solver_progress()
    F = 4.0 * k * m * n + 6.0 * k * k * n + 4.0 * k * k * m
    out[f"{alg}_tflops_last"] = F / times[-1] * 1e-9
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
