#!/bin/bash
# round 2, call D (one GPU): tests after the in-kernel split-R fix-up; split-count sweep on the C2 products.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
for s in 0 9 11 13 15 17 19; do
  SMK_GEMM_SPLITS=$s SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/d_bench_s$s.json 2> gpurun_out/d_bench_s$s.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/d_bench_s$s.json").read().strip().splitlines()[-1])
    print("splits $s:", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in j["roofline"]["launch_ms"].items()}, "parity", j["parity"]["ok"], {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()})
except Exception as ex:
    print("splits $s: failed", ex)
PY
done
