#!/bin/bash
# round 2, call H (one GPU): blocked Cholesky in the wide NNLS kernel, plain GEMM instantiation restored: tests, phases, split sweep.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/h_pytest.log
tail -4 gpurun_out/h_pytest.log
for sg in 1 0; do
SMK_SIDE_GRAM=$sg SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c5 > gpurun_out/h_bench_phases_sg$sg.json 2> gpurun_out/h_bench_phases_sg$sg.err; echo "phases sg=$sg rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/h_bench_phases_sg$sg.json").read().strip().splitlines()[-1])
print("C2 side_gram=$sg:", round(j["value"], 1), round(j["ms_per_step"], 4), j["roofline"]["launch_ms"], j["parity"]["ok"], {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()})
e = j["extra"]["c5"]; print("C5:", e.get("value"), e.get("ms_per_step"), e.get("error"), {k: round(v, 2) for k, v in (e.get("phases_ms_per_step") or {}).items()})
PY
done
timeout 300 python tools/sweep_splits.py c2 2>&1 | cut -c1-700 | grep "world=1"
