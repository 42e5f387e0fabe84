#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c5 > gpurun_out/j_bench_phases.json 2> gpurun_out/j_bench_phases.err; echo "phases rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/j_bench_phases.json").read().strip().splitlines()[-1])
print("C2:", round(j["value"], 1), round(j["ms_per_step"], 4), j["roofline"]["launch_ms"], j["parity"]["ok"], {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()})
e = j["extra"]["c5"]; print("C5:", e.get("value"), e.get("ms_per_step"), e.get("error"), {k: round(v, 2) for k, v in (e.get("phases_ms_per_step") or {}).items()})
PY
