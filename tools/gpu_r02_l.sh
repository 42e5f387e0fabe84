#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/l_pytest.log
tail -6 gpurun_out/l_pytest.log
for g in 1 0; do
SMK_GRAPH=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c1,c4,preprocess > gpurun_out/l_bench_g$g.json 2> gpurun_out/l_bench_g$g.err; echo "bench graph=$g rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/l_bench_g$g.json").read().strip().splitlines()[-1])
print("C2:", round(j["value"], 1), round(j["ms_per_step"], 4))
for nm, e in j["extra"].items():
    print("  ", nm, {k: e.get(k) for k in ("value", "ms_per_step", "iterations_to_converge", "seconds", "error", "parity", "seconds_spent")}, (e.get("e2e") or {}).get("value"), (e.get("e2e") or {}).get("seconds"), (e.get("e2e") or {}).get("profile_s"), (e.get("cpu_baseline") or {}).get("value"))
PY
done
