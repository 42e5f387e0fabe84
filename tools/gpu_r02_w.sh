#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "hals or HALS or golden or scale" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/w_pytest.log; tail -3 gpurun_out/w_pytest.log
SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/w_c3.json 2> gpurun_out/w_c3.err; echo "c3 rc=$?"; tail -2 gpurun_out/w_c3.err
python - <<PY
import json
j = json.loads(open("gpurun_out/w_c3.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"], 3), "ms", j["roofline"]["launch_ms"], {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()}, "metric", j.get("progress_metric_last"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hals_block_sweep|hals_block_outer" -s 4 -c 2 -o gpurun_out/prof_r02_c3_hals_d -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/w_ncu_hals.log 2>&1; echo "ncu hals rc=$?"
