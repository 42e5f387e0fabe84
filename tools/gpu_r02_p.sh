#!/bin/bash
# lean SpMM kernels: parity, then C3 with the round-1 forms (SMK_SPMM_LEAN=0) and the lean ones; gather-rate probe
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 tools/l2_gather_peak > gpurun_out/p_gather_peak.txt 2>&1; echo "gather rc=$?"; tail -4 gpurun_out/p_gather_peak.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_scale.py -m gpu -x -q -k "sparse or hals or c3" > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
for lean in 0 1; do
SMK_SPMM_LEAN=$lean SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p_c3_lean$lean.json 2> gpurun_out/p_c3_lean$lean.err; echo "c3 lean=$lean rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/p_c3_lean$lean.json").read().strip().splitlines()[-1])
print("lean=$lean", round(j["ms_per_step"], 3), "ms", j["roofline"]["launch_ms"], {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()}, "metric", j.get("progress_metric_last"))
PY
done
