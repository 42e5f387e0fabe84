#!/bin/bash
# W-side HALS: outer pass without selects behind its loads; phase B as one cooperative kernel per block. Parity, then C3 phases A/B.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "hals or HALS or golden or scale or host_api or hierclust" > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/t_pytest.log; tail -6 gpurun_out/t_pytest.log
for sw in 0 1; do
SMK_HALS_SWEEP=$sw SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t_c3_sweep$sw.json 2> gpurun_out/t_c3_sweep$sw.err; echo "c3 sweep=$sw rc=$?"; tail -2 gpurun_out/t_c3_sweep$sw.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/t_c3_sweep$sw.json").read().strip().splitlines()[-1])
    print("sweep=$sw", round(j["ms_per_step"], 3), "ms", j["roofline"]["launch_ms"], {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()}, "metric", j.get("progress_metric_last"), (j["roofline"].get("l2_gather_bound") or {}).get("cap_TBs"))
except Exception as ex: print("failed", ex)
PY
done
