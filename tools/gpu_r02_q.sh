#!/bin/bash
# lean SpMM kernels: 16 warps/SM x 8 gathers in flight per lane against 32 warps/SM x 4
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "8 8" "4 8" "8 4" "4 4"; do
set -- $cfg
SMK_SPMM_SLAB_U=$1 SMK_SPMM_WIDE_U=$2 SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/q_c3_$1_$2.json 2> gpurun_out/q_c3_$1_$2.err; echo "c3 slabU=$1 wideU=$2 rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/q_c3_$1_$2.json").read().strip().splitlines()[-1])
print("slabU=$1 wideU=$2", round(j["ms_per_step"], 3), "ms", j["roofline"]["launch_ms"], {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()}, "metric", j.get("progress_metric_last"))
PY
done
SMK_SPMM_SLAB_U=4 SMK_SPMM_WIDE_U=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
