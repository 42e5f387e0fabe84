#!/bin/bash
# ncu --set full of the W-side HALS kernels at C3 (outer pass, cooperative block sweep); C4 with the priority-score section timers
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hals_block_outer|hals_block_sweep" -s 16 -c 4 -o gpurun_out/prof_r02_c3_hals_b -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/u_ncu_hals.log 2>&1; echo "ncu hals rc=$?"
SMK_PRIORITY_PROF=1 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/u_c4.json 2> gpurun_out/u_c4.err; echo "c4 rc=$?"; grep "compute_priority_rows" gpurun_out/u_c4.err | tail -2
python - <<PY
import json
j = json.loads(open("gpurun_out/u_c4.json").read().strip().splitlines()[-1])
print("C4 e2e", j["e2e"]["seconds"], j["e2e"]["profile_s"])
PY
