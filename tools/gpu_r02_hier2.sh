#!/bin/bash
# edge-shape parity tests, hierclust parity, then C4 with the driver profile
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_edges.py tests/test_gpu_hierclust.py tests/test_gpu_host_api.py -m gpu -q > gpurun_out/hier2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/hier2_pytest.log; tail -15 gpurun_out/hier2_pytest.log
SMK_HIER_PROF=1 SMK_PRIORITY_PROF=1 timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/hier2_c4_async.json 2> gpurun_out/hier2_c4_async.err; echo "c4 async rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/hier2_c4_async.json").read().strip().splitlines()[-1])
print(d["e2e"]["seconds"], d["e2e"]["profile_s"], d["value"])
P
tail -n 12 gpurun_out/hier2_c4_async.err
