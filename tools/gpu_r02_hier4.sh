#!/bin/bash
# hierclust driver: both children's scores pending while the next split is made ahead of time
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_hierclust.py tests/test_gpu_host_api.py -m gpu -q > gpurun_out/hier4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/hier4_pytest.log; tail -15 gpurun_out/hier4_pytest.log
SMK_HIER_ASYNC=0 timeout 600 python -m pytest tests/test_gpu_hierclust.py -m gpu -q > gpurun_out/hier4_pytest_sync.log 2>&1; echo "pytest (SMK_HIER_ASYNC=0) rc=$?" | tee -a gpurun_out/hier4_pytest_sync.log; tail -3 gpurun_out/hier4_pytest_sync.log
SMK_HIER_PROF=1 timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/hier4_c4_async.json 2> gpurun_out/hier4_c4_async.err; echo "c4 async rc=$?"
python - <<'P'
import json
for f in ("async",):
    try:
        d = json.loads(open(f"gpurun_out/hier4_c4_{f}.json").read().strip().splitlines()[-1])
        print(f, d["e2e"]["seconds"], d["e2e"]["profile_s"], d["value"], d["config"]["nmf_count"], d["steps"])
    except Exception as ex:
        print(f, "failed", ex)
P
tail -n 7 gpurun_out/hier4_c4_async.err
