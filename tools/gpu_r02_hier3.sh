#!/bin/bash
# hierclust driver with compact node factors, two score workers and the next split started ahead of time
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_hierclust.py tests/test_gpu_host_api.py -m gpu -q > gpurun_out/hier3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/hier3_pytest.log; tail -15 gpurun_out/hier3_pytest.log
SMK_HIER_PROF=1 timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/hier3_c4_async.json 2> gpurun_out/hier3_c4_async.err; echo "c4 async rc=$?"
SMK_HIER_ASYNC=0 SMK_HIER_PROF=1 timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/hier3_c4_sync.json 2> gpurun_out/hier3_c4_sync.err; echo "c4 sync rc=$?"
python - <<'P'
import json
for f in ("async", "sync"):
    try:
        d = json.loads(open(f"gpurun_out/hier3_c4_{f}.json").read().strip().splitlines()[-1])
        print(f, d["e2e"]["seconds"], d["e2e"]["profile_s"], d["value"], d["config"]["nmf_count"], d["steps"])
    except Exception as ex:
        print(f, "failed", ex)
P
tail -n 7 gpurun_out/hier3_c4_async.err
