#!/bin/bash
# G * X - R (k x k x q, reduction of 4-8 chunks) on the TMA kernel or on the cp.async kernel: C3 phases, a C5 column shard
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_trace or gemm" > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s2_pytest.log
for t in 0 1; do
SMK_GEMM_TMA_SHORT=$t SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_c3_$t.json 2> gpurun_out/s2_c3_$t.err
python - <<PY
import json
j = json.loads(open("gpurun_out/s2_c3_$t.json").read().strip().splitlines()[-1])
print("short-tma=$t C3", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in (j.get("phases_ms_per_step") or {}).items()}, "metric", j.get("progress_metric_last"))
PY
SMK_GEMM_TMA_SHORT=$t SMK_PHASES=1 timeout 300 python tools/measure_dense.py 100000 12500 256 BPP 4 > gpurun_out/s2_c5_$t.json 2> gpurun_out/s2_c5_$t.err; python -c "
import json; j=json.loads(open('gpurun_out/s2_c5_$t.json').read().strip().splitlines()[-1]); print('short-tma=$t C5 shard', round(j['ms_per_iter'],3), {k: round(v,3) for k,v in j['phases_ms'].items()})"
done
