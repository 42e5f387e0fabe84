#!/bin/bash
# TMA GEMM: parity, C2 with and without it (phases), the 8-GPU column shard shape on one GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm" > gpurun_out/s_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?" | tee -a gpurun_out/s_pytest_gemm.log; tail -15 gpurun_out/s_pytest_gemm.log
for t in 0 1; do
SMK_GEMM_TMA=$t SMK_PHASES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/s_c2_tma$t.json 2> gpurun_out/s_c2_tma$t.err; echo "c2 tma=$t rc=$?"; tail -3 gpurun_out/s_c2_tma$t.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/s_c2_tma$t.json").read().strip().splitlines()[-1])
    print("C2 tma=$t", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", j["roofline"]["launch_ms"], "frac", round(j["roofline"]["frac"], 4), "parity", j["parity"]["ok"], j["parity"].get("max_rel_err_vs_reference"), {k: round(v, 4) for k, v in (j.get("phases_ms_per_step") or {}).items()})
except Exception as ex: print("failed", ex)
PY
SMK_GEMM_TMA=$t SMK_PHASES=1 timeout 300 python tools/measure_dense.py 20000 2500 64 BPP 20 > gpurun_out/s_shard_tma$t.json 2> gpurun_out/s_shard_tma$t.err; python -c "
import json; j=json.loads(open('gpurun_out/s_shard_tma$t.json').read().strip().splitlines()[-1]); print('shard tma=$t', j['ms_per_iter'], j['phases_ms'])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s_pytest.log; tail -4 gpurun_out/s_pytest.log
