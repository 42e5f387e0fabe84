#!/bin/bash
# NNLS fast kernel (dense dual product, one solver call site): parity, then the C2 phases; full GPU suite
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r_pytest.log
tail -4 gpurun_out/r_pytest.log
SMK_PHASES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r_c2_phases.json 2> gpurun_out/r_c2_phases.err; echo "c2 rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/r_c2_phases.json").read().strip().splitlines()[-1])
print("C2", round(j["value"], 1), "it/s", round(j["ms_per_step"], 4), "ms", j["roofline"]["launch_ms"], "parity", j["parity"]["ok"], {k: round(v, 4) for k, v in (j.get("phases_ms_per_step") or {}).items()})
PY
SMK_PHASES=1 timeout 300 python tools/measure_dense.py 20000 2500 64 BPP 20 > gpurun_out/r_shard_phases.json 2> gpurun_out/r_shard_phases.err; echo "shard rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/r_shard_phases.json').read().strip().splitlines()[-1]); print('shard', j['ms_per_iter'], j['phases_ms'])"
