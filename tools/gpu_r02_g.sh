#!/bin/bash
# round 2, call G (one GPU): tests after the side-stream Grams and the wide-NNLS unrolls; C2 + C5 phases; ncu of the wide NNLS kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c5 > gpurun_out/g_bench_phases.json 2> gpurun_out/g_bench_phases.err; echo "phases rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/g_bench_phases.json").read().strip().splitlines()[-1])
print("C2:", round(j["value"], 1), round(j["ms_per_step"], 4), j["roofline"]["launch_ms"], j["parity"]["ok"], {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()})
e = j["extra"]["c5"]; print("C5:", e.get("value"), e.get("ms_per_step"), e.get("error"), {k: round(v, 2) for k, v in (e.get("phases_ms_per_step") or {}).items()})
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/g_bench_plain.json 2> gpurun_out/g_bench_plain.err; echo "plain rc=$?"; tail -c 300 gpurun_out/g_bench_plain.json | head -c 300; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_wide -s 4 -c 2 -o gpurun_out/prof_r02_nnls_wide -f python tools/measure_dense.py 20000 10000 256 BPP 2 > gpurun_out/g_ncu_wide.log 2>&1; echo "ncu wide rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_fast -s 6 -c 2 -o gpurun_out/prof_r02_nnls_fast -f python tools/measure_dense.py 20000 20000 64 BPP 2 > gpurun_out/g_ncu_fast.log 2>&1; echo "ncu fast rc=$?"
ls -la gpurun_out | tail -6
