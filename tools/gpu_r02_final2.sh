#!/bin/bash
# last check of the round: every GPU test, smoke, the bench line with all extras (refreshes profiles/bench_r02_final_n1*.json)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final_smoke.log
( time timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err ) 2> gpurun_out/final_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/final_bench_n1.time
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/final_bench_n1_k20.json 2> gpurun_out/final_bench_n1_k20.err; echo "bench k20 rc=$?"
