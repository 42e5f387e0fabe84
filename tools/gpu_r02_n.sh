#!/bin/bash
# state check of HEAD: every GPU parity test, smoke, then the default bench line with all extras (one GPU)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/n_pytest.log
tail -4 gpurun_out/n_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/n_smoke.log
( time timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err ) 2> gpurun_out/n_bench.time; echo "bench rc=$?"; tail -3 gpurun_out/n_bench.time
SMK_PHASES=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_c3_phases.json 2> gpurun_out/n_c3_phases.err; echo "c3 rc=$?"
cut -c1-400 gpurun_out/n_bench.json
