#!/bin/bash
# round 2, call C (one GPU): tests after the NNLS inverse change, the default bench line, phases, C4 on its own.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c_bench.json; tail -5 gpurun_out/c_bench.err
SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c5 > gpurun_out/c_bench_phases.json 2> gpurun_out/c_bench_phases.err; echo "phases rc=$?"
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/c_c4_alone.json 2> gpurun_out/c_c4_alone.err; echo "c4 rc=$?"
ls -la gpurun_out | tail -8
