#!/bin/bash
# round 2, call A (one GPU): test suite incl. the at-scale parity cases, the bench line (writes the N = 1 metric trace),
# phase breakdown, A/B of the experimental SpMM kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --write-trace > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
cp tests/golden/bench_c2_trace_n1.json gpurun_out/ 2>/dev/null
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
SMK_PHASES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --extras c5,c3 > gpurun_out/a_bench_phases.json 2> gpurun_out/a_bench_phases.err; echo "phases rc=$?"
timeout 300 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/a_bench_100.json 2> gpurun_out/a_bench_100.err; echo "bench100 rc=$?"
SMK_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or slabs or residency" > gpurun_out/a_pytest_experimental.log 2>&1
echo "experimental pytest rc=$?" >> gpurun_out/a_pytest_experimental.log; tail -3 gpurun_out/a_pytest_experimental.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -20
