#!/bin/bash
# last pass of the round after the host-side changes: hierclust / host-API parity, then the bench line with all extras
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_hierclust.py tests/test_gpu_host_api.py tests/test_gpu_edges.py -m gpu -q > gpurun_out/final3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final3_pytest.log; tail -4 gpurun_out/final3_pytest.log
( time timeout 900 python bench.py > gpurun_out/final3_bench_n1.json 2> gpurun_out/final3_bench_n1.err ) 2> gpurun_out/final3_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/final3_bench_n1.time
