#!/bin/bash
# k > 256 fallback kernels: their parity tests, the backup-rule fixtures, then the bench line with the two new variant lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_bigk.py tests/test_gpu_golden.py -m gpu -q -x --durations=8 -k "any_k or backup" > gpurun_out/bigk_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/bigk_pytest.log; tail -25 gpurun_out/bigk_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --extras c3 > gpurun_out/bigk_bench.json 2> gpurun_out/bigk_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bigk_bench.err
