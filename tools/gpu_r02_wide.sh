#!/bin/bash
# the 128-thread form of the 64 < k <= 128 NNLS kernel: A/B timing, then the parity tests that reach it
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 40 python tools/nnls_wide_ab.py > gpurun_out/wide_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/wide_ab.log
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "nnls or golden or fixture or (dense_trace and BPP)" > gpurun_out/wide_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/wide_pytest.log; tail -5 gpurun_out/wide_pytest.log
