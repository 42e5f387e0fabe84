#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q > gpurun_out/k_pytest_pre.log 2>&1; echo "pytest preprocess rc=$?"; tail -15 gpurun_out/k_pytest_pre.log
timeout 60 tools/cond_test; echo "cond rc=$?"
