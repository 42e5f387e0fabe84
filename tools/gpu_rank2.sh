#!/bin/bash
# Fused rank-2 check: all GPU parity tests, C4 (hierclust) end to end, and where the tree leaves the reference's.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/measure_c3_c4.py c4 > gpurun_out/c4.log 2>&1; grep workload gpurun_out/c4.log | cut -c1-500
timeout 300 python tools/diag_c4.py 40000 250000 16 > gpurun_out/diag_c4small.log 2>&1; grep nodes gpurun_out/diag_c4small.log | cut -c1-1500
timeout 600 python tools/diag_c4.py 320000 2000000 4 > gpurun_out/diag_c4_4.log 2>&1; grep nodes gpurun_out/diag_c4_4.log | cut -c1-1500
