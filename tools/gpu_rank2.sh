#!/bin/bash
# Fused rank-2 check: hierclust / host-API / sparse parity tests, C4 end to end with the reference's tree beside it, C4-small launch list.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "hier or host_api or RANK2 or rank2 or sparse or smoke" ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/measure_c3_c4.py c4 > gpurun_out/c4.log 2>&1; grep workload gpurun_out/c4.log | cut -c1-500
timeout 300 python tests/manual/diag_c4.py 40000 250000 16 1 > gpurun_out/diag_c4small_fused.log 2>&1; grep nodes gpurun_out/diag_c4small_fused.log | cut -c1-500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c4small.csv \
   python tools/measure_c3_c4.py c4small > gpurun_out/c4small_ncu.log 2>&1
