#!/bin/bash
# Tiered SpMM check: sparse parity tests, C3 with and without the residency classes, ncu --set full of both products.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "sparse" > gpurun_out/pytest_sparse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sparse.log
tail -6 gpurun_out/pytest_sparse.log
timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_tiers.log 2>&1; tail -1 gpurun_out/c3_tiers.log | cut -c1-1200
SMK_SPMM_TIERS=0 timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_notiers.log 2>&1; tail -1 gpurun_out/c3_notiers.log | cut -c1-1200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_seg_tier -s 4 -c 2 -o gpurun_out/prof_c3_spmm_tier -f \
   python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_spmm_tier.log 2>&1
ls -la gpurun_out | tail -5
