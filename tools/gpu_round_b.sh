#!/bin/bash
# All GPU tests; C3 with the k-slab SpMM on/off and the outer HALS pass at 2/3 CTAs per SM; C4 with the reference (1 thread) beside it; ncu of the new kernels.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_slab.log 2>&1; tail -1 gpurun_out/c3_slab.log | cut -c1-900
SMK_SPMM_SLAB=0 timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_noslab.log 2>&1; tail -1 gpurun_out/c3_noslab.log | cut -c1-900
SMK_HALS_OUTER_OCC=3 timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_occ3.log 2>&1; tail -1 gpurun_out/c3_occ3.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hals_block_outer|spmm_seg_slab' -s 10 -c 5 -o gpurun_out/prof_c3_outer_slab -f \
   python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_outer_slab.log 2>&1
timeout 900 python tools/measure_c3_c4.py c4 > gpurun_out/c4.log 2>&1; grep workload gpurun_out/c4.log | cut -c1-600
