#!/bin/bash
# A/B of the experimental kernels that are in the tree but off by default: parity first, then C3 with and without.
#   SMK_SPMM_PIPE=1   software-pipelined SpMM gathers (spmm_seg_slab_pipe_kernel, spmm_seg_tier_kernel<.., PIPE = true>)
mkdir -p gpurun_out
SMK_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or slabs or residency" > gpurun_out/pytest_experimental.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_experimental.log; tail -4 gpurun_out/pytest_experimental.log
timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_base.log 2>&1; tail -1 gpurun_out/c3_base.log | cut -c1-700
SMK_SPMM_PIPE=1 timeout 300 python tools/measure_c3_c4.py c3 > gpurun_out/c3_pipe.log 2>&1; tail -1 gpurun_out/c3_pipe.log | cut -c1-700
SMK_SPMM_PIPE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_seg_slab_pipe -s 8 -c 2 -o gpurun_out/prof_c3_spmm_slab_pipe -f \
   python tools/measure_c3_c4.py c3 > gpurun_out/prof_c3_spmm_slab_pipe.log 2>&1
ls -la gpurun_out | tail -6
