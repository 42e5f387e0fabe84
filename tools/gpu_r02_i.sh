#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python tools/sweep_splits.py c2 2>&1 | cut -c1-500 | grep "fixup=1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_wide -s 4 -c 1 -o gpurun_out/prof_r02_nnls_wide_b -f python tools/measure_dense.py 20000 10000 256 BPP 2 > gpurun_out/i_ncu_wide.log 2>&1; echo "ncu wide rc=$?"; tail -2 gpurun_out/i_ncu_wide.log | cut -c1-300
