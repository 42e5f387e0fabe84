#!/bin/bash
# round 2 profile captures (one GPU): launch list of the bench command; ncu --set full of the dominant kernels of C2, C3, C5.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 6 -c 4 -o gpurun_out/prof_r02_c2_gemm -f python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_seg -s 10 -c 5 -o gpurun_out/prof_r02_c3_spmm -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_spmm.log 2>&1; echo "spmm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_wide -s 4 -c 2 -o gpurun_out/prof_r02_c5_nnls_wide -f python tools/measure_dense.py 20000 10000 256 BPP 2 > gpurun_out/p_ncu_wide.log 2>&1; echo "wide rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hals_block -s 40 -c 4 -o gpurun_out/prof_r02_c3_hals -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_hals.log 2>&1; echo "hals rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02_bench_c2.csv
