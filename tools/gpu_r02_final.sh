#!/bin/bash
# Final single-GPU pass of round 2: every GPU test, smoke, the bench line with all extras, the reference arm, and the ncu evidence
# (launch list of the bench command; --set full captures of the dominant kernels of C2 and C3).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final_smoke.log
( time timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err ) 2> gpurun_out/final_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/final_bench_n1.time
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/final_bench_n1_k20.json 2> gpurun_out/final_bench_n1_k20.err; echo "bench k20 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma|reduce_partials" -s 8 -c 4 -o gpurun_out/prof_r02_final_c2_gemm -f python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_fast -s 6 -c 2 -o gpurun_out/prof_r02_final_c2_nnls -f python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_nnls.log 2>&1; echo "nnls rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_seg -s 10 -c 5 -o gpurun_out/prof_r02_final_c3_spmm -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_spmm.log 2>&1; echo "spmm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|hals|gemm|pg_|reduce" -s 120 -c 120 --csv --log-file gpurun_out/final_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_c3.log 2>&1; echo "c3 launch list rc=$?"
ls -la gpurun_out/prof_r02_final* gpurun_out/final_*
