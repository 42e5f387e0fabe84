#!/bin/bash
# One GPU-box pass: parity tests, bench line, reference arm, ncu launch list, ncu full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 6 -c 2 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_fast -s 6 -c 2 -o gpurun_out/prof_nnls -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_nnls.log 2>&1
ls -la gpurun_out
