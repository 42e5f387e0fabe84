#!/bin/bash
# the edge / all-zero-warm-start parity tests, then an ncu capture of the 64 < k <= 256 NNLS kernel inside the sparse-BPP variant of C3
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_edges.py -m gpu -q > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/last_pytest.log; tail -4 gpurun_out/last_pytest.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nnls_bpp_wide -s 2 -c 2 -o gpurun_out/prof_r02_c3_sparse_bpp_nnls -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/last_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r02_c3_sparse_bpp_nnls.ncu-rep 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02_c3_sparse_bpp_nnls.ncu-rep > gpurun_out/ncu_r02_c3_sparse_bpp_nnls_wide.txt 2>/dev/null; head -30 gpurun_out/ncu_r02_c3_sparse_bpp_nnls_wide.txt
