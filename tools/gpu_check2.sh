#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "=== gemm_bench PAIR=1"; timeout 300 python tools/gemm_bench.py 10 > gpurun_out/gemm_bench_pair1.txt 2>&1; tail -28 gpurun_out/gemm_bench_pair1.txt
echo "=== gemm_bench PAIR=0"; TG_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py 10 > gpurun_out/gemm_bench_pair0.txt 2>&1; tail -28 gpurun_out/gemm_bench_pair0.txt
echo "=== roles"; for p in 1; do TG_GEMM_PAIR=$p timeout 200 python tools/gemm_roles.py; done 2>&1 | tee gpurun_out/gemm_roles.txt
