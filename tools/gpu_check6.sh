#!/bin/bash
set -u
echo "=== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "=== gemm_bench STREAMK=1"; timeout 300 python tools/gemm_bench.py 10 2>&1 | tail -26
echo "=== gemm_bench STREAMK=0"; TG_GEMM_STREAMK=0 timeout 300 python tools/gemm_bench.py 10 2>&1 | tail -26
