#!/bin/bash
set -u
echo "=== gemm tests (wide on)"; timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "=== gemm_bench WIDE=1"; timeout 300 python tools/gemm_bench.py 10 2>&1 | grep -E "C=320|C=640|C=1280 M=2560|conv"
echo "=== gemm_bench WIDE=0"; TG_GEMM_WIDE=0 timeout 300 python tools/gemm_bench.py 10 2>&1 | grep -E "C=320|C=640|C=1280 M=2560|conv"
