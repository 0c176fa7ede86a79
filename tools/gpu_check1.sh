#!/bin/bash
# one gpurun call: kernel parity tests, GEMM pair-vs-single A/B, GroupNorm sweep, full GPU suite, bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== gemm + ops tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "=== gemm_bench PAIR=1"; timeout 300 python tools/gemm_bench.py 10 > gpurun_out/gemm_bench_pair1.txt 2>&1; tail -30 gpurun_out/gemm_bench_pair1.txt
echo "=== gemm_bench PAIR=0"; TG_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py 10 > gpurun_out/gemm_bench_pair0.txt 2>&1; tail -30 gpurun_out/gemm_bench_pair0.txt
echo "=== gn_bench"; timeout 600 python tools/gn_bench.py 10 > gpurun_out/gn_bench.txt 2>&1; cat gpurun_out/gn_bench.txt
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
echo "=== bench PAIR=0"; TG_GEMM_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a_pair0.json 2> gpurun_out/bench_a_pair0.err; cat gpurun_out/bench_a_pair0.json
