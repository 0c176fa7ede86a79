#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; cat gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
