#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== controlnet tests"; timeout 900 python -m pytest tests/test_controlnet_gpu.py -x -q -m gpu -s 2>&1 | tail -40
echo "=== pipeline + unet tests"; timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; cat gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
