#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== controlnet + pipeline tests"; timeout 900 python -m pytest tests/test_controlnet_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "=== bench overlap=1"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','graph_replay_ms','clocks')})
PY
tail -3 gpurun_out/bench_d.err
echo "=== bench overlap=0"; TG_CN_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d0.json 2> gpurun_out/bench_d0.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d0.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','graph_replay_ms','clocks')})
PY
