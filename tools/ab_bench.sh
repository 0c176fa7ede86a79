#!/bin/bash
# A/B of an environment knob on the bench's timed legs only: tools/ab_bench.sh VAR val1 val2 ... (alternating, 2 rounds)
var=$1; shift
for r in 1 2; do for v in "$@"; do
  echo -n "$var=$v : "
  env $var="$v" python bench.py --no-cpu-baseline --no-parity --no-eager-reference --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], 'steps/s', d['ms_per_step'], 'ms', 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"
done; done
