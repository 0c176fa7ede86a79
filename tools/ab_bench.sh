#!/bin/bash
# A/B of an environment knob on the bench's timed legs (20-step value + the 250-step sustained, power-limited leg):
#   tools/ab_bench.sh VAR val1 val2 ...        (alternating, 2 rounds)
var=$1; shift
for r in 1 2; do for v in "$@"; do
  echo -n "$var=$v : "
  env $var="$v" python bench.py --no-cpu-baseline --no-parity --no-eager-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['sustained']; print(d['value'], 'steps/s', d['ms_per_step'], 'ms | sustained', s['value'], s['clocks']['sm_mhz'], 'MHz', s['clocks'].get('power_w'), 'W | launches', d['gpu_launches'])"
done; done
