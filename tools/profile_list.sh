#!/bin/bash
# ncu launch list only (cheap): tools/profile_list.sh <tag> [skip] [count]
TAG=${1:-r01}
SKIP=${2:-5000}
CNT=${3:-900}
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s $SKIP -c $CNT --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"
