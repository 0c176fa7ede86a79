#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) one --set full capture of the top kernel.
# Run under gpurun (1 GPU).  Usage: tools/profile.sh <tag> [top-kernel-regex]
set -u
TAG=${1:-r01}
TOP=${2:-gemm_tc_kernel}
mkdir -p gpurun_out
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s 5000 -c 1700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${TOP}" -s 600 -c 3 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_${TAG}.stdout 2> gpurun_out/prof_${TAG}.stderr
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
