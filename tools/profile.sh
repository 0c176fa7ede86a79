#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) one --set full capture of the two top kernels.
# Run under gpurun (1 GPU).  Usage: tools/profile.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to'
# skip the capture warm-ups + 3 warm-up steps, then list ~2 bench steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s ${SKIP:-5400} -c ${CNT:-830} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel" -s 1196 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_gemm_${TAG}.stdout 2> gpurun_out/prof_gemm_${TAG}.stderr
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_tc_kernel" -s 30 -c 1 \
    -o gpurun_out/prof_attn_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_attn_${TAG}.stdout 2> gpurun_out/prof_attn_${TAG}.stderr
echo "attn capture rc=$?"
ls -la gpurun_out | tail -8
