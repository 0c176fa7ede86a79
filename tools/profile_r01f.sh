#!/bin/bash
# Reduced ncu evidence pass (round 1, final tree): launch list of the timed region, DRAM traffic of every GEMM launch,
# and one --set full capture of six GEMM launches.  Run under gpurun (1 GPU).  Usage: tools/profile_r01f.sh <tag>
set -u
TAG=${1:-r01f}
mkdir -p gpurun_out
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "tg_timed/" -k "$MINE" --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"; wc -l gpurun_out/launches_${TAG}.csv
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" -k "regex:gemm_tc_kernel" -s ${GEMM_SKIP:-40} -c 6 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_gemm_${TAG}.stdout 2> gpurun_out/prof_gemm_${TAG}.stderr
echo "gemm capture rc=$?"
ls -la gpurun_out | tail -5
