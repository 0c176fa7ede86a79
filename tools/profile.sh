#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command's TIMED region (NVTX range tg_timed), (2) one
# --set full capture of the two top kernels inside the same region.  Run under gpurun (1 GPU).  Usage: tools/profile.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "tg_timed/" -k "$MINE" --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"; wc -l gpurun_out/launches_${TAG}.csv
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" -k "regex:gemm_tc_kernel" -s ${GEMM_SKIP:-40} -c 6 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_gemm_${TAG}.stdout 2> gpurun_out/prof_gemm_${TAG}.stderr
echo "gemm capture rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" -k "regex:attn_tc_kernel" -c 2 \
    -o gpurun_out/prof_attn_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_attn_${TAG}.stdout 2> gpurun_out/prof_attn_${TAG}.stderr
echo "attn capture rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" -k "regex:gn_stats_kernel|gn_apply_kernel|gn_fused_kernel" -s 8 -c 6 \
    -o gpurun_out/prof_gn_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_gn_${TAG}.stdout 2> gpurun_out/prof_gn_${TAG}.stderr
echo "gn capture rc=$?"
# DRAM traffic of every GEMM launch of the timed region (roofline.traffic in bench.py = mean per launch)
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "tg_timed/" \
    -k "regex:gemm_tc_kernel" --csv --log-file gpurun_out/gemm_dram_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/gemm_dram_${TAG}.stdout 2> gpurun_out/gemm_dram_${TAG}.stderr
echo "gemm dram rc=$?"
ls -la gpurun_out | tail -8
