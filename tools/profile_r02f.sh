#!/bin/bash
# ncu evidence pass of the final round-2 tree: launch list of bench.py's timed region (time + DRAM bytes per launch) and
# --set full captures of the launches that set the headline: the 1280-channel GEMM tile <256,3,pair>, the 64x64 patch-reuse
# convolution <160,3,pair,KY3>, the 64x64 self-attention (packed heads) and the folded cross-attention score GEMM with the
# key-softmax epilogue <192,...>.  Run under gpurun (1 GPU).  Usage: tools/profile_r02f.sh <tag>
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-eager-reference --no-sustained"
MINE='regex:gemm_tc|attn_|gn_|layernorm|conv_in|conv_out|cfg_ddim|guidance_|upsample2x|im2col|timestep_embed|dup_kernel|add_inplace|nchw_to|nhwc_'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "tg_timed/" -k "$MINE" --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
echo "launch list rc=$?"; wc -l gpurun_out/launches_${TAG}.csv
cap() {  # name, mangled-name regex, skip, launches
  timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" --kernel-name-base mangled \
      -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1_${TAG} -f $BENCH > gpurun_out/prof_$1_${TAG}.stdout 2> gpurun_out/prof_$1_${TAG}.stderr
  echo "$1 capture rc=$?"
}
cap gemm256 'gemm_tc_kernelILi256ELi3ELb1ELb0E' 20 4
cap conv160 'gemm_tc_kernelILi160ELi3ELb1ELb1E' 6 4
cap attn64 'attn_tc_kernelILi1ELi0ELb0E' 2 3
cap xscores 'gemm_tc_kernelILi192E' 0 6
ls -la gpurun_out | grep "${TAG}" | tail -14
