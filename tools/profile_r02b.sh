set -u
TAG=r02b
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-eager-reference --no-sustained"
cap() {
  timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "tg_timed/" --kernel-name-base mangled \
      -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1_${TAG} -f $BENCH > gpurun_out/prof_$1_${TAG}.stdout 2> gpurun_out/prof_$1_${TAG}.stderr
  echo "$1 capture rc=$?"
}
cap gemm256 'gemm_tc_kernelILi256ELi3ELb1ELb0E' 20 8
cap conv160 'gemm_tc_kernelILi160ELi3ELb1ELb1E' 6 8
cap attn64 'attn_tc_kernelILi1ELi0ELb0E' 2 4
ls -la gpurun_out | grep r02b
