// Does the tensor core's read of a TMEM-resident A operand (tcgen05.mma .ts form) share the tensor-memory read port with
// tcgen05.ld?  One CTA per SM: warp 1 issues a stream of 128 x N x 16 MMAs (TS or SS form), warps 4..7 stream
// tcgen05.ld.32x32b.x32 (4 KB per warp-instruction, 32 lanes x 32 columns) over the accumulator columns.
//   mode 0: tcgen05.ld only     mode 1: MMA only     mode 2: both
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mix_probe tmem_mix_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../../theatergen_b200/csrc/common.cuh"

using namespace tg;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(256, 1) probe(int N, int iters, int ss, int mode, int ld_warps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 1 && (mode & 1)) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a = smem_u32(smem), b = a + 16 * 1024;
    const long long t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ss) umma_f16_ss(tmem + (it & 1) * 64, make_smem_desc_sw128(a + k * 32), make_smem_desc_sw128(b + k * 32), idesc, 1);
          else umma_ts(tmem + (it & 1) * 64, tmem + 256 + k * 8, make_smem_desc_sw128(b + k * 32), idesc, 1);
        }
      }
      tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (blockIdx.x == 0 && leader) out[0] = clock64() - t0;
  } else if (warp >= 4 && warp < 4 + ld_warps && (mode & 2)) {
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem + (it & 3) * 32 + lane_off, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[7] = acc;
    if (blockIdx.x == 0 && lane == 0) out[1 + (warp - 4)] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  const int iters = 4000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ss = 0; ss < 2; ++ss)
    for (int N : {48, 64}) for (int ldw : {4}) for (int mode : {2, 1, 3}) {
      if (ss && mode == 2) continue;
      cudaMemset(d_out, 0, 64);
      probe<<<148, 256, 50 * 1024>>>(N, iters, ss, mode, ldw, d_out);
      probe<<<148, 256, 50 * 1024>>>(N, iters, ss, mode, ldw, d_out);
      cudaError_t err = cudaDeviceSynchronize();
      long long h[8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%s N=%2d mode=%d (%s%s): ", ss ? "SS" : "TS", N, mode, (mode & 1) ? "mma " : "", (mode & 2) ? "ld" : "");
      if (mode & 1) printf("%6.1f clk/MMA  ", h[0] / (4.0 * iters));
      if (mode & 2) printf("%6.1f clk per x32 tcgen05.ld per warp (%d warps) -> %5.1f B/clk/SM  ", (double)h[1] / iters, ldw,
                           ldw * 4096.0 / ((double)h[1] / iters));
      printf("%s\n", err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  // one warp alone: latency-bound rate of dependent x32 loads
  cudaMemset(d_out, 0, 64);
  probe<<<148, 256, 50 * 1024>>>(64, iters, 0, 2, 1, d_out);
  cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("one warp: %6.1f clk per x32 tcgen05.ld + wait\n", (double)h[1] / iters);
  return 0;
}
