// Why does a TMA-fed smem ring take ~540-600 cycles per stage regardless of stage size and depth?  Variants of the same
// ring (A box 128 x 64 fp16 + optional B box, L2-resident sources, one CTA per SM):
//   mode 0  two threads (producer / consumer), mbarrier.try_wait              (what gemm.cu does)
//   mode 1  as 0 + prefetch.tensormap first
//   mode 2  as 0 with mbarrier.test_wait spinning instead of try_wait
//   mode 3  ONE thread: refill stage s right after it completed (no empty barrier, no second thread)
//   mode 4  as 3 with test_wait spin
//   mode 5  as 0 but the consumer is a full warp (all 32 lanes wait, lane 0 arrives)
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

template <bool SPIN>
__device__ __forceinline__ void waitb(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    if (SPIN)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                               int mode, int a_rows, int bn, int nk, int stages, int total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  const int a_bytes = a_rows * 128, stage_bytes = a_bytes + bn * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
    if (mode == 1) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  }
  __syncthreads();
  auto issue = [&](int it) {
    const int s = it % stages, kb = it % nk;
    uint8_t* sa = smem + s * stage_bytes;
    mbar_arrive_expect_tx(&full[s], stage_bytes);
    tma_load_2d(sa, &tmA, &full[s], kb * 64, blockIdx.x * a_rows);
    if (bn) tma_load_2d(sa + a_bytes, &tmB, &full[s], kb * 64, 0);
  };
  if (mode == 3 || mode == 4) {
    if (threadIdx.x == 0) {
      for (int it = 0; it < stages; ++it) issue(it);
      const long long t0 = clock64();
      for (int it = 0; it < total; ++it) {
        const int s = it % stages;
        if (mode == 4) waitb<true>(&full[s], (it / stages) & 1); else waitb<false>(&full[s], (it / stages) & 1);
        if (it + stages < total) issue(it + stages);
      }
      out[blockIdx.x] = clock64() - t0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    for (int it = 0; it < total; ++it) {
      const int s = it % stages;
      if (mode == 2) waitb<true>(&empty[s], ((it / stages) & 1) ^ 1); else waitb<false>(&empty[s], ((it / stages) & 1) ^ 1);
      issue(it);
    }
  } else if (threadIdx.x >= 32 && (mode == 5 || threadIdx.x == 32)) {
    const long long t0 = clock64();
    for (int it = 0; it < total; ++it) {
      const int s = it % stages;
      if (mode == 2) waitb<true>(&full[s], (it / stages) & 1); else waitb<false>(&full[s], (it / stages) & 1);
      if (mode == 5) __syncwarp();
      if (threadIdx.x == 32) mbar_arrive(&empty[s]);
    }
    if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int ctas = 148, K = 1024, nk = K / 64, total = 4096;
  __half *A, *B;
  cudaMalloc(&A, (size_t)ctas * 256 * K * 2); cudaMalloc(&B, (size_t)256 * K * 2);
  cudaMemset(A, 0, (size_t)ctas * 256 * K * 2); cudaMemset(B, 0, (size_t)256 * K * 2);
  long long* d_out; cudaMalloc(&d_out, ctas * sizeof(long long));
  if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) { printf("attr failed\n"); return 1; }
  printf("%-5s %-7s %-4s %-7s | cycles per stage, B/clk/SM\n", "mode", "a_rows", "bn", "stages");
  struct Cfg { int a_rows, bn, stages; };
  const Cfg cfgs[] = {{128, 0, 4}, {128, 160, 4}, {64, 0, 4}, {256, 0, 4}, {128, 256, 4}, {128, 160, 2}};
  for (int mode = 0; mode <= 5; ++mode) {
    for (const Cfg& c : cfgs) {
      const int stage_bytes = c.a_rows * 128 + c.bn * 128;
      CUtensorMap ma, mb;
      if (make_tmap_2d(&ma, A, (uint64_t)ctas * 256, K, K, c.a_rows) || make_tmap_2d(&mb, B, 256, K, K, c.bn ? c.bn : 64)) { printf("tmap failed: %s\n", get_error()); return 1; }
      cudaMemset(d_out, 0, ctas * sizeof(long long));
      for (int rep = 0; rep < 2; ++rep) probe<<<ctas, 64, c.stages * stage_bytes + 1024>>>(ma, mb, mode, c.a_rows, c.bn, nk, c.stages, total, d_out);
      cudaError_t err = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (err == cudaSuccess) err = e2;
      long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      double clk = 0; for (int i = 0; i < ctas; ++i) clk += h[i]; clk /= ctas;
      printf("%-5d %-7d %-4d %-7d | %7.0f  %6.1f  %s\n", mode, c.a_rows, c.bn, c.stages, clk / total, stage_bytes / (clk / total), err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
