// Timeline of 8 back-to-back TMA tensor loads (128 x 64 fp16 boxes, L2-resident) issued by one thread into 8 stages:
// cycle at which each issue instruction returned, and cycle at which each stage's mbarrier completed (polled by a second
// thread with test_wait, so no suspend/wake-up is involved).  One CTA per SM runs it; CTA 0 reports.
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int rows, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  __shared__ long long t_start;
  const int stage_bytes = rows * 128;
  if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) mbar_init(&full[s], 1); fence_barrier_init(); }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    if (threadIdx.x == 0) {
      long long t[9];
      t[0] = clock64();
      t_start = t[0];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_2d(smem + s * stage_bytes, &tm, &full[s], ((r * 8 + s) % 16) * 64, blockIdx.x * rows);
        t[s + 1] = clock64();
      }
      if (blockIdx.x == 0 && r == reps - 1) for (int s = 0; s < 8; ++s) out[s] = t[s + 1] - t[0];
    } else if (threadIdx.x == 32) {
      long long t[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(r & 1) : "memory");
        t[s] = clock64();
      }
      if (blockIdx.x == 0 && r == reps - 1) for (int s = 0; s < 8; ++s) out[8 + s] = t[s] - *(volatile long long*)&t_start;
    }
    __syncthreads();
  }
}

int main() {
  const int ctas = 148, K = 1024;
  __half* A; cudaMalloc(&A, (size_t)ctas * 256 * K * 2); cudaMemset(A, 0, (size_t)ctas * 256 * K * 2);
  long long* d_out; cudaMalloc(&d_out, 16 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int nctas : {1, 148}) for (int rows : {8, 128}) {
    CUtensorMap tm;
    if (make_tmap_2d(&tm, A, (uint64_t)ctas * 256, K, K, rows)) { printf("tmap failed\n"); return 1; }
    probe<<<nctas, 64, 8 * rows * 128 + 1024>>>(tm, rows, 6, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("ctas=%3d rows=%3d  issue returned at:", nctas, rows); for (int s = 0; s < 8; ++s) printf(" %5lld", h[s]);
    printf("\n                   completed at:     "); for (int s = 0; s < 8; ++s) printf(" %5lld", h[8 + s]);
    printf("   %s\n", cudaGetErrorString(e));
  }
  return 0;
}
