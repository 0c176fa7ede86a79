// Ground truth for the tcgen05.mma operand feed on this part: cycles per MMA (kind::f16, K = 16, fp32 accumulate) when
// one thread issues a long back-to-back stream of MMAs on resident shared-memory tiles (no TMA, no barriers in the loop).
//   ./umma_probe           -> table over tile widths N for cta_group::1 (M = 128) and cta_group::2 (M = 256, B split)
// SS mode (A and B from shared memory) is what gemm.cu uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include "../../theatergen_b200/csrc/common.cuh"

using namespace tg;

__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// TS mode: A (128 x 64 fp16 = 32 TMEM columns) lives in tensor memory, B in shared memory
__global__ void __launch_bounds__(128, 1) probe_ts(int N, int iters, int b_mn_major, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    uint32_t idesc = make_idesc_f16(128, N);
    if (b_mn_major) idesc |= 1u << 16;
    const uint32_t b = smem_u32(smem);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ts(tmem + (it & 1) * 128, tmem + 256 + k * 8, make_smem_desc_sw128(b + (b_mn_major ? k * 2048 : k * 32)), idesc, 1);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    if (blockIdx.x == 0) out[1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// iters k-blocks of 4 MMAs (K = 64 per k-block), rotating over `stages` resident smem stages
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int stages, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)ctarank() : 0;
  const int b_rows = PAIR ? N / 2 : N;
  const int stage_bytes = 128 * 128 + b_rows * 128;
  for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&slot, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (PAIR) csync();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = smem_u32(smem + (it % stages) * stage_bytes), b = a + 128 * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) umma2(tmem + (it & 1) * 256, make_smem_desc_sw128(a + k * 32), make_smem_desc_sw128(b + k * 32), idesc, 1);
        else umma_f16_ss(tmem + (it & 1) * 256, make_smem_desc_sw128(a + k * 32), make_smem_desc_sw128(b + k * 32), idesc, 1);
      }
    }
    if (PAIR) commit2(&bar, 0x3); else tc_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (PAIR && warp == 1 && lane == 0) {
    mbar_wait(&bar, 0);  // the multicast commit also lands here
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) csync();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}

template <bool PAIR>
static void run(int N, int iters, int stages, int ctas, long long* d_out) {
  const int b_rows = PAIR ? N / 2 : N;
  const size_t smem = (size_t)stages * (128 * 128 + b_rows * 128) + 1024;
  auto kern = probe<PAIR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {PAIR ? 2u : 1u, 1, 1};
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, kern, N, iters, stages, d_out);  // warm-up
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, kern, N, iters, stages, d_out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * iters, flop = 2.0 * (PAIR ? 256 : 128) * N * 16 * mmas * (PAIR ? ctas / 2 : ctas);
  const int opnd = 128 * 32 + b_rows * 32;  // operand bytes one SM's shared memory serves per MMA
  printf("%-6s M=%3d N=%3d stages=%d ctas=%3d : %7.1f clk/MMA (issue %6.1f)  smem operand %5d B/MMA/SM -> %5.1f B/clk/SM | %7.1f TFLOP/s by events  %s\n",
         PAIR ? "pair" : "single", PAIR ? 256 : 128, N, stages, ctas, h[1] / mmas, h[0] / mmas, opnd, opnd / (h[1] / mmas),
         flop / (ms * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  const int iters = 4000;
  for (int ctas : {1, 148}) {
    for (int N : {64, 128, 160, 256}) run<false>(N, iters, 4, ctas, d_out);
    for (int N : {64, 128, 160, 256}) run<true>(N, iters, 4, ctas == 1 ? 2 : 148, d_out);
  }
  run<false>(256, iters, 1, 148, d_out);
  run<true>(256, iters, 1, 148, d_out);
  for (int N : {48, 64, 128, 256}) run<false>(N, iters, 4, 148, d_out);
  cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mn = 0; mn < 2; ++mn) for (int N : {16, 48, 64, 128, 256}) {
    if (mn && N > 64) continue;
    probe_ts<<<148, 128, 40 * 1024>>>(N, iters, mn, d_out);
    probe_ts<<<148, 128, 40 * 1024>>>(N, iters, mn, d_out);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("TS (A in TMEM) M=128 N=%3d B %s : %7.1f clk/MMA  %s\n", N, mn ? "MN-major" : "K-major ", h[1] / (4.0 * iters), err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}
