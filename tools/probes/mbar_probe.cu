// mbarrier hand-off latency between two warps of one CTA (the producer <-> MMA-issuer <-> epilogue hand-offs of the
// warp-specialised kernels): round-trip cycles of a ping-pong through two mbarriers for different wait flavours.
//   0: mbarrier.try_wait (default suspend hint)   1: try_wait with suspendTimeHint = 32 ns   2: test_wait spin
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

template <int MODE>
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    if (MODE == 0)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else if (MODE == 1)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(32) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

template <int MODE>
__global__ void pingpong(int iters, long long* out) {
  __shared__ uint64_t a, b;
  if (threadIdx.x == 0) { mbar_init(&a, 1); mbar_init(&b, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive(&a); wait<MODE>(&b, i & 1); }
    out[MODE] = (clock64() - t0) / iters;
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < iters; ++i) { wait<MODE>(&a, i & 1); mbar_arrive(&b); }
  }
}

// the same ring as the GEMM main loop without any data: producer arrives on full[s] after empty[s]; consumer issues a
// tcgen05.commit (no MMAs pending -> immediate arrive) on empty[s] after full[s].  cycles per k-block step.
template <int MODE>
__global__ void ring(int iters, int stages, long long* out) {
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) { const int s = it % stages; wait<MODE>(&empty[s], ((it / stages) & 1) ^ 1); mbar_arrive(&full[s]); }
  } else if (threadIdx.x == 32) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) { const int s = it % stages; wait<MODE>(&full[s], (it / stages) & 1); tc_commit(&empty[s]); }
    out[4 + MODE] = (clock64() - t0) / iters;
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  pingpong<0><<<1, 64>>>(20000, d); pingpong<1><<<1, 64>>>(20000, d); pingpong<2><<<1, 64>>>(20000, d);
  ring<0><<<1, 64>>>(20000, 4, d); ring<1><<<1, 64>>>(20000, 4, d); ring<2><<<1, 64>>>(20000, 4, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("mbarrier ping-pong round trip (2 hand-offs), cycles: try_wait %lld | try_wait(hint 32ns) %lld | test_wait spin %lld   %s\n", h[0], h[1], h[2], cudaGetErrorString(e));
  printf("empty/full ring (4 stages, commit-arrive), cycles per step: try_wait %lld | hint %lld | test_wait %lld\n", h[4], h[5], h[6]);
  return 0;
}
