// Per-instruction cost (cycles, single thread, back-to-back) of the synchronisation primitives a warp-specialised
// pipeline executes once per k-block.
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

__global__ void probe(const __half* src, long long* out) {
  __shared__ uint64_t bar_done, bar_big, bar_tx;
  __shared__ __align__(128) uint8_t buf[4096];
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1); mbar_init(&bar_big, 100000); mbar_init(&bar_tx, 1);
    fence_barrier_init();
    mbar_arrive(&bar_done);  // phase 0 complete: waiting for parity 0 succeeds immediately
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int N = 2000;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar_done)), "r"(0) : "memory");
    acc += ok;
    if (!ok) break;
  }
  out[0] = (clock64() - t0) / N;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar_done)), "r"(0) : "memory");
    acc += ok;
    if (!ok) break;
  }
  out[1] = (clock64() - t0) / N;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(smem_u32(&bar_done)) : "memory");
    acc += (uint32_t)(v >> 63);
    if (acc == 0xffffffffu) break;
  }
  out[2] = (clock64() - t0) / N;
  t0 = clock64();
  for (int i = 0; i < N; ++i) mbar_arrive(&bar_big);
  out[3] = (clock64() - t0) / N;
  t0 = clock64();
  for (int i = 0; i < N; ++i) mbar_arrive_expect_tx(&bar_big, 0);
  out[4] = (clock64() - t0) / N;
  // expect_tx + one 1 KB bulk copy + wait for it, strictly serial: issue->complete latency
  t0 = clock64();
  for (int i = 0; i < 200; ++i) {
    mbar_arrive_expect_tx(&bar_tx, 1024);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf)), "l"(src), "r"(1024), "r"(smem_u32(&bar_tx)) : "memory");
    mbar_wait(&bar_tx, i & 1);
  }
  out[5] = (clock64() - t0) / 200;
  // tcgen05.commit with nothing pending -> arrive; wait for it
  t0 = clock64();
  for (int i = 0; i < 200; ++i) {
    tc_commit(&bar_tx);
    mbar_wait(&bar_tx, i & 1);
  }
  out[6] = (clock64() - t0) / 200;
  // barrier word before/after for format inspection
  out[8] = *reinterpret_cast<volatile long long*>(&bar_done);
  out[9] = *reinterpret_cast<volatile long long*>(&bar_big);
  out[7] = acc;
}

int main() {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  __half* src; cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  probe<<<1, 64>>>(src, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  printf("cycles per op: try_wait(complete) %lld | test_wait(complete) %lld | ld.shared.b64 %lld | arrive %lld | arrive.expect_tx %lld\n", h[0], h[1], h[2], h[3], h[4]);
  printf("serial expect_tx + 1 KB bulk copy (L2 hit) + wait: %lld | tcgen05.commit (nothing pending) + wait: %lld   %s\n", h[5], h[6], cudaGetErrorString(e));
  printf("barrier words: completed-phase barrier 0x%016llx, pending barrier 0x%016llx\n", (unsigned long long)h[8], (unsigned long long)h[9]);
  return 0;
}
