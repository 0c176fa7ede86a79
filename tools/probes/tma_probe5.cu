// Bursts: one thread issues N TMA loads (128 x 64 fp16 boxes, L2-resident) that all signal ONE mbarrier, waits for it,
// repeats.  Cycles per burst vs N tells whether the TMA unit overlaps the loads of a burst (time ~ latency + N*small) or
// serialises them (time ~ N * latency).  Second table: same, but two independent bursts (two barriers) kept in flight.
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int rows, int N, int depth, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  const int box_bytes = rows * 128;
  if (threadIdx.x == 0) { for (int s = 0; s < 4; ++s) mbar_init(&bar[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  auto burst = [&](int it) {
    const int d = it % depth;
    mbar_arrive_expect_tx(&bar[d], N * box_bytes);
    for (int i = 0; i < N; ++i)
      tma_load_2d(smem + (d * N + i) * box_bytes, &tm, &bar[d], ((it * N + i) % 16) * 64, blockIdx.x * rows);
  };
  for (int it = 0; it < depth; ++it) burst(it);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    mbar_wait(&bar[it % depth], (it / depth) & 1);
    if (it + depth < iters) burst(it + depth);
  }
  out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int ctas = 148, K = 1024, iters = 2000;
  __half* A; cudaMalloc(&A, (size_t)ctas * 256 * K * 2); cudaMemset(A, 0, (size_t)ctas * 256 * K * 2);
  long long* d_out; cudaMalloc(&d_out, ctas * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%-5s %-3s %-6s | cycles per burst | per load | B/clk/SM\n", "rows", "N", "depth");
  for (int rows : {8, 128}) for (int depth : {1, 2, 4}) for (int N : {1, 2, 4, 8}) {
    if ((size_t)depth * N * rows * 128 + 1024 > 200 * 1024) continue;
    CUtensorMap tm;
    if (make_tmap_2d(&tm, A, (uint64_t)ctas * 256, K, K, rows)) { printf("tmap failed\n"); return 1; }
    cudaMemset(d_out, 0, ctas * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) probe<<<ctas, 64, depth * N * rows * 128 + 1024>>>(tm, rows, N, depth, iters, d_out);
    cudaError_t err = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (err == cudaSuccess) err = e2;
    long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double clk = 0; for (int i = 0; i < ctas; ++i) clk += h[i]; clk /= ctas;
    printf("%-5d %-3d %-6d | %8.0f | %7.0f | %6.1f  %s\n", rows, N, depth, clk / iters, clk / iters / N, (double)N * rows * 128 / (clk / iters), err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}
