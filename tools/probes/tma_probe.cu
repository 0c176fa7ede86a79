// TMA ingest ceiling: bytes per clock per SM that cp.async.bulk.tensor delivers into shared memory when nothing
// consumes the data (a consumer thread only recycles the stages).  Access pattern = the GEMM's: per k-block one
// 128 x 64 A box (rows of the CTA's own M-tile, streamed from HBM) and one BN x 64 B box (weights, shared by all CTAs -> L2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tma_probe.cu -L../../theatergen_b200 -ltheatergen -Xlinker -rpath=...
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                               int bn, int nk, int stages, int passes, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  const int stage_bytes = 128 * 128 + bn * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int total = nk * passes;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < total; ++it) {
      const int s = it % stages, kb = it % nk;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
      uint8_t* sa = smem + s * stage_bytes;
      mbar_arrive_expect_tx(&full[s], stage_bytes);
      tma_load_2d(sa, &tmA, &full[s], kb * 64, blockIdx.x * 128);
      if (bn) tma_load_2d(sa + 128 * 128, &tmB, &full[s], kb * 64, 0);
    }
    out[2 * blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    const long long t0 = clock64();
    for (int it = 0; it < total; ++it) {
      const int s = it % stages;
      mbar_wait(&full[s], (it / stages) & 1);
      mbar_arrive(&empty[s]);
    }
    out[2 * blockIdx.x + 1] = clock64() - t0;
  }
}

int main() {
  const int ctas = 148, K = 1024, nk = K / 64;  // A = 148*128 x 1024 fp16 = 38 MB: L2 resident after the warm-up
  __half *A, *B;
  cudaMalloc(&A, (size_t)ctas * 128 * K * 2);
  cudaMalloc(&B, (size_t)256 * K * 2);
  cudaMemset(A, 0, (size_t)ctas * 128 * K * 2);
  cudaMemset(B, 0, (size_t)256 * K * 2);
  long long* d_out; cudaMalloc(&d_out, 2 * ctas * sizeof(long long)); cudaMemset(d_out, 0, 2 * ctas * sizeof(long long));
  if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) { printf("attr failed\n"); return 1; }
  printf("TMA ingest, %d CTAs (one per SM), A (%d MB) and B both L2-resident\n", ctas, (int)((size_t)ctas * 128 * K * 2 >> 20));
  for (int bn : {0, 64, 128, 160, 256}) {
    for (int stages : {2, 4, 6, 8, 12}) {
      const int stage_bytes = 128 * 128 + bn * 128;
      if ((size_t)stages * stage_bytes + 1024 > 220 * 1024) continue;
      CUtensorMap ma, mb;
      if (make_tmap_2d(&ma, A, (uint64_t)ctas * 128, K, K, 128) || make_tmap_2d(&mb, B, 256, K, K, bn ? bn : 64)) { printf("tmap failed: %s\n", get_error()); return 1; }
      const int passes = 16;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaMemset(d_out, 0, 2 * ctas * sizeof(long long));
      probe<<<ctas, 64, stages * stage_bytes + 1024>>>(ma, mb, bn, nk, stages, passes, d_out);
      cudaEventRecord(e0);
      probe<<<ctas, 64, stages * stage_bytes + 1024>>>(ma, mb, bn, nk, stages, passes, d_out);
      cudaEventRecord(e1);
      cudaError_t err = cudaGetLastError(); cudaError_t err2 = cudaDeviceSynchronize(); if (err == cudaSuccess) err = err2;
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[2 * 148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      double clk = 0; for (int i = 0; i < ctas; ++i) clk += h[2 * i + 1]; clk /= ctas;
      const double bytes = (double)stage_bytes * nk * passes;
      printf("BN=%3d stages=%d (%3d KB in flight/SM): %6.0f clk per k-block, %5.1f B/clk/SM, aggregate %6.2f TB/s by events %s\n", bn, stages,
             stages * stage_bytes >> 10, clk / (nk * passes), bytes / clk, bytes * ctas / (ms * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
