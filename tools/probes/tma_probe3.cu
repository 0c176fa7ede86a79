// Follow-up to tma_probe2: what does the ~530-cycle-per-stage floor of a TMA-fed ring depend on?
//   kind 0: 2D tensor box [rows x 64 fp16] SWIZZLE_128B (gemm.cu's loads), P producer warps taking stages round-robin
//   kind 1: 1D bulk copy (cp.async.bulk, contiguous bytes) of the same size
//   kind 2: 3D tensor box [kc x rows x 64] (kc k-blocks of one row block in ONE instruction)
// One CTA per SM, sources L2-resident, a consumer warp recycles stages.
#include <cstdio>
#include "../../theatergen_b200/csrc/common.cuh"
using namespace tg;

__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(192, 1) probe(const __grid_constant__ CUtensorMap tm, const __half* src, int kind, int P,
                                                int rows, int kc, int nk, int stages, int total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  const int stage_bytes = rows * 128 * kc;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < P && lane == 0) {
    for (int it = warp; it < total; it += P) {
      const int s = it % stages, kb = (it * kc) % nk;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
      uint8_t* sa = smem + s * stage_bytes;
      mbar_arrive_expect_tx(&full[s], stage_bytes);
      if (kind == 0) tma_load_2d(sa, &tm, &full[s], kb * 64, blockIdx.x * rows);
      else if (kind == 1) bulk_1d(sa, src + ((size_t)blockIdx.x * 16 + (it % 16)) * 32768 / 2, stage_bytes, &full[s]);
      else tma_load_3d(sa, &tm, &full[s], 0, blockIdx.x * rows, kb);
    }
  } else if (warp == 5 && lane == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < total; ++it) {
      const int s = it % stages;
      mbar_wait(&full[s], (it / stages) & 1);
      mbar_arrive(&empty[s]);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

static int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t kc);

int main() {
  const int ctas = 148, K = 1024, nk = K / 64, total = 4096;
  __half* A; cudaMalloc(&A, (size_t)ctas * 256 * K * 2); cudaMemset(A, 0, (size_t)ctas * 256 * K * 2);
  long long* d_out; cudaMalloc(&d_out, ctas * sizeof(long long));
  if (cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) { printf("attr failed\n"); return 1; }
  printf("%-5s %-3s %-5s %-3s %-7s %-9s | cycles per stage, B/clk/SM\n", "kind", "P", "rows", "kc", "stages", "bytes");
  struct Cfg { int kind, P, rows, kc, stages; };
  const Cfg cfgs[] = {{0, 1, 128, 1, 4}, {0, 2, 128, 1, 4}, {0, 4, 128, 1, 4}, {0, 4, 128, 1, 8}, {0, 1, 32, 1, 4}, {0, 1, 8, 1, 4},
                      {0, 4, 256, 1, 4}, {1, 1, 128, 1, 4}, {1, 1, 256, 1, 4}, {1, 2, 128, 1, 4}, {1, 4, 128, 1, 8},
                      {2, 1, 128, 2, 4}, {2, 1, 128, 4, 3}, {2, 2, 128, 2, 4}, {2, 1, 64, 4, 4}};
  for (const Cfg& c : cfgs) {
    const int stage_bytes = c.rows * 128 * c.kc;
    CUtensorMap tm;
    int rc = c.kind == 2 ? make_tmap_3d(&tm, A, (uint64_t)ctas * 256, K, c.rows, c.kc) : make_tmap_2d(&tm, A, (uint64_t)ctas * 256, K, K, c.rows);
    if (rc) { printf("tmap failed: %s\n", get_error()); continue; }
    cudaMemset(d_out, 0, ctas * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) probe<<<ctas, 192, c.stages * stage_bytes + 1024>>>(tm, A, c.kind, c.P, c.rows, c.kc, nk, c.stages, total, d_out);
    cudaError_t err = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (err == cudaSuccess) err = e2;
    long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double clk = 0; for (int i = 0; i < ctas; ++i) clk += h[i]; clk /= ctas;
    printf("%-5d %-3d %-5d %-3d %-7d %-9d | %7.0f  %6.1f  %s\n", c.kind, c.P, c.rows, c.kc, c.stages, stage_bytes, clk / total, stage_bytes / (clk / total), err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}

// [k_outer = K/64][rows][64] view of a row-major [rows, K] fp16 matrix; box = [kc][box_rows][64], SWIZZLE_128B
#include <cudaTypedefs.h>
static int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t kc) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  cuuint64_t dims[3] = {64, rows, K / 64};
  cuuint64_t strides[2] = {K * 2, 128};
  cuuint32_t box[3] = {64, box_rows, kc};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode3d failed %d\n", (int)r); return -1; }
  return 0;
}
