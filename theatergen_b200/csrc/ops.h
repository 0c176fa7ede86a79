// Internal C++ declarations of the kernel launchers (the public boundary is include/theatergen.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace tg {

struct EpilogueSpec {
  const __half* bias = nullptr;      // [N]
  const __half* rowbias = nullptr;   // [images, N] broadcast over the rows of an image (ResBlock temb add)
  int rows_per_image = 1;
  int rowbias_ld = 0;
  const __half* residual = nullptr;  // [M, N] with pitch ldr
  int ldr = 0;
  int act = 0;                       // 0 none, 1 SiLU, 2 GEGLU (packed weight layout)
  float out_scale = 1.f;
};

// out[M,N] = epi( [A1 | A2][M,K] * Wt[N,K]^T )
int gemm_linear(const __half* A1, int lda1, int K1, const __half* A2, int lda2, const __half* Wt, int M, int N,
                int K, const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st);
// out[B*H*W, Cout] = epi( conv3x3_pad1(X[B,H,W,Cin]) ), Wt = [Cout, 3, 3, Cin]
int conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, const EpilogueSpec& e,
                 __half* out, int ldc, cudaStream_t st);

}  // namespace tg
