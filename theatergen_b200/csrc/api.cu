// extern "C" boundary: validates arguments and forwards to the launchers in namespace tg.
#include "../../include/theatergen.h"

#include "common.cuh"
#include "ops.h"

#define H(p) reinterpret_cast<const __half*>(p)
#define HM(p) reinterpret_cast<__half*>(p)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* tg_last_error(void) { return tg::get_error(); }
int tg_version(void) { return 100; }
unsigned long long tg_launch_count(void) { return tg::g_launch_count; }

int tg_gemm(const tg_half* A1, int lda1, int K1, const tg_half* A2, int lda2, const tg_half* Wt, int M, int N,
            int K, const tg_half* bias, const tg_half* rowbias, int rows_per_image, int rowbias_ld,
            const tg_half* residual, int ldr, int act, float out_scale, tg_half* out, int ldc, void* stream) {
  TG_CHECK(A1 && Wt && out, "tg_gemm: null pointer");
  tg::EpilogueSpec e;
  e.bias = H(bias); e.rowbias = H(rowbias); e.rows_per_image = rows_per_image; e.rowbias_ld = rowbias_ld;
  e.residual = H(residual); e.ldr = ldr; e.act = act; e.out_scale = out_scale;
  return tg::gemm_linear(H(A1), lda1, K1, H(A2), lda2, H(Wt), M, N, K, e, HM(out), ldc, ST(stream));
}

int tg_conv3x3(const tg_half* X, int B, int Hh, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias,
               const tg_half* rowbias, int rowbias_ld, const tg_half* residual, int ldr, float out_scale,
               tg_half* out, int ldc, void* stream) {
  TG_CHECK(X && Wt && out, "tg_conv3x3: null pointer");
  tg::EpilogueSpec e;
  e.bias = H(bias); e.rowbias = H(rowbias); e.rows_per_image = Hh * W; e.rowbias_ld = rowbias_ld;
  e.residual = H(residual); e.ldr = ldr; e.act = 0; e.out_scale = out_scale;
  return tg::conv3x3_nhwc(H(X), B, Hh, W, Cin, H(Wt), Cout, e, HM(out), ldc, ST(stream));
}

}  // extern "C"
