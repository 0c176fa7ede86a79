/* theatergen.h -- C-ABI boundary of libtheatergen.so (B200 / sm_100a).
 *
 * The reference (donahowe/TheaterGen) is pure Python on stock diffusers: its "plugin API" for the
 * denoise hot path is the diffusers attention-processor protocol plus the UNet module call.  It has no
 * FFI of its own, so every entry point below cites the reference Python interface whose arithmetic it
 * replaces; theatergen_b200/ (Python, ctypes) mirrors the reference names on top of these.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; tg_last_error() gives a thread-local message.
 *   - all pointers are DEVICE pointers owned by the caller (fp16 = IEEE binary16 unless stated);
 *     nothing is retained after the call returns; all work is enqueued on `stream` (a cudaStream_t
 *     passed as void*), never synchronises, and is CUDA-graph capturable.
 *   - activations are channels-last: an image tensor is [B, H, W, C]; a token tensor is [B*N, C].
 *   - there is no CPU fallback: without an sm_100a device every call fails.
 */
#ifndef THEATERGEN_H_
#define THEATERGEN_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t tg_half; /* raw IEEE fp16 bits */

const char* tg_last_error(void);
int tg_version(void);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
unsigned long long tg_launch_count(void);

/* activation codes of the GEMM epilogue */
#define TG_ACT_NONE 0
#define TG_ACT_SILU 1
#define TG_ACT_GEGLU 2 /* weight rows packed per 64: [32 value | 32 gate]; out has N/2 columns */

/* out[M,N] = act( [A1|A2][M,K] * Wt[N,K]^T + bias[N] + rowbias[m/rows_per_image, N] ) + residual[M,N]
 * tcgen05 tensor-core GEMM (TMA-fed, TMEM accumulators).  A2 may be NULL (K1 == K); bias / rowbias /
 * residual may be NULL.  M arbitrary; N, K, K1 multiples of 64.
 * Replaces nn.Linear / 1x1 conv call sites: ip_adapter/attention_processor.py:333,340-341,359,461,475-476,
 * 497-498,520; ip_adapter/attention.py:310,318,365-367 (GEGLU); ip_adapter/transformer_2d.py:290,339;
 * ip_adapter/resampler.py:13-20,59-76,127-147; ip_adapter/ip_adapter.py:41-47; diffusers ResnetBlock2D
 * conv_shortcut / time_emb_proj and TimestepEmbedding. */
int tg_gemm(const tg_half* A1, int lda1, int K1, const tg_half* A2, int lda2, const tg_half* Wt, int M, int N,
            int K, const tg_half* bias, const tg_half* rowbias, int rows_per_image, int rowbias_ld,
            const tg_half* residual, int ldr, int act, float out_scale, tg_half* out, int ldc, void* stream);

/* out[B*H*W, Cout] = conv3x3(pad=1, stride=1)(X[B,H,W,Cin]) + bias + rowbias[b] (+ residual) * out_scale
 * implicit GEMM on tcgen05; padding comes from TMA out-of-bounds zero fill.  Wt = [Cout,3,3,Cin].
 * Replaces the 3x3 convolutions of diffusers ResnetBlock2D / Upsample2D (call sites
 * ip_adapter/unet_2d_blocks.py:668,709,1093,1192,2239,2328,2184,2251,2298,2332). */
int tg_conv3x3(const tg_half* X, int B, int H, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias,
               const tg_half* rowbias, int rowbias_ld, const tg_half* residual, int ldr, float out_scale,
               tg_half* out, int ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THEATERGEN_H_ */
