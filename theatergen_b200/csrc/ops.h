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
  // optional GroupNorm statistics of the stored output, accumulated by the epilogue (api: tg_gemm_set_stats):
  // stats[image = row / stats_hw][column pair][2] += 2^24 x (sum, sum of squares) as 64-bit integers; zeroed by the caller
  long long* stats = nullptr;
  int stats_hw = 0;
  // per-image B operand + key-softmax epilogue (act 5): the folded cross-attention GEMMs (gemm.cu GemmArgs)
  int b_img_rows = 0;
  int sm_text = 0, sm_ip = 0;
  const float* sm_ipscale = nullptr;
};

// out[M,N] = epi( [A1 | A2][M,K] * Wt[N,K]^T )
int gemm_linear(const __half* A1, int lda1, int K1, const __half* A2, int lda2, const __half* Wt, int M, int N,
                int K, const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st);
// out[B*H*W, Cout] = epi( conv3x3_pad1(X[B,H,W,Cin]) ), Wt = [Cout, 3, 3, Cin]
int conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, const EpilogueSpec& e,
                 __half* out, int ldc, cudaStream_t st);
// out[B*(H/2)*(W/2), Cout] = epi( conv3x3_stride2(X[B,H,W,Cin]) ), padding (pad_lo, 1)
int conv3x3_s2_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, int pad_lo,
                    const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st);
// out[B,2H,2W,Cout] = epi( conv3x3_pad1(nearest_x2(X[B,H,W,Cin])) ) as four 2x2 convolutions; Wt4 = [4][Cout,2,2,Cin]
int upsample2x_conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt4, int Cout, const EpilogueSpec& e,
                            __half* out, cudaStream_t st);


// allocates the per-device split-K workspace (call once outside stream capture; lazily done otherwise)
int gemm_init();
// selects the split-K workspace used by GEMMs enqueued from this host thread (one slot per concurrently running stream)
int gemm_set_scratch_slot(int slot);
// measurement hook: when non-null every GEMM CTA writes 8 counters {main-loop cycles, producer wait, MMA wait(full),
// MMA wait(accumulator), epilogue wait, k-blocks, tiles, -} at dev_buf[blockIdx.x*8..] (tools/gemm_probe.py)
void gemm_set_prof(long long* dev_buf);
void attn_set_prof(long long* dev_buf);
void attn_set_variant(int v);
void attn_cross_set_impl(int mode);

// GroupNorm(32 groups)(+SiLU) over the channel-concat [x1 | x2] (x2 may be null, C2 = 0); fp32 statistics.
// workspace: >= B*64*32*2 floats.
int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, int silu, __half* out, float* workspace, cudaStream_t st);
// the same normalisation from statistics accumulated by the producing GEMMs' epilogues (EpilogueSpec::stats): one pass
int groupnorm_apply_nhwc(const __half* x1, int C1, const long long* st1, const long long* st1_alt, const __half* x2, int C2,
                         const long long* st2, const long long* st2_alt, int alt_b0, int alt_b1, int B, int HW,
                         const __half* gamma, const __half* beta, float eps, int silu, __half* out, cudaStream_t st);
// launch-geometry knobs of groupnorm_nhwc (<= 0 / < 0 keep the current value); used by tools/gn_bench.py
void groupnorm_tune(int stats_slabs, int apply_ctas_per_sm, int fused);
int layernorm_rows(const __half* x, int ldx, const __half* gamma, const __half* beta, float eps, __half* out,
                   int ldo, int M, int C, cudaStream_t st);

// O[b, n, h*d..] = softmax(scale * Q K^T) V per (b, h); row pitches in elements
int attention_flash(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                    int B, int heads, int Nq, int Nk, int d, float scale, int causal, cudaStream_t st);
// tcgen05 flash self-attention; Q/K/V head-padded to 64 columns (head h at column h*64), N % 128 == 0, d <= 64
int attention_tc(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                 int B, int heads, int N, int d, float scale, int packed, cudaStream_t st);
// text + image-token cross attention (see attn.cu)
int attention_cross_ip(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki,
                       const __half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                       const __half* add_mask, __half* O, int ldo, __half* probs, int probs_b0, int B, int heads,
                       int Nq, int d, float scale, cudaStream_t st);
int attention_cross_ip_tc(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki,
                       const __half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                       const __half* add_mask, __half* O, int ldo, __half* probs, int probs_b0, int B, int heads,
                       int Nq, int d, float scale, cudaStream_t st);


int upsample2x_nhwc(const __half* x, int B, int H, int W, int C, __half* y, cudaStream_t st);
int im2col_s2_nhwc(const __half* x, int B, int H, int W, int C, __half* a, int pad_lo, cudaStream_t st);
int im2col_patch(const __half* x, __half* out, int B, int Cin, int H, int W, int P, int Kpad, cudaStream_t st);
int nhwc_head_to_nchw(const __half* x, __half* y, int B, int Cs, int Cout, int HW, cudaStream_t st);
int gemm_set_tuned(int M, int N, int K, int act, int conv, int bn, int split, int pair);
int gemm_clear_tuned();
int gemm_force_tile(int bn, int split, int pair);
// backward.cu: input gradients of the guidance step (models/pipelines.py:62-128)
int layernorm_bwd(const __half* x, int ldx, const __half* gamma, const __half* dy, int lddy, __half* dx, int lddx, int M,
                  int C, float eps, cudaStream_t st);
int groupnorm_bwd(const __half* x, const __half* gamma, const __half* beta, const __half* dy, __half* dx, int B, int HW,
                  int C, float eps, int silu, cudaStream_t st);
int softmax_bwd_rows(const __half* p, int ldp, const __half* dp, int lddp, __half* ds, int ldds, int M, int N,
                     float scale, cudaStream_t st);
int geglu_bwd(const __half* pre, const __half* dy, __half* dpre, int M, int F, cudaStream_t st);
int cross_ip_bwd_q(const __half* q, int ldq, const __half* kt, const __half* vt, const __half* ki, const __half* vi,
                   int ldkv, int n_text, int n_ip, float ip_scale, const __half* dout, int lddo, const __half* dprobs,
                   __half* dq, int lddq, int Nq, int heads, int d, float alpha, cudaStream_t st);
int upsample2x_bwd(const __half* dy, int B, int H, int W, int C, __half* dx, cudaStream_t st);
int col2im_s2(const __half* dcols, int B, int H, int W, int C, __half* dx, cudaStream_t st);
int guidance_update(__half* latents, const __half* grad, float inv_scale, float step, float* grad_out, int* bad, int n,
                    cudaStream_t st);
int embed_tokens(const long long* ids, const __half* table, const __half* pos, __half* out, int B, int T, int C,
                 int vocab, cudaStream_t st);
int gaussian_sample(const __half* moments, const __half* noise, __half* z, int B, int L, int HW, float out_scale,
                    cudaStream_t st);
int conv_in_nchw_to_nhwc(const __half* x, const __half* w, const __half* bias, __half* y, int B, int Cin, int H, int W,
                         int Cout, int silu, cudaStream_t st);
// [B,Cin,H,W] NCHW -> [B*H*W, 64]: 3x3 patches flattened as k = ci*9 + tap, zero padded (Cin*9 <= 64)
int im2col_small_nchw(const __half* x, __half* out, int B, int Cin, int H, int W, cudaStream_t st);
int conv_out_nhwc_to_nchw(const __half* x, const __half* w, const __half* bias, __half* y, int B, int H, int W, int C,
                          int Cout, cudaStream_t st);
int pointwise_nchw(const __half* x, const __half* w, const __half* bias, __half* y, int B, int Cin, int Cout, int HW,
                   cudaStream_t st);
int softmax_rows(const __half* x, int ldx, __half* out, int ldo, int M, int N, float scale, cudaStream_t st);
int timestep_embedding(const float* t_dev, __half* out, int B, int dim, cudaStream_t st);
int add_inplace(__half* y, const __half* x, size_t n, cudaStream_t st);
int nchw_to_nhwc(const __half* x, __half* y, int B, int C, int HW, cudaStream_t st);
int cfg_ddim_composite(const __half* eps, __half* latents, const float* coef_dev, const __half* frozen,
                       const __half* mask, int K, int C, int HW, __half* latents_copy, int pair_stride,
                       cudaStream_t st);
int duplicate_latents(const __half* lat, __half* out, int n, cudaStream_t st);

// ---- guidance energy (guidance.cu); the whole description travels as a kernel parameter (graph capturable) ----
static constexpr int GUIDANCE_MAX_KEYS = 8, GUIDANCE_MAX_BOXES = 16, GUIDANCE_MAX_ITEMS = 128,
                     GUIDANCE_MAX_REF_ITEMS = 32, GUIDANCE_MAX_HW = 256;
struct GuidanceItem {
  unsigned char key, obj, box, ref;
  int tok;
  float weight;
};
struct GuidanceParams {
  const __half* maps[GUIDANCE_MAX_KEYS];      // [heads, HW, n_tok] each
  __half* dmaps[GUIDANCE_MAX_KEYS];           // nullable, same shape: d(energy*grad_scale)/d(map), accumulated
  const __half* ref_maps[GUIDANCE_MAX_REF_ITEMS];  // [heads, HW] each
  int map_hw[GUIDANCE_MAX_KEYS];
  double boxes[GUIDANCE_MAX_BOXES][4];        // normalised x0,y0,x1,y1 (python floats)
  int box_obj[GUIDANCE_MAX_BOXES];
  GuidanceItem items[GUIDANCE_MAX_ITEMS];
  GuidanceItem ref_items[GUIDANCE_MAX_REF_ITEMS];
  int n_keys, n_box, n_items, n_ref_items, heads, n_tok;
  float fg_top_p, bg_top_p, fg_weight, bg_weight, grad_scale, ref_eps;
};
int guidance_energy(const GuidanceParams& P, float* loss_out, float* workspace, cudaStream_t st);

}  // namespace tg
