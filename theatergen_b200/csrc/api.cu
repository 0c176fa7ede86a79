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
int tg_init(void) { return tg::gemm_init(); }

// one-shot GroupNorm-statistics target of the next tg_gemm / tg_conv3x3* call of this host thread
static thread_local long long* t_stats = nullptr;
static thread_local int t_stats_hw = 0;
int tg_gemm_set_stats(long long* stats, int rows_per_image) {
  TG_CHECK(stats == nullptr || rows_per_image > 0, "tg_gemm_set_stats: rows_per_image must be positive");
  t_stats = stats;
  t_stats_hw = rows_per_image;
  return 0;
}
static void take_stats(tg::EpilogueSpec& e) {
  e.stats = t_stats;
  e.stats_hw = t_stats_hw;
  t_stats = nullptr;
  t_stats_hw = 0;
}

int tg_gemm(const tg_half* A1, int lda1, int K1, const tg_half* A2, int lda2, const tg_half* Wt, int M, int N,
            int K, const tg_half* bias, const tg_half* rowbias, int rows_per_image, int rowbias_ld,
            const tg_half* residual, int ldr, int act, float out_scale, tg_half* out, int ldc, void* stream) {
  TG_CHECK(A1 && Wt && out, "tg_gemm: null pointer");
  tg::EpilogueSpec e;
  e.bias = H(bias); e.rowbias = H(rowbias); e.rows_per_image = rows_per_image; e.rowbias_ld = rowbias_ld;
  e.residual = H(residual); e.ldr = ldr; e.act = act; e.out_scale = out_scale;
  take_stats(e);
  return tg::gemm_linear(H(A1), lda1, K1, H(A2), lda2, H(Wt), M, N, K, e, HM(out), ldc, ST(stream));
}

int tg_cross_scores(const tg_half* X, int ldx, const tg_half* Kf, int B, int HW, int C, int heads, int n_text, int n_ip,
                    const float* ip_scale, tg_half* P, int ldp, void* stream) {
  TG_CHECK(X && Kf && P, "tg_cross_scores: null pointer");
  TG_CHECK(B > 0 && HW > 0 && heads > 0 && heads % 2 == 0 && C % 64 == 0, "tg_cross_scores: B=%d HW=%d heads=%d C=%d", B, HW, heads, C);
  tg::EpilogueSpec e;
  e.act = 5; e.b_img_rows = HW; e.sm_text = n_text; e.sm_ip = n_ip; e.sm_ipscale = ip_scale;
  return tg::gemm_linear(H(X), ldx, C, nullptr, 0, H(Kf), B * HW, heads * 96, C, e, HM(P), ldp, ST(stream));
}

int tg_cross_apply(const tg_half* P, int ldp, const tg_half* Vf, int B, int HW, int C, int heads, const tg_half* bias,
                   const tg_half* residual, int ldr, tg_half* out, int ldc, void* stream) {
  TG_CHECK(P && Vf && out, "tg_cross_apply: null pointer");
  TG_CHECK(B > 0 && HW > 0 && heads > 0 && C % 64 == 0, "tg_cross_apply: B=%d HW=%d heads=%d C=%d", B, HW, heads, C);
  tg::EpilogueSpec e;
  e.bias = H(bias); e.residual = H(residual); e.ldr = ldr; e.b_img_rows = HW;
  return tg::gemm_linear(H(P), ldp, heads * 96, nullptr, 0, H(Vf), B * HW, C, heads * 96, e, HM(out), ldc, ST(stream));
}

int tg_conv3x3(const tg_half* X, int B, int Hh, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias,
               const tg_half* rowbias, int rowbias_ld, const tg_half* residual, int ldr, int act, float out_scale,
               tg_half* out, int ldc, void* stream) {
  TG_CHECK(X && Wt && out, "tg_conv3x3: null pointer");
  TG_CHECK(act == 0 || act == 1 || act == 3, "tg_conv3x3: act must be none / silu / gelu");
  tg::EpilogueSpec e;
  e.bias = H(bias); e.rowbias = H(rowbias); e.rows_per_image = Hh * W; e.rowbias_ld = rowbias_ld;
  e.residual = H(residual); e.ldr = ldr; e.act = act; e.out_scale = out_scale;
  take_stats(e);
  return tg::conv3x3_nhwc(H(X), B, Hh, W, Cin, H(Wt), Cout, e, HM(out), ldc, ST(stream));
}

int tg_conv3x3_s2(const tg_half* X, int B, int Hh, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias, int pad_lo,
                  int act, tg_half* out, int ldc, void* stream) {
  TG_CHECK(X && Wt && out, "tg_conv3x3_s2: null pointer");
  TG_CHECK(pad_lo == 0 || pad_lo == 1, "tg_conv3x3_s2: pad_lo must be 0 or 1");
  TG_CHECK(act == 0 || act == 1, "tg_conv3x3_s2: act must be none / silu");
  tg::EpilogueSpec e;
  e.bias = H(bias); e.act = act;
  take_stats(e);
  return tg::conv3x3_s2_nhwc(H(X), B, Hh, W, Cin, H(Wt), Cout, pad_lo, e, HM(out), ldc, ST(stream));
}

int tg_upsample2x_conv3x3(const tg_half* X, int B, int Hh, int W, int Cin, const tg_half* Wt4, int Cout, const tg_half* bias,
                          tg_half* out, void* stream) {
  TG_CHECK(X && Wt4 && out, "tg_upsample2x_conv3x3: null pointer");
  tg::EpilogueSpec e;
  e.bias = H(bias);
  take_stats(e);
  return tg::upsample2x_conv3x3_nhwc(H(X), B, Hh, W, Cin, H(Wt4), Cout, e, HM(out), ST(stream));
}


int tg_groupnorm(const tg_half* x1, int C1, const tg_half* x2, int C2, int B, int HW, const tg_half* gamma,
                 const tg_half* beta, float eps, int silu, tg_half* out, float* workspace, void* stream) {
  TG_CHECK(x1 && gamma && beta && out && (C2 == 0 || x2), "tg_groupnorm: null pointer");
  return tg::groupnorm_nhwc(H(x1), C1, H(x2), C2, B, HW, H(gamma), H(beta), eps, silu, HM(out), workspace, ST(stream));
}

int tg_gemm_set_scratch_slot(int slot) { return tg::gemm_set_scratch_slot(slot); }
int tg_set_pdl(int mode) {
  tg::set_pdl(mode);
  return 0;
}

int tg_attn_cross_set_impl(int mode) {
  TG_CHECK(mode >= -1 && mode <= 1, "tg_attn_cross_set_impl: %d (-1 auto, 0 mma.sync, 1 tcgen05)", mode);
  tg::attn_cross_set_impl(mode);
  return 0;
}
int tg_attn_set_variant(int v) {
  TG_CHECK(v == 0 || v == 2 || v == 3, "tg_attn_set_variant: %d (0, 2 or 3)", v);
  tg::attn_set_variant(v);
  return 0;
}
int tg_attn_set_prof(void* dev_buf) {
  tg::attn_set_prof(static_cast<long long*>(dev_buf));
  return 0;
}
int tg_gemm_set_prof(void* dev_buf) {
  tg::gemm_set_prof(static_cast<long long*>(dev_buf));
  return 0;
}

int tg_groupnorm_apply(const tg_half* x1, int C1, const long long* stats1, const long long* stats1_alt, const tg_half* x2, int C2,
                       const long long* stats2, const long long* stats2_alt, int alt_b0, int alt_b1, int B, int HW,
                       const tg_half* gamma, const tg_half* beta, float eps, int silu, tg_half* out, void* stream) {
  TG_CHECK(x1 && gamma && beta && out && (C2 == 0 || x2), "tg_groupnorm_apply: null pointer");
  return tg::groupnorm_apply_nhwc(H(x1), C1, stats1, stats1_alt, H(x2), C2, stats2, stats2_alt, alt_b0, alt_b1, B, HW, H(gamma),
                                  H(beta), eps, silu, HM(out), ST(stream));
}

int tg_groupnorm_tune(int stats_slabs, int apply_ctas_per_sm, int fused) {
  tg::groupnorm_tune(stats_slabs, apply_ctas_per_sm, fused);
  return 0;
}

int tg_layernorm(const tg_half* x, int ldx, const tg_half* gamma, const tg_half* beta, float eps, tg_half* out,
                 int ldo, int M, int C, void* stream) {
  TG_CHECK(x && gamma && beta && out, "tg_layernorm: null pointer");
  return tg::layernorm_rows(H(x), ldx, H(gamma), H(beta), eps, HM(out), ldo, M, C, ST(stream));
}

int tg_attention(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                 int ldo, int B, int heads, int Nq, int Nk, int d, float scale, int causal, void* stream) {
  TG_CHECK(Q && K && V && O && Nq > 0 && Nk > 0, "tg_attention: bad arguments");
  return tg::attention_flash(H(Q), H(K), H(V), ldq, ldk, ldv, HM(O), ldo, B, heads, Nq, Nk, d, scale, causal, ST(stream));
}

int tg_attention_tc(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                    int ldo, int B, int heads, int N, int d, float scale, void* stream) {
  TG_CHECK(Q && K && V && O && N > 0, "tg_attention_tc: bad arguments");
  return tg::attention_tc(H(Q), H(K), H(V), ldq, ldk, ldv, HM(O), ldo, B, heads, N, d, scale, 0, ST(stream));
}

int tg_attention_tc_packed(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                           int ldo, int B, int heads, int N, int d, float scale, void* stream) {
  TG_CHECK(Q && K && V && O && N > 0, "tg_attention_tc_packed: bad arguments");
  return tg::attention_tc(H(Q), H(K), H(V), ldq, ldk, ldv, HM(O), ldo, B, heads, N, d, scale, 1, ST(stream));
}

int tg_attention_cross_ip(const tg_half* Q, int ldq, const tg_half* Kt, const tg_half* Vt, const tg_half* Ki,
                          const tg_half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                          const tg_half* add_mask, tg_half* O, int ldo, tg_half* probs, int probs_b0, int B,
                          int heads, int Nq, int d, float scale, void* stream) {
  TG_CHECK(Q && Kt && Vt && O && (n_ip == 0 || (Ki && Vi)), "tg_attention_cross_ip: null pointer");
  return tg::attention_cross_ip(H(Q), ldq, H(Kt), H(Vt), H(Ki), H(Vi), ldkv, n_text, n_ip, ip_scale, H(add_mask),
                                HM(O), ldo, HM(probs), probs_b0, B, heads, Nq, d, scale, ST(stream));
}

int tg_upsample2x(const tg_half* x, int B, int Hh, int W, int C, tg_half* y, void* stream) {
  TG_CHECK(x && y, "tg_upsample2x: null pointer");
  return tg::upsample2x_nhwc(H(x), B, Hh, W, C, HM(y), ST(stream));
}
int tg_im2col_s2(const tg_half* x, int B, int Hh, int W, int C, tg_half* a, int pad_lo, void* stream) {
  TG_CHECK(x && a, "tg_im2col_s2: null pointer");
  return tg::im2col_s2_nhwc(H(x), B, Hh, W, C, HM(a), pad_lo, ST(stream));
}
int tg_nhwc_head_to_nchw(const tg_half* x, tg_half* y, int B, int Cs, int Cout, int HW, void* stream) {
  TG_CHECK(x && y, "tg_nhwc_head_to_nchw: null pointer");
  return tg::nhwc_head_to_nchw(H(x), HM(y), B, Cs, Cout, HW, ST(stream));
}
int tg_gemm_set_tuned(int M, int N, int K, int act, int conv, int bn, int split, int pair) {
  return tg::gemm_set_tuned(M, N, K, act, conv, bn, split, pair);
}
int tg_gemm_clear_tuned(void) { return tg::gemm_clear_tuned(); }
int tg_gemm_force_tile(int bn, int split, int pair) { return tg::gemm_force_tile(bn, split, pair); }
int tg_layernorm_bwd(const tg_half* x, int ldx, const tg_half* gamma, const tg_half* dy, int lddy, tg_half* dx, int lddx,
                     int M, int C, float eps, void* stream) {
  TG_CHECK(x && gamma && dy && dx, "tg_layernorm_bwd: null pointer");
  return tg::layernorm_bwd(H(x), ldx, H(gamma), H(dy), lddy, HM(dx), lddx, M, C, eps, ST(stream));
}
int tg_groupnorm_bwd(const tg_half* x, const tg_half* gamma, const tg_half* beta, const tg_half* dy, tg_half* dx, int B,
                     int HW, int C, float eps, int silu, void* stream) {
  TG_CHECK(x && gamma && beta && dy && dx, "tg_groupnorm_bwd: null pointer");
  return tg::groupnorm_bwd(H(x), H(gamma), H(beta), H(dy), HM(dx), B, HW, C, eps, silu, ST(stream));
}
int tg_softmax_bwd_rows(const tg_half* p, int ldp, const tg_half* dp, int lddp, tg_half* ds, int ldds, int M, int N,
                        float scale, void* stream) {
  TG_CHECK(p && dp && ds, "tg_softmax_bwd_rows: null pointer");
  return tg::softmax_bwd_rows(H(p), ldp, H(dp), lddp, HM(ds), ldds, M, N, scale, ST(stream));
}
int tg_geglu_bwd(const tg_half* pre, const tg_half* dy, tg_half* dpre, int M, int F, void* stream) {
  TG_CHECK(pre && dy && dpre, "tg_geglu_bwd: null pointer");
  return tg::geglu_bwd(H(pre), H(dy), HM(dpre), M, F, ST(stream));
}
int tg_attention_cross_ip_bwd_q(const tg_half* q, int ldq, const tg_half* kt, const tg_half* vt, const tg_half* ki,
                                const tg_half* vi, int ldkv, int n_text, int n_ip, float ip_scale, const tg_half* dout,
                                int lddo, const tg_half* dprobs, tg_half* dq, int lddq, int Nq, int heads, int d,
                                float alpha, void* stream) {
  TG_CHECK(q && kt && vt && dout && dq, "tg_attention_cross_ip_bwd_q: null pointer");
  return tg::cross_ip_bwd_q(H(q), ldq, H(kt), H(vt), H(ki), H(vi), ldkv, n_text, n_ip, ip_scale, H(dout), lddo, H(dprobs),
                            HM(dq), lddq, Nq, heads, d, alpha, ST(stream));
}
int tg_upsample2x_bwd(const tg_half* dy, int B, int Hh, int W, int C, tg_half* dx, void* stream) {
  TG_CHECK(dy && dx, "tg_upsample2x_bwd: null pointer");
  return tg::upsample2x_bwd(H(dy), B, Hh, W, C, HM(dx), ST(stream));
}
int tg_col2im_s2(const tg_half* dcols, int B, int Hh, int W, int C, tg_half* dx, void* stream) {
  TG_CHECK(dcols && dx, "tg_col2im_s2: null pointer");
  return tg::col2im_s2(H(dcols), B, Hh, W, C, HM(dx), ST(stream));
}
int tg_guidance_update(tg_half* latents, const tg_half* grad, float inv_scale, float step, float* grad_out, int* bad,
                       int n, void* stream) {
  TG_CHECK(latents && grad, "tg_guidance_update: null pointer");
  return tg::guidance_update(HM(latents), H(grad), inv_scale, step, grad_out, bad, n, ST(stream));
}
int tg_embed_tokens(const long long* ids, const tg_half* table, const tg_half* pos, tg_half* out, int B, int T, int C,
                    int vocab, void* stream) {
  TG_CHECK(ids && table && pos && out, "tg_embed_tokens: null pointer");
  return tg::embed_tokens(ids, H(table), H(pos), HM(out), B, T, C, vocab, ST(stream));
}
int tg_im2col_patch(const tg_half* x, tg_half* out, int B, int Cin, int Hh, int W, int P, int Kpad, void* stream) {
  TG_CHECK(x && out, "tg_im2col_patch: null pointer");
  return tg::im2col_patch(H(x), HM(out), B, Cin, Hh, W, P, Kpad, ST(stream));
}
int tg_gaussian_sample(const tg_half* moments, const tg_half* noise, tg_half* z, int B, int L, int HW, float out_scale,
                       void* stream) {
  TG_CHECK(moments && z, "tg_gaussian_sample: null pointer");
  return tg::gaussian_sample(H(moments), H(noise), HM(z), B, L, HW, out_scale, ST(stream));
}
int tg_conv_in(const tg_half* x, const tg_half* w, const tg_half* bias, tg_half* y, int B, int Cin, int Hh, int W,
               int Cout, int silu, void* stream) {
  TG_CHECK(x && w && bias && y, "tg_conv_in: null pointer");
  return tg::conv_in_nchw_to_nhwc(H(x), H(w), H(bias), HM(y), B, Cin, Hh, W, Cout, silu, ST(stream));
}
int tg_im2col_small(const tg_half* x, tg_half* out, int B, int Cin, int Hh, int W, void* stream) {
  TG_CHECK(x && out, "tg_im2col_small: null pointer");
  return tg::im2col_small_nchw(H(x), HM(out), B, Cin, Hh, W, ST(stream));
}
int tg_conv_out(const tg_half* x, const tg_half* w, const tg_half* bias, tg_half* y, int B, int Hh, int W, int C,
                int Cout, void* stream) {
  TG_CHECK(x && w && bias && y, "tg_conv_out: null pointer");
  return tg::conv_out_nhwc_to_nchw(H(x), H(w), H(bias), HM(y), B, Hh, W, C, Cout, ST(stream));
}
int tg_pointwise_nchw(const tg_half* x, const tg_half* w, const tg_half* bias, tg_half* y, int B, int Cin, int Cout, int HW,
                      void* stream) {
  TG_CHECK(x && w && bias && y, "tg_pointwise_nchw: null pointer");
  return tg::pointwise_nchw(H(x), H(w), H(bias), HM(y), B, Cin, Cout, HW, ST(stream));
}
int tg_softmax_rows(const tg_half* x, int ldx, tg_half* out, int ldo, int M, int N, float scale, void* stream) {
  TG_CHECK(x && out && M > 0, "tg_softmax_rows: bad arguments");
  return tg::softmax_rows(H(x), ldx, HM(out), ldo, M, N, scale, ST(stream));
}
int tg_timestep_embedding(const float* t_dev, tg_half* out, int B, int dim, void* stream) {
  TG_CHECK(t_dev && out && dim % 2 == 0, "tg_timestep_embedding: bad arguments");
  return tg::timestep_embedding(t_dev, HM(out), B, dim, ST(stream));
}
int tg_add_inplace(tg_half* y, const tg_half* x, long long n, void* stream) {
  TG_CHECK(x && y && n >= 0, "tg_add_inplace: bad arguments");
  return tg::add_inplace(HM(y), H(x), (size_t)n, ST(stream));
}
int tg_nchw_to_nhwc(const tg_half* x, tg_half* y, int B, int C, int HW, void* stream) {
  TG_CHECK(x && y, "tg_nchw_to_nhwc: null pointer");
  return tg::nchw_to_nhwc(H(x), HM(y), B, C, HW, ST(stream));
}
int tg_duplicate_latents(const tg_half* lat, tg_half* out, int n, void* stream) {
  TG_CHECK(lat && out, "tg_duplicate_latents: null pointer");
  return tg::duplicate_latents(H(lat), HM(out), n, ST(stream));
}
int tg_cfg_ddim_composite(const tg_half* eps, tg_half* latents, const float* coef_dev, const tg_half* frozen,
                          const tg_half* mask, int K, int C, int HW, tg_half* latents_copy, int pair_stride,
                          void* stream) {
  TG_CHECK(eps && latents && coef_dev && ((frozen == nullptr) == (mask == nullptr)), "tg_cfg_ddim_composite: bad arguments");
  return tg::cfg_ddim_composite(H(eps), HM(latents), coef_dev, H(frozen), H(mask), K, C, HW, HM(latents_copy), pair_stride,
                                ST(stream));
}

int tg_guidance_energy(const tg_half* const* maps, tg_half* const* dmaps, const int* map_hw, int n_keys,
                       int heads, int n_tok, const double* boxes, const int* box_obj, int n_box,
                       const tg_guidance_item* items, int n_items, const tg_guidance_item* ref_items,
                       const tg_half* const* ref_maps, int n_ref_items, float fg_top_p, float bg_top_p,
                       float fg_weight, float bg_weight, float grad_scale, float ref_eps, float* loss_out,
                       float* workspace, void* stream) {
  TG_CHECK(maps && map_hw && loss_out && workspace, "tg_guidance_energy: null pointer");
  TG_CHECK(n_keys <= tg::GUIDANCE_MAX_KEYS && n_box <= tg::GUIDANCE_MAX_BOXES && n_items <= tg::GUIDANCE_MAX_ITEMS &&
               n_ref_items <= tg::GUIDANCE_MAX_REF_ITEMS,
           "tg_guidance_energy: limits are %d keys, %d boxes, %d items, %d ref items", tg::GUIDANCE_MAX_KEYS,
           tg::GUIDANCE_MAX_BOXES, tg::GUIDANCE_MAX_ITEMS, tg::GUIDANCE_MAX_REF_ITEMS);
  tg::GuidanceParams P{};
  for (int k = 0; k < n_keys; ++k) {
    P.maps[k] = H(maps[k]);
    P.dmaps[k] = dmaps ? HM(dmaps[k]) : nullptr;
    P.map_hw[k] = map_hw[k];
  }
  for (int b = 0; b < n_box; ++b) {
    for (int j = 0; j < 4; ++j) P.boxes[b][j] = boxes[b * 4 + j];
    P.box_obj[b] = box_obj[b];
  }
  for (int i = 0; i < n_items; ++i) {
    P.items[i].key = items[i].key; P.items[i].obj = items[i].obj; P.items[i].box = items[i].box;
    P.items[i].ref = items[i].ref; P.items[i].tok = items[i].tok; P.items[i].weight = items[i].weight;
    TG_CHECK(items[i].key < n_keys && items[i].tok < n_tok, "tg_guidance_energy: item %d out of range", i);
  }
  for (int i = 0; i < n_ref_items; ++i) {
    P.ref_items[i].key = ref_items[i].key; P.ref_items[i].obj = ref_items[i].obj; P.ref_items[i].box = ref_items[i].box;
    P.ref_items[i].ref = ref_items[i].ref; P.ref_items[i].tok = ref_items[i].tok; P.ref_items[i].weight = ref_items[i].weight;
    TG_CHECK(ref_items[i].key < n_keys && ref_items[i].tok < n_tok && ref_items[i].box < n_box && ref_maps,
             "tg_guidance_energy: ref item %d out of range", i);
    P.ref_maps[i] = H(ref_maps[ref_items[i].ref]);
    P.ref_items[i].ref = (unsigned char)i;
  }
  P.n_keys = n_keys; P.n_box = n_box; P.n_items = n_items; P.n_ref_items = n_ref_items; P.heads = heads; P.n_tok = n_tok;
  P.fg_top_p = fg_top_p; P.bg_top_p = bg_top_p; P.fg_weight = fg_weight; P.bg_weight = bg_weight;
  P.grad_scale = grad_scale; P.ref_eps = ref_eps;
  return tg::guidance_energy(P, loss_out, workspace, ST(stream));
}

}  // extern "C"
