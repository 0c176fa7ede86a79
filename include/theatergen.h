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
 *   - process model: ONE PROCESS PER GPU (the launch model of bench.py / torch.distributed), one host thread issuing
 *     calls per device at a time.  The library keeps per-device state only where stated (the split-K workspace of
 *     tg_gemm / tg_conv3x3, keyed by (device, slot); kernel attributes are set once per (kernel, device)), so a process
 *     that drives several devices works as long as each device is driven by one thread at a time; concurrent host
 *     threads on the SAME device are not supported (tg_gemm_set_scratch_slot is per calling thread, the workspace
 *     of a slot is not locked).
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

/* allocate per-device scratch (64 MB split-K workspace) on the CURRENT device; call once before capturing CUDA
 * graphs.  Done lazily by the first GEMM otherwise.  Scratch is shared by all GEMM calls of the process on that
 * device: issue them on one stream at a time. */
int tg_init(void);

/* activation codes of the GEMM epilogue */
#define TG_ACT_NONE 0
#define TG_ACT_SILU 1
#define TG_ACT_GEGLU 2 /* weight rows packed per 64: [32 value | 32 gate]; out has N/2 columns */
#define TG_ACT_GELU 3  /* exact (erf) GELU, ip_adapter/resampler.py:18 */
#define TG_ACT_QUICK_GELU 4 /* x * sigmoid(1.702 x): the CLIP text encoder's MLP (generate.py:60 loads it) */

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

/* out[B*H*W, Cout] = (act(conv3x3(pad=1, stride=1)(X[B,H,W,Cin]) + bias + rowbias[b]) + residual) * out_scale
 * implicit GEMM on tcgen05; padding comes from TMA out-of-bounds zero fill.  Wt = [Cout,3,3,Cin]; Cin, Cout multiples
 * of 64; W divides 128 or is a multiple of 128.  act: 0 none, 1 SiLU.
 * Replaces the 3x3 convolutions of diffusers ResnetBlock2D / Upsample2D (call sites
 * ip_adapter/unet_2d_blocks.py:668,709,1093,1192,2239,2328,2184,2251,2298,2332) and of the stage-2 ControlNet's
 * conditioning embedder (stock diffusers ControlNetModel built at generate.py:91, called models/pipelines.py:770-778). */
int tg_conv3x3(const tg_half* X, int B, int H, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias,
               const tg_half* rowbias, int rowbias_ld, const tg_half* residual, int ldr, int act, float out_scale,
               tg_half* out, int ldc, void* stream);

/* out[B*(H/2)*(W/2), Cout] = act(conv3x3(stride 2)(X[B,H,W,Cin]) + bias): diffusers Downsample2D (call sites
 * ip_adapter/unet_2d_blocks.py:1035,1112,1161,1198; the ControlNet's conditioning embedder and encoder; pad_lo = 0: the VAE
 * encoder's one-sided padding).  Implicit GEMM: the TMA gathers every other pixel (element stride 2), no im2col buffer.
 * Wt = [Cout,3,3,Cin]; H, W even; the output width must divide 128 or be a multiple of it. */
int tg_conv3x3_s2(const tg_half* X, int B, int H, int W, int Cin, const tg_half* Wt, int Cout, const tg_half* bias, int pad_lo,
                  int act, tg_half* out, int ldc, void* stream);

/* out[B,2H,2W,Cout] = conv3x3(pad 1)(nearest_x2(X[B,H,W,Cin])) + bias: diffusers Upsample2D (ip_adapter/unet_2d_blocks.py:
 * 2184,2251,2298,2332) without the upsampled tensor: four 2x2 convolutions on the low-resolution grid, one per output
 * pixel parity (py, px), 2.25x fewer multiply-adds.  Wt4 = [4][Cout,2,2,Cin]: for phase p = 2 py + px, tap (a, b) is the sum
 * of the 3x3 taps (ky, kx) that read input row i + py - 1 + a and column j + px - 1 + b. */
int tg_upsample2x_conv3x3(const tg_half* X, int B, int H, int W, int Cin, const tg_half* Wt4, int Cout, const tg_half* bias,
                          tg_half* out, void* stream);

/* out = [SiLU](GroupNorm32([x1|x2])) for channels-last x1 [B,HW,C1], x2 [B,HW,C2] (x2 NULL when C2 == 0).
 * The up-path torch.cat([hidden, skip], 1) (ip_adapter/unet_2d_blocks.py:2210-2212, 2318-2320) is folded in.
 * workspace: >= 16640 + B*16384 floats, ZERO-INITIALISED once (holds self-resetting arrival counters); B <= 256.
 * Replaces nn.GroupNorm(+SiLU) of diffusers ResnetBlock2D (eps 1e-5),
 * Transformer2DModel.norm (ip_adapter/transformer_2d.py:288, eps 1e-6), conv_norm_out
 * (ip_adapter/unet_2d_condition.py:1117-1119). */
int tg_groupnorm(const tg_half* x1, int C1, const tg_half* x2, int C2, int B, int HW, const tg_half* gamma,
                 const tg_half* beta, float eps, int silu, tg_half* out, float* workspace, void* stream);

/* Folded cross-attention (no reference counterpart as an op; replaces, per layer, attn2.to_q + the two softmaxes + the
 * P.V products + attn2.to_out of IPAttnProcessor.__call__, ip_adapter/attention_processor.py:461-553, for levels whose
 * images have a multiple of 256 tokens).  The context is constant along a trajectory, so the caller folds it ONCE into
 *   Kf[b][h*96 + j][c] = scale*log2(e) * sum_dd K[b][j][h*d+dd] * Wq[h*d+dd][c]      (keys: slots 0..n_text-1 text,
 *   Vf[b][c][h*96 + j] =                 sum_dd V[b][j][h*d+dd] * Wo[c][h*d+dd]        80..80+n_ip-1 image tokens, rest 0)
 * and a step needs two GEMMs with a per-image B operand:
 *   tg_cross_scores: P[B*HW, heads*96] = per head { softmax over the text slots | ip_scale[b] * softmax over the image
 *                    slots | 0 elsewhere } of X[B*HW, C] . Kf[b]^T          (softmax in the GEMM epilogue, fp32)
 *   tg_cross_apply : out[B*HW, C] = P . Vf[b]^T + bias + residual
 * P holds the probabilities the reference's get_attention_scores returns (fp16), so attention-map capture is a slice. */
int tg_cross_scores(const tg_half* X, int ldx, const tg_half* Kf, int B, int HW, int C, int heads, int n_text, int n_ip,
                    const float* ip_scale, tg_half* P, int ldp, void* stream);
int tg_cross_apply(const tg_half* P, int ldp, const tg_half* Vf, int B, int HW, int C, int heads, const tg_half* bias,
                   const tg_half* residual, int ldr, tg_half* out, int ldc, void* stream);

/* tg_gemm / tg_conv3x3 keep a split-K workspace (fp32 partials + arrival counters) per device and SLOT.  Launches that
 * may run concurrently on different streams must use different slots (0..3); the slot is a property of the calling
 * host thread and applies to the launches it enqueues afterwards.  Default slot 0.  (No reference counterpart: the
 * reference runs everything on one stream.) */
int tg_gemm_set_scratch_slot(int slot);

/* Programmatic dependent launch (every kernel of the library is launched with programmatic stream serialisation: its
 * prologue overlaps the previous kernel's tail) for the launches the calling host thread enqueues next: 0 = off, 1 = on,
 * -1 = the process default (on; TG_PDL=0 turns it off).  A pre-launched dependent holds the SMs that the previous kernel's
 * early-finishing CTAs free; while a second branch runs on another stream (the ControlNet encoder beside the UNet encoder)
 * those SMs are better given to that branch, so the module-level path switches it off for the UNet encoder of a stage-2
 * step.  (No reference counterpart.) */
int tg_set_pdl(int mode);

/* Measurement hook (no reference counterpart): with a non-NULL device buffer of >= 8*grid int64, every tg_gemm /
 * tg_conv3x3 CTA records {main-loop cycles, producer wait, MMA wait on data, MMA wait on accumulator, epilogue wait,
 * k-blocks, tiles} so a run can tell which role of the pipeline the others wait for.  NULL switches it off. */
int tg_gemm_set_prof(void* dev_buf);
/* Same for tg_attention_tc: with a non-NULL device buffer of >= 8 int64 per softmax warp (4 per CTA, CTAs in grid order
 * (x fastest)), every softmax warp records {wait for S, wait for tcgen05.ld, max (+ lazy rescale), exponentials + P store
 * issue, store wait + arrive, tiles, SM id, end clock}.  NULL switches it off. */
int tg_attn_set_prof(void* dev_buf);
/* Measurement hook: schedule of tg_attention_tc's kernels -- 0 (default): P overlays S, one query tile per CTA; 2: P in its
 * own tensor-memory columns, score tiles issued two tiles ahead; 3: two query tiles per CTA with alternating MUFU turns
 * (d <= 56, N % 256 == 0; else as 2).  Same results within the kernel's tolerance; process-wide. */
int tg_attn_set_variant(int v);
/* Measurement hook: implementation behind tg_attention_cross_ip -- -1 (default) the tcgen05 / TMEM kernel for head dims
 * <= 64 and the mma.sync kernel above (each where it measured faster), 0 mma.sync everywhere, 1 tcgen05 everywhere. */
int tg_attn_cross_set_impl(int mode);

/* GroupNorm without the statistics pass.  tg_gemm_set_stats(stats, rows_per_image) arms the NEXT tg_gemm / tg_conv3x3 /
 * tg_conv3x3_s2 / tg_upsample2x_conv3x3 call of the calling host thread (one shot; NULL disarms): its epilogue also adds, per
 * (image = row / rows_per_image, pair of adjacent output columns), 2^24 x (sum, sum of squares) of the fp16 values it
 * stores to stats[image][column / 2][0..1] (64-bit integers, caller-zeroed; integer atomics, so the totals are
 * independent of the arrival order -- bit-reproducible).  rows_per_image: a multiple of 32 dividing M; not with GEGLU.
 * tg_groupnorm_apply normalises (+SiLU) [x1 | x2] from such totals in one pass: the same function as tg_groupnorm (diffusers
 * ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out; ip_adapter/unet_2d_blocks.py, transformer_2d.py:288,
 * unet_2d_condition.py:1117-1119).  statsN_alt (may be NULL): images [alt_b0, alt_b1) of that input take their totals from
 * statsN_alt[image - alt_b0] -- rows a later in-place GEMM (the ControlNet residual add) rewrote and re-measured. */
int tg_gemm_set_stats(long long* stats, int rows_per_image);
int tg_groupnorm_apply(const tg_half* x1, int C1, const long long* stats1, const long long* stats1_alt, const tg_half* x2, int C2,
                       const long long* stats2, const long long* stats2_alt, int alt_b0, int alt_b1, int B, int HW,
                       const tg_half* gamma, const tg_half* beta, float eps, int silu, tg_half* out, void* stream);

/* Launch-geometry knobs of tg_groupnorm for measurements (values <= 0 keep the current setting; fused: 1 allows the
 * single-kernel cluster path for tensors whose per-image slab fits on chip, 0 forces the two-pass kernels, <0 keeps).
 * No reference counterpart (the reference leaves this to ATen). */
int tg_groupnorm_tune(int stats_slabs, int apply_ctas_per_sm, int fused);

/* row-wise LayerNorm (ip_adapter/attention.py:198,226,238; resampler.py:16,43-44,100; ip_adapter.py:39) */
int tg_layernorm(const tg_half* x, int ldx, const tg_half* gamma, const tg_half* beta, float eps, tg_half* out,
                 int ldo, int M, int C, void* stream);

/* Flash attention: O[b,n,h*d:(h+1)*d] = softmax(scale Q K^T) V per (b,h).  Q/K/V/O are token-major with row
 * pitches ld* (elements) and head h at column offset h*d (pass pointers into a fused QKV buffer).
 * d in {40,64,80,160}; causal != 0 masks keys behind the query (Nq == Nk).  Replaces AttnProcessor.__call__
 * (ip_adapter/attention_processor.py:343-356), PerceiverAttention's softmax(QK^T)V (ip_adapter/resampler.py:65-74) and the
 * attention of the CLIP encoders (transformers CLIPAttention; ip_adapter/ip_adapter.py:78-81, models/pipelines.py:238). */
int tg_attention(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                 int ldo, int B, int heads, int Nq, int Nk, int d, float scale, int causal, void* stream);

/* Same contract on the tcgen05 path (TMA-fed tiles, S/O accumulators in TMEM) for head dims <= 64: Q/K/V must be
 * head-padded to 64 columns (head h at column h*64, columns [d,64) zero -- obtained for free by zero rows in the
 * fused QKV projection weight), Nq == Nk == N with N % 128 == 0.  O is written unpadded ([B*N, heads*d]).
 * d <= 56: 64 columns per head; d <= 120: 128 columns per head; 128 <= d <= 176 (multiple of 16; SD-1.5's d = 160):
 * heads PACKED (head h at column h*d, no padding) -- three 64-column chunks per head. */
int tg_attention_tc(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                    int ldo, int B, int heads, int N, int d, float scale, void* stream);
/* The same kernels on PACKED heads for every head dim (Q/K/V are [B*N, heads*d] views, head h at column h*d, d % 8 == 0,
 * K / V 16-byte aligned): the fused QKV projection then has 3*heads*d output columns instead of 3*heads*64 (d = 40) or
 * 3*heads*128 (d = 80).  The 64-column TMA boxes of a head's last chunk reach into the next head (past the tensor: zero
 * fill); those columns meet zeros of the Q operand in Q K^T and lie beyond the output columns read after P V. */
int tg_attention_tc_packed(const tg_half* Q, const tg_half* K, const tg_half* V, int ldq, int ldk, int ldv, tg_half* O,
                           int ldo, int B, int heads, int N, int d, float scale, void* stream);

/* IP-Adapter decoupled cross attention (IPAttnProcessor.__call__, ip_adapter/attention_processor.py:478-516):
 *   O = softmax(scale Q Kt^T + add_mask) Vt + ip_scale[b] * softmax(scale Q Ki^T) Vi
 * Kt/Vt/Ki/Vi point into a K/V projection buffer with (n_text + n_ip) rows per batch element and row pitch
 * ldkv; text keys are rows [0,n_text), image-token keys rows [n_text, n_text+n_ip).  n_text <= 80, n_ip <= 16
 * (n_ip == 0: plain AttnProcessor cross attention).  ip_scale: device float[B] (NULL = 1.0): the reference's
 * python attribute `IPAttnProcessor.scale`, per sample so characters with different scales can share a batch.
 * add_mask: device fp16 [B, n_text] additive mask or NULL.  probs: optional capture of the text
 * probabilities, fp16 [B - probs_b0, heads, Nq, n_text] (save_attn_to_dict, :532-549; probs_b0 = B/2 is
 * return_cond_ca_only). */
int tg_attention_cross_ip(const tg_half* Q, int ldq, const tg_half* Kt, const tg_half* Vt, const tg_half* Ki,
                          const tg_half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                          const tg_half* add_mask, tg_half* O, int ldo, tg_half* probs, int probs_b0, int B,
                          int heads, int Nq, int d, float scale, void* stream);

/* layout / small ops of the UNet forward (ip_adapter/unet_2d_condition.py:897-902, 999, 1044-1053, 1117-1120;
 * diffusers Upsample2D / Downsample2D) */
int tg_upsample2x(const tg_half* x, int B, int H, int W, int C, tg_half* y, void* stream);
/* pad_lo = 1: symmetric padding 1 (UNet / ControlNet Downsample2D); pad_lo = 0: the VAE encoder's F.pad(x,(0,1,0,1)) */
int tg_im2col_s2(const tg_half* x, int B, int H, int W, int C, tg_half* a, int pad_lo, void* stream);
/* non-overlapping P x P patches of an NCHW image as GEMM rows [B*(H/P)*(W/P), Kpad], column k = c*P*P + ky*P + kx (the
 * CLIP ViT patch embedding, transformers CLIPVisionEmbeddings; ip_adapter/ip_adapter.py:78-81, 142-153) */
int tg_im2col_patch(const tg_half* x_nchw, tg_half* out, int B, int Cin, int H, int W, int P, int Kpad, void* stream);
/* out[b][t][:] = table[ids[b][t]][:] + pos[t][:]: transformers CLIPTextEmbeddings, the first step of
 * `text_encoder(input_ids)[0]` (models/models.py:69-70).  ids are int64 token ids in [0, vocab), device memory. */
int tg_embed_tokens(const long long* ids, const tg_half* table, const tg_half* pos, tg_half* out, int B, int T, int C,
                    int vocab, void* stream);
/* diffusers DiagonalGaussianDistribution.sample on NCHW moments [B,2L,HW]: z = (mean + exp(0.5 clamp(logvar,-30,20)) *
 * noise) * out_scale; noise NULL = the mode (models/pipelines.py:624-626) */
int tg_gaussian_sample(const tg_half* moments, const tg_half* noise, tg_half* z, int B, int L, int HW, float out_scale,
                       void* stream);
/* conv_in: NCHW [B,Cin,H,W] (Cin = 4 latents, or 3 for the ControlNet hint image) -> NHWC [B,H,W,Cout], optional SiLU */
int tg_conv_in(const tg_half* x_nchw, const tg_half* w, const tg_half* bias, tg_half* y_nhwc, int B, int Cin, int H,
               int W, int Cout, int silu, void* stream);
/* 3x3 patches of a small-channel NCHW image as GEMM rows: out[B*H*W, 64], column k = ci*9 + ky*3 + kx (a torch
 * [Cout][Cin][3][3] weight flattened), zero beyond Cin*9.  conv_in (unet_2d_condition.py:999) and the ControlNet hint
 * embedder's first convolution then run as tg_gemm with K = 64. */
int tg_im2col_small(const tg_half* x_nchw, tg_half* out, int B, int Cin, int H, int W, void* stream);
/* conv_out: NHWC [B,H,W,C] -> NCHW [B,Cout,H,W], 3x3 pad 1, w = [Cout,3,3,C]; Cout = 4 (UNet eps,
 * unet_2d_condition.py:1120) or 3 (VAE decoder RGB, models/pipelines.py:468, 162-173) */
int tg_conv_out(const tg_half* x_nhwc, const tg_half* w, const tg_half* bias, tg_half* y_nchw, int B, int H,
                int W, int C, int Cout, void* stream);
/* 1x1 convolution between tiny channel counts on NCHW (diffusers AutoencoderKL.post_quant_conv on the latents) */
int tg_pointwise_nchw(const tg_half* x, const tg_half* w, const tg_half* bias, tg_half* y, int B, int Cin, int Cout,
                      int HW, void* stream);
/* out[r,:] = softmax(scale * x[r,:]) in fp32, one row per CTA, N <= 8192 (the single-head 512-channel attention of the
 * VAE mid block runs as tg_gemm -> tg_softmax_rows -> tg_gemm; diffusers Attention with upcast_softmax) */
int tg_softmax_rows(const tg_half* x, int ldx, tg_half* out, int ldo, int M, int N, float scale, void* stream);
int tg_timestep_embedding(const float* t_dev, tg_half* out, int B, int dim, void* stream);
int tg_add_inplace(tg_half* y, const tg_half* x, long long n, void* stream);
int tg_nchw_to_nhwc(const tg_half* x, tg_half* y, int B, int C, int HW, void* stream);
int tg_duplicate_latents(const tg_half* lat, tg_half* out, int n, void* stream);

/* One launch per denoise step: classifier-free guidance + DDIM update (+ masked latent composite).
 *   eps [2K,C,H,W] (rows [0,K) unconditional), latents [K,C,H,W] in/out, coef_dev = device float[5]
 *   {guidance_scale, sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)}; frozen [K,C,H,W] and mask [H,W]
 *   may be NULL (stage 1); latents_copy (nullable) receives a second copy (the latents_all trace).
 *   pair_stride: rows between a trajectory's unconditional and conditional eps (0 = K).  A call may cover a sub-range
 *   of a larger co-batch: pass eps + k0*C*HW, latents + k0*C*HW, K = count, pair_stride = K_total.
 * Replaces models/pipelines.py:441-447 and :820-834 + diffusers DDIMScheduler.step. */
int tg_cfg_ddim_composite(const tg_half* eps, tg_half* latents, const float* coef_dev, const tg_half* frozen,
                          const tg_half* mask, int K, int C, int HW, tg_half* latents_copy, int pair_stride,
                          void* stream);

/* Cross-attention guidance energy, utils/guidance.py:244-286 (compute_ca_lossv3) with :91-148 (top-k box
 * energy) and :150-242 (reference-attention transfer).  All descriptor arrays are HOST memory and are passed
 * to the kernel by value (graph capturable).
 *   maps[k]   device fp16 [heads, map_hw[k], n_tok]   (cond half of the captured probabilities)
 *   dmaps[k]  nullable device fp16, same shape, caller-zeroed: receives d(grad_scale*energy)/d(map)
 *   boxes     [n_box][4] normalised x0,y0,x1,y1 (python floats); box_obj[n_box] object index of each box
 *   items     n_items x {key, obj, tok, weight}: one top-k term; weight = 1/(n_tok_of_obj * n_obj * n_keys)
 *   ref_items n_ref x {key, box, tok, ref, weight}: one transfer term against ref_maps[ref] [heads, map_hw]
 *   loss_out  device float; workspace >= (n_items + n_ref) floats */
typedef struct {
  unsigned char key, obj, box, ref;
  int tok;
  float weight;
} tg_guidance_item;
int tg_guidance_energy(const tg_half* const* maps, tg_half* const* dmaps, const int* map_hw, int n_keys,
                       int heads, int n_tok, const double* boxes, const int* box_obj, int n_box,
                       const tg_guidance_item* items, int n_items, const tg_guidance_item* ref_items,
                       const tg_half* const* ref_maps, int n_ref_items, float fg_top_p, float bg_top_p,
                       float fg_weight, float bg_weight, float grad_scale, float ref_eps, float* loss_out,
                       float* workspace, void* stream);

/* y [B,Cout,HW] (NCHW planes) = the first Cout (<= 8) channels of x [B,HW,Cs]: the tail of the UNet's conv_out
 * (ip_adapter/unet_2d_condition.py:1120) when it runs as tg_conv3x3 with the 4 output channels padded to 64 */
int tg_nhwc_head_to_nchw(const tg_half* x, tg_half* y, int B, int Cs, int Cout, int HW, void* stream);

/* Tile selection of tg_gemm / tg_conv3x3.  A measured table (tools/gemm_autotune.py -> theatergen_b200/gemm_tuned.json,
 * loaded when the library is opened) overrides the built-in cost model per problem shape: tile width bn in
 * {64,128,160,192,256}, split-K factor 1..8, CTA pairing 0/1; conv != 0 for the implicit-GEMM convolution (K = 9 Cin).
 * tg_gemm_force_tile(bn,...) forces one configuration for every following call (bn = 0: off) -- tuner / tests only. */
int tg_gemm_set_tuned(int M, int N, int K, int act, int conv, int bn, int split, int pair);
int tg_gemm_clear_tuned(void);
int tg_gemm_force_tile(int bn, int split, int pair);

/* ---- input gradients of the guidance step: `torch.autograd.grad(loss, [latents])` of latent_backward_guidance
 * (models/pipelines.py:62-128, :102).  The weights are frozen, so GEMM / convolution backward is tg_gemm / tg_conv3x3 on
 * transposed / rotated weight copies; these entry points are the remaining operator backwards.  fp16 tensors, fp32 math. */
/* LayerNorm: dx = rstd (dy gamma - mean(dy gamma) - xhat mean(dy gamma xhat))  (ip_adapter/attention.py:121,142,158) */
int tg_layernorm_bwd(const tg_half* x, int ldx, const tg_half* gamma, const tg_half* dy, int lddy, tg_half* dx, int lddx,
                     int M, int C, float eps, void* stream);
/* GroupNorm(32) (+SiLU when silu != 0) over channels-last [B,HW,C] (diffusers ResnetBlock2D / Transformer2DModel norms) */
int tg_groupnorm_bwd(const tg_half* x, const tg_half* gamma, const tg_half* beta, const tg_half* dy, tg_half* dx, int B,
                     int HW, int C, float eps, int silu, void* stream);
/* dS = scale P o (dP - rowsum(P o dP)), P = softmax(scale S) [M,N]; ds may alias dp (AttnProcessor's softmax,
 * ip_adapter/attention_processor.py:211-217) */
int tg_softmax_bwd_rows(const tg_half* p, int ldp, const tg_half* dp, int lddp, tg_half* ds, int ldds, int M, int N,
                        float scale, void* stream);
/* GEGLU on the packed pre-activation [M,2F] (per 64 columns: 32 value | 32 gate, bias included), dy [M,F] -> dpre [M,2F] */
int tg_geglu_bwd(const tg_half* pre, const tg_half* dy, tg_half* dpre, int M, int F, void* stream);
/* Query gradient of tg_attention_cross_ip for ONE image (keys / values are constants of the context): dprobs (nullable,
 * [heads,Nq,n_text]) is the energy's gradient on the captured text probabilities (tg_guidance_energy's dmaps) */
int tg_attention_cross_ip_bwd_q(const tg_half* q, int ldq, const tg_half* kt, const tg_half* vt, const tg_half* ki,
                                const tg_half* vi, int ldkv, int n_text, int n_ip, float ip_scale, const tg_half* dout,
                                int lddo, const tg_half* dprobs, tg_half* dq, int lddq, int Nq, int heads, int d,
                                float alpha, void* stream);
/* nearest-x2 upsampling: dx [B,H,W,C] = 2x2 block sums of dy [B,2H,2W,C] */
int tg_upsample2x_bwd(const tg_half* dy, int B, int H, int W, int C, tg_half* dx, void* stream);
/* backward of tg_im2col_s2 (pad_lo = 1): dcols [B*(H/2)*(W/2), 9C] -> dx [B,H,W,C] */
int tg_col2im_s2(const tg_half* dcols, int B, int H, int W, int C, tg_half* dx, void* stream);
/* latents -= (grad * inv_scale) * step in fp32 (models/pipelines.py:106-115); grad_out (nullable, float[n]) receives the
 * unscaled gradient; *bad (nullable, device int) is set when a gradient value is not finite (that element is skipped) */
int tg_guidance_update(tg_half* latents, const tg_half* grad, float inv_scale, float step, float* grad_out, int* bad,
                       int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* THEATERGEN_H_ */
