// tcgen05 cross-attention with the IP-Adapter image-token branch (reference: IPAttnProcessor.__call__,
// ip_adapter/attention_processor.py:461-516): text keys (<= 80, SD-1.5: 77) and image-token keys (<= 16, IP-Adapter: 4,
// IP-Adapter-Plus: 16) of one (image, head) form ONE 96-row key / value tile,
//     rows  0..79  text  K_t / V_t   (rows >= n_text zero)
//     rows 80..95  image K_i / V_i   (rows >= n_ip zero)
// so a 128-query tile costs one group of S = Q Kcat^T MMAs (M = 128, N = 96) and one group of O = Pcat Vcat MMAs (K = 96
// keys): the two softmaxes of the reference (text and image tokens separately) are taken over columns 0..n_text-1 and
// 80..80+n_ip-1 of the same score row, each normalised in registers, the image part multiplied by ip_scale[b], and the
// concatenated probabilities go back to tensor memory as the A operand of the second product:
//     out = P_t V_t + ip_scale[b] * P_i V_i    in one accumulator (fp32), rounded to fp16 once.
// (The reference rounds P_t V_t and P_i V_i to fp16 separately before the scaled sum; the single rounding is closer to
// exact arithmetic and inside the golden-vector tolerance.)  The additive attention mask and the capture of the text
// probabilities (save_attn_to_dict, :532-551) are applied / written from the same registers.
//
// Per CTA: (image b, head h) and `qtiles` consecutive 128-query tiles; the 96 x d K / V rows (15-61 KB) are staged ONCE by
// the four softmax warps straight into the canonical 128-byte-swizzled UMMA layout (they are head slices of the fused
// context projection, 80-320 B per row: too narrow for a swizzled TMA box, and 15 KB per CTA is not worth one);
// Q rows go to tensor memory (TS-form MMAs: the A operand of both products lives in TMEM, 34-42 cycles per MMA).
//   warps 0..3  staging, softmax (one query row per thread == TMEM lane), epilogue
//   warp 4      TMEM allocator + MMA issuer (warp-uniform loop, one elected lane)
// Tensor memory: S 96 columns | P 48 (96 fp16) | O 64 per head chunk | Q 8 per K-step.
// Head dims 40 / 80 / 160 -> CH = 1 / 2 / 3 chunks of 64 columns; CH = 1 runs two CTAs per SM.
#include "common.cuh"
#include "ops.h"
#include "tmem.cuh"

namespace tg {

static constexpr int XQ = 128;            // queries per tile
static constexpr int XTXT = 80;           // text rows of the key tile (n_text <= 80)
static constexpr int XIP = 16;            // image-token rows (n_ip <= 16)
static constexpr int XKEYS = XTXT + XIP;  // 96
static constexpr int X_SOFTMAX_WARPS = 4;
static constexpr int X_THREADS = (X_SOFTMAX_WARPS + 1) * 32;
static constexpr int X_WARP_MMA = X_SOFTMAX_WARPS;
static constexpr int X_CHUNK_BYTES = XKEYS * 128;  // 96 rows x 128 B: one 64-column chunk of K (or V)

template <int CH>
struct CrossTcCfg {
  static constexpr int KV_BYTES = CH * X_CHUNK_BYTES;
  static constexpr int BAR_OFFSET = 2 * KV_BYTES;
  static constexpr int SMEM = BAR_OFFSET + 4 * 8 + 16 + 1024;
  static constexpr uint32_t TS = 0, TP = 96, TO = 144, TQ = 144 + 64 * CH;
  static constexpr uint32_t TCOLS = (TQ + 8 * 4 * CH <= 256) ? 256 : 512;  // Q: up to 4 K-steps per chunk
};

// byte offset of element (row, col) inside a [rows][64 x CH] tile in the canonical SWIZZLE_128B layout: 64-column chunks
// of 128-byte rows, 16-byte unit u of row r stored at unit u ^ (r & 7)
__device__ __forceinline__ int sw128_offset(int row, int col) {
  const int c = col & 63;
  return (col >> 6) * X_CHUNK_BYTES + row * 128 + ((((c >> 3) ^ (row & 7)) << 4) | ((c & 7) << 1));
}

template <int CH>
__global__ void __launch_bounds__(X_THREADS, CH == 1 ? 2 : 1)
attn_cross_tc_kernel(const __half* __restrict__ Q, int ldq, const __half* __restrict__ Kt, const __half* __restrict__ Vt,
                     const __half* __restrict__ Ki, const __half* __restrict__ Vi, int ldkv, int n_text, int n_ip,
                     const float* __restrict__ ip_scale, const __half* __restrict__ add_mask, __half* __restrict__ O,
                     int ldo, __half* __restrict__ probs, int probs_b0, int Nq, int d, float scale_log2, int qtiles) {
  using Cfg = CrossTcCfg<CH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* q_ready = bars;      // Q rows of the tile are in tensor memory (4 warps)
  uint64_t* s_full = bars + 1;   // S complete
  uint64_t* p_full = bars + 2;   // P stored (4 warps)
  uint64_t* o_full = bars + 3;   // O complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, tile0 = blockIdx.x * qtiles;
  const int ntiles_total = (Nq + XQ - 1) / XQ;
  const int ntiles = min(qtiles, ntiles_total - tile0);
  const int T = n_text + n_ip;           // context rows per image in the K/V projection buffer
  const int ksteps = (d + 15) / 16;      // K extent of Q Kcat^T in MMA steps (columns >= d are zero in Q and K)

  if (threadIdx.x == 0) {
    mbar_init(q_ready, X_SOFTMAX_WARPS);
    mbar_init(s_full, 1);
    mbar_init(p_full, X_SOFTMAX_WARPS);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == X_WARP_MMA) {
    tmem_alloc(tmem_slot, Cfg::TCOLS);
    tmem_relinquish();
  }
  // zero the key / value tile (pad rows and pad columns must be finite: P is 0 there, 0 x NaN would poison O)
  for (int i = threadIdx.x; i < 2 * Cfg::KV_BYTES / 16; i += X_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  grid_dep_launch();
  grid_dep_wait();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tS = tmem_base + Cfg::TS, tP = tmem_base + Cfg::TP, tO = tmem_base + Cfg::TO, tQ = tmem_base + Cfg::TQ;

  if (warp < X_SOFTMAX_WARPS) {
    // ---- stage K / V of this (image, head): 16-byte units, rows 0..n_text-1 -> tile rows 0.., rows n_text.. -> 80.. ----
    const int upr = d >> 3;  // 16-byte units per row (d is a multiple of 8)
    const size_t row0 = (size_t)b * T;
    const int total = T * upr * 2;
    constexpr int NT = X_SOFTMAX_WARPS * 32, UNR = 4;  // four independent 16-byte loads in flight per thread
    for (int i0 = threadIdx.x; i0 < total; i0 += NT * UNR) {
      uint4 v[UNR];
      int dst[UNR];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const int i = i0 + j * NT;
        dst[j] = -1;
        if (i < total) {
          const int which = i / (T * upr), rem = i - which * (T * upr);
          const int r = rem / upr, u = rem - r * upr;
          const bool text = r < n_text;
          const __half* src = which == 0 ? (text ? Kt : Ki) : (text ? Vt : Vi);
          v[j] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ldkv + h * d) + u);
          dst[j] = which * Cfg::KV_BYTES + sw128_offset(text ? r : XTXT + (r - n_text), u * 8);
        }
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j)
        if (dst[j] >= 0) *reinterpret_cast<uint4*>(smem + dst[j]) = v[j];
    }
    fence_proxy_async();  // generic-proxy shared-memory writes -> visible to the tensor core
    asm volatile("bar.sync 1, %0;" ::"n"(X_SOFTMAX_WARPS * 32) : "memory");
  }

  if (warp == X_WARP_MMA) {
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_f16(XQ, XKEYS);
    uint32_t idesc_o[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int nc = (c + 1 < CH) ? 64 : ((d - 64 * c + 15) / 16) * 16;
      idesc_o[c] = make_idesc_f16(XQ, nc) | (1u << 16);  // V is the MN-major B operand
    }
    const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV);
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(q_ready, t & 1);  // also orders the staged K / V (the softmax warps fenced before arriving)
      tc_fence_after();
      if (leader) {
        for (int k = 0; k < ksteps; ++k)
          umma_f16_ts(tS, tQ + k * 8, make_smem_desc_sw128(k_base + (k >> 2) * X_CHUNK_BYTES + (k & 3) * 32), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      mbar_wait(p_full, t & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int k = 0; k < XKEYS / 16; ++k)
            umma_f16_ts(tO + c * 64, tP + k * 8, make_smem_desc_sw128_mn(v_base + c * X_CHUNK_BYTES + k * 2048), idesc_o[c],
                        k > 0 ? 1u : 0u);
        tc_commit(o_full);
      }
    }
  } else {
    const int q = warp;                         // TMEM lane quarter (hardware rule: warp id % 4)
    const int r = q * 32 + lane;                // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float isc = (n_ip > 0 && ip_scale != nullptr) ? ip_scale[b] : 1.f;
    constexpr float LOG2E = 1.4426950408889634f;
    const __half* mrow = add_mask ? add_mask + (size_t)b * n_text : nullptr;
    __half* pout_base = nullptr;
    if (probs != nullptr && b >= probs_b0) pout_base = probs + (((size_t)(b - probs_b0) * gridDim.y + h) * Nq) * n_text;

    auto load_q = [&](int tile, uint4 (&qv)[8 * CH]) {  // this thread's query row, zero beyond d / beyond Nq
      const int row = tile * XQ + r;
      const uint4* qp = reinterpret_cast<const uint4*>(Q + ((size_t)b * Nq + row) * ldq + h * d);
#pragma unroll
      for (int u = 0; u < 8 * CH; ++u) qv[u] = (row < Nq && u * 8 < d) ? __ldg(qp + u) : make_uint4(0, 0, 0, 0);
    };
    uint4 qv[8 * CH];
    if (ntiles > 0) load_q(tile0, qv);

    for (int t = 0; t < ntiles; ++t) {
      const int row = (tile0 + t) * XQ + r;
      // Q row -> tensor memory: 16 fp16 = 8 packed columns per MMA K-step.  The previous tile's S MMAs (the readers of
      // these columns) retired before its P was computed, i.e. long ago.
#pragma unroll
      for (int k = 0; k < 4 * CH; ++k)  // static indices keep the row in registers
        if (k < ksteps) tmem_st_32x32b_x8(tQ + k * 8 + lane_off, reinterpret_cast<const uint32_t*>(&qv[2 * k]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
      if (t + 1 < ntiles) load_q(tile0 + t + 1, qv);  // next tile's row: the loads travel during this tile's math

      mbar_wait(s_full, t & 1);
      tc_fence_after();
      uint32_t s[96];
      tmem_ld_32x32b_x32(tS + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32b_x32(tS + 32 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_32x32b_x32(tS + 64 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      tmem_ld_wait();

      // ---- text softmax over columns [0, n_text) ----
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < XTXT; ++i) {
        float x = __uint_as_float(s[i]) * scale_log2;
        if (mrow != nullptr && i < n_text) x += __half2float(mrow[i]) * LOG2E;
        x = i < n_text ? x : -INFINITY;
        s[i] = __float_as_uint(x);
        mx = fmaxf(mx, x);
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < XTXT; ++i) {
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(__uint_as_float(s[i]) - mx));
        s[i] = __float_as_uint(p);
        sum += p;
      }
      const float inv_t = 1.f / sum;
      // ---- image-token softmax over columns [80, 80 + n_ip) ----
      float mi = -INFINITY;
#pragma unroll
      for (int i = 0; i < XIP; ++i) {
        const float x = i < n_ip ? __uint_as_float(s[XTXT + i]) * scale_log2 : -INFINITY;
        s[XTXT + i] = __float_as_uint(x);
        mi = fmaxf(mi, x);
      }
      float sumi = 0.f;
#pragma unroll
      for (int i = 0; i < XIP; ++i) {
        float p = 0.f;
        if (n_ip > 0) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(__uint_as_float(s[XTXT + i]) - mi));
        s[XTXT + i] = __float_as_uint(p);
        sumi += p;
      }
      const float inv_i = n_ip > 0 ? isc / sumi : 0.f;

      // Pcat row: 96 fp16, two per 32-bit tensor-memory column, packed in place (element i comes from 2i, 2i+1 >= i)
#pragma unroll
      for (int i = 0; i < XTXT / 2; ++i) {
        const __half2 hp = __floats2half2_rn(__uint_as_float(s[2 * i]) * inv_t, __uint_as_float(s[2 * i + 1]) * inv_t);
        s[i] = *reinterpret_cast<const uint32_t*>(&hp);
      }
#pragma unroll
      for (int i = 0; i < XIP / 2; ++i) {
        const __half2 hp = __floats2half2_rn(__uint_as_float(s[XTXT + 2 * i]) * inv_i, __uint_as_float(s[XTXT + 2 * i + 1]) * inv_i);
        s[XTXT / 2 + i] = *reinterpret_cast<const uint32_t*>(&hp);
      }
      // P of the previous tile fed the previous P.V group, which retired before the previous tile's epilogue read O
      tmem_st_32x32b_x32(tP + lane_off, *reinterpret_cast<const uint32_t(*)[32]>(&s[0]));
      tmem_st_32x32b_x16(tP + 32 + lane_off, &s[32]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);

      if (pout_base != nullptr && row < Nq) {  // captured text probabilities (fp16, [.., Nq, n_text])
        __half* po = pout_base + (size_t)row * n_text;
#pragma unroll
        for (int i = 0; i < XTXT / 2; ++i) {  // static register indices; two probabilities per packed word
          const __half2 hp = *reinterpret_cast<const __half2*>(&s[i]);
          if (2 * i < n_text) po[2 * i] = __low2half(hp);
          if (2 * i + 1 < n_text) po[2 * i + 1] = __high2half(hp);
        }
      }

      mbar_wait(o_full, t & 1);
      tc_fence_after();
      __half* orow = O + ((size_t)b * Nq + row) * ldo + h * d;
#pragma unroll
      for (int c0 = 0; c0 < CH * 64; c0 += 32) {
        if (c0 < d) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO + c0 + lane_off, o);
          tmem_ld_wait();
          if (row < Nq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (c0 + 8 * i < d) {  // d is a multiple of 8
                uint4 v;
                __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  hh[k] = __floats2half2_rn(__uint_as_float(o[8 * i + 2 * k]), __uint_as_float(o[8 * i + 2 * k + 1]));
                *reinterpret_cast<uint4*>(orow + c0 + 8 * i) = v;
              }
            }
          }
        }
      }
      tc_fence_before();  // O / S / P reads of this tile are complete before the next tile's MMAs are issued
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == X_WARP_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TCOLS);
  }
}

template <int CH>
static int launch_cross_tc(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki, const __half* Vi,
                           int ldkv, int n_text, int n_ip, const float* ip_scale, const __half* add_mask, __half* O, int ldo,
                           __half* probs, int probs_b0, int B, int heads, int Nq, int d, float scale, cudaStream_t st) {
  using Cfg = CrossTcCfg<CH>;
  auto kern = attn_cross_tc_kernel<CH>;
  TG_SMEM_ATTR_ONCE(kern, Cfg::SMEM);
  // query tiles per CTA: the K / V staging (15-61 KB) is shared by more queries while enough CTAs remain to fill the GPU
  const int tiles = (Nq + XQ - 1) / XQ;
  const int per_sm = CH == 1 ? 2 : 1;
  int qtiles = 1;
  while (qtiles < 8 && ((tiles + 2 * qtiles - 1) / (2 * qtiles)) * heads * B >= 148 * per_sm * 2) qtiles *= 2;
  dim3 grid((tiles + qtiles - 1) / qtiles, heads, B);
  TG_CUDA(launch_kernel(kern, grid, dim3(X_THREADS), Cfg::SMEM, st, 1, Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale,
                        add_mask, O, ldo, probs, probs_b0, Nq, d, scale * 1.4426950408889634f, qtiles));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

int attention_cross_ip_tc(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki, const __half* Vi,
                          int ldkv, int n_text, int n_ip, const float* ip_scale, const __half* add_mask, __half* O, int ldo,
                          __half* probs, int probs_b0, int B, int heads, int Nq, int d, float scale, cudaStream_t st) {
  TG_CHECK(n_text >= 1 && n_text <= XTXT && n_ip >= 0 && n_ip <= XIP, "cross attention: n_text=%d (<=80) n_ip=%d (<=16)", n_text,
           n_ip);
  TG_CHECK(d % 8 == 0 && d >= 8 && d <= 192, "cross attention: head dim %d (multiple of 8, <= 192)", d);
  TG_CHECK(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "cross attention: pitches must be multiples of 8 elements");
  TG_CHECK(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(Kt) | reinterpret_cast<uintptr_t>(Vt) |
             reinterpret_cast<uintptr_t>(O)) & 15) == 0 && (n_ip == 0 || ((reinterpret_cast<uintptr_t>(Ki) |
             reinterpret_cast<uintptr_t>(Vi)) & 15) == 0), "cross attention: pointers must be 16-byte aligned");
  if (d <= 64) return launch_cross_tc<1>(Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo, probs, probs_b0, B, heads, Nq, d, scale, st);
  if (d <= 128) return launch_cross_tc<2>(Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo, probs, probs_b0, B, heads, Nq, d, scale, st);
  return launch_cross_tc<3>(Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo, probs, probs_b0, B, heads, Nq, d, scale, st);
}

}  // namespace tg
