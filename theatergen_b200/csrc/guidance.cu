// Cross-attention guidance energy (warp-reduction kernels; <= 1 MB touched per evaluation).
//
// Restates utils/guidance.py:
//   add_ca_loss_per_attn_map_to_loss (:91-148, max-based branch)   top-k foreground / background box energy
//   add_ref_ca_loss_per_attn_map_to_lossv2 (:150-242)              L1 between masked-normalised maps
//   compute_ca_lossv3 (:244-286)                                    normalisation by (n_obj * n_keys)
// plus the analytic d(energy)/d(attention map) needed by latent_backward_guidance (models/pipelines.py:62-128).
//
// One CTA per (attention key, object, token) item, one warp per head.  The exact top-k SUM is obtained without a
// sort: a 16-step radix select on the (non-negative) fp16 bit patterns finds the k-th largest value v_k, then
// sum = sum(x > v_k) + (k - count(x > v_k)) * v_k.  Deterministic: per-item partials + a fixed-order final sum.
#include "common.cuh"
#include "ops.h"

namespace tg {

__device__ __forceinline__ int warp_isum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// values: fp16 bit patterns of non-negative numbers held 'PER' per lane (0 for unused slots).
// returns (sum of the k largest, bit pattern of the k-th largest)
template <int PER>
__device__ __forceinline__ float warp_topk_sum(const uint32_t (&bits)[PER], int k, uint32_t* kth_bits) {
  uint32_t prefix = 0;
  // find the largest threshold T such that count(x >= T) >= k
#pragma unroll 1
  for (int bit = 15; bit >= 0; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) c += (bits[i] >= cand) ? 1 : 0;
    c = warp_isum(c);
    if (c >= k) prefix = cand;
  }
  float s = 0.f;
  int cg = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (bits[i] > prefix) {
      s += __half2float(__ushort_as_half((unsigned short)bits[i]));
      ++cg;
    }
  }
  s = warp_sum(s);
  cg = warp_isum(cg);
  *kth_bits = prefix;
  return s + (float)(k - cg) * __half2float(__ushort_as_half((unsigned short)prefix));
}

__global__ void __launch_bounds__(256)
guidance_energy_kernel(const GuidanceParams P, float* __restrict__ partial) {
  grid_dep_launch();
  grid_dep_wait();
  const GuidanceItem it = P.items[blockIdx.x];
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = P.map_hw[it.key];
  const int side = (int)(sqrtf((float)HW) + 0.5f);
  __shared__ unsigned char s_mask[GUIDANCE_MAX_HW];
  __shared__ float s_part[8];
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int y = p / side, x = p % side;
    unsigned char m = 0;
    for (int bi = 0; bi < P.n_box; ++bi) {
      if (P.box_obj[bi] != it.obj) continue;
      // utils.scale_proportion (utils/utils.py:55-68); python round() == round-half-even == rint
      const double* bx = P.boxes[bi];
      int x0 = (int)rint(bx[0] * side), y0 = (int)rint(bx[1] * side);
      const int bw = (int)rint((bx[2] - bx[0]) * side), bh = (int)rint((bx[3] - bx[1]) * side);
      int x1 = x0 + bw, y1 = y0 + bh;
      x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, side); y1 = min(y1, side);
      if (x >= x0 && x < x1 && y >= y0 && y < y1) m = 1;
    }
    s_mask[p] = m;
  }
  __syncthreads();
  int nfg = 0;
  for (int p = lane; p < HW; p += 32) nfg += s_mask[p];
  nfg = warp_isum(nfg);
  // k = (mask.sum() * top_p).long().clamp_(min=1)   (fp32 product, truncation)  utils/guidance.py:133-134
  const int k_fg = max(1, (int)((float)nfg * P.fg_top_p));
  const int k_bg = max(1, (int)((float)(HW - nfg) * P.bg_top_p));

  float contrib = 0.f;
  if (head < P.heads) {
    const __half* map = P.maps[it.key] + ((size_t)head * HW) * P.n_tok + it.tok;
    constexpr int PER = GUIDANCE_MAX_HW / 32;
    uint32_t fg[PER], bg[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int p = lane + 32 * i;
      uint32_t v = 0;
      unsigned char m = 0;
      if (p < HW) {
        v = __half_as_ushort(map[(size_t)p * P.n_tok]);
        m = s_mask[p];
      }
      fg[i] = m ? v : 0u;
      bg[i] = (p < HW && !m) ? v : 0u;
    }
    uint32_t kth_fg, kth_bg;
    const float sfg = warp_topk_sum<PER>(fg, k_fg, &kth_fg);
    const float sbg = warp_topk_sum<PER>(bg, k_bg, &kth_bg);
    contrib = (1.f - sfg / (float)k_fg) * P.fg_weight + (sbg / (float)k_bg) * P.bg_weight;

    if (P.dmaps[it.key] != nullptr) {
      // d(energy)/dA: -fg_w/k_fg on the top-k foreground pixels, +bg_w/k_bg on the top-k background pixels.
      // Ties at the k-th value: the first (k - count_greater) tied pixels in pixel order take the gradient
      // (torch.topk picks an arbitrary tied subset; the energy itself is tie-invariant).
      __half* dmap = P.dmaps[it.key] + ((size_t)head * HW) * P.n_tok + it.tok;
      const float w = it.weight * P.grad_scale;
      int cg_fg = 0, cg_bg = 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) { cg_fg += fg[i] > kth_fg; cg_bg += bg[i] > kth_bg; }
      int ties_fg = k_fg - warp_isum(cg_fg), ties_bg = k_bg - warp_isum(cg_bg);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int p = lane + 32 * i;
        const bool valid = p < HW;
        const unsigned char m = valid ? s_mask[p] : 0;
        const bool tie_f = valid && m && fg[i] == kth_fg, tie_b = valid && !m && bg[i] == kth_bg;
        const unsigned bf = __ballot_sync(0xffffffffu, tie_f), bb = __ballot_sync(0xffffffffu, tie_b);
        const int rank_f = __popc(bf & ((1u << lane) - 1)), rank_b = __popc(bb & ((1u << lane) - 1));
        float g = 0.f;
        if (valid && m && (fg[i] > kth_fg || (tie_f && rank_f < ties_fg))) g = -P.fg_weight / (float)k_fg * w;
        if (valid && !m && (bg[i] > kth_bg || (tie_b && rank_b < ties_bg))) g = P.bg_weight / (float)k_bg * w;
        ties_fg -= __popc(bf);
        ties_bg -= __popc(bb);
        if (valid && g != 0.f) dmap[(size_t)p * P.n_tok] = __float2half_rn(__half2float(dmap[(size_t)p * P.n_tok]) + g);
      }
    }
  }
  if (lane == 0) s_part[head] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int h = 0; h < P.heads; ++h) s += s_part[h];
    partial[blockIdx.x] = s * it.weight;
  }
}

// reference-attention transfer term: mean_heads sum_px | A m / (sum A m + eps) - R m / (sum R m + eps) |
__global__ void __launch_bounds__(256)
guidance_ref_kernel(const GuidanceParams P, float* __restrict__ partial) {
  grid_dep_launch();
  grid_dep_wait();
  const GuidanceItem it = P.ref_items[blockIdx.x];
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = P.map_hw[it.key];
  const int side = (int)(sqrtf((float)HW) + 0.5f);
  __shared__ float s_part[8];
  float contrib = 0.f;
  if (head < P.heads) {
    const double* bx = P.boxes[it.box];
    int x0 = (int)rint(bx[0] * side), y0 = (int)rint(bx[1] * side);
    const int bw = (int)rint((bx[2] - bx[0]) * side), bh = (int)rint((bx[3] - bx[1]) * side);
    int x1 = min(x0 + bw, side), y1 = min(y0 + bh, side);
    x0 = max(x0, 0); y0 = max(y0, 0);
    const __half* map = P.maps[it.key] + ((size_t)head * HW) * P.n_tok + it.tok;
    const __half* ref = P.ref_maps[it.ref] + (size_t)head * HW;  // [heads, HW] (single token)
    float sa = 0.f, sr = 0.f;
    for (int p = lane; p < HW; p += 32) {
      const int y = p / side, x = p % side;
      if (x >= x0 && x < x1 && y >= y0 && y < y1) {
        sa += __half2float(map[(size_t)p * P.n_tok]);
        sr += __half2float(ref[p]);
      }
    }
    sa = warp_sum(sa) + P.ref_eps;
    sr = warp_sum(sr) + P.ref_eps;
    float l1 = 0.f;
    for (int p = lane; p < HW; p += 32) {
      const int y = p / side, x = p % side;
      if (x >= x0 && x < x1 && y >= y0 && y < y1)
        l1 += fabsf(__half2float(map[(size_t)p * P.n_tok]) / sa - __half2float(ref[p]) / sr);
    }
    contrib = warp_sum(l1);
  }
  if (lane == 0) s_part[head] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int h = 0; h < P.heads; ++h) s += s_part[h];
    partial[P.n_items + blockIdx.x] = s / (float)P.heads * it.weight;
  }
}

__global__ void guidance_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ loss) {
  grid_dep_launch();
  grid_dep_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partial[i];
    *loss = s;
  }
}

int guidance_energy(const GuidanceParams& P, float* loss_out, float* workspace, cudaStream_t st) {
  TG_CHECK(P.heads <= 8 && P.n_keys <= GUIDANCE_MAX_KEYS && P.n_box <= GUIDANCE_MAX_BOXES &&
               P.n_items <= GUIDANCE_MAX_ITEMS && P.n_ref_items <= GUIDANCE_MAX_REF_ITEMS,
           "guidance: too many keys/boxes/items");
  for (int k = 0; k < P.n_keys; ++k) TG_CHECK(P.map_hw[k] <= GUIDANCE_MAX_HW, "guidance: map HW %d > %d", P.map_hw[k], GUIDANCE_MAX_HW);
  if (P.n_items > 0) {
    TG_CUDA(launch_kernel(guidance_energy_kernel, dim3(P.n_items), dim3(256), 0, st, 1, P, workspace));
    TG_COUNT_LAUNCH();
    TG_LAUNCH_CHECK();
  }
  if (P.n_ref_items > 0) {
    TG_CUDA(launch_kernel(guidance_ref_kernel, dim3(P.n_ref_items), dim3(256), 0, st, 1, P, workspace));
    TG_COUNT_LAUNCH();
    TG_LAUNCH_CHECK();
  }
  TG_CUDA(launch_kernel(guidance_final_kernel, dim3(1), dim3(32), 0, st, 1, workspace, P.n_items + P.n_ref_items, loss_out));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

}  // namespace tg
