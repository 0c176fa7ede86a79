// GroupNorm(+SiLU) and LayerNorm for channels-last activations (HBM-bound, vectorised, warp reductions).
//
// GroupNorm: x is the channel-concat of up to two NHWC tensors ([B,HW,C1] | [B,HW,C2]) -- the up-path
// torch.cat([hidden, skip], 1) of the reference (ip_adapter/unet_2d_blocks.py:2210-2212, 2318-2320) is never
// materialised; the normalised (and optionally SiLU-ed) result is written as one [B,HW,C1+C2] tensor that the
// implicit-GEMM conv consumes.  Statistics in fp32 (two-level, deterministic: per-slab partials then a
// per-CTA final reduce).  Replaces nn.GroupNorm + SiLU of diffusers ResnetBlock2D (eps 1e-5),
// Transformer2DModel.norm (eps 1e-6, ip_adapter/transformer_2d.py:288) and conv_norm_out
// (ip_adapter/unet_2d_condition.py:1117-1119).
// LayerNorm: one warp per row (ip_adapter/attention.py:198,226,238; eps 1e-5).
#include "common.cuh"
#include "ops.h"

namespace tg {

static constexpr int GN_GROUPS = 32;
static constexpr int GN_MAX_SLABS = 64;
static constexpr int GN_MAX_C = 2560;

// partial[b][slab][g][2] = (sum, sumsq) over the slab's pixels of group g
__global__ void __launch_bounds__(256)
gn_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int slabs,
                float* __restrict__ partial) {
  const int b = blockIdx.y, slab = blockIdx.x;
  const int C = C1 + C2, cpg = C / GN_GROUPS;
  const int rows_per = (HW + slabs - 1) / slabs;
  const int r0 = slab * rows_per, r1 = min(HW, r0 + rows_per);
  // per-channel partial sums land in shared memory, then one thread per group adds its channels in a fixed
  // order: no atomics, so the statistics (and everything downstream) are bit-reproducible run to run.
  __shared__ float s_csum[GN_MAX_C], s_csq[GN_MAX_C];
  const int c2n = C / 2;  // half2 columns
  for (int cc = threadIdx.x; cc < c2n; cc += blockDim.x) {
    const int c = cc * 2;
    const __half* src;
    int ld, cs;
    if (c < C1) { src = x1; ld = C1; cs = c; } else { src = x2; ld = C2; cs = c - C1; }
    const __half* p = src + ((size_t)b * HW + r0) * ld + cs;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r, p += ld) {
      float2 f = __half22float2(*reinterpret_cast<const __half2*>(p));
      s0 += f.x; q0 += f.x * f.x;
      s1 += f.y; q1 += f.y * f.y;
    }
    s_csum[c] = s0; s_csq[c] = q0;
    s_csum[c + 1] = s1; s_csq[c + 1] = q1;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    float s = 0.f, q = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { s += s_csum[c]; q += s_csq[c]; }
    float* o = partial + (((size_t)b * slabs + slab) * GN_GROUPS + threadIdx.x) * 2;
    o[0] = s;
    o[1] = q;
  }
}

// Coalescing note for gn_stats_kernel: consecutive threads read consecutive half2 of the same pixel row, so a
// warp reads 128 contiguous bytes per row step.

__global__ void __launch_bounds__(256)
gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int slabs,
                const float* __restrict__ partial, const __half* __restrict__ gamma,
                const __half* __restrict__ beta, float eps, int silu, __half* __restrict__ out, int rows_per_cta) {
  const int b = blockIdx.y;
  const int C = C1 + C2, cpg = C / GN_GROUPS;
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  if (threadIdx.x < GN_GROUPS) {
    float s = 0.f, q = 0.f;
    const float* p = partial + ((size_t)b * slabs * GN_GROUPS + threadIdx.x) * 2;
    for (int i = 0; i < slabs; ++i, p += GN_GROUPS * 2) { s += p[0]; q += p[1]; }
    const float n = (float)HW * cpg;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const int c8n = C / 8;  // 16-byte vectors per row (C1, C2 are multiples of 8)
  const int total = (r1 - r0) * c8n;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = r0 + i / c8n, c = (i % c8n) * 8;
    const __half* src = (c < C1) ? x1 + ((size_t)b * HW + r) * C1 + c : x2 + ((size_t)b * HW + r) * C2 + (c - C1);
    uint4 v = *reinterpret_cast<const uint4*>(src);
    uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + c));
    uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + c));
    const __half* hv = reinterpret_cast<const __half*>(&v);
    const __half* hg = reinterpret_cast<const __half*>(&gm);
    const __half* hb = reinterpret_cast<const __half*>(&bt);
    uint4 o;
    __half* ho = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      float y = (__half2float(hv[j]) - s_mean[g]) * s_rstd[g] * __half2float(hg[j]) + __half2float(hb[j]);
      if (silu) y = silu_f(y);
      ho[j] = __float2half_rn(y);
    }
    *reinterpret_cast<uint4*>(out + ((size_t)b * HW + r) * C + c) = o;
  }
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, int silu, __half* out, float* workspace, cudaStream_t st) {
  const int C = C1 + C2;
  TG_CHECK(C % GN_GROUPS == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C <= GN_MAX_C, "groupnorm: C1=%d C2=%d unsupported", C1, C2);
  TG_CHECK(workspace != nullptr, "groupnorm: workspace (B*64*32*2 floats) required");
  int slabs = HW >= 4096 ? 64 : HW >= 1024 ? 32 : HW >= 256 ? 8 : 2;
  if (slabs > GN_MAX_SLABS) slabs = GN_MAX_SLABS;
  gn_stats_kernel<<<dim3(slabs, B), 256, 0, st>>>(x1, C1, x2, C2, HW, slabs, workspace);
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  // ~2 waves of 148 SMs when the tensor is big, never less than 8 rows per CTA
  int ctas_per_img = max(1, min(HW / 8, (2 * 148 + B - 1) / B));
  int rows_per_cta = (HW + ctas_per_img - 1) / ctas_per_img;
  ctas_per_img = (HW + rows_per_cta - 1) / rows_per_cta;
  gn_apply_kernel<<<dim3(ctas_per_img, B), 256, 0, st>>>(x1, C1, x2, C2, HW, slabs, workspace, gamma, beta, eps, silu,
                                                         out, rows_per_cta);
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, values kept in registers (two-pass variance)
// ---------------------------------------------------------------------------------------------------------
template <int PER_LANE>  // half2 per lane: C = 64 * PER_LANE
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, int ldx, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, float eps, __half* __restrict__ out, int ldo, int M) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int C = 64 * PER_LANE;
  const __half2* px = reinterpret_cast<const __half2*>(x + (size_t)row * ldx);
  float2 v[PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    v[i] = __half22float2(px[lane + 32 * i]);
    s += v[i].x + v[i].y;
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean;
    q += a * a + b * b;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  const __half2* pg = reinterpret_cast<const __half2*>(gamma);
  const __half2* pb = reinterpret_cast<const __half2*>(beta);
  __half2* po = reinterpret_cast<__half2*>(out + (size_t)row * ldo);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const float2 g = __half22float2(__ldg(pg + lane + 32 * i));
    const float2 b = __half22float2(__ldg(pb + lane + 32 * i));
    po[lane + 32 * i] = __floats2half2_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
  }
}

int layernorm_rows(const __half* x, int ldx, const __half* gamma, const __half* beta, float eps, __half* out,
                   int ldo, int M, int C, cudaStream_t st) {
  TG_CHECK(C % 64 == 0 && C / 64 <= 24, "layernorm: C=%d unsupported (multiple of 64, <= 1536)", C);
  const int warps = 8;
  dim3 grid((M + warps - 1) / warps);
#define TG_LN(P)                                                                                    \
  case P:                                                                                           \
    layernorm_kernel<P><<<grid, warps * 32, 0, st>>>(x, ldx, gamma, beta, eps, out, ldo, M);        \
    break;
  switch (C / 64) {
    TG_LN(5) TG_LN(10) TG_LN(12) TG_LN(16) TG_LN(20) TG_LN(24)
    default:
      TG_CHECK(false, "layernorm: C=%d not instantiated", C);
  }
#undef TG_LN
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

}  // namespace tg
