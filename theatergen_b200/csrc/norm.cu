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
static constexpr int GN_THREADS = 320;  // >= C/8 for C <= 2560
static constexpr int GN_VEC = 4;
static constexpr int GN_MAX_BATCH = 256;

// partial[b][slab][g][2] = (sum, sumsq) over the slab's pixels of group g.
// Thread (r, chunk) owns one 16-byte column chunk (8 channels) and walks rows r, r+R, ... of the slab with four
// independent 16-byte loads in flight; row-lanes are then combined through shared memory in a fixed order and one
// thread per group adds its channels in a fixed order: no atomics -> bit-reproducible statistics.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int slabs,
                float* __restrict__ partial, float* __restrict__ stats, int* __restrict__ counters, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.y, slab = blockIdx.x;
  const int C = C1 + C2, cpg = C / GN_GROUPS, c8n = C / 8;
  const int rows_per = (HW + slabs - 1) / slabs;
  const int r0 = slab * rows_per, r1 = min(HW, r0 + rows_per);
  const int R = GN_THREADS / c8n;  // row lanes (>= 1 because c8n <= GN_THREADS)
  extern __shared__ float s_gn[];  // [R][C] sums, then [R][C] sumsq
  float* s_sum = s_gn;
  float* s_sq = s_gn + R * C;
  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  if (rl < R) {
    const int c = chunk * 8;
    const __half* src;
    int ld;
    if (c < C1) { src = x1 + c; ld = C1; } else { src = x2 + (c - C1); ld = C2; }
    src += (size_t)b * HW * ld;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    int r = r0 + rl;
    for (; r + 3 * R < r1; r += 4 * R) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * R) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          s[2 * j] += f.x; q[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
        }
      }
    }
    for (; r < r1; r += R) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * ld);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_sum[rl * C + c + j] = s[j];
      s_sq[rl * C + c + j] = q[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    float s = 0.f, q = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c)
      for (int k = 0; k < R; ++k) { s += s_sum[k * C + c]; q += s_sq[k * C + c]; }
    float* o = partial + (((size_t)b * slabs + slab) * GN_GROUPS + threadIdx.x) * 2;
    o[0] = s;
    o[1] = q;
  }
  // the last slab-CTA of this image to arrive folds the partials (fixed order) into mean / rstd
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(counters + b, 1);
    s_last = old == slabs - 1;
    if (s_last) counters[b] = 0;  // self-reset for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < GN_GROUPS) {
    float s = 0.f, q = 0.f;
    const float* p = partial + ((size_t)b * slabs * GN_GROUPS + threadIdx.x) * 2;
    for (int i = 0; i < slabs; ++i, p += GN_GROUPS * 2) { s += __ldcg(p); q += __ldcg(p + 1); }
    const float n = (float)HW * cpg;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    stats[((size_t)b * GN_GROUPS + threadIdx.x) * 2] = mean;
    stats[((size_t)b * GN_GROUPS + threadIdx.x) * 2 + 1] = rsqrtf(var + eps);
  }
}

// normalise (+SiLU): thread (r, chunk) owns one 16-byte column chunk; its 8 scale/shift pairs
// (a = rstd*gamma, s = beta - mean*a) live in registers, rows are streamed with four 16-byte loads in flight.
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int slabs,
                const float* __restrict__ stats, const __half* __restrict__ gamma,
                const __half* __restrict__ beta, int silu, __half* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.y, slab = blockIdx.x;
  const int C = C1 + C2, cpg = C / GN_GROUPS, c8n = C / 8;
  const int R = GN_THREADS / c8n;
  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  if (rl >= R) return;
  const int c = chunk * 8;
  float sa[8], sb[8];
  {
    const uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + c));
    const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + c));
    const __half* hg = reinterpret_cast<const __half*>(&gm);
    const __half* hb = reinterpret_cast<const __half*>(&bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      const float mean = stats[((size_t)b * GN_GROUPS + g) * 2], rstd = stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
      sa[j] = rstd * __half2float(hg[j]);
      sb[j] = __half2float(hb[j]) - mean * sa[j];
    }
  }
  const __half* src;
  int ld;
  if (c < C1) { src = x1 + c; ld = C1; } else { src = x2 + (c - C1); ld = C2; }
  src += (size_t)b * HW * ld;
  __half* dst = out + (size_t)b * HW * C + c;
  const int rows_per = (HW + slabs - 1) / slabs;
  const int r0 = slab * rows_per, r1 = min(HW, r0 + rows_per);
  auto emit = [&](const uint4& v, int r) {
    const __half* hv = reinterpret_cast<const __half*>(&v);
    uint4 o;
    __half* ho = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(__half2float(hv[j]), sa[j], sb[j]);
      if (silu) y = silu_f(y);
      ho[j] = __float2half_rn(y);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)r * C) = o;
  };
  int r = r0 + rl;
  for (; r + 3 * R < r1; r += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * R) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(v[u], r + u * R);
  }
  for (; r < r1; r += R) emit(*reinterpret_cast<const uint4*>(src + (size_t)r * ld), r);
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, int silu, __half* out, float* workspace, cudaStream_t st) {
  const int C = C1 + C2;
  TG_CHECK(C % GN_GROUPS == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C / 8 <= GN_THREADS, "groupnorm: C1=%d C2=%d unsupported", C1, C2);
  TG_CHECK(workspace != nullptr, "groupnorm: zero-initialised workspace (16640 + B*4096 floats) required");
  // enough slabs for >= ~4 CTAs per SM, at least 8 rows per row-lane pass
  int slabs = (4 * 148 + B - 1) / B;
  const int R = GN_THREADS / (C / 8);
  slabs = max(1, min(min(slabs, GN_MAX_SLABS), HW / max(1, 2 * R)));
  const size_t smem = (size_t)2 * R * C * sizeof(float);
  static bool attr = false;
  if (!attr) {
    TG_CUDA(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  TG_CHECK(smem <= 64 * 1024, "groupnorm: smem %zu too large", smem);
  // workspace layout (fixed offsets, independent of B): [256 arrival counters][256 x 32 x (mean, rstd)][partials]
  TG_CHECK(B <= GN_MAX_BATCH, "groupnorm: batch %d > %d", B, GN_MAX_BATCH);
  int* counters = reinterpret_cast<int*>(workspace);
  float* stats = workspace + GN_MAX_BATCH;
  float* partial = stats + (size_t)GN_MAX_BATCH * GN_GROUPS * 2;
  TG_CUDA(launch_kernel(gn_stats_kernel, dim3(slabs, B), dim3(GN_THREADS), smem, st, 1, x1, C1, x2, C2, HW, slabs, partial, stats, counters, eps));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  int aslabs = (8 * 148 + B - 1) / B;  // ~8 CTAs per SM in flight, >= 4 rows per row-lane
  aslabs = max(1, min(aslabs, HW / max(1, 4 * R)));
  TG_CUDA(launch_kernel(gn_apply_kernel, dim3(aslabs, B), dim3(GN_THREADS), 0, st, 1, x1, C1, x2, C2, HW, aslabs, stats, gamma, beta, silu, out));
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
  grid_dep_launch();
  grid_dep_wait();
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
    TG_CUDA(launch_kernel(layernorm_kernel<P>, dim3(grid), dim3(warps * 32), 0, st, 1, x, ldx, gamma, beta, eps, out, ldo, M));        \
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
