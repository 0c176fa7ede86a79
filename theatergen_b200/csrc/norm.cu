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
static constexpr int GN_MAX_SLABS = 256;
static constexpr int GN_THREADS = 320;  // two-pass kernels: >= C/8 for C <= 2560
static constexpr int GN_FUSED_THREADS = 512;
static constexpr int GN_FUSED_MAX_CLUSTER = 16;
static constexpr int GN_FUSED_MAX_SMEM = 220 * 1024;   // slab + reduction scratch of one CTA (+ 4.7 KB static)
static constexpr int GN_MAX_BATCH = 256;

// tunables (tools/gn_bench.py sweeps them through tg_groupnorm_tune): slabs per image of the statistics kernel (a
// function of the tensor shape only -- never of the batch -- so an image's statistics are summed in the same order
// whatever batch it travels in), CTAs per SM of the apply kernel, and whether the single-kernel cluster path may be used
static int g_gn_stats_slabs = 32, g_gn_apply_per_sm = 8, g_gn_fused = 1;  // measured best (tools/gn_bench.py, B=10)
static bool g_gn_cluster16_ok = false;
void groupnorm_tune(int stats_slabs, int apply_per_sm, int fused) {
  if (stats_slabs > 0) g_gn_stats_slabs = stats_slabs < GN_MAX_SLABS ? stats_slabs : GN_MAX_SLABS;
  if (apply_per_sm > 0) g_gn_apply_per_sm = apply_per_sm;
  if (fused >= 0) g_gn_fused = fused;
}

__device__ __forceinline__ void gn_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void gn_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
  uint32_t a;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}

__device__ __forceinline__ void gn_acc8(const uint4& v, float (&s)[8], float (&q)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    s[2 * j] += f.x; q[2 * j] += f.x * f.x;
    s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
  }
}

// Thread (rl, chunk) owns one 16-byte column chunk (8 channels) and walks rows r0+rl, r0+rl+R, ... with four
// independent 16-byte loads in flight.  STAGE: every loaded vector is also parked in shared memory at the thread's own
// slot (read back by the same thread in the apply pass -> no synchronisation, conflict-free).
template <bool STAGE>
__device__ __forceinline__ void gn_accumulate(const __half* __restrict__ src, int ld, int r0, int r1, int R, int rl,
                                              float (&s)[8], float (&q)[8], uint4* stage, int c8n, int chunk) {
  int r = r0 + rl;
  for (; r + 3 * R < r1; r += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * R) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      gn_acc8(v[u], s, q);
      if (STAGE) stage[(size_t)(r + u * R - r0) * c8n + chunk] = v[u];
    }
  }
  for (; r < r1; r += R) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * ld);
    gn_acc8(v, s, q);
    if (STAGE) stage[(size_t)(r - r0) * c8n + chunk] = v;
  }
}

// CTA-wide reduction of the per-thread channel sums to the 32 groups, in a FIXED order (no atomics -> bit-reproducible):
// row lanes are folded per channel by all threads, then 64 threads (group, sum|sumsq) add their cpg channels.
// s_lane: [2][R][C] floats.  Result in s_grp[64] = (sum, sumsq) per group.
__device__ __forceinline__ void gn_reduce_groups(const float (&s)[8], const float (&q)[8], bool active, int C, int cpg,
                                                 int R, int rl, int c, float* s_lane, float* s_grp) {
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_lane[rl * C + c + j] = s[j];
      s_lane[(R + rl) * C + c + j] = q[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i >= C, ch = i - which * C;
    float* p = s_lane + (size_t)which * R * C + ch;
    float a = p[0];
    for (int k = 1; k < R; ++k) a += p[k * C];
    p[0] = a;
  }
  __syncthreads();
  if (threadIdx.x < 2 * GN_GROUPS) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    const float* p = s_lane + (size_t)which * R * C + g * cpg;
    float a = 0.f;
    for (int k = 0; k < cpg; ++k) a += p[k];
    s_grp[threadIdx.x] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ void gn_scale_shift(const float* __restrict__ stat /*[32][2] mean,rstd*/,
                                               const __half* __restrict__ gamma, const __half* __restrict__ beta, int c,
                                               int cpg, float (&sa)[8], float (&sb)[8]) {
  const uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + c));
  const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + c));
  const __half* hg = reinterpret_cast<const __half*>(&gm);
  const __half* hb = reinterpret_cast<const __half*>(&bt);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    const float mean = stat[2 * g], rstd = stat[2 * g + 1];
    sa[j] = rstd * __half2float(hg[j]);
    sb[j] = __half2float(hb[j]) - mean * sa[j];
  }
}

__device__ __forceinline__ uint4 gn_emit(const uint4& v, const float (&sa)[8], const float (&sb)[8], int silu) {
  const __half* hv = reinterpret_cast<const __half*>(&v);
  uint4 o;
  __half* ho = reinterpret_cast<__half*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = fmaf(__half2float(hv[j]), sa[j], sb[j]);
    if (silu) y = silu_f(y);
    ho[j] = __float2half_rn(y);
  }
  return o;
}

// ---------------------------------------------------------------------------------------------------------
// single-kernel path: one cluster of CS (2..16) CTAs per image.  One warp streams the CTA's slab of rows ONCE from global
// memory into shared memory with bulk copies (cp.async.bulk: a row per copy, or the whole slab in one copy when it is
// contiguous -- far more bytes in flight than register-staged loads can keep), all threads accumulate the statistics
// out of shared memory, the CTAs exchange their 32 x (sum, sumsq) through distributed shared memory (one parallel
// gather), and the slab is normalised (+SiLU) straight out of shared memory: one HBM read, one HBM write, one launch.
// Possible whenever a slab (HW/CS rows x C) fits on chip next to the reduction scratch.  The cluster size is a
// function of the tensor shape only, so an image's result does not depend on the batch it travels in.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(GN_FUSED_THREADS, 1)
gn_fused_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps, int silu,
                __half* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  __shared__ float s_grp[2 * GN_GROUPS];
  __shared__ float s_gather[GN_FUSED_MAX_CLUSTER * 2 * GN_GROUPS];
  __shared__ float s_stat[2 * GN_GROUPS];
  __shared__ __align__(8) uint64_t s_bar;
  const int b = blockIdx.y, rank = blockIdx.x, CS = gridDim.x;  // gridDim.x == cluster size
  const int C = C1 + C2, cpg = C / GN_GROUPS, c8n = C / 8;
  const int rows_per = HW / CS;
  const int r0 = rank * rows_per;
  const int R = GN_FUSED_THREADS / c8n;
  uint4* stage = reinterpret_cast<uint4*>(s_raw);  // [rows_per][C] fp16
  float* s_lane = reinterpret_cast<float*>(s_raw + (size_t)rows_per * C * 2);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  grid_dep_launch();
  grid_dep_wait();
  if (threadIdx.x < 32) {  // one warp issues the bulk copies
    const uint32_t slab_bytes = (uint32_t)rows_per * C * 2;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar, slab_bytes);
    __syncwarp();
    if (C2 == 0) {  // contiguous slab: a few large copies
      const uint8_t* src = reinterpret_cast<const uint8_t*>(x1 + ((size_t)b * HW + r0) * C);
      const uint32_t piece = ((slab_bytes / 32) + 15) & ~15u;
      const uint32_t off = threadIdx.x * piece;
      if (off < slab_bytes) gn_bulk_load(s_raw + off, src + off, min(piece, slab_bytes - off), &s_bar);
    } else {        // channel-concat of two tensors: one copy per row and source
      for (int r = threadIdx.x; r < rows_per; r += 32) {
        uint8_t* dst = s_raw + (size_t)r * C * 2;
        gn_bulk_load(dst, x1 + ((size_t)b * HW + r0 + r) * C1, C1 * 2, &s_bar);
        gn_bulk_load(dst + C1 * 2, x2 + ((size_t)b * HW + r0 + r) * C2, C2 * 2, &s_bar);
      }
    }
  }
  mbar_wait(&s_bar, 0);

  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  const bool active = rl < R;
  const int c = chunk * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active)
    for (int r = rl; r < rows_per; r += R) gn_acc8(stage[(size_t)r * c8n + chunk], s, q);
  gn_reduce_groups(s, q, active, C, cpg, R, rl, c, s_lane, s_grp);
  gn_cluster_arrive();
  gn_cluster_wait();  // every CTA's s_grp is complete and visible
  for (int i = threadIdx.x; i < CS * 2 * GN_GROUPS; i += GN_FUSED_THREADS)  // one parallel gather of all ranks' partials
    s_gather[i] = ld_dsmem_f32(&s_grp[i & (2 * GN_GROUPS - 1)], (uint32_t)(i / (2 * GN_GROUPS)));
  gn_cluster_arrive();  // my remote reads are issued and complete (peers may exit once everybody has arrived)
  __syncthreads();
  if (threadIdx.x < 2 * GN_GROUPS) {
    float a = 0.f;
    for (int k = 0; k < CS; ++k) a += s_gather[k * 2 * GN_GROUPS + threadIdx.x];  // fixed order
    const float other = __shfl_xor_sync(0xffffffffu, a, 1);  // lanes (2g, 2g+1) = (sum, sumsq)
    const float n = (float)HW * cpg;
    const float sum = (threadIdx.x & 1) ? other : a, sq = (threadIdx.x & 1) ? a : other;
    const float mean = sum / n;
    const float var = fmaxf(sq / n - mean * mean, 0.f);
    s_stat[threadIdx.x] = (threadIdx.x & 1) ? rsqrtf(var + eps) : mean;
  }
  __syncthreads();
  if (active) {
    float sa[8], sb[8];
    gn_scale_shift(s_stat, gamma, beta, c, cpg, sa, sb);
    __half* dst = out + ((size_t)b * HW + r0) * C + c;
    for (int r = rl; r < rows_per; r += R)
      *reinterpret_cast<uint4*>(dst + (size_t)r * C) = gn_emit(stage[(size_t)r * c8n + chunk], sa, sb, silu);
  }
  gn_cluster_wait();
}

// ---------------------------------------------------------------------------------------------------------
// two-pass path (tensors too large to park on chip): statistics, then normalise (second read served by L2).
// partial[b][slab][64] = (sum, sumsq) per group over the slab's pixels; the last slab-CTA of an image to arrive folds
// the partials in a fixed order into mean / rstd.
// ---------------------------------------------------------------------------------------------------------
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
  extern __shared__ float s_gn[];  // [2][R][C]
  __shared__ float s_grp[2 * GN_GROUPS];
  __shared__ int s_last;
  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  const bool active = rl < R;
  const int c = chunk * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    const __half* src;
    int ld;
    if (c < C1) { src = x1 + c; ld = C1; } else { src = x2 + (c - C1); ld = C2; }
    src += (size_t)b * HW * ld;
    gn_accumulate<false>(src, ld, r0, r1, R, rl, s, q, nullptr, c8n, chunk);
  }
  gn_reduce_groups(s, q, active, C, cpg, R, rl, c, s_gn, s_grp);
  if (threadIdx.x < 2 * GN_GROUPS)
    partial[((size_t)b * slabs + slab) * 2 * GN_GROUPS + threadIdx.x] = s_grp[threadIdx.x];
  // the last slab-CTA of this image to arrive folds the partials (fixed order) into mean / rstd
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
  // 256 threads: value v = (group, sum|sumsq), four interleaved slab lanes with independent loads in flight, folded
  // (l0 + l1) + (l2 + l3)
  if (threadIdx.x < 8 * GN_GROUPS) {
    const int v = threadIdx.x >> 2, part = threadIdx.x & 3;
    const float* p = partial + (size_t)b * slabs * 2 * GN_GROUPS + v;
    float a = 0.f;
    for (int i = part; i < slabs; i += 4) a += __ldcg(p + (size_t)i * 2 * GN_GROUPS);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (part == 0) s_grp[v] = a;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const float n = (float)HW * cpg;
    const float mean = s_grp[2 * threadIdx.x] / n;
    const float var = fmaxf(s_grp[2 * threadIdx.x + 1] / n - mean * mean, 0.f);
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
  gn_scale_shift(stats + (size_t)b * GN_GROUPS * 2, gamma, beta, c, cpg, sa, sb);
  const __half* src;
  int ld;
  if (c < C1) { src = x1 + c; ld = C1; } else { src = x2 + (c - C1); ld = C2; }
  src += (size_t)b * HW * ld;
  __half* dst = out + (size_t)b * HW * C + c;
  const int rows_per = (HW + slabs - 1) / slabs;
  const int r0 = slab * rows_per, r1 = min(HW, r0 + rows_per);
  int r = r0 + rl;
  for (; r + 3 * R < r1; r += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * R) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(dst + (size_t)(r + u * R) * C) = gn_emit(v[u], sa, sb, silu);
  }
  for (; r < r1; r += R)
    *reinterpret_cast<uint4*>(dst + (size_t)r * C) = gn_emit(*reinterpret_cast<const uint4*>(src + (size_t)r * ld), sa, sb, silu);
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm whose statistics came out of the producers' epilogues (gemm.cu::chunk_stats): per (image, channel pair) fixed-
// point (sum, sumsq) totals.  Each CTA folds them to the 32 (mean, rstd) of its image (exact integer sums, moments in
// double), then streams its slab like gn_apply_kernel: the activation is read ONCE and written once -- the statistics
// pass over the tensor is gone.  An input may take the rows [b0, b1) of its statistics from a second block (`alt`, indexed
// from b0): the rows the stage-2 ControlNet residuals were added to afterwards, by a GEMM that produced fresh totals.
// ---------------------------------------------------------------------------------------------------------
struct GnStatsSrc {
  const long long* main;  // [B][C/2][2]
  const long long* alt;   // [b1 - b0][C/2][2] or null
  int b0, b1;
};
__device__ __forceinline__ const long long* gn_src_of(const GnStatsSrc& s, int b, int C) {
  return (s.alt != nullptr && b >= s.b0 && b < s.b1) ? s.alt + (size_t)(b - s.b0) * C : s.main + (size_t)b * C;
}

__global__ void __launch_bounds__(GN_THREADS)
gn_apply_pre_kernel(const __half* __restrict__ x1, int C1, GnStatsSrc s1, const __half* __restrict__ x2, int C2, GnStatsSrc s2,
                    int HW, int slabs, const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps, int silu,
                    __half* __restrict__ out) {
  __shared__ float s_stat[2 * GN_GROUPS];
  __shared__ long long s_tot[2 * (GN_THREADS * 8 / 2)];  // (sum, sumsq) per channel pair, C <= 2560
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.y, slab = blockIdx.x;
  const int C = C1 + C2, cpg = C / GN_GROUPS, c8n = C / 8;
  {
    // one L2 round trip: every pair's totals are fetched in parallel, then 64 threads (group, sum | sumsq) fold their pairs
    const long long* p1 = gn_src_of(s1, b, C1);
    const long long* p2 = C2 ? gn_src_of(s2, b, C2) : nullptr;
    for (int k = threadIdx.x; k < C / 2; k += GN_THREADS) {
      const int c = 2 * k;
      const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(c < C1 ? p1 + c : p2 + (c - C1)));
      s_tot[2 * k] = v.x;
      s_tot[2 * k + 1] = v.y;
    }
    __syncthreads();
    if (threadIdx.x < 2 * GN_GROUPS) {
      const int g = threadIdx.x >> 1, which = threadIdx.x & 1, ppg = cpg >> 1;
      long long a = 0;
      for (int k = 0; k < ppg; ++k) a += s_tot[2 * (g * ppg + k) + which];
      const double n = (double)HW * cpg;
      const double m1 = (double)a * (1.0 / 16777216.0) / n;  // mean (which = 0) or E[x^2] (which = 1)
      const double other = __shfl_xor_sync(0xffffffffu, m1, 1);
      if (which == 0) {
        s_stat[2 * g] = (float)m1;
        s_stat[2 * g + 1] = rsqrtf((float)fmax(other - m1 * m1, 0.0) + eps);
      }
    }
  }
  __syncthreads();
  const int R = GN_THREADS / c8n;
  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  if (rl >= R) return;
  const int c = chunk * 8;
  float sa[8], sb[8];
  gn_scale_shift(s_stat, gamma, beta, c, cpg, sa, sb);
  const __half* src;
  int ld;
  if (c < C1) { src = x1 + c; ld = C1; } else { src = x2 + (c - C1); ld = C2; }
  src += (size_t)b * HW * ld;
  __half* dst = out + (size_t)b * HW * C + c;
  const int rows_per = (HW + slabs - 1) / slabs;
  const int r0 = slab * rows_per, r1 = min(HW, r0 + rows_per);
  int r = r0 + rl;
  for (; r + 3 * R < r1; r += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(r + u * R) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(dst + (size_t)(r + u * R) * C) = gn_emit(v[u], sa, sb, silu);
  }
  for (; r < r1; r += R)
    *reinterpret_cast<uint4*>(dst + (size_t)r * C) = gn_emit(*reinterpret_cast<const uint4*>(src + (size_t)r * ld), sa, sb, silu);
}

int groupnorm_apply_nhwc(const __half* x1, int C1, const long long* st1, const long long* st1_alt, const __half* x2, int C2,
                         const long long* st2, const long long* st2_alt, int alt_b0, int alt_b1, int B, int HW,
                         const __half* gamma, const __half* beta, float eps, int silu, __half* out, cudaStream_t st) {
  const int C = C1 + C2;
  TG_CHECK(C % (2 * GN_GROUPS) == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C / 8 <= GN_THREADS, "groupnorm_apply: C1=%d C2=%d unsupported", C1, C2);
  TG_CHECK(st1 != nullptr && (C2 == 0 || st2 != nullptr), "groupnorm_apply: statistics missing");
  const int R = GN_THREADS / (C / 8);
  static int per_sm = -1;  // CTAs per SM the grid is sized for (TG_GN_PRE_PER_SM: tools/gn_stats_bench.py sweeps it)
  if (per_sm < 0) {
    const char* e = getenv("TG_GN_PRE_PER_SM");
    per_sm = e ? atoi(e) : 3;  // measured best at every level (B = 10)
    if (per_sm < 1) per_sm = 1;
  }
  int aslabs = (per_sm * 148) / B;
  aslabs = max(1, min(aslabs, HW / max(1, 4 * R)));
  GnStatsSrc s1{st1, st1_alt, alt_b0, alt_b1}, s2{st2, st2_alt, alt_b0, alt_b1};
  TG_CUDA(launch_kernel(gn_apply_pre_kernel, dim3(aslabs, B), dim3(GN_THREADS), 0, st, 1, x1, C1, s1, x2, C2, s2, HW, aslabs, gamma,
                        beta, eps, silu, out));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, int silu, __half* out, float* workspace, cudaStream_t st) {
  const int C = C1 + C2;
  TG_CHECK(C % GN_GROUPS == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C / 8 <= GN_THREADS, "groupnorm: C1=%d C2=%d unsupported", C1, C2);
  TG_CHECK(B <= GN_MAX_BATCH, "groupnorm: batch %d > %d", B, GN_MAX_BATCH);
  static bool attr_dev[TG_MAX_DEVICES] = {};
  int cur_dev = 0;
  TG_CUDA(cudaGetDevice(&cur_dev));
  TG_CHECK(cur_dev >= 0 && cur_dev < TG_MAX_DEVICES, "device index %d out of range", cur_dev);
  bool& attr = attr_dev[cur_dev];
  if (!attr) {
    TG_CUDA(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    TG_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_FUSED_MAX_SMEM));
    // clusters of 16 are beyond the portable size: opt in and ask the occupancy calculator whether one fits at all
    g_gn_cluster16_ok = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (g_gn_cluster16_ok) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(16, 1);
      cfg.blockDim = dim3(GN_FUSED_THREADS);
      cfg.dynamicSmemBytes = GN_FUSED_MAX_SMEM - 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      g_gn_cluster16_ok = cudaOccupancyMaxActiveClusters(&nclusters, gn_fused_kernel, &cfg) == cudaSuccess && nclusters > 0;
    }
    (void)cudaGetLastError();
    attr = true;
  }
  // ---- single-kernel cluster path ----
  // g_gn_fused: 0 = never (two-pass kernels); otherwise wherever a slab fits on chip with a cluster of 8 or 16 CTAs per
  // image -- measured faster than the two-pass kernels at every such shape (tools/gn_bench.py, B = 10: 25.5 vs 27 us at
  // 64x64x320, 11.7 vs 15 at 32x32x640, 9.2 vs 14 at 16x16x1280, 7 vs 12 at 8x8) once SiLU costs one MUFU operation.
  if (g_gn_fused) {
    const int Rf = GN_FUSED_THREADS / (C / 8);
    const size_t scratch = (size_t)2 * Rf * C * sizeof(float);
    int cs = 0;
    for (int cand = 8; cand <= GN_FUSED_MAX_CLUSTER; cand *= 2) {  // >= 8 CTAs per image for parallelism
      if (HW % cand) break;
      if ((size_t)(HW / cand) * C * 2 + scratch <= GN_FUSED_MAX_SMEM - 1024) { cs = cand; break; }
    }
    if (cs == 16 && !g_gn_cluster16_ok) cs = 0;
    if (cs) {
      const size_t smem = (size_t)(HW / cs) * C * 2 + scratch;
      TG_CUDA(launch_kernel(gn_fused_kernel, dim3(cs, B), dim3(GN_FUSED_THREADS), smem, st, cs, x1, C1, x2, C2, HW, gamma, beta,
                            eps, silu, out));
      TG_COUNT_LAUNCH();
      TG_LAUNCH_CHECK();
      return 0;
    }
  }
  // ---- two-pass path ----
  TG_CHECK(workspace != nullptr, "groupnorm: zero-initialised workspace (16640 + B*16384 floats) required");
  const int R = GN_THREADS / (C / 8);
  // slabs per image: a function of the tensor shape only; the VAE decoder's 256x256 / 512x512 tensors (HW = 64 K / 256 K
  // at batch 1) get up to 256
  const int want = max(g_gn_stats_slabs, HW / 512);
  const int slabs = max(1, min(min(want, GN_MAX_SLABS), HW / max(1, 2 * R)));
  const size_t smem = (size_t)2 * R * C * sizeof(float);
  TG_CHECK(smem <= 64 * 1024, "groupnorm: smem %zu too large", smem);
  // workspace layout (fixed offsets, independent of B): [256 arrival counters][256 x 32 x (mean, rstd)][partials]
  int* counters = reinterpret_cast<int*>(workspace);
  float* stats = workspace + GN_MAX_BATCH;
  float* partial = stats + (size_t)GN_MAX_BATCH * GN_GROUPS * 2;
  TG_CUDA(launch_kernel(gn_stats_kernel, dim3(slabs, B), dim3(GN_THREADS), smem, st, 1, x1, C1, x2, C2, HW, slabs, partial, stats, counters, eps));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  int aslabs = (g_gn_apply_per_sm * 148 + B - 1) / B;
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
