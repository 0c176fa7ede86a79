// Fused attention kernels (flash-style: the N x N probability matrix never reaches HBM).
//
//  attn_self_kernel     : softmax(alpha Q K^T) V per (batch, head); online softmax, fp32 statistics.
//                         Replaces AttnProcessor.__call__ (ip_adapter/attention_processor.py:343-356) whose
//                         baddbmm -> softmax -> bmm materialises [B*8, N, N] fp16 (537 MB per layer at 64x64).
//  attn_cross_ip_kernel : text cross-attention + IP-Adapter image-token branch in one pass:
//                           out = softmax(alpha Q Kt^T [+mask]) Vt + ip_scale[b] * softmax(alpha Q Ki^T) Vi
//                         with optional capture of the text probabilities [B,8,N,n_text]
//                         (IPAttnProcessor.__call__, ip_adapter/attention_processor.py:478-516, 532-549).
//
// Tensor-core path of this first version: mma.sync.m16n8k16 (fp16 in, fp32 accumulate) fed by
// ldmatrix from cp.async-staged shared memory.  Head dims 40 / 80 / 160 (SD-1.5: C/8), 64 (resampler).
#include "common.cuh"
#include "ops.h"

namespace tg {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// cooperative tile load: ROWS x D halfs from global (row pitch ld) into smem (row pitch PITCH); rows >= valid
// are zero-filled.  D*2 bytes is a multiple of 16.
template <int D, int PITCH, int NTHREADS>
__device__ __forceinline__ void load_tile(__half* s, const __half* g, int ld, int rows, int valid, int tid) {
  constexpr int CH = D / 8;
  for (int i = tid; i < rows * CH; i += NTHREADS) {
    const int r = i / CH, c = (i % CH) * 8;
    if (r < valid)
      cp_async16(s + r * PITCH + c, g + (size_t)r * ld + c);
    else
      *reinterpret_cast<uint4*>(s + r * PITCH + c) = make_uint4(0, 0, 0, 0);
  }
}

// =========================================================================================================
// self attention
// =========================================================================================================
template <int D, int NWARPS>
struct SelfCfg {
  static constexpr int DP = (D + 15) / 16 * 16;  // MMA-k padded head dim
  static constexpr int PITCH = DP + 8;           // halfs; (PITCH*2/16) odd -> conflict-free ldmatrix
  static constexpr int BQ = NWARPS * 16;
  static constexpr int BKV = 64;
  static constexpr int SMEM = (BQ + 4 * BKV) * PITCH * 2;
};

template <int D, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_self_kernel(const __half* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V, int ldq,
                 int ldk, int ldv, __half* __restrict__ O, int ldo, int Nq, int Nk, float scale_log2, int causal) {
  grid_dep_launch();
  grid_dep_wait();
  using Cfg = SelfCfg<D, NWARPS>;
  constexpr int DP = Cfg::DP, PITCH = Cfg::PITCH, BQ = Cfg::BQ, BKV = Cfg::BKV, NT = NWARPS * 32;
  constexpr int KSTEPS = DP / 16;     // k16 steps over the head dim for S = Q K^T
  constexpr int DT = (D + 7) / 8;     // n8 output tiles over the head dim
  constexpr int DT2 = (DT + 1) / 2;   // ldmatrix.x4.trans covers two of them
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __half* sQ = reinterpret_cast<__half*>(smem_attn);
  __half* sK = sQ + BQ * PITCH;       // [2][BKV][PITCH]
  __half* sV = sK + 2 * BKV * PITCH;  // [2][BKV][PITCH]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const __half* gQ = Q + ((size_t)b * Nq + q0) * ldq + h * D;
  const __half* gK = K + (size_t)b * Nk * ldk + h * D;
  const __half* gV = V + (size_t)b * Nk * ldv + h * D;

  // zero the pad columns once (cp.async never writes them)
  if constexpr (DP != D) {
    constexpr int PADC = DP - D;  // 8 for D=40
    for (int i = tid; i < (BQ + 4 * BKV) * (PADC / 8); i += NT) {
      const int r = i / (PADC / 8), c = D + (i % (PADC / 8)) * 8;
      *reinterpret_cast<uint4*>(sQ + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    }
  }
  load_tile<D, PITCH, NT>(sQ, gQ, ldq, BQ, min(BQ, Nq - q0), tid);
  load_tile<D, PITCH, NT>(sK, gK, ldk, BKV, min(BKV, Nk), tid);
  load_tile<D, PITCH, NT>(sV, gV, ldv, BKV, min(BKV, Nk), tid);
  cp_async_commit();

  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];

  const int nkv = (Nk + BKV - 1) / BKV;
  for (int it = 0; it < nkv; ++it) {
    const int buf = it & 1;
    if (it + 1 < nkv) {
      const int k1 = (it + 1) * BKV;
      load_tile<D, PITCH, NT>(sK + (buf ^ 1) * BKV * PITCH, gK + (size_t)k1 * ldk, ldk, BKV, min(BKV, Nk - k1), tid);
      load_tile<D, PITCH, NT>(sV + (buf ^ 1) * BKV * PITCH, gV + (size_t)k1 * ldv, ldv, BKV, min(BKV, Nk - k1), tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks)
        ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8);
    }
    const __half* cK = sK + buf * BKV * PITCH;
    const __half* cV = sV + buf * BKV * PITCH;

    // ---- S = Q K^T  (16 x 64 per warp) ----
    float s[BKV / 8][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int j2 = 0; j2 < BKV / 16; ++j2) {
        uint32_t kb[4];
        ldsm_x4(kb, cK + (j2 * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8);
        mma16816(s[2 * j2], qf[ks], kb[0], kb[1]);
        mma16816(s[2 * j2 + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // mask keys beyond Nk in the last tile
    const int kbase = it * BKV;
    if (kbase + BKV > Nk) {
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
        const int c = kbase + j * 8 + (lane & 3) * 2;
        if (c >= Nk) s[j][0] = s[j][2] = -INFINITY;
        if (c + 1 >= Nk) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    // causal mask (CLIP text encoder): key index > query index.  Key 0 is visible to every row, so the running maximum
    // is finite after the first tile and fully masked later tiles contribute exp2(-inf) = 0.
    if (causal) {
      const int qr0 = q0 + warp * 16 + (lane >> 2), qr1 = qr0 + 8;
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
        const int c = kbase + j * 8 + (lane & 3) * 2;
        if (c > qr0) s[j][0] = -INFINITY;
        if (c + 1 > qr0) s[j][1] = -INFINITY;
        if (c > qr1) s[j][2] = -INFINITY;
        if (c + 1 > qr1) s[j][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g and g+8 of the warp's 16) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r]);
      corr[r] = exp2f((m_run[r] - mnew) * scale_log2);
      m_run[r] = mnew;
      msc[r] = mnew * scale_log2;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      s[j][0] = exp2f(s[j][0] * scale_log2 - msc[0]);
      s[j][1] = exp2f(s[j][1] * scale_log2 - msc[0]);
      s[j][2] = exp2f(s[j][2] * scale_log2 - msc[1]);
      s[j][3] = exp2f(s[j][3] * scale_log2 - msc[1]);
      rs[0] += s[j][0] + s[j][1];
      rs[1] += s[j][2] + s[j][3];
    }
    l_run[0] = l_run[0] * corr[0] + rs[0];
    l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < DT2; ++d2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, cV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + d2 * 16 + (lane >> 4) * 8);
        mma16816(o[2 * d2], pa, vb[0], vb[1]);
        if (2 * d2 + 1 < DT) mma16816(o[2 * d2 + 1], pa, vb[2], vb[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled next iteration
  }

  // ---- finalise ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
  __half* gO = O + (size_t)b * Nq * ldo + h * D;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const int c = i * 8 + (lane & 3) * 2;
    if (row0 < Nq) *reinterpret_cast<__half2*>(gO + (size_t)row0 * ldo + c) = __floats2half2_rn(o[i][0] * inv0, o[i][1] * inv0);
    if (row1 < Nq) *reinterpret_cast<__half2*>(gO + (size_t)row1 * ldo + c) = __floats2half2_rn(o[i][2] * inv1, o[i][3] * inv1);
  }
}

template <int D, int NWARPS>
static int launch_self(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O,
                       int ldo, int B, int heads, int Nq, int Nk, float scale, int causal, cudaStream_t st) {
  using Cfg = SelfCfg<D, NWARPS>;
  auto kern = attn_self_kernel<D, NWARPS>;
  TG_SMEM_ATTR_ONCE(kern, Cfg::SMEM);
  dim3 grid((Nq + Cfg::BQ - 1) / Cfg::BQ, heads, B);
  TG_CUDA(launch_kernel(kern, dim3(grid), dim3(NWARPS * 32), Cfg::SMEM, st, 1, Q, K, V, ldq, ldk, ldv, O, ldo, Nq, Nk, scale * 1.4426950408889634f, causal));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

int attention_flash(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                    int B, int heads, int Nq, int Nk, int d, float scale, int causal, cudaStream_t st) {
  TG_CHECK(!causal || Nq == Nk, "attention: the causal mask needs Nq == Nk");
  TG_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: pitches must be 16B aligned");
  const bool big = Nq >= 128;
#define TG_SELF(DD)                                                                                         \
  case DD:                                                                                                  \
    return big ? launch_self<DD, 8>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, Nq, Nk, scale, causal, st)    \
               : launch_self<DD, 4>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, Nq, Nk, scale, causal, st);
  switch (d) {
    TG_SELF(40) TG_SELF(64) TG_SELF(80) TG_SELF(160)
    default:
      TG_CHECK(false, "attention: head dim %d not instantiated (40, 64, 80, 160)", d);
  }
#undef TG_SELF
  return 0;
}

// =========================================================================================================
// cross attention with the IP-Adapter image-token branch
// =========================================================================================================
static constexpr int XT = 80;  // text keys padded (n_text <= 80; CLIP = 77)
static constexpr int XI = 16;  // image-token keys padded (n_ip <= 16; IPAdapter = 4, Plus = 16)

template <int D, int NWARPS>
struct CrossCfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int PITCH = DP + 8;
  static constexpr int BQ = NWARPS * 16;
  static constexpr int SMEM = (BQ + 2 * XT + 2 * XI) * PITCH * 2;
};

template <int NKEYS, int D, int DP, int PITCH>
__device__ __forceinline__ void cross_branch(const uint32_t (&qf)[DP / 16][4], const __half* sK, const __half* sV,
                                             int nvalid, float scale_log2, const __half* add_mask, float w,
                                             float (&o)[(D + 7) / 8][4], int lane, __half* probs_out, int ldp,
                                             int row0, int row1, int Nq) {
  constexpr int NTILES = NKEYS / 8, KSTEPS = DP / 16, DT = (D + 7) / 8, DT2 = (DT + 1) / 2;
  float s[NTILES][4];
#pragma unroll
  for (int j = 0; j < NTILES; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
    for (int j2 = 0; j2 < NKEYS / 16; ++j2) {
      uint32_t kb[4];
      ldsm_x4(kb, sK + (j2 * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8);
      mma16816(s[2 * j2], qf[ks], kb[0], kb[1]);
      mma16816(s[2 * j2 + 1], qf[ks], kb[2], kb[3]);
    }
  }
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < NTILES; ++j) {
    const int c = j * 8 + (lane & 3) * 2;
    // scores in log2 units: alpha*s*log2e (+ mask*log2e); padded keys -> -inf
    float m0 = 0.f, m1 = 0.f;
    if (add_mask != nullptr) {
      if (c < nvalid) m0 = __half2float(add_mask[c]) * 1.4426950408889634f;
      if (c + 1 < nvalid) m1 = __half2float(add_mask[c + 1]) * 1.4426950408889634f;
    }
    s[j][0] = c < nvalid ? s[j][0] * scale_log2 + m0 : -INFINITY;
    s[j][2] = c < nvalid ? s[j][2] * scale_log2 + m0 : -INFINITY;
    s[j][1] = c + 1 < nvalid ? s[j][1] * scale_log2 + m1 : -INFINITY;
    s[j][3] = c + 1 < nvalid ? s[j][3] * scale_log2 + m1 : -INFINITY;
    mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
    mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
#pragma unroll
  for (int j = 0; j < NTILES; ++j) {
    s[j][0] = exp2f(s[j][0] - mx[0]); s[j][1] = exp2f(s[j][1] - mx[0]);
    s[j][2] = exp2f(s[j][2] - mx[1]); s[j][3] = exp2f(s[j][3] - mx[1]);
    sum[0] += s[j][0] + s[j][1];
    sum[1] += s[j][2] + s[j][3];
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
  }
  const float inv0 = 1.f / sum[0], inv1 = 1.f / sum[1];
#pragma unroll
  for (int j = 0; j < NTILES; ++j) {
    s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;
  }
  if (probs_out != nullptr) {  // capture [.., N, n_text] fp16 (attention_processor.py:534)
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
      const int c = j * 8 + (lane & 3) * 2;
      if (row0 < Nq) {
        if (c < nvalid) probs_out[(size_t)row0 * ldp + c] = __float2half_rn(s[j][0]);
        if (c + 1 < nvalid) probs_out[(size_t)row0 * ldp + c + 1] = __float2half_rn(s[j][1]);
      }
      if (row1 < Nq) {
        if (c < nvalid) probs_out[(size_t)row1 * ldp + c] = __float2half_rn(s[j][2]);
        if (c + 1 < nvalid) probs_out[(size_t)row1 * ldp + c + 1] = __float2half_rn(s[j][3]);
      }
    }
  }
  // the reference rounds probabilities to fp16 before P.V (attention_processor.py:217) -- pack_h2 does the same;
  // the branch weight w (1 for text, ip_scale for image tokens) is applied to the fp32 partial product.
  float acc[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NKEYS / 16; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int d2 = 0; d2 < DT2; ++d2) {
      uint32_t vb[4];
      ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + d2 * 16 + (lane >> 4) * 8);
      mma16816(acc[2 * d2], pa, vb[0], vb[1]);
      if (2 * d2 + 1 < DT) mma16816(acc[2 * d2 + 1], pa, vb[2], vb[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    // the reference rounds each branch's P.V to fp16 before the scaled sum (attention_processor.py:492,513-516)
    o[i][0] += w * __half2float(__float2half_rn(acc[i][0]));
    o[i][1] += w * __half2float(__float2half_rn(acc[i][1]));
    o[i][2] += w * __half2float(__float2half_rn(acc[i][2]));
    o[i][3] += w * __half2float(__float2half_rn(acc[i][3]));
  }
}

template <int D, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_cross_ip_kernel(const __half* __restrict__ Q, int ldq, const __half* __restrict__ Kt,
                     const __half* __restrict__ Vt, const __half* __restrict__ Ki, const __half* __restrict__ Vi,
                     int ldkv, int n_text, int n_ip, const float* __restrict__ ip_scale,
                     const __half* __restrict__ add_mask, __half* __restrict__ O, int ldo,
                     __half* __restrict__ probs, int probs_b0, int Nq, float scale_log2, int qtiles) {
  grid_dep_launch();
  grid_dep_wait();
  using Cfg = CrossCfg<D, NWARPS>;
  constexpr int DP = Cfg::DP, PITCH = Cfg::PITCH, BQ = Cfg::BQ, NT = NWARPS * 32, DT = (D + 7) / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __half* sQ = reinterpret_cast<__half*>(smem_attn);
  __half* sKt = sQ + BQ * PITCH;
  __half* sVt = sKt + XT * PITCH;
  __half* sKi = sVt + XT * PITCH;
  __half* sVi = sKi + XI * PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A CTA walks `qtiles` consecutive query tiles of its (image, head): the text / image-token K and V (13 KB, more than
  // a query tile) are staged once, and the next query tile streams in while the current one is computed.
  const int b = blockIdx.z, h = blockIdx.y, qfirst = blockIdx.x * qtiles * BQ;
  const int T = n_text + n_ip;  // context rows per batch element in the K/V projection buffer

  if constexpr (DP != D) {
    constexpr int PADC = DP - D;
    for (int i = tid; i < (BQ + 2 * XT + 2 * XI) * (PADC / 8); i += NT) {
      const int r = i / (PADC / 8), c = D + (i % (PADC / 8)) * 8;
      *reinterpret_cast<uint4*>(sQ + r * PITCH + c) = make_uint4(0, 0, 0, 0);
    }
  }
  load_tile<D, PITCH, NT>(sQ, Q + ((size_t)b * Nq + qfirst) * ldq + h * D, ldq, BQ, min(BQ, Nq - qfirst), tid);
  load_tile<D, PITCH, NT>(sKt, Kt + (size_t)b * T * ldkv + h * D, ldkv, XT, n_text, tid);
  load_tile<D, PITCH, NT>(sVt, Vt + (size_t)b * T * ldkv + h * D, ldkv, XT, n_text, tid);
  if (n_ip > 0) {
    load_tile<D, PITCH, NT>(sKi, Ki + ((size_t)b * T + n_text) * ldkv + h * D, ldkv, XI, n_ip, tid);
    load_tile<D, PITCH, NT>(sVi, Vi + ((size_t)b * T + n_text) * ldkv + h * D, ldkv, XI, n_ip, tid);
  }
  cp_async_commit();

  for (int t = 0; t < qtiles; ++t) {
  const int q0 = qfirst + t * BQ;
  if (q0 >= Nq) break;
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qf[DP / 16][4];
#pragma unroll
  for (int ks = 0; ks < DP / 16; ++ks)
    ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8);
  if (t + 1 < qtiles && q0 + BQ < Nq) {  // every warp holds its Q fragments: the staging buffer takes the next tile
    __syncthreads();
    load_tile<D, PITCH, NT>(sQ, Q + ((size_t)b * Nq + q0 + BQ) * ldq + h * D, ldq, BQ, min(BQ, Nq - q0 - BQ), tid);
    cp_async_commit();
  }

  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;

  __half* pout = nullptr;
  if (probs != nullptr && b >= probs_b0)
    pout = probs + (((size_t)(b - probs_b0) * gridDim.y + h) * Nq) * n_text;
  cross_branch<XT, D, DP, PITCH>(qf, sKt, sVt, n_text, scale_log2, add_mask ? add_mask + (size_t)b * n_text : nullptr,
                                 1.f, o, lane, pout, n_text, row0, row1, Nq);
  if (n_ip > 0)
    cross_branch<XI, D, DP, PITCH>(qf, sKi, sVi, n_ip, scale_log2, nullptr, ip_scale ? ip_scale[b] : 1.f, o, lane,
                                   nullptr, 0, row0, row1, Nq);

  __half* gO = O + (size_t)b * Nq * ldo + h * D;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const int c = i * 8 + (lane & 3) * 2;
    if (row0 < Nq) *reinterpret_cast<__half2*>(gO + (size_t)row0 * ldo + c) = __floats2half2_rn(o[i][0], o[i][1]);
    if (row1 < Nq) *reinterpret_cast<__half2*>(gO + (size_t)row1 * ldo + c) = __floats2half2_rn(o[i][2], o[i][3]);
  }
  }  // query tiles of this CTA
}

template <int D, int NWARPS>
static int launch_cross(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki,
                        const __half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                        const __half* add_mask, __half* O, int ldo, __half* probs, int probs_b0, int B, int heads,
                        int Nq, float scale, cudaStream_t st) {
  using Cfg = CrossCfg<D, NWARPS>;
  auto kern = attn_cross_ip_kernel<D, NWARPS>;
  TG_SMEM_ATTR_ONCE(kern, Cfg::SMEM);
  // query tiles per CTA: 2 or 4 while at least ~8 CTAs per SM remain (the K/V staging is then shared by more queries)
  const int tiles = (Nq + Cfg::BQ - 1) / Cfg::BQ;
  int qtiles = 1;
  while (qtiles < 4 && tiles % (2 * qtiles) == 0 && (tiles / (2 * qtiles)) * heads * B >= 148 * 8) qtiles *= 2;
  dim3 grid((tiles + qtiles - 1) / qtiles, heads, B);
  TG_CUDA(launch_kernel(kern, dim3(grid), dim3(NWARPS * 32), Cfg::SMEM, st, 1, Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo,
                                             probs, probs_b0, Nq, scale * 1.4426950408889634f, qtiles));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

static int g_cross_impl = -2;  // -2: not read from the environment yet
void attn_cross_set_impl(int mode) { g_cross_impl = mode; }

int attention_cross_ip(const __half* Q, int ldq, const __half* Kt, const __half* Vt, const __half* Ki,
                       const __half* Vi, int ldkv, int n_text, int n_ip, const float* ip_scale,
                       const __half* add_mask, __half* O, int ldo, __half* probs, int probs_b0, int B, int heads,
                       int Nq, int d, float scale, cudaStream_t st) {
  // Two implementations with the same results (tests/test_ops_gpu.py runs both on every shape):
  //   attn_cross_tc.cu  tcgen05 / TMEM: one 96-key tile, TS-form MMAs -- used for head dims <= 64 (the 64x64 level, 78 %
  //                     of the cross-attention bytes), where it matches the mma.sync kernel (38.4 vs 38.9 us at B = 10)
  //   this file         mma.sync, several small CTAs per SM -- still faster for d = 80 / 160 (20.7 vs 31.0 us, 15.6 vs
  //                     23.2 us): one 128-query tile per tensor-memory allocation is a strictly sequential
  //                     Q -> S -> softmax -> P -> O chain and 512 columns do not hold two of them at these head dims
  // tg_attn_cross_set_impl / TG_CROSS_TC: -1 auto (default), 0 mma.sync everywhere, 1 tcgen05 everywhere.
  if (g_cross_impl == -2) g_cross_impl = getenv("TG_CROSS_TC") ? atoi(getenv("TG_CROSS_TC")) : -1;
  if (g_cross_impl == 1 || (g_cross_impl == -1 && d <= 64))
    return attention_cross_ip_tc(Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo, probs, probs_b0, B, heads,
                                 Nq, d, scale, st);
  TG_CHECK(n_text >= 1 && n_text <= XT && n_ip >= 0 && n_ip <= XI, "cross attention: n_text=%d (<=80) n_ip=%d (<=16)",
           n_text, n_ip);
  TG_CHECK(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0, "cross attention: pitches must be 16B aligned");
#define TG_X(DD)                                                                                                    \
  case DD:                                                                                                          \
    return launch_cross<DD, 4>(Q, ldq, Kt, Vt, Ki, Vi, ldkv, n_text, n_ip, ip_scale, add_mask, O, ldo, probs,       \
                               probs_b0, B, heads, Nq, scale, st);
  switch (d) {
    TG_X(40) TG_X(80) TG_X(160)
    default:
      TG_CHECK(false, "cross attention: head dim %d not instantiated (40, 80, 160)", d);
  }
#undef TG_X
  return 0;
}

}  // namespace tg
