// tcgen05 + TMA GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] * Wt[N,K]^T )        fp16 operands, fp32 accumulate in TMEM, fp16 out
//
// A-operand modes
//   LINEAR : A is row-major [M,K] (optionally the channel-concat of two tensors: K = K1 + K2), 2D TMA
//   CONV3  : A is an NHWC activation [B,H,W,Cin]; K = 9*Cin ordered (ky,kx,c).  Each K-block of 64
//            channels of one filter tap is one 4D TMA box whose out-of-bounds rows/cols are zero-filled by
//            the TMA unit -> the conv padding costs nothing and no im2col buffer exists.
// Both land in shared memory as the canonical K-major SWIZZLE_128B layout that tcgen05.mma consumes.
//
// Roles (192 threads): warp0 = TMA producer, warp1 = TMEM allocator + single-thread MMA issuer,
// warps2..5 = epilogue (TMEM -> registers -> fused bias/temb/residual/activation -> global).
// Replaces, for the hot path, every cuBLAS/cuDNN call the reference issues through
// nn.Linear / nn.Conv2d (reference: ip_adapter/attention_processor.py:333-359,461-520,
// ip_adapter/attention.py:269-367, ip_adapter/transformer_2d.py:288-342, diffusers ResnetBlock2D).
#include "common.cuh"
#include "ops.h"

namespace tg {

static constexpr int BM = 128;
static constexpr int BK = 64;

struct GemmArgs {
  int M, N, K, K1;
  int mode;  // 0 linear, 1 conv3x3
  int H, W, Cin, box_h, box_b;
  const __half* bias;
  const __half* rowbias;
  int rows_per_image, rowbias_ld;
  const __half* residual;
  int ldr;
  __half* out;
  int ldc;
  int act;  // 0 none, 1 silu, 2 geglu (tile columns packed [32 value | 32 gate] per 64), 3 gelu(erf)
  float out_scale;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ void epilogue_store32(const GemmArgs& g, float (&v)[32], int m, int n, bool row_ok) {
  // v: 32 consecutive output columns [n, n+32) of row m (already scaled / activated except bias etc.)
  if (g.bias) {
    const uint4* bp = reinterpret_cast<const uint4*>(g.bias + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 b = __ldg(bp + i);
      const __half2* h = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        v[i * 8 + j * 2] += f.x;
        v[i * 8 + j * 2 + 1] += f.y;
      }
    }
  }
  if (!row_ok) return;
  if (g.rowbias) {
    const uint4* bp =
        reinterpret_cast<const uint4*>(g.rowbias + (size_t)(m / g.rows_per_image) * g.rowbias_ld + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 b = __ldg(bp + i);
      const __half2* h = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        v[i * 8 + j * 2] += f.x;
        v[i * 8 + j * 2 + 1] += f.y;
      }
    }
  }
  if (g.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
  } else if (g.act == 3) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
  }
  if (g.residual) {
    const uint4* rp = reinterpret_cast<const uint4*>(g.residual + (size_t)m * g.ldr + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 b = __ldg(rp + i);
      const __half2* h = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        v[i * 8 + j * 2] += f.x;
        v[i * 8 + j * 2 + 1] += f.y;
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(g.out + (size_t)m * g.ldc + n);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      h[j] = __floats2half2_rn(v[i * 8 + j * 2] * g.out_scale, v[i * 8 + j * 2 + 1] * g.out_scale);
    op[i] = o;
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int nk = g.K / BK;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int conv_b0 = 0, conv_y0 = 0, cblk = 1;
      if (g.mode == 1) {
        const int hw = g.H * g.W;
        conv_b0 = m0 / hw;
        conv_y0 = (m0 % hw) / g.W;
        cblk = g.Cin / BK;
      }
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if (g.mode == 0) {
          const int k0 = it * BK;
          if (k0 < g.K1)
            tma_load_2d(sa, &tmA1, &full_bar[s], k0, m0);
          else
            tma_load_2d(sa, &tmA2, &full_bar[s], k0 - g.K1, m0);
        } else {
          const int tap = it / cblk;
          const int c0 = (it - tap * cblk) * BK;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(sa, &tmA1, &full_bar[s], c0, kx - 1, conv_y0 + ky - 1, conv_b0);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], it * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * 32);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * 32);
          umma_f16_ss(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      tc_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < g.M;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (g.act != 2) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + c, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        epilogue_store32(g, v, m, n0 + c, row_ok);
      }
    } else {
      // GEGLU: per 64 packed columns -> 32 outputs value*gelu(gate); bias is packed the same way
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        uint32_t ra[32], rg[32];
        tmem_ld_32x32b_x32(lane_addr + c, ra);
        tmem_ld_32x32b_x32(lane_addr + c + 32, rg);
        tmem_ld_wait();
        float v[32];
        const uint4* bpa = reinterpret_cast<const uint4*>(g.bias + n0 + c);
        const uint4* bpg = reinterpret_cast<const uint4*>(g.bias + n0 + c + 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 ba = __ldg(bpa + i), bg = __ldg(bpg + i);
          const __half2* ha = reinterpret_cast<const __half2*>(&ba);
          const __half2* hg = reinterpret_cast<const __half2*>(&bg);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 fa = __half22float2(ha[j]), fg = __half22float2(hg[j]);
            const int e = i * 8 + j * 2;
            // the reference rounds both halves of the projection to fp16 before the gated product
            // (ip_adapter/attention.py:359-367); keep that rounding so GELU sees the same gate
            float a0 = __half2float(__float2half_rn(__uint_as_float(ra[e]) + fa.x));
            float a1 = __half2float(__float2half_rn(__uint_as_float(ra[e + 1]) + fa.y));
            float g0 = __half2float(__float2half_rn(__uint_as_float(rg[e]) + fg.x));
            float g1 = __half2float(__float2half_rn(__uint_as_float(rg[e + 1]) + fg.y));
            v[e] = a0 * gelu_erf_f(g0);
            v[e + 1] = a1 * gelu_erf_f(g1);
          }
        }
        if (row_ok) {
          uint4* op = reinterpret_cast<uint4*>(g.out + (size_t)m * g.ldc + (n0 + c) / 2);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[i * 8 + j * 2], v[i * 8 + j * 2 + 1]);
            op[i] = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b, const GemmArgs& g,
                       cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, STAGES>;
  if (!attr_set) {
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  dim3 grid(g.N / BN, (g.M + BM - 1) / BM);
  kern<<<grid, 192, L::TOTAL, st>>>(a1, a2, b, g);
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

static int dispatch(const CUtensorMap& a1, const CUtensorMap& a2, const GemmArgs& g, const __half* Wt,
                    cudaStream_t st) {
  int bn = (g.N % 128 == 0) ? 128 : 64;
  if (g.act == 2) TG_CHECK(g.N % 64 == 0 && g.bias != nullptr, "geglu epilogue needs N%%64==0 and a bias");
  CUtensorMap b;
  if (make_tmap_2d(&b, Wt, g.N, g.K, g.K, bn)) return -1;
  if (bn == 128) return launch_gemm<128, 3>(a1, a2, b, g, st);
  return launch_gemm<64, 4>(a1, a2, b, g, st);
}

int gemm_linear(const __half* A1, int lda1, int K1, const __half* A2, int lda2, const __half* Wt, int M, int N,
                int K, const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st) {
  TG_CHECK(M > 0 && N % 64 == 0 && K % 64 == 0 && K1 % 64 == 0 && K1 <= K, "gemm: bad dims M=%d N=%d K=%d K1=%d", M,
           N, K, K1);
  TG_CHECK(A2 != nullptr || K1 == K, "gemm: K1<K needs a second A source");
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.K1 = K1; g.mode = 0;
  g.bias = e.bias; g.rowbias = e.rowbias; g.rows_per_image = e.rows_per_image > 0 ? e.rows_per_image : 1;
  g.rowbias_ld = e.rowbias_ld; g.residual = e.residual; g.ldr = e.ldr; g.out = out; g.ldc = ldc; g.act = e.act;
  g.out_scale = e.out_scale;
  CUtensorMap a1, a2;
  if (make_tmap_2d(&a1, A1, M, K1, lda1, BM)) return -1;
  if (K1 < K) {
    if (make_tmap_2d(&a2, A2, M, K - K1, lda2, BM)) return -1;
  } else {
    a2 = a1;
  }
  return dispatch(a1, a2, g, Wt, st);
}

int conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, const EpilogueSpec& e,
                 __half* out, int ldc, cudaStream_t st) {
  TG_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "conv3x3: Cin/Cout must be multiples of 64 (got %d/%d)", Cin, Cout);
  TG_CHECK(W <= 128 && 128 % W == 0, "conv3x3: W must divide 128 (got %d)", W);
  int box_h = 128 / W, box_b = 1;
  if (box_h > H) {
    box_b = box_h / H;
    box_h = H;
  }
  TG_CHECK(H % box_h == 0 && B % box_b == 0 && box_b * box_h * W == 128, "conv3x3: tile %dx%dx%d does not cover 128",
           box_b, box_h, W);
  GemmArgs g{};
  g.M = B * H * W; g.N = Cout; g.K = 9 * Cin; g.K1 = g.K; g.mode = 1;
  g.H = H; g.W = W; g.Cin = Cin; g.box_h = box_h; g.box_b = box_b;
  g.bias = e.bias; g.rowbias = e.rowbias; g.rows_per_image = e.rows_per_image > 0 ? e.rows_per_image : 1;
  g.rowbias_ld = e.rowbias_ld; g.residual = e.residual; g.ldr = e.ldr; g.out = out; g.ldc = ldc; g.act = e.act;
  g.out_scale = e.out_scale;
  CUtensorMap a1;
  if (make_tmap_nhwc(&a1, X, B, H, W, Cin, box_b, box_h, W)) return -1;
  return dispatch(a1, a1, g, Wt, st);
}

}  // namespace tg
