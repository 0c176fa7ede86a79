// tcgen05 + TMA GEMM / implicit-GEMM 3x3 convolution for sm_100a (persistent, warp-specialised, split-K).
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
// One CTA per SM loops over work items (128 x BN output tile, K-split); N fastest so co-running CTAs share A
// rows in L2.  Roles (192 threads): warp0 = TMA producer (STAGES-deep smem ring), warp1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps2..5 = epilogue.  The accumulator is double-buffered in TMEM
// (2 x BN columns), so the epilogue of item i (TMEM -> registers -> bias/temb/activation/residual -> swizzled
// smem -> TMA store) overlaps the main loop of item i+1.
// CTA pairs (PAIR = true, used whenever there are >= 2 M-tiles): two CTAs of one TPC run as a cluster and own a
// 256 x BN tile through tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages its own 128 rows of A and only HALF of
// the B (weight) tile (BN/2 rows); the tensor cores of both SMs read both halves, so the L2->smem traffic per SM and
// k-block drops from 16 KB + 128*BN to 16 KB + 64*BN bytes -- the smem ingest, not the tensor pipe, is what bounds a
// single-CTA tile (measured ~51 B/clk/SM).  The leader CTA (cluster rank 0) issues every MMA; the TMA loads of both CTAs
// complete on the leader's full barrier (cp.async.bulk.tensor .cta_group::2), a multicast tcgen05.commit releases the
// smem slot / publishes the accumulator in BOTH CTAs, and the epilogue warps of both CTAs arrive on the leader's
// accumulator-empty barrier through its shared::cluster address.  Each CTA runs the epilogue of its own 128 rows.
// Split-K (small-M, long-K layers: the 8x8 / 16x16 levels stream 30-60 MB of weights through <= 40 tiles):
// every split writes its fp32 partial tile to an L2-resident workspace, a per-tile semaphore elects the last
// arriving CTA, which sums the partials in a FIXED order (bit-reproducible) and runs the fused epilogue.
// Replaces, for the hot path, every cuBLAS/cuDNN call the reference issues through
// nn.Linear / nn.Conv2d (reference: ip_adapter/attention_processor.py:333-359,461-520,
// ip_adapter/attention.py:269-367, ip_adapter/transformer_2d.py:288-342, diffusers ResnetBlock2D).
#include <stdlib.h>

#include <unordered_map>

#include "common.cuh"
#include "ops.h"

namespace tg {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int KPS = 2;                    // k-blocks (64-wide, one SWIZZLE_128B tile each) per pipeline stage
static constexpr int EPI_WARPS = 8;              // two warps per TMEM lane quarter, alternating column chunks
static constexpr int EPI_BUF_BYTES = 2 * 32 * 64;  // per warp: two staging buffers of 32 rows x 32 fp16
static constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;
static constexpr int MAX_SPLIT = 8;
static constexpr size_t SPLITK_WS_BYTES = (size_t)64 << 20;
static constexpr int SPLITK_MAX_TILES = 4096;

struct GemmArgs {
  int M, N, K, K1;
  int mode;  // 0 linear, 1 convolution (implicit GEMM over a taps_w x taps_w window of the NHWC input)
  int H, W, Cin, box_h, box_b;  // H, W: OUTPUT image (= the M decomposition); the input map may gather with pixel stride 2
  // convolution geometry: tap (ky, kx) of output pixel (y, x) reads input pixel (y * in_stride + ky + off_y,
  // x * in_stride + kx + off_x).  3x3 pad 1: taps_w 3, off -1, stride 1.  Stride-2 3x3 (Downsample2D): in_stride 2.
  // Phase (py, px) of nearest-x2 + 3x3 on the low-resolution grid: taps_w 2, off (py - 1, px - 1).
  int taps_w, off_y, off_x, in_stride;
  int out3d_w;  // > 0: the output is a phase image of width out3d_w stored through a 3D map (make_tmap_phase_out)
  const __half* bias;
  const __half* rowbias;
  int rows_per_image, rowbias_ld;
  const __half* residual;
  int ldr;
  int act;  // 0 none, 1 silu, 2 geglu (tile columns packed [32 value | 32 gate] per 64), 3 gelu(erf), 5 key softmax
  // per-image B operand (folded cross-attention, tg_cross_scores / tg_cross_apply): image i = row / img_rows multiplies
  // rows [i * N, (i + 1) * N) of the stacked B matrix; 0 = one weight matrix for every row
  int img_rows;
  // act 5: the tile's columns are heads of 96 key slots (BN = 192: two heads); every output row becomes, per head, the
  // text softmax over slots [0, sm_text) and the image-token softmax over slots [80, 80 + sm_ip) times sm_ipscale[image];
  // the other slots are written as zeros.  Scores arrive already multiplied by scale * log2(e).
  int sm_text, sm_ip;
  const float* sm_ipscale;
  float out_scale;
  int num_m, num_n, splits;
  int patch_bytes, ky_stride;  // KY3 convolution: bytes of the (box_h + 2)-row input patch, bytes of one image row in it
  float* ws;   // split-K partials: [tile][split][BN/4][128][4] fp32
  int* sem;    // split-K arrival counters, one per tile (self-resetting)
  // GroupNorm statistics of the stored tile (EpilogueSpec::stats): [M / stats_hw][n_out / 2][2] fixed-point (sum, sumsq)
  unsigned long long* stats;
  int stats_hw;
  long long* prof;  // measurement only (tg_gemm_set_prof): per CTA {total, producer wait, MMA wait full, MMA wait acc, epilogue wait} cycles
};

// KY3 (3x3 convolution, image rows no wider than 64 pixels, one image per tile): a stage holds ONE input patch of
// box_h + 2 image rows (<= 32 KB) and the weight chunks of the three taps (ky = 0,1,2) of one (kx, channel block).  The
// three taps read the same patch at offsets of one image row (W x 128 B, a multiple of the 1024-byte swizzle period), so
// the activation bytes an SM ingests per three k-blocks drop from 48 KB to 16 KB + 256 B x W.
static constexpr int KY3_PATCH_MAX = 32768;
// Wider tiles would fit only two such stages (80 / 68 KB): measured slower than the tap-per-k-block ring with cold weights
// (16x16 level, N = 1280: 83 vs 76 us), so tiles of 192 / 256 columns keep the three-stage ring of two-k-block stages.
static constexpr int KY3_MAX_BN = 160;
template <int BN, int STAGES, bool PAIR, bool KY3 = false>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;                    // one k-block of A
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;  // one k-block of B; a CTA of a pair stages half of the weight tile
  static constexpr int A_STAGE_BYTES = KY3 ? KY3_PATCH_MAX : KPS * A_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + (KY3 ? 3 : KPS) * B_BYTES;  // [A kb0 | A kb1 | B kb0 | B kb1]
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_WARPS * EPI_BUF_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 32 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 32 staged rows x 32 columns -> global: rows [m, m + 32) of the [M, N] output, or -- phase image of an upsampling
// convolution -- 32 consecutive phase pixels (j fastest, then the image row) through the 3D map
__device__ __forceinline__ void store_chunk(const CUtensorMap* m, const void* smem_src, int n, int row, int out3d_w) {
  if (out3d_w > 0) tma_store_3d(m, smem_src, n, row % out3d_w, row / out3d_w);
  else tma_store_2d(m, smem_src, n, row);
}
// ---- CTA-pair (cta_group::2) flavours ----
// shared::cluster address of the same shared-memory location in the cluster's CTA `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
// TMA loads whose completion bytes are signalled on an mbarrier that may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive (count 1) on the same barrier in every CTA of cta_mask once all previously issued pair-MMAs have retired
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void add_h8(float* v, uint4 b) {
  const __half2* h = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
  }
}

// 32 output columns of one row: bias / temb / activation / residual, then fp16 into the swizzled staging buffer
// (32 rows x 64 bytes per warp, SWIZZLE_64B: 16-byte chunk i of row r lives at chunk i ^ ((r >> 1) & 3) -- must match
// the C tensor map).
__device__ __forceinline__ void epilogue_row(const GemmArgs& g, float (&v)[32], const uint4 (&res)[4], int m, int n,
                                             bool row_ok, uint8_t* stage_row, int row_in_warp) {
  if (g.bias && g.act != 2) {
    const uint4* bp = reinterpret_cast<const uint4*>(g.bias + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) add_h8(&v[i * 8], __ldg(bp + i));
  }
  if (row_ok && g.rowbias) {
    const uint4* bp = reinterpret_cast<const uint4*>(g.rowbias + (size_t)(m / g.rows_per_image) * g.rowbias_ld + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) add_h8(&v[i * 8], __ldg(bp + i));
  }
  if (g.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
  } else if (g.act == 3) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
  } else if (g.act == 4) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = quick_gelu_f(v[i]);
  }
  if (row_ok && g.residual) {
#pragma unroll
    for (int i = 0; i < 4; ++i) add_h8(&v[i * 8], res[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      h[j] = __floats2half2_rn(v[i * 8 + j * 2] * g.out_scale, v[i * 8 + j * 2 + 1] * g.out_scale);
    *reinterpret_cast<uint4*>(stage_row + ((i ^ ((row_in_warp >> 1) & 3)) << 4)) = o;
  }
}

// GroupNorm statistics of the chunk AS STORED (fp16-rounded values read back from the staging buffer): per (image, channel
// pair) the sum and the sum of squares over the warp's 32 rows x 2 columns, added to stats[image][pair] in 2^-24 fixed
// point with 64-bit integer atomics.  Integer addition is associative, so the totals do not depend on the order the CTAs
// arrive in: bit-reproducible, which float atomics would not be.  Lane (h, p): rows h, h+2, .. of column pair p -- the two
// half-warps read different rows (different bank halves), conflict-free.  The consumer is norm.cu::gn_apply_pre_kernel.
__device__ __forceinline__ void chunk_stats(const GemmArgs& g, const uint8_t* buf, int m_first, int n, int n_total, int lane) {
  const int p = lane & 15, h = lane >> 4;
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int r = h; r < 32; r += 2) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(buf + r * 64 + (((p >> 2) ^ ((r >> 1) & 3)) << 4) + ((p & 3) << 2));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    s += f.x + f.y;
    q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
  }
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  q += __shfl_xor_sync(0xffffffffu, q, 16);
  if (h == 0) {
    unsigned long long* dst = g.stats + ((size_t)(m_first / g.stats_hw) * (n_total >> 1) + ((n >> 1) + p)) * 2;
    atomicAdd(dst, (unsigned long long)__float2ll_rn(s * 16777216.f));
    atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(q * 16777216.f));
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(x));
  return p;
}

// residual prefetch (global, row-contiguous 64 bytes per thread): issued one chunk ahead so its latency overlaps the
// accumulator wait / the previous chunk's math
__device__ __forceinline__ void load_residual(const GemmArgs& g, uint4 (&res)[4], int m, int n, bool row_ok) {
  if (row_ok && g.residual) {
    const uint4* rp = reinterpret_cast<const uint4*>(g.residual + (size_t)m * g.ldr + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) res[i] = __ldg(rp + i);
  }
}

template <int BN, int STAGES, bool PAIR, bool KY3 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using L = GemmSmem<BN, STAGES, PAIR, KY3>;
  static_assert(!KY3 || PAIR, "the KY3 convolution path exists for CTA pairs only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches non-divergent and
  // the operands of tcgen05.mma / TMA instructions in uniform registers.  With `threadIdx.x >> 5` and a lane-0 branch
  // every tcgen05.mma was wrapped in an ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY loop (21 instructions, measured in the
  // attention kernel at ~115-170 cycles per MMA issued) -- more than the 64-128 cycles the MMA occupies the tensor pipe.
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int nk = KY3 ? g.K / (3 * BK) : g.K / BK;  // KY3: units of (kx, channel block) = three k-blocks each
  const int S = g.splits;
  // PAIR: a cluster of two CTAs owns the M-tile pair (2*pm, 2*pm + 1); rank selects the CTA's 128 rows and the half
  // of the weight tile it stages.  Rank 0 is the leader (MMA issuer; owner of the full / accumulator-empty barriers).
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int rows_m = PAIR ? (g.num_m + 1) / 2 : g.num_m;  // M-tile rows of the work grid
  const int num_work = rows_m * g.num_n * S;
  constexpr uint32_t ACC_STRIDE = BN <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // PAIR: the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {  // the same warp of both CTAs: one allocation spanning the pair's two tensor memories
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  grid_dep_launch();
  grid_dep_wait();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop (all lanes wait), one elected lane issues =====
    {
      const bool leader = elect_one();
      int it = 0;  // k-iteration counter across work items (ring position)
      long long p_wait = 0;
      const int hw = g.H * g.W;
      const int cblk = g.mode == 1 ? g.Cin / BK : 1;
      for (int w = worker; w < num_work; w += num_workers) {
        const int wt = w / S, split = w - wt * S;
        const int m0 = ((wt / g.num_n) * (PAIR ? 2 : 1) + rank) * BM, n0 = (wt % g.num_n) * BN;
        const int kb0 = (nk * split) / S, kb1 = (nk * (split + 1)) / S;
        int conv_b0 = 0, conv_y0 = 0, conv_x0 = 0;
        if (g.mode == 1) {
          conv_b0 = m0 / hw;
          conv_y0 = (m0 % hw) / g.W;
          conv_x0 = m0 % g.W;  // non-zero only for rows wider than a tile (W > 128: a tile is a 128-pixel row segment)
        }
        // One stage = KPS consecutive k-blocks: the loads of a stage are issued back to back after ONE wait.  Measured on
        // B200 (tools/probes): a group of TMA loads issued after a barrier wait costs ~570 cycles + ~70 per extra load,
        // whatever its size, and groups do not overlap each other -- so a stage must carry as much MMA work as possible.
        if constexpr (KY3) {
          // one patch (A) + one 4D weight box (the chunks of ky = 0,1,2) per stage: two TMA instructions for three k-blocks
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (g.prof) {
              const long long t0 = clock64();
              mbar_wait(&empty_bar[s], ph ^ 1);
              p_wait += clock64() - t0;
            } else {
              mbar_wait(&empty_bar[s], ph ^ 1);
            }
            uint8_t* sa = smem + s * L::STAGE_BYTES;
            uint8_t* sb = sa + L::A_STAGE_BYTES;
            const uint32_t bytes = (uint32_t)g.patch_bytes + 3 * L::B_BYTES;
            const uint32_t fb = mapa_rank(smem_u32(&full_bar[s]), 0);
            const int kx = kb / cblk, cb = kb - kx * cblk;
            if (leader) {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * bytes);
              tma_load_4d_2sm(sa, &tmA2, fb, cb * BK, kx - 1, conv_y0 - 1, conv_b0);
              tma_load_4d_2sm(sb, &tmB, fb, 0, n0 + rank * (BN / 2), kb, 0);
            }
          }
          continue;
        }
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (g.prof) {
            const long long t0 = clock64();
            mbar_wait(&empty_bar[s], ph ^ 1);
            p_wait += clock64() - t0;
          } else {
            mbar_wait(&empty_bar[s], ph ^ 1);
          }
          const int nkb = (kb1 - kb) < KPS ? (kb1 - kb) : KPS;
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + KPS * L::A_BYTES;
          // the weight box always brings KPS chunks (the one past K is zero-filled by the TMA unit and never multiplied)
          const uint32_t bytes = nkb * L::A_BYTES + KPS * L::B_BYTES;
          uint32_t fb = 0;
          if (PAIR) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            fb = mapa_rank(smem_u32(&full_bar[s]), 0);
            if (rank == 0 && leader) mbar_arrive_expect_tx(&full_bar[s], 2 * bytes);
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], bytes);
          }
          if (leader)
          for (int j = 0; j < nkb; ++j) {
            uint8_t* dst = sa + j * L::A_BYTES;
            if (g.mode == 0) {
              const int k0 = (kb + j) * BK;
              const CUtensorMap* tm = k0 < g.K1 ? &tmA1 : &tmA2;
              const int kc = k0 < g.K1 ? k0 : k0 - g.K1;
              if (PAIR) tma_load_2d_2sm(dst, tm, fb, kc, m0); else tma_load_2d(dst, tm, &full_bar[s], kc, m0);
            } else {
              const int tap = (kb + j) / cblk;
              const int c0 = (kb + j - tap * cblk) * BK;
              const int ky = tap / g.taps_w, kx = tap - ky * g.taps_w;
              const int ix = conv_x0 * g.in_stride + kx + g.off_x, iy = conv_y0 * g.in_stride + ky + g.off_y;
              if (PAIR) tma_load_4d_2sm(dst, &tmA1, fb, c0, ix, iy, conv_b0);
              else tma_load_4d(dst, &tmA1, &full_bar[s], c0, ix, iy, conv_b0);
            }
          }
          if (leader) {
            const int nb = g.img_rows > 0 ? n0 + (m0 / g.img_rows) * g.N : n0;  // per-image B: the image's block of rows
            if (PAIR) tma_load_3d_2sm(sb, &tmB, fb, 0, nb + rank * (BN / 2), kb);  // my half of the weight tile
            else tma_load_3d(sb, &tmB, &full_bar[s], 0, nb, kb);
          }
        }
      }
      if (g.prof && leader) g.prof[blockIdx.x * 8 + 1] = p_wait;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (waits included) so every address stays warp-uniform; one elected
    // lane issues the MMAs and commits =====
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * BM : BM, BN);
      int it = 0, j = 0;
      long long m_wait_full = 0, m_wait_acc = 0;
      const long long m_t0 = clock64();
      for (int w = worker; w < num_work; w += num_workers, ++j) {
        const int wt = w / S, split = w - wt * S;
        const int kb0 = (nk * split) / S, kb1 = (nk * (split + 1)) / S;
        const int as = j & 1;
        {
          const long long t0 = g.prof ? clock64() : 0;
          mbar_wait(&acc_empty[as], ((j >> 1) & 1) ^ 1);  // epilogue drained this accumulator
          if (g.prof) m_wait_acc += clock64() - t0;
        }
        tc_fence_after();
        const uint32_t d_addr = tmem_base + as * ACC_STRIDE;
        if constexpr (KY3) {
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (g.prof) {
              const long long t0 = clock64();
              mbar_wait(&full_bar[s], ph);
              m_wait_full += clock64() - t0;
            } else {
              mbar_wait(&full_bar[s], ph);
            }
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
            const uint32_t b_base = a_base + L::A_STAGE_BYTES;
            if (leader) {
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const uint32_t a_addr = a_base + ky * g.ky_stride, b_addr = b_base + ky * L::B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t da = make_smem_desc_sw128(a_addr + k * 32);
                  const uint64_t db = make_smem_desc_sw128(b_addr + k * 32);
                  umma_f16_ss_2sm(d_addr, da, db, idesc, (kb > kb0 || ky > 0 || k > 0) ? 1u : 0u);
                }
              }
              tc_commit_2sm(&empty_bar[s], 0x3);
            }
          }
          if (leader) tc_commit_2sm(&acc_full[as], 0x3);
          continue;
        }
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (g.prof) {
            const long long t0 = clock64();
            mbar_wait(&full_bar[s], ph);
            m_wait_full += clock64() - t0;
          } else {
            mbar_wait(&full_bar[s], ph);
          }
          tc_fence_after();
          const int nkb = (kb1 - kb) < KPS ? (kb1 - kb) : KPS;
          const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_base = a_base + KPS * L::A_BYTES;
          if (leader) {
            for (int j = 0; j < nkb; ++j) {
              const uint32_t a_addr = a_base + j * L::A_BYTES, b_addr = b_base + j * L::B_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t da = make_smem_desc_sw128(a_addr + k * 32);
                const uint64_t db = make_smem_desc_sw128(b_addr + k * 32);
                const uint32_t acc = (kb + j > kb0 || k > 0) ? 1u : 0u;
                if (PAIR)
                  umma_f16_ss_2sm(d_addr, da, db, idesc, acc);
                else
                  umma_f16_ss(d_addr, da, db, idesc, acc);
              }
            }
            if (PAIR)
              tc_commit_2sm(&empty_bar[s], 0x3);  // frees the stage in both CTAs when these MMAs retire
            else
              tc_commit(&empty_bar[s]);
          }
        }
        if (leader) {
          if (PAIR)
            tc_commit_2sm(&acc_full[as], 0x3);  // accumulator halves complete in both CTAs
          else
            tc_commit(&acc_full[as]);
        }
      }
      if (g.prof && leader) {
        g.prof[blockIdx.x * 8 + 0] = clock64() - m_t0;
        g.prof[blockIdx.x * 8 + 2] = m_wait_full;
        g.prof[blockIdx.x * 8 + 3] = m_wait_acc;
        g.prof[blockIdx.x * 8 + 5] = it;
        g.prof[blockIdx.x * 8 + 6] = j;
      }
    }
  } else {
    // ===== epilogue warps =====
    // warp (2 + e): TMEM lane quarter q = warp & 3 (hardware rule), column-chunk parity half = e >> 2.
    // Work unit: 32 output columns (64 accumulator columns for GEGLU) of 32 rows -> one 2 KB TMA store; two staging
    // buffers per warp so the store of chunk i overlaps the math of chunk i+1.
    constexpr int EC = 32;
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int row_in_tile = q * 32 + lane;
    uint8_t* my_stage = smem + L::EPI_OFFSET + e * EPI_BUF_BYTES;  // 2 x 2 KB
    int j = 0, chunk_ctr = 0;
    long long e_wait = 0;
    auto release_acc = [&](uint64_t* bar) {  // hand the accumulator back to the (leader's) MMA warp
      if (PAIR)
        mbar_arrive_cluster(mapa_rank(smem_u32(bar), 0));
      else
        mbar_arrive(bar);
    };
    for (int w = worker; w < num_work; w += num_workers, ++j) {
      const int wt = w / S, split = w - wt * S;
      const int m_blk = (wt / g.num_n) * (PAIR ? 2 : 1) + rank, n_blk = wt % g.num_n;
      const int tile = m_blk * g.num_n + n_blk;  // output-tile index (split-K workspace / semaphore slot)
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int as = j & 1;
      const int m = m0 + row_in_tile;
      const bool row_ok = m < g.M;
      const uint32_t lane_addr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);

      if (S > 1) {
        // ---- split-K: spill the fp32 partial tile, elect the last arriving CTA of this tile ----
        mbar_wait(&acc_full[as], (j >> 1) & 1);
        tc_fence_after();
        if (m0 >= g.M) {  // padding tile of an odd M-tile pair: nothing to reduce, just recycle the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(&acc_empty[as]);
          continue;
        }
        float4* wsp = reinterpret_cast<float4*>(g.ws) + ((size_t)(tile * S + split) * (BN / 4)) * BM + row_in_tile;
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            __stcg(wsp + (size_t)(c / 4 + i) * BM,
                   make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                               __uint_as_float(r[4 * i + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&acc_empty[as]);  // accumulator is free again
        __threadfence();
        epi_bar_sync();
        if (threadIdx.x == 64) {
          const int old = atomicAdd(g.sem + tile, 1);
          const int last = old == S - 1;
          if (last) g.sem[tile] = 0;  // self-reset: every split of this tile has arrived
          *last_flag = last;
        }
        epi_bar_sync();
        const bool is_last = *last_flag != 0;
        epi_bar_sync();  // everyone has read the flag before the next item may overwrite it
        if (!is_last) continue;
        __threadfence();
        const float4* rsp = reinterpret_cast<const float4*>(g.ws) + ((size_t)tile * S * (BN / 4)) * BM + row_in_tile;
#pragma unroll 1
        for (int c = half * EC; c < BN; c += 2 * EC, ++chunk_ctr) {
          uint8_t* buf = my_stage + (chunk_ctr & 1) * (EPI_BUF_BYTES / 2);
          uint4 res[4];
          load_residual(g, res, m, n0 + c, row_ok);
          float v[EC];
#pragma unroll
          for (int i = 0; i < EC; ++i) v[i] = 0.f;
          for (int s = 0; s < S; ++s) {  // fixed summation order -> deterministic
            const float4* p = rsp + ((size_t)s * (BN / 4) + c / 4) * BM;
#pragma unroll
            for (int i = 0; i < EC / 4; ++i) {
              const float4 f = __ldcg(p + (size_t)i * BM);
              v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
            }
          }
          if (lane == 0) bulk_wait_read<1>();  // the store issued two chunks ago has finished reading this buffer
          __syncwarp();
          epilogue_row(g, v, res, m, n0 + c, row_ok, buf + lane * 64, lane);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            store_chunk(&tmC, buf, n0 + c, m0 + q * 32, g.out3d_w);
            bulk_commit();
          }
          if (g.stats && m0 + q * 32 < g.M) chunk_stats(g, buf, m0 + q * 32, n0 + c, g.N, lane);
        }
        continue;
      }

      if constexpr (BN == 192) {
        if (g.act == 5) {
          // ---- key softmax (folded cross-attention scores): this warp owns head `half` of the tile = 96 key slots of
          // its 32 rows; the whole score row of the head sits in registers, so the accumulator is released right after
          // the three tcgen05.ld and both softmaxes are plain per-thread loops ----
          mbar_wait(&acc_full[as], (j >> 1) & 1);
          tc_fence_after();
          float sc[96];
          {
            uint32_t r0[32], r1[32], r2[32];
            tmem_ld_32x32b_x32(lane_addr + half * 96, r0);
            tmem_ld_32x32b_x32(lane_addr + half * 96 + 32, r1);
            tmem_ld_32x32b_x32(lane_addr + half * 96 + 64, r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              sc[i] = __uint_as_float(r0[i]);
              sc[32 + i] = __uint_as_float(r1[i]);
              sc[64 + i] = __uint_as_float(r2[i]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(&acc_empty[as]);
          float mt = -INFINITY, mi = -INFINITY;
#pragma unroll
          for (int i = 0; i < 80; ++i) mt = fmaxf(mt, i < g.sm_text ? sc[i] : -INFINITY);
#pragma unroll
          for (int i = 80; i < 96; ++i) mi = fmaxf(mi, i - 80 < g.sm_ip ? sc[i] : -INFINITY);
          float st = 0.f, si = 0.f;
#pragma unroll
          for (int i = 0; i < 80; ++i) {
            sc[i] = i < g.sm_text ? ex2_approx(sc[i] - mt) : 0.f;
            st += sc[i];
          }
#pragma unroll
          for (int i = 80; i < 96; ++i) {
            sc[i] = i - 80 < g.sm_ip ? ex2_approx(sc[i] - mi) : 0.f;
            si += sc[i];
          }
          const float it_ = 1.f / st;
          const float ii_ = (g.sm_ip > 0 && row_ok) ? __ldg(g.sm_ipscale + m / g.img_rows) / si : 0.f;
#pragma unroll
          for (int i = 0; i < 80; ++i) sc[i] *= it_;
#pragma unroll
          for (int i = 80; i < 96; ++i) sc[i] *= ii_;
#pragma unroll
          for (int k = 0; k < 3; ++k, ++chunk_ctr) {
            uint8_t* buf = my_stage + (chunk_ctr & 1) * (EPI_BUF_BYTES / 2);
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            uint8_t* stage_row = buf + lane * 64;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                h[jj] = __floats2half2_rn(sc[k * 32 + i * 8 + jj * 2], sc[k * 32 + i * 8 + jj * 2 + 1]);
              *reinterpret_cast<uint4*>(stage_row + ((i ^ ((lane >> 1) & 3)) << 4)) = o;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              store_chunk(&tmC, buf, n0 + half * 96 + k * 32, m0 + q * 32, 0);
              bulk_commit();
            }
          }
          continue;
        }
      }
      const int step = g.act == 2 ? 2 * EC : EC;  // accumulator columns consumed per EC output columns
      const int c_first = half * step;
      uint4 res_next[4];
      if (c_first < BN) load_residual(g, res_next, m, g.act == 2 ? (n0 + c_first) >> 1 : n0 + c_first, row_ok);
      if (g.prof && e == 0 && lane == 0) {
        const long long t0 = clock64();
        mbar_wait(&acc_full[as], (j >> 1) & 1);
        e_wait += clock64() - t0;
      } else {
        mbar_wait(&acc_full[as], (j >> 1) & 1);
      }
      tc_fence_after();
#pragma unroll 1
      for (int c = c_first; c < BN; c += 2 * step, ++chunk_ctr) {
        uint8_t* buf = my_stage + (chunk_ctr & 1) * (EPI_BUF_BYTES / 2);
        const int n_out = g.act == 2 ? (n0 + c) >> 1 : n0 + c;
        uint4 res[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) res[i] = res_next[i];
        const int c_nxt = c + 2 * step;
        if (c_nxt < BN) load_residual(g, res_next, m, g.act == 2 ? (n0 + c_nxt) >> 1 : n0 + c_nxt, row_ok);
        float v[EC];
        if (g.act != 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        } else {
          // GEGLU: accumulator columns packed per 64 as [32 value | 32 gate]; bias packed the same way
          // (ip_adapter/attention.py:359-367).  value * gelu_erf(gate) is formed in fp32 and rounded once.
          uint32_t ra[32], rg[32];
          tmem_ld_32x32b_x32(lane_addr + c, ra);
          tmem_ld_32x32b_x32(lane_addr + c + 32, rg);
          tmem_ld_wait();
          const uint4* bpa = reinterpret_cast<const uint4*>(g.bias + n0 + c);
          const uint4* bpg = bpa + 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // eight outputs at a time, their GELU polynomials evaluated side by side
            const uint4 ba = __ldg(bpa + i), bg = __ldg(bpg + i);
            const __half2* ha = reinterpret_cast<const __half2*>(&ba);
            const __half2* hg = reinterpret_cast<const __half2*>(&bg);
            float gate[8], ge[8];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float2 fg = __half22float2(hg[jj]);
              gate[2 * jj] = __uint_as_float(rg[i * 8 + 2 * jj]) + fg.x;
              gate[2 * jj + 1] = __uint_as_float(rg[i * 8 + 2 * jj + 1]) + fg.y;
            }
            gelu_erf_fast_n<8>(gate, ge);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float2 fa = __half22float2(ha[jj]);
              const int ee = i * 8 + jj * 2;
              v[ee] = (__uint_as_float(ra[ee]) + fa.x) * ge[2 * jj];
              v[ee + 1] = (__uint_as_float(ra[ee + 1]) + fa.y) * ge[2 * jj + 1];
            }
          }
        }
        if (lane == 0) bulk_wait_read<1>();  // the store issued two chunks ago has finished reading this buffer
        __syncwarp();
        epilogue_row(g, v, res, m, n_out, row_ok, buf + lane * 64, lane);
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          store_chunk(&tmC, buf, n_out, m0 + q * 32, g.out3d_w);  // rows >= M are clipped by the tensor map
          bulk_commit();
        }
        if (g.stats && m0 + q * 32 < g.M) chunk_stats(g, buf, m0 + q * 32, n_out, g.N, lane);
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&acc_empty[as]);
    }
    if (lane == 0) bulk_wait<0>();
    if (g.prof && e == 0 && lane == 0) g.prof[blockIdx.x * 8 + 4] = e_wait;
    tc_fence_before();
  }
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal / use it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else
      tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct DeviceScratch {
  float* ws = nullptr;
  int* sem = nullptr;
  int num_sms = 0;
};
// Split-K workspace (partials + per-tile arrival counters) per (device, slot).  GEMMs enqueued on different streams may
// run concurrently (the stage-2 ControlNet encoder runs beside the UNet encoder) and must not share semaphores: the
// caller selects a slot per stream with gemm_set_scratch_slot (slot 0 unless told otherwise).
static constexpr int MAX_SLOTS = 4;
static DeviceScratch g_scratch[16][MAX_SLOTS];
static thread_local int g_slot = 0;
int gemm_set_scratch_slot(int slot) {
  TG_CHECK(slot >= 0 && slot < MAX_SLOTS, "scratch slot %d out of range [0,%d)", slot, MAX_SLOTS);
  g_slot = slot;
  return 0;
}

static int get_scratch(DeviceScratch** out) {
  int dev = 0;
  TG_CUDA(cudaGetDevice(&dev));
  TG_CHECK(dev >= 0 && dev < 16, "device index %d out of range", dev);
  DeviceScratch& s = g_scratch[dev][g_slot];
  if (s.ws == nullptr) {
    // allocated once per (device, slot), outside any stream capture (every captured path is warmed up eagerly first)
    TG_CUDA(cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, dev));
    TG_CUDA(cudaMalloc(&s.ws, SPLITK_WS_BYTES));
    TG_CUDA(cudaMalloc(&s.sem, SPLITK_MAX_TILES * sizeof(int)));
    TG_CUDA(cudaMemset(s.sem, 0, SPLITK_MAX_TILES * sizeof(int)));
  }
  *out = &s;
  return 0;
}

static long long* g_prof = nullptr;
void gemm_set_prof(long long* dev_buf) { g_prof = dev_buf; }

int gemm_init() {
  DeviceScratch* s;
  return get_scratch(&s);
}

template <int BN, int STAGES, bool PAIR, bool KY3 = false>
static int launch_gemm(const CUtensorMap& a1, const CUtensorMap& a2, const __half* Wt, const CUtensorMap& c,
                       GemmArgs g, const DeviceScratch& sc, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES, PAIR, KY3>;
  auto kern = gemm_tc_kernel<BN, STAGES, PAIR, KY3>;
  TG_SMEM_ATTR_ONCE(kern, L::TOTAL);
  CUtensorMap b;
  if (KY3) {
    if (make_tmap_w_ky3(&b, Wt, g.N, g.Cin, BN / 2)) return -1;
  } else {
    const uint64_t b_rows = g.img_rows > 0 ? (uint64_t)g.N * (g.M / g.img_rows) : (uint64_t)g.N;  // per-image B: stacked blocks
    if (make_tmap_kchunks(&b, Wt, b_rows, g.K, PAIR ? BN / 2 : BN, KPS)) return -1;
  }
  const int rows_m = PAIR ? (g.num_m + 1) / 2 : g.num_m;
  const int work = rows_m * g.num_n * g.splits;
  const int max_workers = PAIR ? sc.num_sms / 2 : sc.num_sms;
  const int workers = work < max_workers ? work : max_workers;
  TG_CUDA(launch_kernel(kern, dim3(PAIR ? 2 * workers : workers), dim3(NUM_THREADS), L::TOTAL, st, PAIR ? 2 : 1, a1, a2, b, c, g));
  TG_COUNT_LAUNCH();
  return 0;
}

// CTA pairs (cta_group::2) are the default whenever a problem has >= 2 M-tiles; TG_GEMM_PAIR=0 forces single-CTA tiles
// (A/B measurements).
static bool conv_ky3_enabled() {  // TG_CONV_KY3=0: every convolution on the tap-per-k-block path (A/B measurements)
  static int use = -1;
  if (use < 0) {
    const char* e = getenv("TG_CONV_KY3");
    use = (e && e[0] == '0') ? 0 : 1;
  }
  return use == 1;
}
static bool pair_enabled() {
  static int use = -1;
  if (use < 0) {
    const char* e = getenv("TG_GEMM_PAIR");
    use = (e && e[0] == '0') ? 0 : 1;
  }
  return use == 1;
}

// Tile width / pairing / split choice: minimise   waves(work items over the workers) x (stages per item x per-stage
// cost + epilogue).  Per-stage cost model, measured on B200 with tools/probes (umma_probe, tma_probe3/5, gemm_roles):
//   * tensor pipe: a 128(x2) x BN x 16 MMA takes max(88, BN/2) cycles -> KPS*4 of them per stage;
//   * TMA: the loads issued after one barrier wait cost ~570 cycles + ~70 per extra instruction whatever their size, and
//     such groups do not overlap; a stage also cannot beat ~74 B/clk of per-SM TMA bandwidth;
//   whichever is larger.  Wider tiles are cheaper per flop, pairs halve the weight bytes per SM, wave quantisation decides
//   the rest; split-K fills the SMs when M is small and K long.
// Measured choices take precedence over the model: tools/gemm_autotune.py times every valid (tile width, split, pairing)
// per problem shape on the GPU (cold weights, warm activations) and theatergen_b200/gemm_tuned.json is loaded into this
// table when the library is opened.  Shapes without an entry fall through to the cost model below.
struct TileCfg {
  int bn, split, pair;
};
static std::unordered_map<unsigned long long, TileCfg> g_tuned;
static TileCfg g_forced = {0, 1, 0};  // bn == 0: no override (tuner only)
static unsigned long long tile_key(int M, int N, int K, int act, int conv) {
  return (unsigned long long)M | ((unsigned long long)N << 24) | ((unsigned long long)K << 40) |
         ((unsigned long long)(act == 2) << 56) | ((unsigned long long)(conv != 0) << 57);
}
static bool tile_cfg_valid(int M, int N, int K, int act, int bn, int split, int pair) {
  if (bn != 256 && bn != 192 && bn != 160 && bn != 128 && bn != 64) return false;
  if (N % bn || (act == 2 && (bn % 128 || split != 1))) return false;
  if (act == 5 && (bn != 192 || split != 1)) return false;
  const int num_m = (M + BM - 1) / BM, nk = K / BK;
  if (pair != 0 && pair != 1) return false;
  if (pair && num_m < 2) return false;
  if (split < 1 || split > MAX_SPLIT || split > nk) return false;
  if (split > 1 && ((size_t)num_m * (N / bn) * split * BM * bn * 4 > SPLITK_WS_BYTES || num_m * (N / bn) > SPLITK_MAX_TILES))
    return false;
  return true;
}
int gemm_set_tuned(int M, int N, int K, int act, int conv, int bn, int split, int pair) {
  TG_CHECK(M > 0 && M < (1 << 24) && N > 0 && N < (1 << 16) && K > 0 && K < (1 << 16), "gemm_set_tuned: shape out of range");
  TG_CHECK(tile_cfg_valid(M, N, K, act, bn, split, pair), "gemm_set_tuned: invalid tile %d/%d/%d for M=%d N=%d K=%d act=%d",
           bn, split, pair, M, N, K, act);
  g_tuned[tile_key(M, N, K, act, conv)] = TileCfg{bn, split, pair};
  return 0;
}
int gemm_clear_tuned() {
  g_tuned.clear();
  return 0;
}
int gemm_force_tile(int bn, int split, int pair) {  // bn = 0 switches the override off
  g_forced = TileCfg{bn, split, pair};
  return 0;
}

static void pick_tile(int M, int N, int K, int act, int conv, int patch_bytes, int num_sms, int* bn_out, int* split_out,
                      int* pair_out) {
  const int num_m = (M + BM - 1) / BM, nk = K / BK;
  if (act == 5) {  // key softmax: a tile is two whole heads (2 x 96 key slots), never split along K
    *bn_out = 192;
    *split_out = 1;
    *pair_out = (pair_enabled() && num_m >= 2 && nk >= 10) ? 1 : 0;
    return;
  }
  if (!g_tuned.empty()) {
    auto it = g_tuned.find(tile_key(M, N, K, act, conv));
    if (it != g_tuned.end()) {
      *bn_out = it->second.bn;
      *split_out = it->second.split;
      *pair_out = it->second.pair;
      return;
    }
  }
  const int cands[5] = {256, 192, 160, 128, 64};
  double best_cost = 1e30;
  *bn_out = 64;
  *split_out = 1;
  *pair_out = 0;
  // short-K problems (K < 640: the 320-channel linears) are paced by the epilogue, where a pair only adds a cross-CTA
  // hand-off per tile (measured: 574 vs 688 TFLOP/s on the 64x64 qkv projection)
  for (int pair = (pair_enabled() && num_m >= 2 && nk >= 10) ? 1 : 0; pair >= 0; --pair) {
    const int workers = pair ? num_sms / 2 : num_sms;
    const int rows_m = pair ? (num_m + 1) / 2 : num_m;
    for (int i = 0; i < 5; ++i) {
      const int bn = cands[i];
      if (N % bn) continue;
      if (act == 2 && bn % 128) continue;
      const int tiles = rows_m * (N / bn);
      int s = 1;
      if (act != 2 && tiles * 2 <= workers && nk >= 16) {
        s = workers / tiles;
        if (s > MAX_SPLIT) s = MAX_SPLIT;
        if (s > nk / 8) s = nk / 8;
        if (s < 1) s = 1;
        while (s > 1 && ((size_t)num_m * (N / bn) * s * BM * bn * 4 > SPLITK_WS_BYTES || num_m * (N / bn) > SPLITK_MAX_TILES)) --s;
      }
      const int waves = (tiles * s + workers - 1) / workers;
      // KY3 convolution stages (pairs only): three k-blocks share one input patch -> one A load of patch_bytes
      const bool ky3 = pair && patch_bytes > 0 && bn <= KY3_MAX_BN;
      const int kps = ky3 ? 3 : KPS;
      const double mma = kps * 4.0 * (bn / 2.0 > 88.0 ? bn / 2.0 : 88.0);
      const double stage_bytes = ky3 ? patch_bytes + 3 * 64.0 * bn : KPS * (16384.0 + (pair ? 64.0 : 128.0) * bn);
      double tma = 570.0 + 70.0 * (ky3 ? 1 : KPS);  // A loads + 1 weight load
      if (stage_bytes / 74.0 > tma) tma = stage_bytes / 74.0;
      const double per_stage = mma > tma ? mma : tma;
      const int kb_per_item = (nk + s - 1) / s;
      const double epi = 1500.0 + 8.0 * bn + (s > 1 ? 6.0 * bn * s : 0.0);
      const double cost = waves * (per_stage * ((kb_per_item + kps - 1) / kps) + epi);
      if (cost < best_cost * 0.999) {
        best_cost = cost;
        *bn_out = bn;
        *split_out = s;
        *pair_out = pair;
      }
    }
  }
}

static int dispatch(const CUtensorMap& a1, const CUtensorMap& a2, GemmArgs g, const __half* Wt, __half* out, int ldc,
                    cudaStream_t st) {
  if (g.act == 2) TG_CHECK(g.N % 128 == 0 && g.bias != nullptr, "geglu epilogue needs N%%128==0 and a bias");
  if (g.stats)
    TG_CHECK(g.act != 2 && g.stats_hw > 0 && g.stats_hw % 32 == 0 && g.M % g.stats_hw == 0,
             "GroupNorm statistics epilogue: rows per image %d must be a multiple of 32 dividing M=%d (no GEGLU)", g.stats_hw, g.M);
  if (g.img_rows > 0)
    TG_CHECK(g.mode == 0 && g.M % g.img_rows == 0 && g.img_rows % (2 * BM) == 0,
             "per-image B operand: rows per image %d must be a multiple of %d dividing M=%d (linear mode)", g.img_rows, 2 * BM, g.M);
  if (g.act == 5)
    TG_CHECK(g.N % 192 == 0 && g.img_rows > 0 && g.sm_text > 0 && g.sm_text <= 80 && g.sm_ip >= 0 && g.sm_ip <= 16 &&
                 (g.sm_ip == 0 || g.sm_ipscale != nullptr) && !g.bias && !g.rowbias && !g.residual && !g.stats,
             "key-softmax epilogue: N=%d must be heads x 96, text keys %d <= 80, image keys %d <= 16, no other epilogue terms",
             g.N, g.sm_text, g.sm_ip);
  DeviceScratch* sc;
  if (get_scratch(&sc)) return -1;
  const int n_out = g.act == 2 ? g.N / 2 : g.N;
  int bn, splits, pair;
  if (g_forced.bn) {
    TG_CHECK(tile_cfg_valid(g.M, g.N, g.K, g.act, g_forced.bn, g_forced.split, g_forced.pair),
             "forced tile %d/%d/%d is not valid for M=%d N=%d K=%d act=%d", g_forced.bn, g_forced.split, g_forced.pair, g.M,
             g.N, g.K, g.act);
    bn = g_forced.bn, splits = g_forced.split, pair = g_forced.pair;
  } else {
    pick_tile(g.M, g.N, g.K, g.act, g.mode, g.patch_bytes, sc->num_sms, &bn, &splits, &pair);
  }
  // KY3 stages for pair-tiled convolutions whose tile lies inside one image (conv3x3_nhwc prepared the patch map as a2)
  const bool ky3 = pair && g.patch_bytes > 0 && bn <= KY3_MAX_BN && splits <= g.K / (3 * BK);
  CUtensorMap c;  // epilogue chunks: 32 rows x 32 columns (64-byte rows, 64B swizzle)
  if (g.out3d_w > 0) {
    const uint32_t bw = g.out3d_w < 32 ? g.out3d_w : 32;
    TG_CHECK(32 % bw == 0 && g.out3d_w % bw == 0 && g.M % g.out3d_w == 0 && ldc == n_out, "phase output: W=%d M=%d ldc=%d", g.out3d_w,
             g.M, ldc);
    if (make_tmap_phase_out(&c, out, g.M / g.out3d_w, g.out3d_w, n_out, bw, 32 / bw)) return -1;
  } else if (make_tmap_2d(&c, out, g.M, n_out, ldc, 32, 32)) {
    return -1;
  }
  g.num_m = (g.M + BM - 1) / BM;
  g.num_n = g.N / bn;
  g.splits = splits;
  g.ws = sc->ws;
  g.sem = sc->sem;
  g.prof = g_prof;
  // stages of KPS k-blocks; a CTA of a pair stages half of the weight tile -> deeper rings in the same shared memory
  if (ky3) {  // stage = patch (<= 32 KB) + 3 weight half-chunks
    switch (bn) {
      case 160: return launch_gemm<160, 3, true, true>(a1, a2, Wt, c, g, *sc, st);   // 62 KB
      case 128: return launch_gemm<128, 3, true, true>(a1, a2, Wt, c, g, *sc, st);   // 56 KB
      default: return launch_gemm<64, 4, true, true>(a1, a2, Wt, c, g, *sc, st);     // 44 KB
    }
  }
  if (pair) {
    switch (bn) {
      case 256: return launch_gemm<256, 3, true>(a1, a2, Wt, c, g, *sc, st);   // 64 KB stages
      case 192: return launch_gemm<192, 3, true>(a1, a2, Wt, c, g, *sc, st);   // 56 KB
      case 160: return launch_gemm<160, 3, true>(a1, a2, Wt, c, g, *sc, st);   // 52 KB
      case 128: return launch_gemm<128, 4, true>(a1, a2, Wt, c, g, *sc, st);   // 48 KB
      default: return launch_gemm<64, 4, true>(a1, a2, Wt, c, g, *sc, st);     // 40 KB
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256, 2, false>(a1, a2, Wt, c, g, *sc, st);    // 96 KB stages
    case 192: return launch_gemm<192, 2, false>(a1, a2, Wt, c, g, *sc, st);    // 80 KB
    case 160: return launch_gemm<160, 2, false>(a1, a2, Wt, c, g, *sc, st);    // 72 KB
    case 128: return launch_gemm<128, 3, false>(a1, a2, Wt, c, g, *sc, st);    // 64 KB
    default: return launch_gemm<64, 4, false>(a1, a2, Wt, c, g, *sc, st);      // 48 KB
  }
}

static void fill_args(GemmArgs& g, const EpilogueSpec& e) {
  g.bias = e.bias;
  g.rowbias = e.rowbias;
  g.rows_per_image = e.rows_per_image > 0 ? e.rows_per_image : 1;
  g.rowbias_ld = e.rowbias_ld;
  g.residual = e.residual;
  g.ldr = e.ldr;
  g.act = e.act;
  g.out_scale = e.out_scale;
  g.stats = reinterpret_cast<unsigned long long*>(e.stats);
  g.stats_hw = e.stats_hw;
  g.img_rows = e.b_img_rows;
  g.sm_text = e.sm_text;
  g.sm_ip = e.sm_ip;
  g.sm_ipscale = e.sm_ipscale;
}

int gemm_linear(const __half* A1, int lda1, int K1, const __half* A2, int lda2, const __half* Wt, int M, int N,
                int K, const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st) {
  TG_CHECK(M > 0 && N % 64 == 0 && K % 64 == 0 && K1 % 64 == 0 && K1 <= K, "gemm: bad dims M=%d N=%d K=%d K1=%d", M,
           N, K, K1);
  TG_CHECK(A2 != nullptr || K1 == K, "gemm: K1<K needs a second A source");
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.K1 = K1; g.mode = 0;
  fill_args(g, e);
  CUtensorMap a1, a2;
  if (make_tmap_2d(&a1, A1, M, K1, lda1, BM)) return -1;
  if (K1 < K) {
    if (make_tmap_2d(&a2, A2, M, K - K1, lda2, BM)) return -1;
  } else {
    a2 = a1;
  }
  return dispatch(a1, a2, g, Wt, out, ldc, st);
}

int conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, const EpilogueSpec& e,
                 __half* out, int ldc, cudaStream_t st) {
  TG_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "conv3x3: Cin/Cout must be multiples of 64 (got %d/%d)", Cin, Cout);
  TG_CHECK((W <= 128 && 128 % W == 0) || W % 128 == 0, "conv3x3: W must divide 128 or be a multiple of it (got %d)", W);
  const int box_w = W < 128 ? W : 128;
  int box_h = 128 / box_w, box_b = 1;
  if (box_h > H) {
    box_b = box_h / H;
    box_h = H;
  }
  TG_CHECK(H % box_h == 0 && B % box_b == 0 && box_b * box_h * box_w == 128, "conv3x3: tile %dx%dx%d does not cover 128",
           box_b, box_h, box_w);
  GemmArgs g{};
  g.M = B * H * W; g.N = Cout; g.K = 9 * Cin; g.K1 = g.K; g.mode = 1;
  g.H = H; g.W = W; g.Cin = Cin; g.box_h = box_h; g.box_b = box_b;
  g.taps_w = 3; g.off_y = g.off_x = -1; g.in_stride = 1;
  fill_args(g, e);
  CUtensorMap a1, a2;
  if (make_tmap_nhwc(&a1, X, B, H, W, Cin, box_b, box_h, box_w)) return -1;
  a2 = a1;
  // KY3: the tile is box_h whole image rows of one image -> the three ky taps of a (kx, channel block) read one patch of
  // box_h + 2 rows (TMA zero-fills the rows / columns outside the image)
  if (conv_ky3_enabled() && box_b == 1 && box_w == W && W % 8 == 0 && (box_h + 2) * W * 128 <= KY3_PATCH_MAX) {
    if (make_tmap_nhwc(&a2, X, B, H, W, Cin, 1, box_h + 2, W)) return -1;
    g.patch_bytes = (box_h + 2) * W * 128;
    g.ky_stride = W * 128;
  }
  return dispatch(a1, a2, g, Wt, out, ldc, st);
}

// tile geometry of an implicit-GEMM convolution whose OUTPUT image is Ho x Wo: a 128-row tile = box_b images x box_h rows
// x box_w pixels
static int conv_boxes(int B, int Ho, int Wo, int* box_b, int* box_h, int* box_w) {
  TG_CHECK((Wo <= 128 && 128 % Wo == 0) || Wo % 128 == 0, "conv: output width must divide 128 or be a multiple of it (got %d)", Wo);
  *box_w = Wo < 128 ? Wo : 128;
  *box_h = 128 / *box_w;
  *box_b = 1;
  if (*box_h > Ho) {
    *box_b = *box_h / Ho;
    *box_h = Ho;
  }
  TG_CHECK(Ho % *box_h == 0 && B % *box_b == 0 && *box_b * *box_h * *box_w == 128, "conv: tile %dx%dx%d does not cover 128",
           *box_b, *box_h, *box_w);
  return 0;
}

// 3x3 convolution with stride 2 (diffusers Downsample2D): the TMA gathers every other input pixel (traversal stride 2 on the
// x / y dimensions of the NHWC map), so the operand never exists as an im2col buffer.  pad_lo = 1: padding 1 (UNet /
// ControlNet); pad_lo = 0: the VAE encoder's one-sided (0, 1) padding.
int conv3x3_s2_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt, int Cout, int pad_lo,
                    const EpilogueSpec& e, __half* out, int ldc, cudaStream_t st) {
  TG_CHECK(Cin % 64 == 0 && Cout % 64 == 0 && H % 2 == 0 && W % 2 == 0, "conv3x3 s2: Cin/Cout multiples of 64, even H/W (got %d/%d %dx%d)",
           Cin, Cout, H, W);
  const int Ho = H / 2, Wo = W / 2;
  int box_b, box_h, box_w;
  if (conv_boxes(B, Ho, Wo, &box_b, &box_h, &box_w)) return -1;
  GemmArgs g{};
  g.M = B * Ho * Wo; g.N = Cout; g.K = 9 * Cin; g.K1 = g.K; g.mode = 1;
  g.H = Ho; g.W = Wo; g.Cin = Cin; g.box_h = box_h; g.box_b = box_b;
  g.taps_w = 3; g.off_y = g.off_x = -pad_lo; g.in_stride = 2;
  fill_args(g, e);
  CUtensorMap a1;
  if (make_tmap_nhwc(&a1, X, B, H, W, Cin, box_b, box_h, box_w, 2)) return -1;
  return dispatch(a1, a1, g, Wt, out, ldc, st);
}

// nearest x2 upsample -> 3x3 convolution (diffusers Upsample2D) WITHOUT the upsampled tensor and with 2.25x fewer MACs:
// output pixel (2i + py, 2j + px) only sees a 2 x 2 window of the low-resolution image, rows {i + py - 1, i + py} and
// columns {j + px - 1, j + px}, with the 3x3 taps that land on the same input pixel summed (Wt4[py * 2 + px] =
// [Cout, 2, 2, Cin], packed once by the host).  Four implicit GEMMs on the low-resolution grid, each storing its phase of
// the [B, 2H, 2W, Cout] image through a 3D tensor map.
int upsample2x_conv3x3_nhwc(const __half* X, int B, int H, int W, int Cin, const __half* Wt4, int Cout, const EpilogueSpec& e,
                            __half* out, cudaStream_t st) {
  TG_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "upsample conv: Cin/Cout must be multiples of 64 (got %d/%d)", Cin, Cout);
  TG_CHECK(e.residual == nullptr && e.rowbias == nullptr, "upsample conv: residual / row bias epilogues are not wired for phase outputs");
  int box_b, box_h, box_w;
  if (conv_boxes(B, H, W, &box_b, &box_h, &box_w)) return -1;
  CUtensorMap a1;
  if (make_tmap_nhwc(&a1, X, B, H, W, Cin, box_b, box_h, box_w)) return -1;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      GemmArgs g{};
      g.M = B * H * W; g.N = Cout; g.K = 4 * Cin; g.K1 = g.K; g.mode = 1;
      g.H = H; g.W = W; g.Cin = Cin; g.box_h = box_h; g.box_b = box_b;
      g.taps_w = 2; g.off_y = py - 1; g.off_x = px - 1; g.in_stride = 1;
      g.out3d_w = W;
      fill_args(g, e);
      const __half* wt = Wt4 + (size_t)(py * 2 + px) * Cout * 4 * Cin;
      __half* o = out + ((size_t)py * 2 * W + px) * Cout;
      if (dispatch(a1, a1, g, wt, o, Cout, st)) return -1;
    }
  return 0;
}

}  // namespace tg
