// tcgen05 flash self-attention for head dims 8..120 (SD-1.5's 64x64 level: N = 4096, d = 40 -- 90 % of the
// self-attention time -- and the 32x32 level: N = 1024, d = 80).  Q/K/V arrive head-padded to CH x 64 columns (CH = 1 for
// d <= 56, 2 for d <= 120; zero weight rows in the fused QKV projection), so every K/V tile is CH canonical 128-byte-row
// SWIZZLE_128B chunks that TMA writes and tcgen05.mma reads.  CH = 1 runs two CTAs per SM (256 TMEM columns each),
// CH = 2 one (512 columns, 192 KB of K/V stages).
//
// Per CTA: one (batch, head, 128-query tile); keys stream through a 3-stage TMA ring, 128 keys (two 64-key score
// tiles) per stage -- a group of TMA loads issued after a barrier wait costs ~570 cycles whatever its size (measured,
// tools/probes), so a stage carries two tiles.
//   warp 4      TMA producer (K_2s..2s+1 | V_2s..2s+1 per stage: two 16 KB boxes)
//   warp 5      TMEM allocator + MMA issuer (warp-uniform loop, one elected lane issues), BOTH products with the A operand in TENSOR MEMORY (.ts form):
//                   S_j = Q K_j^T      (A = Q in TMEM, M=128, N=64, K=ceil(d/16)*16)               -> TMEM S[j&1]
//                   O  += P_j V_j      (A = P_j in TMEM, M=128, N=64; V is the MN-major B operand)  -> TMEM O
//               Measured on B200: a 128 x N x 16 MMA with N <= 64 takes 88 cycles with A in shared memory (bound by the
//               4 KB A read) but 45 cycles with A in tensor memory.
//   warps 0..3  one query row per thread (row == TMEM lane): Q row -> TMEM once; per tile S_j -> registers, softmax
//               numerators, P_j -> fp16 pairs -> tcgen05.st INTO THE COLUMNS OF S_j (P overlays S: the tensor pipe executes
//               in issue order, so S_{j+2} cannot overwrite P_j before P_j V_j has read it -- no extra barrier).
//               Measured alternatives that did NOT help (B = 10, N = 4096, d = 40: 504 us as shipped): 8 softmax warps
//               splitting a row's 64 columns with the row max exchanged through shared memory (530 us); S loaded in four
//               16-column chunks with speculative exponentials against the stale reference so a chunk's TMEM load
//               overlaps the previous chunk's MUFU work (531 us); 1/6, 1/4, 1/2 of the exponentials on the FMA pipes by a
//               degree-4 polynomial (531 / 563 / 629 us).  Every added instruction costs time and removing MUFU work does
//               not pay: the softmax warps are bound by their own instruction stream (about 275 instructions per row
//               and tile at two warps per scheduler), not by the MUFU or TMEM ports.  A packed-half2 softmax (HMNMX2 /
//               HFMA2 / h2exp2) is no shorter in SASS (h2exp2 unpacks to two fp32 MUFU.EX2 and repacks: 540 us) and
//               fails the 2e-3 parity tolerance.
// What bounds this kernel is TMEM->register bandwidth and the MUFU pipe, so the design minimises both:
//   * O is accumulated IN TMEM across all KV tiles (accumulate flag) and is read exactly once, at the end.
//     The running max is only a reference exponent: it is raised -- and O rescaled in TMEM (tcgen05.ld / st)
//     -- only when a row's new maximum exceeds it by more than 2^8 (lazy rescaling; P stays <= 256 in fp16),
//     which after the first tiles practically never happens.  Per tile only S (128 x 64 fp32) crosses TMEM->RF.
//   * the softmax denominator is not summed by CUDA cores: column d of each V tile is set to 1.0 in shared
//     memory, so the tensor core accumulates sum_k P[k] in O[:, d] with exactly the same rescaling as O.
// S is double-buffered in TMEM so the tensor core computes S_{j+1} while the softmax warps work on S_j; two CTAs
// fit per SM (97 KB smem, 256 TMEM columns each).  The N x N probabilities never leave the SM
// (reference: AttnProcessor.__call__, ip_adapter/attention_processor.py:343-356, materialises them in HBM).
#include "common.cuh"
#include "ops.h"
#include "tmem.cuh"

namespace tg {

static constexpr int ATQ = 128;   // queries per CTA
static constexpr int ATK = 64;    // keys per score tile
static constexpr int ATD = 64;    // columns of one head chunk
static constexpr int AT_TPS = 2;  // score tiles per pipeline stage
static constexpr int AT_STAGES = 3;
static constexpr int AT_TILE_BYTES = ATK * ATD * 2;             // 8 KB: 64 keys of one chunk
static constexpr int AT_CHUNK_BYTES = AT_TPS * AT_TILE_BYTES;   // 16 KB: 128 keys of one chunk of K (or V)
static constexpr int AT_SOFTMAX_WARPS = 4;
static constexpr int AT_THREADS = 64 + AT_SOFTMAX_WARPS * 32;
static constexpr int AT_NUM_BARS = 2 * AT_STAGES + 2 + 2 + 2 + 2;   // kv full/empty, s_full, p_full, o_full, s_free
template <int CH>
struct AttnCfg {
  static constexpr int KV_BYTES = CH * AT_CHUNK_BYTES;          // 128 keys of K (or V), CH chunks
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;              // K chunks | V chunks
  static constexpr int STAGES = CH == 3 ? 2 : AT_STAGES;        // CH = 3: 96 KB per stage
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int SMEM = BAR_OFFSET + AT_NUM_BARS * 8 + 16 + 1024;
  // tensor memory columns, VAR 0: S0 | S1 (P_j overlays the first 32 columns of S_j) | O (64 per chunk) | Q (8 per K-step)
  //                        VAR 2: S0 | S1 | P (32) | O (64 per chunk) | Q (8 per K-step; 24 for d = 40 -> 248 of 256)
  static constexpr uint32_t TS = 0, TO = 128, TQ = 128 + 64 * CH, TCOLS = CH == 1 ? 256 : 512;
  static constexpr uint32_t TP2 = 128, TO2 = 160, TQ2 = 160 + 64 * CH;
};

// One half (32 scores) of a score row -> 32 fp16 softmax numerators packed into 16 tensor-memory columns.
// fp32 exp2 (one MUFU each -- ex2.f16x2 is split into two MUFU ops on sm_100 anyway, and with a stale reference exponent
// the dominant terms sit at |x| up to 8 where an fp16 argument would cost 0.3 %)
__device__ __forceinline__ void softmax_half(const uint32_t (&s)[32], uint32_t (&pk)[16], float scale_log2, float msc) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = __uint_as_float(s[2 * i]) * scale_log2 - msc;
    const float x1 = __uint_as_float(s[2 * i + 1]) * scale_log2 - msc;
    float p0, p1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
    const __half2 hp = __floats2half2_rn(p0, p1);
    pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
  }
}

// VAR selects the softmax warps' schedule:
//   0  load S_j (two x32 tcgen05.ld), wait, max, 64 exponentials, one x32 tcgen05.st           (round-1 kernel)
//   2  decoupled: P_j has its own tensor-memory columns (single buffer) instead of overlaying S_j, so the issuer may
//      overwrite S_j with S_{j+2} as soon as the four softmax warps have loaded it (s_free) -- see the issuer's comment.
//      (Tried and measured slower, removed: prefetching the halves of S_{j+1} with tcgen05.ld while the exponentials of
//      tile j run -- tcgen05.ld costs ~77 cycles, there is nothing to hide: 509 vs 490 us.)
// PROF: cycle stamps per phase into `prof` (measurement hook, tg_attn_set_prof).
// Warp roles.  The softmax warps take the LOW warp ids, the TMA producer and the MMA issuer the HIGH ones: the warp
// scheduler of an SM sub-partition favours the higher warp id, and the issuer -- a short, latency-critical instruction
// stream that shares its sub-partition with a softmax warp of each resident CTA -- must not wait for issue slots behind
// 64-instruction FFMA / MUFU bursts (measured with the tg_attn_set_prof trace: with the issuer as warp 1 the four P.V MMAs
// + commit took ~300-350 cycles to issue and S_{j+2} became visible ~1400 cycles after P_j was complete).
static constexpr int AT_WARP_TMA = AT_SOFTMAX_WARPS, AT_WARP_MMA = AT_SOFTMAX_WARPS + 1;

template <int CH, int VAR, bool PROF>
__global__ void __launch_bounds__(AT_THREADS, CH == 1 ? 2 : 1)
attn_tc_kernel(const __half* __restrict__ Q, int ldq, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, __half* __restrict__ O, int ldo, int N, int d,
               float scale_log2, long long* __restrict__ prof, int hstride) {
  // hstride: columns between consecutive heads in Q / K / V (CH * 64 for the zero-padded layouts, d for packed heads)
  using Cfg = AttnCfg<CH>;
  constexpr int AT_KV_BYTES = Cfg::KV_BYTES, AT_STAGE_BYTES = Cfg::STAGE_BYTES, AT_STAGES = Cfg::STAGES;
  constexpr uint32_t AT_TS = Cfg::TS, AT_TO = VAR == 2 ? Cfg::TO2 : Cfg::TO, AT_TQ = VAR == 2 ? Cfg::TQ2 : Cfg::TQ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = kv_full + AT_STAGES;
  uint64_t* s_full = kv_empty + AT_STAGES;  // [2] S_j complete in TMEM
  uint64_t* p_full = s_full + 2;            // [2] P_j (and the ones column of V_j) written by the 4 softmax warps
  uint64_t* o_full = p_full + 2;            // [2] P_j V_j retired (lazy rescale / final read of O)
  uint64_t* s_free = o_full + 2;            // [2] VAR 2: the softmax warps hold S_j in registers -> S_{j+2} may overwrite it
  // VAR 2 uses p_full[0] = P_j stored, p_full[1] = Q rows stored (once), o_full[0] = P_j V_j retired (one phase per tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AT_NUM_BARS);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches non-divergent and the
  // MMA issuer's operands in uniform registers (otherwise every tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST loop:
  // measured ~115-170 cycles per MMA issue, which made the single issuing thread the critical path of the kernel)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATQ;
  const int T = N / ATK;             // score tiles
  const int NS = T / AT_TPS;         // pipeline stages' worth of keys (N is a multiple of 128)
  const int row_base = b * N;
  const int ksteps = (d + 15) / 16;  // K extent of Q K^T in MMA steps (Q columns >= d are zero)

  if (warp == AT_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < AT_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], AT_SOFTMAX_WARPS);
      mbar_init(&o_full[s], 1);
      mbar_init(&s_free[s], AT_SOFTMAX_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == AT_WARP_MMA) {
    tmem_alloc(tmem_slot, Cfg::TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  grid_dep_launch();  // PDL: the prologue above overlapped the previous kernel's tail
  grid_dep_wait();
  // PROF trace of CTA (0,0,0): per tile {softmax warp 2 arrives on p_full, MMA thread past the p_full wait, MMA thread has
  // issued + committed S_{j+1}, softmax warp 2 past the s_full wait, MMA thread past the kv_full wait} -> prof + records
  long long* trace = nullptr;
  if constexpr (PROF) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
      trace = prof + (size_t)gridDim.x * gridDim.y * gridDim.z * AT_SOFTMAX_WARPS * 8;
  }
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tS = tmem_base + AT_TS, tO = tmem_base + AT_TO, tQ = tmem_base + AT_TQ, tP = tmem_base + Cfg::TP2;

  if (warp == AT_WARP_TMA) {
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) {
        const int s = i % AT_STAGES;
        mbar_wait(&kv_empty[s], ((i / AT_STAGES) & 1) ^ 1);
        uint8_t* dst = sKV + s * AT_STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], AT_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          tma_load_2d(dst + c * AT_CHUNK_BYTES, &tmK, &kv_full[s], h * hstride + c * ATD, row_base + i * AT_TPS * ATK);
          tma_load_2d(dst + AT_KV_BYTES + c * AT_CHUNK_BYTES, &tmV, &kv_full[s], h * hstride + c * ATD, row_base + i * AT_TPS * ATK);
        }
      }
    }
  } else if (warp == AT_WARP_MMA) {
    // The whole warp walks the loop (waits included) so that every address stays warp-uniform; one elected lane issues.
    const bool leader = elect_one();
    {
      constexpr uint32_t idesc_s = make_idesc_f16(ATQ, ATK);                     // B (= K tile) K-major
      // P.V: one MMA per head chunk, N = 64 for full chunks and (d + 1 (the ones column) - 64 c) rounded up to 16 for the
      // last one; B (= V chunk) is MN-major
      uint32_t idesc_o[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int nc = (c + 1 < CH) ? 64 : ((d + 1 - 64 * c + 15) / 16) * 16;
        idesc_o[c] = make_idesc_f16(ATQ, nc) | (1u << 16);
      }
      const uint32_t kv_base = smem_u32(sKV);
      auto issue_s = [&](int j) {
        const int i = j / AT_TPS, s = i % AT_STAGES;
        if (j % AT_TPS == 0) mbar_wait(&kv_full[s], (i / AT_STAGES) & 1);
        if (PROF && trace && leader) trace[5 * j + 4] = clock64();
        tc_fence_after();
        // S buffer j&1 held S_{j-2} / P_{j-2}: the softmax warps are done with it (p_full[j-2] was waited before
        // P_{j-2} V_{j-2} was issued) and the tensor pipe runs in issue order, so P_{j-2} V_{j-2} reads P before this
        // MMA overwrites it
        const uint32_t k_addr = kv_base + s * AT_STAGE_BYTES + (j % AT_TPS) * AT_TILE_BYTES;
        if (leader) {
          for (int k = 0; k < ksteps; ++k)  // K-step k: chunk k / 4, 32 bytes per step inside the chunk's 128-byte rows
            umma_f16_ts(tS + (j & 1) * 64, tQ + k * 8, make_smem_desc_sw128(k_addr + (k >> 2) * AT_CHUNK_BYTES + (k & 3) * 32),
                        idesc_s, k > 0 ? 1u : 0u);
          tc_commit(&s_full[j & 1]);
          if (PROF && trace) trace[5 * j + 2] = clock64();
        }
      };
      if constexpr (VAR == 2) {
        // Decoupled schedule: S_{j+2} is issued as soon as the softmax warps have READ S_j (s_free), not after P_j V_j --
        // P lives in its own columns.  The score tiles then run two tiles ahead of the softmax and the chain
        // "P_j complete -> P_j V_j -> S_{j+2} -> visible" (measured ~1400 cycles) is off the critical path.
        mbar_wait(&p_full[1], 0);  // Q rows in tensor memory
        issue_s(0);
        if (T > 1) issue_s(1);
        for (int j = 0; j < T; ++j) {
          if (j + 2 < T) {
            mbar_wait(&s_free[j & 1], (j >> 1) & 1);
            issue_s(j + 2);
          }
          const int i = j / AT_TPS, s = i % AT_STAGES;
          mbar_wait(&p_full[0], j & 1);
          if (PROF && trace && leader) trace[5 * j + 1] = clock64();
          tc_fence_after();
          const uint32_t v_addr = kv_base + s * AT_STAGE_BYTES + AT_KV_BYTES + (j % AT_TPS) * AT_TILE_BYTES;
          if (leader) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
              for (int k = 0; k < ATK / 16; ++k)
                umma_f16_ts(tO + c * 64, tP + k * 8, make_smem_desc_sw128_mn(v_addr + c * AT_CHUNK_BYTES + k * 2048),
                            idesc_o[c], (j > 0 || k > 0) ? 1u : 0u);
            tc_commit(&o_full[0]);
            if (j % AT_TPS == AT_TPS - 1) tc_commit(&kv_empty[s]);
          }
        }
      } else {
      mbar_wait(&p_full[0], 0);  // phase 0 of p_full[0]: the Q rows are in tensor memory (see the softmax warps)
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int i = j / AT_TPS, s = i % AT_STAGES;
        mbar_wait(&p_full[j & 1], ((j >> 1) & 1) ^ (j & 1 ? 0 : 1));  // buffer 0 spent one extra phase on Q
        if (PROF && trace && leader) trace[5 * j + 1] = clock64();
        tc_fence_after();
        const uint32_t v_addr = kv_base + s * AT_STAGE_BYTES + AT_KV_BYTES + (j % AT_TPS) * AT_TILE_BYTES;
        if (leader) {
#pragma unroll
          for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int k = 0; k < ATK / 16; ++k)
              umma_f16_ts(tO + c * 64, tS + (j & 1) * 64 + k * 8, make_smem_desc_sw128_mn(v_addr + c * AT_CHUNK_BYTES + k * 2048),
                          idesc_o[c], (j > 0 || k > 0) ? 1u : 0u);  // O accumulates in TMEM over all tiles
          tc_commit(&o_full[j & 1]);
          if (j % AT_TPS == AT_TPS - 1) tc_commit(&kv_empty[s]);
        }
      }
      }  // VAR
    }
  } else {
    const int q = warp;                   // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int r = q * 32 + lane;          // query row inside the tile == TMEM lane
    const int tid_s = threadIdx.x;        // 0..127 among the softmax threads
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int nacc = ((d + 1 + 15) / 16) * 16;  // accumulator columns in use: d outputs + the denominator
    float m_ref = -INFINITY;  // reference exponent (a stale running max is fine: P <= 2^8)
    // position of element (key k, column d) inside the swizzled V tile of head chunk d / 64: 16-byte chunk
    // ((d % 64) / 8) ^ (k & 7) of row k
    const int dl = d & 63;
    const int ones_off = (d >> 6) * AT_CHUNK_BYTES + (tid_s & 63) * 128 + ((((dl >> 3) ^ (tid_s & 7)) << 4) | ((dl & 7) << 1));

    // Q row -> tensor memory (A operand of every S MMA): 16 fp16 = 8 packed columns per MMA K-step
    {
      const uint4* qrow = reinterpret_cast<const uint4*>(Q + (size_t)(row_base + q0 + r) * ldq + h * hstride);
      for (int k = 0; k < ksteps; ++k) {
        uint4 v[2];
        v[0] = __ldg(qrow + 2 * k);
        // packed heads: the columns from d up to the next multiple of 16 belong to the next head (or lie past the row) --
        // never read, zero in the operand, so whatever the K tile holds there contributes nothing
        v[1] = (2 * k + 1) * 8 < d ? __ldg(qrow + 2 * k + 1) : make_uint4(0u, 0u, 0u, 0u);
        tmem_st_32x32b_x8(tQ + k * 8 + lane_off, reinterpret_cast<const uint32_t*>(v));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(VAR == 2 ? &p_full[1] : &p_full[0]);
    }

    long long pf[6] = {0, 0, 0, 0, 0, 0};  // PROF: wait S | wait ld | max (+rescale) | exps + st issue | wait st + arrive | tiles
    auto stamp = [&](long long& t_prev, int slot) {
      if constexpr (PROF) {
        const long long now = clock64();
        pf[slot] += now - t_prev;
        t_prev = now;
      }
    };
    long long tprev = 0;
    if constexpr (PROF) tprev = clock64();
    if constexpr (VAR == 2) {
      for (int j = 0; j < T; ++j) {
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        if (PROF && trace && warp == 0 && lane == 0) trace[5 * j + 3] = clock64();
        tc_fence_after();
        stamp(tprev, 0);
        uint32_t sr[64];
        {
          uint32_t t0[32], t1[32];
          tmem_ld_32x32b_x32(tS + (j & 1) * 64 + lane_off, t0);
          tmem_ld_32x32b_x32(tS + (j & 1) * 64 + 32 + lane_off, t1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { sr[i] = t0[i]; sr[32 + i] = t1[i]; }
        }
        // S_j is in registers: its tensor-memory buffer may take S_{j+2}
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[j & 1]);
        stamp(tprev, 1);
        if (tid_s < 64)  // the ones column of this tile's V (its TMA landed before S_j was issued)
          *reinterpret_cast<__half*>(sKV + ((j / AT_TPS) % AT_STAGES) * AT_STAGE_BYTES + AT_KV_BYTES +
                                     (j % AT_TPS) * AT_TILE_BYTES + ones_off) = __float2half(1.f);
        float m4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m4[i] = __uint_as_float(sr[i]);
#pragma unroll
        for (int i = 4; i < 64; i += 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (i + 4 + c < 64)
              m4[c] = fmax3f(m4[c], __uint_as_float(sr[i + c]), __uint_as_float(sr[i + 4 + c]));
            else
              m4[c] = fmaxf(m4[c], __uint_as_float(sr[i + c]));
          }
        }
        const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
        bool o_waited = false;  // P_{j-1} V_{j-1} known retired
        if (j == 0) {
          m_ref = mx;  // O is still empty (the first P.V overwrites it)
        } else {
          const bool need = (mx - m_ref) * scale_log2 > 8.f;
          if (__any_sync(0xffffffffu, need)) {
            // lazy rescale of this warp's 32 O rows in TMEM: only legal while no P.V is in flight -> P.V_{j-1} retired
            // (P_j V_j cannot start before this warp has arrived on p_full)
            mbar_wait(&o_full[0], (j - 1) & 1);
            o_waited = true;
            tc_fence_after();
            const float f = need ? exp2f((m_ref - mx) * scale_log2) : 1.f;
            if (need) m_ref = mx;
            for (int c = 0; c < nacc; c += 16) {
              uint32_t t[16];
              tmem_ld_32x32b_x16(tO + c + lane_off, t);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
              tmem_st_32x32b_x16(tO + c + lane_off, t);
            }
            tmem_st_wait();
          }
        }
        const float msc = m_ref * scale_log2;
        stamp(tprev, 2);
        uint32_t pk[32];
        {
          uint32_t lo[16], hi[16];
          softmax_half(*reinterpret_cast<const uint32_t(*)[32]>(&sr[0]), lo, scale_log2, msc);
          softmax_half(*reinterpret_cast<const uint32_t(*)[32]>(&sr[32]), hi, scale_log2, msc);
#pragma unroll
          for (int i = 0; i < 16; ++i) { pk[i] = lo[i]; pk[16 + i] = hi[i]; }
        }
        // the single P buffer still feeds P_{j-1} V_{j-1} until that MMA group has retired (issued a whole softmax
        // period ago: practically always complete)
        if (j > 0 && !o_waited) {
          mbar_wait(&o_full[0], (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32b_x32(tP + lane_off, pk);
        stamp(tprev, 3);
        tmem_st_wait();
        fence_proxy_async();  // the ones column (generic-proxy smem store) -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[0]);
        if (PROF && trace && warp == 0 && lane == 0) trace[5 * j + 0] = clock64();
        stamp(tprev, 4);
      }
    } else {
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      if (PROF && trace && warp == 0 && lane == 0) trace[5 * j + 3] = clock64();
      tc_fence_after();
      stamp(tprev, 0);
      uint32_t sr[64];
      {
        uint32_t t0[32], t1[32];
        tmem_ld_32x32b_x32(tS + (j & 1) * 64 + lane_off, t0);
        tmem_ld_32x32b_x32(tS + (j & 1) * 64 + 32 + lane_off, t1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { sr[i] = t0[i]; sr[32 + i] = t1[i]; }
      }
      stamp(tprev, 1);

      // the ones column of this tile's V (its TMA landed before S_j was issued); 64 threads, one store each
      if (tid_s < 64)
        *reinterpret_cast<__half*>(sKV + ((j / AT_TPS) % AT_STAGES) * AT_STAGE_BYTES + AT_KV_BYTES +
                                   (j % AT_TPS) * AT_TILE_BYTES + ones_off) = __float2half(1.f);

      // row max: four independent chains instead of one 64-deep dependent chain, two scores per instruction
      // (3-input max, FMNMX3 on sm_100: 32 instead of 63 instructions per row and tile; exact, so results do not change.
      // Measured: 503 us per B=10 64x64 layer with either form -- the maximum is not on the critical path.)
      float m4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) m4[i] = __uint_as_float(sr[i]);
#pragma unroll
      for (int i = 4; i < 64; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (i + 4 + c < 64)
            m4[c] = fmax3f(m4[c], __uint_as_float(sr[i + c]), __uint_as_float(sr[i + 4 + c]));
          else
            m4[c] = fmaxf(m4[c], __uint_as_float(sr[i + c]));
        }
      }
      const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      if (j == 0) {
        m_ref = mx;  // O is still empty (the first P.V overwrites it)
      } else {
        const bool need = (mx - m_ref) * scale_log2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          // lazy rescale of this warp's 32 O rows in TMEM: only legal while no P.V is in flight -> P.V_{j-1} retired
          mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float f = need ? exp2f((m_ref - mx) * scale_log2) : 1.f;
          if (need) m_ref = mx;
          for (int c = 0; c < nacc; c += 32) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(tO + c + lane_off, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
            tmem_st_32x32b_x32(tO + c + lane_off, t);
          }
          tmem_st_wait();
        }
      }
      const float msc = m_ref * scale_log2;
      stamp(tprev, 2);

      uint32_t pk[32];  // P row: 64 fp16 numerators, two per 32-bit tensor-memory column
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = __uint_as_float(sr[2 * i]) * scale_log2 - msc;
        const float x1 = __uint_as_float(sr[2 * i + 1]) * scale_log2 - msc;
        // fp32 exp2 (one MUFU each -- ex2.f16x2 is split into two MUFU ops on sm_100 anyway, and with a stale
        // reference exponent the dominant terms sit at |x| up to 8 where an fp16 argument would cost 0.3 %)
        float p0, p1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
        const __half2 hp = __floats2half2_rn(p0, p1);
        pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
      }
      tmem_st_32x32b_x32(tS + (j & 1) * 64 + lane_off, pk);  // P_j overlays the first half of S_j
      stamp(tprev, 3);
      tmem_st_wait();
      fence_proxy_async();  // the ones column (generic-proxy smem store) -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      if (PROF && trace && warp == 0 && lane == 0) trace[5 * j + 0] = clock64();
      stamp(tprev, 4);
    }
    }  // VAR
    if constexpr (PROF) {
      if (lane == 0) {
        long long* dst = prof + ((size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * AT_SOFTMAX_WARPS + warp) * 8;
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = pf[i];
        dst[5] = T;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        dst[6] = smid;
        dst[7] = clock64();
      }
    }
    // epilogue: O (and the denominator in column d) is read from TMEM exactly once, 32 columns at a time
    if constexpr (VAR == 2) mbar_wait(&o_full[0], (T - 1) & 1);
    else mbar_wait(&o_full[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const uint32_t lden = tmem_ld_32x32b_x1(tO + d + lane_off);
    tmem_ld_wait();
    const float inv = 1.f / __uint_as_float(lden);
    __half* orow = O + (size_t)(row_base + q0 + r) * ldo + h * d;
#pragma unroll
    for (int c0 = 0; c0 < CH * 64; c0 += 32) {
      if (c0 < d) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(tO + c0 + lane_off, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + 8 * i < d) {  // d is a multiple of 8
            uint4 o;
            __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hh[k] = __floats2half2_rn(__uint_as_float(t[8 * i + 2 * k]) * inv, __uint_as_float(t[8 * i + 2 * k + 1]) * inv);
            *reinterpret_cast<uint4*>(orow + c0 + 8 * i) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == AT_WARP_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TCOLS);
  }
}

// =====================================================================================================================
// attn_tc2_kernel: d <= 56 (one 64-column head chunk), N a multiple of 256.  ONE CTA per SM owns TWO 128-query tiles
// (A: softmax warps 0..3, B: warps 4..7) that share every K/V stage, and the exponentials of the two tiles are forced to
// ALTERNATE on the MUFU pipe.
// Why (measured with the tg_attn_set_prof hook, B = 10, N = 4096, d = 40): with two independent CTAs per SM the softmax
// warp of CTA 0 and of CTA 1 on a scheduler start together and stay phase-ALIGNED: both run their 64 MUFU.EX2 + 32 F2FP at
// the same time (928 cycles instead of 608 alone) and then both leave the MUFU pipe idle for ~575 cycles of
// wait / tcgen05.ld / max / tcgen05.st -- 1500 cycles per tile and warp, 68 % MUFU utilisation, 490 us per layer.
// Here warp w (tile A) and warp w + 4 (tile B) sit on the same scheduler and hand the MUFU pipe to each other through two
// named barriers (bar.sync / bar.arrive, 64 threads): while one runs its exponentials the other does everything else.
// The score tiles run two tiles ahead of the softmax (P_j has its own tensor-memory columns; S_{j+2} is issued when the
// four warps of a tile have loaded S_j), so no softmax warp ever waits for the tensor core in steady state.
//   warps 0..3 / 4..7  softmax of query tile A / B, one row per thread (row == TMEM lane)
//   warp 8             TMA producer: K | V of 128 keys per stage (two 16 KB boxes), 4 stages, shared by both query tiles
//   warp 9             TMEM allocator (512 columns) + MMA issuer (warp-uniform loop, one elected lane)
// Tensor memory per query tile t (base 256 t): S0 | S1 | P (32) | O (64) | Q (<= 32).
// =====================================================================================================================
static constexpr int A2_STAGES = 4;
static constexpr int A2_THREADS = 10 * 32;
static constexpr int A2_WARP_TMA = 8, A2_WARP_MMA = 9;
static constexpr int A2_STAGE_BYTES = 2 * AT_CHUNK_BYTES;                 // K (16 KB) | V (16 KB) of 128 keys
static constexpr int A2_BAR_OFFSET = A2_STAGES * A2_STAGE_BYTES;
static constexpr int A2_NUM_BARS = 2 * A2_STAGES + 2 * 7;                 // kv full/empty; per tile: s_full[2] s_free[2] p_full q_ready o_full
static constexpr int A2_SMEM = A2_BAR_OFFSET + A2_NUM_BARS * 8 + 16 + 1024;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <bool PROF>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_tc2_kernel(const __half* __restrict__ Q, int ldq, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, __half* __restrict__ O, int ldo, int N, int d,
                float scale_log2, long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_BAR_OFFSET);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = kv_full + A2_STAGES;
  uint64_t* tile_bars = kv_empty + A2_STAGES;  // [2][7]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + A2_NUM_BARS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 2 * ATQ;
  const int T = N / ATK;             // score tiles per query tile
  const int NS = T / AT_TPS;         // K/V stages
  const int row_base = b * N;
  const int ksteps = (d + 15) / 16;

  if (warp == A2_WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < A2_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      uint64_t* tb = tile_bars + 7 * t;
      mbar_init(&tb[0], 1);                    // s_full[0]
      mbar_init(&tb[1], 1);                    // s_full[1]
      mbar_init(&tb[2], AT_SOFTMAX_WARPS);     // s_free[0]
      mbar_init(&tb[3], AT_SOFTMAX_WARPS);     // s_free[1]
      mbar_init(&tb[4], AT_SOFTMAX_WARPS);     // p_full
      mbar_init(&tb[5], AT_SOFTMAX_WARPS);     // q_ready
      mbar_init(&tb[6], 1);                    // o_full
    }
    fence_barrier_init();
  }
  if (warp == A2_WARP_MMA) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  grid_dep_launch();
  grid_dep_wait();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  constexpr uint32_t C_S = 0, C_P = 128, C_O = 160, C_Q = 224, C_TILE = 256;

  if (warp == A2_WARP_TMA) {
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) {
        const int s = i % A2_STAGES;
        mbar_wait(&kv_empty[s], ((i / A2_STAGES) & 1) ^ 1);
        uint8_t* dst = sKV + s * A2_STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], A2_STAGE_BYTES);
        tma_load_2d(dst, &tmK, &kv_full[s], h * ATD, row_base + i * AT_TPS * ATK);
        tma_load_2d(dst + AT_CHUNK_BYTES, &tmV, &kv_full[s], h * ATD, row_base + i * AT_TPS * ATK);
      }
    }
  } else if (warp == A2_WARP_MMA) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_f16(ATQ, ATK);
    const uint32_t idesc_o = make_idesc_f16(ATQ, ((d + 1 + 15) / 16) * 16) | (1u << 16);  // V is the MN-major B operand
    const uint32_t kv_base = smem_u32(sKV);
    auto issue_s = [&](int t, int j) {
      const int i = j / AT_TPS, s = i % A2_STAGES;
      if (t == 0 && j % AT_TPS == 0) mbar_wait(&kv_full[s], (i / A2_STAGES) & 1);  // tile A meets every stage first
      tc_fence_after();
      const uint32_t k_addr = kv_base + s * A2_STAGE_BYTES + (j % AT_TPS) * AT_TILE_BYTES;
      const uint32_t tb = tmem_base + t * C_TILE;
      if (leader) {
        for (int k = 0; k < ksteps; ++k)
          umma_f16_ts(tb + C_S + (j & 1) * 64, tb + C_Q + k * 8, make_smem_desc_sw128(k_addr + k * 32), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(&tile_bars[7 * t + (j & 1)]);
      }
    };
    for (int t = 0; t < 2; ++t) {
      mbar_wait(&tile_bars[7 * t + 5], 0);  // Q rows of tile t in tensor memory
      issue_s(t, 0);
      issue_s(t, 1);
    }
    for (int j = 0; j < T; ++j) {
      const int i = j / AT_TPS, s = i % A2_STAGES;
      const uint32_t v_addr = kv_base + s * A2_STAGE_BYTES + AT_CHUNK_BYTES + (j % AT_TPS) * AT_TILE_BYTES;
      for (int t = 0; t < 2; ++t) {  // the ping-pong makes P_j of tile A complete half a period before tile B's
        uint64_t* tb = tile_bars + 7 * t;
        if (j + 2 < T) {
          // S_{j+2} of this tile as soon as its buffer has been read (s_free fires at the top of the tile's softmax of
          // S_j, half a period ago): the softmax prefetches S_{j+1} in the middle of tile j, S must run well ahead
          mbar_wait(&tb[2 + (j & 1)], (j >> 1) & 1);
          issue_s(t, j + 2);
        }
        mbar_wait(&tb[4], j & 1);
        tc_fence_after();
        const uint32_t tc = tmem_base + t * C_TILE;
        if (leader) {
#pragma unroll
          for (int k = 0; k < ATK / 16; ++k)
            umma_f16_ts(tc + C_O, tc + C_P + k * 8, make_smem_desc_sw128_mn(v_addr + k * 2048), idesc_o,
                        (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(&tb[6]);
          if (t == 1 && j % AT_TPS == AT_TPS - 1) tc_commit(&kv_empty[s]);
        }
      }
    }
  } else {
    const int t = warp >> 2, q = warp & 3;       // query tile, TMEM lane quarter (hardware rule: warp id % 4)
    const int r = q * 32 + lane;                 // query row inside the tile == TMEM lane
    const int tid_t = threadIdx.x & 127;         // 0..127 inside the tile's softmax threads
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tc = tmem_base + t * C_TILE;
    const uint32_t tS = tc + C_S, tP = tc + C_P, tO = tc + C_O, tQ = tc + C_Q;
    uint64_t* tb = tile_bars + 7 * t;
    uint64_t* s_full = tb;
    uint64_t* s_free = tb + 2;
    uint64_t* p_full = tb + 4;
    uint64_t* q_ready = tb + 5;
    uint64_t* o_full = tb + 6;
    // MUFU hand-over between warp q of tile A and warp q of tile B (same scheduler): `my_turn` is awaited by this warp,
    // `other_turn` is signalled when this warp's exponentials have been issued
    const int my_turn = 1 + t * 4 + q, other_turn = 1 + (t ^ 1) * 4 + q;
    const int nacc = ((d + 1 + 15) / 16) * 16;
    float m_ref = -INFINITY;
    const int dl = d & 63;
    const int ones_off = (tid_t & 63) * 128 + ((((dl >> 3) ^ (tid_t & 7)) << 4) | ((dl & 7) << 1));
    const int qrow = row_base + q0 + t * ATQ + r;

    {
      const uint4* qp = reinterpret_cast<const uint4*>(Q + (size_t)qrow * ldq + h * ATD);
      for (int k = 0; k < ksteps; ++k) {
        uint4 v[2];
        v[0] = __ldg(qp + 2 * k);
        v[1] = __ldg(qp + 2 * k + 1);
        tmem_st_32x32b_x8(tQ + k * 8 + lane_off, reinterpret_cast<const uint32_t*>(v));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    if (t == 1) named_bar_arrive(other_turn, 64);  // tile A goes first

    long long pf[6] = {0, 0, 0, 0, 0, 0};  // PROF: wait S + ld | max (+rescale) + scale | wait turn | exps | store + arrive
    long long tprev = 0;
    auto stamp = [&](int slot) {
      if constexpr (PROF) {
        const long long now = clock64();
        pf[slot] += now - tprev;
        tprev = now;
      }
    };
    if constexpr (PROF) tprev = clock64();

    // One tile: `cur` holds the raw scores S_j (tcgen05.ld issued by the previous tile, or by the prologue), `nxt` takes
    // S_{j+1}: its tensor-memory read travels during this tile's MUFU turn, so the fixed cost outside the turn -- which
    // has to fit into the partner warp's turn -- loses the load latency (~100 cycles) and the mbarrier probe.
    auto tile = [&](int j, uint32_t (&cur)[64], uint32_t (&nxt)[64]) {
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[j & 1]);  // S_j is in registers: its buffer may take S_{j+2}
      stamp(0);
      if (tid_t < 64)  // the ones column of this tile's V (both query tiles write the same value)
        *reinterpret_cast<__half*>(sKV + ((j / AT_TPS) % A2_STAGES) * A2_STAGE_BYTES + AT_CHUNK_BYTES +
                                   (j % AT_TPS) * AT_TILE_BYTES + ones_off) = __float2half(1.f);
      float m4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) m4[i] = __uint_as_float(cur[i]);
#pragma unroll
      for (int i = 4; i < 64; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (i + 4 + c < 64) m4[c] = fmax3f(m4[c], __uint_as_float(cur[i + c]), __uint_as_float(cur[i + 4 + c]));
          else m4[c] = fmaxf(m4[c], __uint_as_float(cur[i + c]));
        }
      }
      const float mx = fmaxf(fmax3f(m4[0], m4[1], m4[2]), m4[3]);
      bool o_waited = false;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = (mx - m_ref) * scale_log2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          // lazy rescale of this warp's 32 O rows: legal while no P.V of this tile is in flight -> P_{j-1} V_{j-1} retired
          mbar_wait(o_full, (j - 1) & 1);
          o_waited = true;
          tc_fence_after();
          const float f = need ? exp2f((m_ref - mx) * scale_log2) : 1.f;
          if (need) m_ref = mx;
          for (int c = 0; c < nacc; c += 16) {
            uint32_t tt[16];
            tmem_ld_32x32b_x16(tO + c + lane_off, tt);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * f);
            tmem_st_32x32b_x16(tO + c + lane_off, tt);
          }
          tmem_st_wait();
        }
      }
      const float msc = m_ref * scale_log2;
#pragma unroll
      for (int i = 0; i < 64; ++i)  // exponent arguments: FMA pipe, outside the MUFU turn
        cur[i] = __float_as_uint(__uint_as_float(cur[i]) * scale_log2 - msc);
      if (j + 1 < T) {  // S_{j+1} has been complete for about a tile (the issuer runs two tiles ahead)
        mbar_wait(&s_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32(tS + ((j + 1) & 1) * 64 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&nxt[0]));
        tmem_ld_32x32b_x32(tS + ((j + 1) & 1) * 64 + 32 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&nxt[32]));
      }
      stamp(1);
      named_bar_sync(my_turn, 64);
      stamp(2);
      // the MUFU turn: 64 independent exponentials back to back (volatile keeps them between the two barriers; the
      // fp16 packing depends on them and waits until the pipe has been handed over)
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(cur[i]));
      if (j + 1 < T || t == 0) named_bar_arrive(other_turn, 64);  // (tile B's last hand-over has no taker)
      stamp(3);
#pragma unroll
      for (int i = 0; i < 32; ++i) {  // packed in place: element i comes from elements 2i, 2i+1 >= i
        const __half2 hp = __floats2half2_rn(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
        cur[i] = *reinterpret_cast<const uint32_t*>(&hp);
      }
      if (j > 0 && !o_waited) {  // the single P buffer feeds P_{j-1} V_{j-1} until that group has retired (long ago)
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_32x32b_x32(tP + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&cur[0]));
      tmem_st_wait();
      fence_proxy_async();  // the ones column (generic-proxy smem store) -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      stamp(4);
    };
    {
      uint32_t xa[64], xb[64];
      mbar_wait(&s_full[0], 0);
      tc_fence_after();
      tmem_ld_32x32b_x32(tS + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&xa[0]));
      tmem_ld_32x32b_x32(tS + 32 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&xa[32]));
      for (int j = 0; j < T; j += 2) {  // T is even (N is a multiple of 128)
        tile(j, xa, xb);
        tile(j + 1, xb, xa);
      }
    }
    if constexpr (PROF) {
      if (lane == 0) {
        long long* dst = prof + ((size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + warp) * 8;
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = pf[i];
        dst[5] = T;
        dst[7] = clock64();
      }
    }
    mbar_wait(o_full, (T - 1) & 1);
    tc_fence_after();
    const uint32_t lden = tmem_ld_32x32b_x1(tO + d + lane_off);
    tmem_ld_wait();
    const float inv = 1.f / __uint_as_float(lden);
    __half* orow = O + (size_t)qrow * ldo + h * d;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      if (c0 < d) {
        uint32_t tt[32];
        tmem_ld_32x32b_x32(tO + c0 + lane_off, tt);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + 8 * i < d) {  // d is a multiple of 8
            uint4 o;
            __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hh[k] = __floats2half2_rn(__uint_as_float(tt[8 * i + 2 * k]) * inv, __uint_as_float(tt[8 * i + 2 * k + 1]) * inv);
            *reinterpret_cast<uint4*>(orow + c0 + 8 * i) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == A2_WARP_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static long long* g_attn_prof = nullptr;
void attn_set_prof(long long* dev_buf) { g_attn_prof = dev_buf; }
// Softmax / issue schedule of the tcgen05 self-attention (tg_attn_set_variant; TG_ATTN_VARIANT sets the initial value):
//   0  (default) P_j overlays S_j, one 128-query tile per CTA, two CTAs per SM (d <= 56) -- fastest measured at every size
//   2  P_j in its own tensor-memory columns, S tiles issued two tiles ahead of the softmax
//   3  d <= 56 and N % 256 == 0: attn_tc2_kernel (two query tiles per CTA, alternating MUFU turns); otherwise like 2
// All three are parity-tested (tests/test_ops_gpu.py); 2 and 3 are kept as measured alternatives (profiles/r02_attention.md).
static int g_attn_variant = -1;
void attn_set_variant(int v) { g_attn_variant = v; }
static int attn_variant() {
  if (g_attn_variant < 0) {
    const char* e = getenv("TG_ATTN_VARIANT");
    g_attn_variant = e ? atoi(e) : 0;
  }
  return g_attn_variant;
}

// heads of Q / K / V packed at stride d (no zero columns between them) also for d <= 120: set by attention_tc for the launch
// it is about to enqueue
static thread_local bool g_attn_packed = false;

template <int CH, int VAR, bool PROF>
static int launch_attn_tc_v(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O,
                            int ldo, int B, int heads, int N, int d, float scale, cudaStream_t st) {
  const bool packed = g_attn_packed;
  using Cfg = AttnCfg<CH>;
  auto kern = attn_tc_kernel<CH, VAR, PROF>;
  TG_SMEM_ATTR_ONCE(kern, Cfg::SMEM);
  CUtensorMap tk, tv;
  // CH = 3 (d = 160): heads packed at stride d -- the 64-column boxes of a head's last chunk reach into the next head (or
  // past the tensor: zero fill); those columns lie beyond the K extent of Q K^T and beyond the output columns that are read
  const int hstride = (CH == 3 || packed) ? d : CH * ATD;
  const uint64_t rows = (uint64_t)B * N, cols = (uint64_t)heads * hstride;
  if (make_tmap_2d(&tk, K, rows, cols, ldk, AT_TPS * ATK)) return -1;
  if (make_tmap_2d(&tv, V, rows, cols, ldv, AT_TPS * ATK)) return -1;
  static int smem_pad = -1;  // TG_ATTN_ONE_CTA=1 (measurement): request > half of the SM's shared memory -> one CTA per SM
  if (smem_pad < 0) smem_pad = (getenv("TG_ATTN_ONE_CTA") && atoi(getenv("TG_ATTN_ONE_CTA"))) ? 120 * 1024 : 0;
  const int smem = Cfg::SMEM > smem_pad ? Cfg::SMEM : smem_pad;
  if (smem_pad) TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  TG_CUDA(launch_kernel(kern, dim3(N / ATQ, heads, B), dim3(AT_THREADS), smem, st, 1, Q, ldq, tk, tv, O, ldo, N, d,
                        scale * 1.4426950408889634f, PROF ? g_attn_prof : (long long*)nullptr, hstride));
  TG_COUNT_LAUNCH();
  return 0;
}

template <bool PROF>
static int launch_attn_tc2(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                           int B, int heads, int N, int d, float scale, cudaStream_t st) {
  auto kern = attn_tc2_kernel<PROF>;
  TG_SMEM_ATTR_ONCE(kern, A2_SMEM);
  CUtensorMap tk, tv;
  const uint64_t rows = (uint64_t)B * N, cols = (uint64_t)heads * ATD;
  if (make_tmap_2d(&tk, K, rows, cols, ldk, AT_TPS * ATK)) return -1;
  if (make_tmap_2d(&tv, V, rows, cols, ldv, AT_TPS * ATK)) return -1;
  TG_CUDA(launch_kernel(kern, dim3(N / (2 * ATQ), heads, B), dim3(A2_THREADS), A2_SMEM, st, 1, Q, ldq, tk, tv, O, ldo, N, d,
                        scale * 1.4426950408889634f, PROF ? g_attn_prof : (long long*)nullptr));
  TG_COUNT_LAUNCH();
  return 0;
}

template <int CH>
static int launch_attn_tc(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                          int B, int heads, int N, int d, float scale, cudaStream_t st) {
  if constexpr (CH == 3) return launch_attn_tc_v<3, 0, false>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  const int var = attn_variant();
  if (CH == 1 && var == 3 && N % (2 * ATQ) == 0 && !g_attn_packed) {  // two query tiles per CTA, alternating exponentials
    if (g_attn_prof != nullptr) return launch_attn_tc2<true>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
    return launch_attn_tc2<false>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  }
  if (g_attn_prof != nullptr) {  // measurement only
    if (var == 0) return launch_attn_tc_v<CH, 0, true>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
    return launch_attn_tc_v<CH, 2, true>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  }
  if (var == 0) return launch_attn_tc_v<CH, 0, false>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  return launch_attn_tc_v<CH, 2, false>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
}

// Q/K/V: [B*N, heads * CH * 64] views (head h at column h * CH * 64, columns >= d zero), CH = 1 for d <= 56, 2 for d <= 120
// packed != 0: heads at stride d for every d (Q / K / V are [B*N, heads * d] views): the 64-column boxes of a head's last chunk
// reach into the next head (or past the tensor: zero fill); those columns meet zeros of the Q operand in Q K^T and lie
// beyond the output columns that are read in P V.
int attention_tc(const __half* Q, const __half* K, const __half* V, int ldq, int ldk, int ldv, __half* O, int ldo,
                 int B, int heads, int N, int d, float scale, int packed, cudaStream_t st) {
  struct PackedScope {
    explicit PackedScope(bool p) { g_attn_packed = p; }
    ~PackedScope() { g_attn_packed = false; }
  } scope(packed != 0);
  if (packed)
    TG_CHECK((reinterpret_cast<uintptr_t>(K) & 15) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (d * 2) % 16 == 0,
             "attention_tc: packed heads need 16-byte aligned K / V and d %% 8 == 0");
  TG_CHECK(N % ATQ == 0 && d % 8 == 0 && d >= 8 && (d + 1 <= 2 * ATD || (d % 16 == 0 && d <= 176)),
           "attention_tc: N=%d (multiple of 128) d=%d (8..120 multiple of 8, or 128..176 multiple of 16)", N, d);
  TG_CHECK(ldo % 8 == 0 && (d * 2) % 16 == 0, "attention_tc: output pitch/head offset must be 16B aligned");
  TG_CHECK(ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0, "attention_tc: Q must be 16-byte aligned with a pitch multiple of 8");
  if (d + 1 <= ATD) return launch_attn_tc<1>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  if (d + 1 <= 2 * ATD) return launch_attn_tc<2>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);
  TG_CHECK((reinterpret_cast<uintptr_t>(K) & 15) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (d * 2) % 16 == 0,
           "attention_tc: packed heads need 16-byte aligned K / V");
  return launch_attn_tc<3>(Q, K, V, ldq, ldk, ldv, O, ldo, B, heads, N, d, scale, st);  // packed heads (see the header)
}

}  // namespace tg
