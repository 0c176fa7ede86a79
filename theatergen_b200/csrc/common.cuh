// Shared device/host helpers for the TheaterGen B200 kernels (sm_100a only).
// Thin inline-PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA/TMEM).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tg {

// ----------------------------------------------------------------------------------------------
// error reporting (thread-local message, C-ABI friendly)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define TG_CHECK(cond, ...)                \
  do {                                     \
    if (!(cond)) {                         \
      tg::set_error(__VA_ARGS__);          \
      return -1;                           \
    }                                      \
  } while (0)

#define TG_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      tg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                    __FILE__, __LINE__);                                           \
      return -2;                                                                   \
    }                                                                              \
  } while (0)

// cudaFuncSetAttribute applies to the CURRENT device's copy of a kernel: opt in to the dynamic shared memory once per
// (kernel instantiation, device) -- a process that drives a second device must not inherit the first one's flag.
#define TG_MAX_DEVICES 16
#define TG_SMEM_ATTR_ONCE(kern, bytes)                                                                    \
  do {                                                                                                    \
    static bool _done[TG_MAX_DEVICES] = {};                                                               \
    int _dev = 0;                                                                                         \
    TG_CUDA(cudaGetDevice(&_dev));                                                                        \
    TG_CHECK(_dev >= 0 && _dev < TG_MAX_DEVICES, "device index %d out of range", _dev);                    \
    if (!_done[_dev]) {                                                                                   \
      TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      _done[_dev] = true;                                                                                 \
    }                                                                                                     \
  } while (0)

#define TG_LAUNCH_CHECK()                                                          \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      tg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                    __FILE__, __LINE__);                                           \
      return -3;                                                                   \
    }                                                                              \
  } while (0)

// counts kernel launches issued by this library (bench.py reports gpu_launches from it)
extern unsigned long long g_launch_count;
#define TG_COUNT_LAUNCH() (++tg::g_launch_count)

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor / byte-count bug must trap, never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {  // try_wait itself sleeps ~us per probe; this is seconds
      printf("tg: mbarrier wait timeout block(%d,%d,%d) thread %d parity %u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, parity);
      __trap();
    }
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16 (fp16/bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 64 fp16 = 128 B, 8-row atoms of 1024 B).
//   start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (=1024>>4)
//   version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: fp16 A/B (K-major both), fp32 accumulate, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                 // c_format = F32
         | (0u << 7) | (0u << 10)  // a/b format = F16
         | (0u << 15) | (0u << 16) // a/b K-major
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Programmatic dependent launch: let the next kernel in the stream get scheduled (and run its prologue) while this
// one is still running; every kernel waits for its prerequisite grids before touching global memory.
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// SiLU with ONE MUFU op: x * sigmoid(x) = x * (0.5 + 0.5 tanh(x / 2)); tanh.approx.f32 has ~2^-11 relative error, below
// the fp16 resolution of the result.  (x / (1 + exp(-x)) costs an ex2 and a reciprocal: the GroupNorm+SiLU apply pass over
// a 26 MB tensor was bound by those 2 x 13 M MUFU operations as much as by memory.)
__device__ __forceinline__ float silu_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
// x * sigmoid(1.702 x), the activation of OpenAI CLIP's MLPs (the SD-1.5 text encoder); one MUFU operation.
__device__ __forceinline__ float quick_gelu_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// exact-GELU without MUFU: erf(z) = z * P(z^2) on [0, 3.2] (degree-8 least-squares fit at Chebyshev nodes, |err| < 9e-5,
// i.e. below fp16 resolution of the result), clamped to 1 beyond.  ~14 FMA-pipe instructions, no SFU traffic, so a GEGLU
// epilogue does not fight the tile's own MUFU budget.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752f, 3.2f);
  const float t = z * z;
  float p = 3.4405886140120856e-08f;
  p = fmaf(p, t, -1.7486363503849134e-06f);
  p = fmaf(p, t, 3.880867370753549e-05f);
  p = fmaf(p, t, -0.0005002432153560221f);
  p = fmaf(p, t, 0.004227499943226576f);
  p = fmaf(p, t, -0.025213463231921196f);
  p = fmaf(p, t, 0.11144663393497467f);
  p = fmaf(p, t, -0.3756941556930542f);
  p = fmaf(p, t, 1.1283652782440186f);
  const float e = fminf(p * z, 1.f);
  return 0.5f * x * (1.f + copysignf(e, x));
}

// The same polynomial on W values at once, Horner step by Horner step: W independent dependency chains side by side, so the
// 4-cycle FMA latency of one chain is covered by the others even when the register allocator keeps few values in flight
// (a GEGLU epilogue evaluated element by element ran ~70 cycles per element: one dependent chain at a time).
template <int W>
__device__ __forceinline__ void gelu_erf_fast_n(const float (&x)[W], float (&y)[W]) {
  float z[W], t[W], p[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    z[j] = fminf(fabsf(x[j]) * 0.70710678118654752f, 3.2f);
    t[j] = z[j] * z[j];
    p[j] = 3.4405886140120856e-08f;
  }
  const float c[8] = {-1.7486363503849134e-06f, 3.880867370753549e-05f, -0.0005002432153560221f, 0.004227499943226576f,
                      -0.025213463231921196f, 0.11144663393497467f, -0.3756941556930542f, 1.1283652782440186f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int j = 0; j < W; ++j) p[j] = fmaf(p[j], t[j], c[k]);
  }
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const float e = fminf(p[j] * z[j], 1.f);
    y[j] = 0.5f * x[j] * (1.f + copysignf(e, x[j]));
  }
}

// 3-input maximum (PTX ISA 8.6, sm_100: one FMNMX3 instead of two FMNMX)
__device__ __forceinline__ float fmax3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

// ----------------------------------------------------------------------------------------------
// host: kernel launch with programmatic stream serialization (PDL) and optional thread-block clusters
// ----------------------------------------------------------------------------------------------
bool pdl_enabled();  // TG_PDL=0 disables (A/B testing)
void set_pdl(int mode);  // per host thread: -1 process default, 0 off, 1 on (tg_set_pdl)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ----------------------------------------------------------------------------------------------
// host: TMA descriptor encoding through the driver entry point (no link-time libcuda dependency)
// ----------------------------------------------------------------------------------------------
// 2D row-major fp16 [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], SWIZZLE_128B
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols = 64);
// row-major fp16 [rows, K] seen as [K/64 chunks][rows][64]; box = [kc chunks][box_rows][64], SWIZZLE_128B: ONE instruction
// lands kc consecutive K-major [box_rows x 64] tiles (chunks past K/64 are zero-filled)
int make_tmap_kchunks(CUtensorMap* out, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t kc);
int make_tmap_w_ky3(CUtensorMap* out, const void* base, uint64_t rows, uint64_t Cin, uint32_t box_rows);
// 4D NHWC fp16 [B, H, W, C]; box = [bb, bh, bw, 64], SWIZZLE_128B, OOB zero fill (conv padding)
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t bb,
                   uint32_t bh, uint32_t bw, uint32_t pixel_stride = 1);
int make_tmap_phase_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t W, uint64_t C, uint32_t box_w,
                        uint32_t box_r);

}  // namespace tg
