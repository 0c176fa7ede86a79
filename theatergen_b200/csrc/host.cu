// Host-side plumbing: error strings, launch counter, TMA descriptor encoding.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tg {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// per host thread: -1 = the process default (TG_PDL), 0 / 1 = off / on for the launches this thread enqueues next
static thread_local int g_pdl_override = -1;
void set_pdl(int mode) { g_pdl_override = mode < 0 ? -1 : (mode != 0); }

bool pdl_enabled() {
  if (g_pdl_override >= 0) return g_pdl_override != 0;
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  TG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
           "tmap2d: base/pitch must be 16-byte aligned (base=%p ld=%llu)", base, (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u) -> %d",
           (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, (int)r);
  return 0;
}

int make_tmap_kchunks(CUtensorMap* out, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t kc) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  TG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && K % 64 == 0, "tmap kchunks: alignment (base=%p K=%llu)", base,
           (unsigned long long)K);
  cuuint64_t dims[3] = {64, rows, K / 64};
  cuuint64_t strides[2] = {K * 2, 128};
  cuuint32_t box[3] = {64, box_rows, kc};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(kchunks rows=%llu K=%llu box=%ux%u) -> %d", (unsigned long long)rows,
           (unsigned long long)K, box_rows, kc, (int)r);
  return 0;
}

// Weights of a 3x3 convolution [N][ky][kx][Cin] seen as [ky][kx*Cin/64 + cb][N][64]: one box {64, box_rows, 1, 3} brings
// the 64-channel chunks of the three taps ky = 0,1,2 of one (kx, channel block) -- the KY3 stage of gemm_tc_kernel.
int make_tmap_w_ky3(CUtensorMap* out, const void* base, uint64_t rows, uint64_t Cin, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  TG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && Cin % 64 == 0, "tmap w_ky3: alignment (base=%p Cin=%llu)", base,
           (unsigned long long)Cin);
  cuuint64_t dims[4] = {64, rows, 3 * Cin / 64, 3};
  cuuint64_t strides[3] = {9 * Cin * 2, 128, 3 * Cin * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 3};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(w_ky3 rows=%llu Cin=%llu box_rows=%u) -> %d", (unsigned long long)rows,
           (unsigned long long)Cin, box_rows, (int)r);
  return 0;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t bb,
                   uint32_t bh, uint32_t bw, uint32_t pixel_stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  TG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (C * 2) % 16 == 0, "tmap nhwc: alignment");
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  // pixel_stride 2: the box gathers every other pixel in x and y (the traversal stride of the TMA unit) -- the operand
  // of a stride-2 convolution without an im2col buffer.  The driver API counts boxDim in INPUT elements: to load n
  // elements along a dimension with traversal stride s, boxDim = n * s.
  cuuint32_t box[4] = {64, bw * pixel_stride, bh * pixel_stride, bb};
  TG_CHECK(box[1] <= 256 && box[2] <= 256, "tmap nhwc: box %u x %u exceeds 256 with pixel stride %u", box[1], box[2], pixel_stride);
  cuuint32_t estr[4] = {1, pixel_stride, pixel_stride, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc B=%llu H=%llu W=%llu C=%llu box=%u,%u,%u) -> %d",
           (unsigned long long)B, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, bb, bh, bw,
           (int)r);
  return 0;
}

// Output of one (py, px) phase of "nearest x2 upsample -> 3x3 convolution" computed on the LOW-resolution grid: phase pixel
// (b, i, j) lives at pixel (2i + py, 2j + px) of the [B, 2H, 2W, C] image.  `base` already points at pixel (py, px); seen
// from there the phase image is [rows = B*H][W][C] with strides (4 W C, 2 C, 1) elements.  Store box: 32 channels x
// box_w x box_r phase pixels = the 32 consecutive tile rows an epilogue warp stages (64-byte rows, 64B swizzle).
int make_tmap_phase_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t W, uint64_t C, uint32_t box_w,
                        uint32_t box_r) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  TG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (C * 2) % 16 == 0, "tmap phase out: alignment");
  cuuint64_t dims[3] = {C, W, rows};
  cuuint64_t strides[2] = {2 * C * 2, 4 * W * C * 2};
  cuuint32_t box[3] = {32, box_w, box_r};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(phase out rows=%llu W=%llu C=%llu box=%u,%u) -> %d",
           (unsigned long long)rows, (unsigned long long)W, (unsigned long long)C, box_w, box_r, (int)r);
  return 0;
}

}  // namespace tg
