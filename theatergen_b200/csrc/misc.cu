// Small HBM/latency-bound kernels around the tensor-core ones.
#include "common.cuh"
#include "ops.h"

namespace tg {

// ---------------------------------------------------------------------------------------------------------
// nearest x2 upsample, NHWC (diffusers Upsample2D: F.interpolate(scale_factor=2, mode="nearest") before its conv;
// call sites ip_adapter/unet_2d_blocks.py:2251,2332).  16-byte vectors, each input vector written 4 times.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C8, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t p = i / C8;
    const int ox = (int)(p % (2 * W));
    p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const int b = (int)(p / (2 * H));
    y[i] = x[(((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * C8 + c];
  }
}

int upsample2x_nhwc(const __half* x, int B, int H, int W, int C, __half* y, cudaStream_t st) {
  TG_CHECK(C % 8 == 0, "upsample: C %% 8 != 0");
  const size_t total = (size_t)B * 4 * H * W * (C / 8);
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 8) ? ((total + 255) / 256) : (size_t)(148 * 8));
  TG_CUDA(launch_kernel(upsample2x_kernel, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W, C / 8,
                                          total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// im2col for the stride-2 3x3 convs (diffusers Downsample2D).  pad_lo = 1: symmetric padding 1 (the UNet / ControlNet,
// ip_adapter/unet_2d_blocks.py:1112,1198); pad_lo = 0: the VAE encoder's F.pad(x, (0, 1, 0, 1)) + conv(padding=0)
// A[b*Ho*Wo + oy*Wo + ox][(ky*3+kx)*C + c] = x[b][2oy+ky-pad_lo][2ox+kx-pad_lo][c]  (0 outside)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ a, int H, int W, int C8, size_t total, int pad_lo) {
  grid_dep_launch();
  grid_dep_wait();
  const int Ho = H / 2, Wo = W / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t p = i / C8;
    const int tap = (int)(p % 9);
    p /= 9;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int iy = 2 * oy + tap / 3 - pad_lo, ix = 2 * ox + tap % 3 - pad_lo;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((size_t)b * H + iy) * W + ix) * C8 + c];
    a[i] = v;
  }
}

int im2col_s2_nhwc(const __half* x, int B, int H, int W, int C, __half* a, int pad_lo, cudaStream_t st) {
  TG_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && (pad_lo == 0 || pad_lo == 1), "im2col_s2: bad dims");
  const size_t total = (size_t)B * (H / 2) * (W / 2) * 9 * (C / 8);
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(im2col_s2_kernel, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(a), H, W, C / 8,
                                         total, pad_lo));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// conv_in: NCHW [B,4,H,W] -> NHWC [B,H,W,320], 3x3 pad 1 (ip_adapter/unet_2d_condition.py:999)
// weights kept in torch layout [Cout][Cin][3][3]; K = 36 -> CUDA cores, weights in shared memory.
// ---------------------------------------------------------------------------------------------------------
template <int CIN, int PIX>
__global__ void __launch_bounds__(256)
conv_in_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
               __half* __restrict__ y, int H, int W, int Cout, int silu) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float sm_ci[];
  float* sw = sm_ci;                        // [CIN*9][Cout]
  float* sp = sw + CIN * 9 * Cout;          // [PIX][CIN*9]
  const int b = blockIdx.y, p0 = blockIdx.x * PIX, HW = H * W;
  for (int i = threadIdx.x; i < CIN * 9 * Cout; i += blockDim.x) {
    const int co = i / (CIN * 9), k = i % (CIN * 9);
    sw[k * Cout + co] = __half2float(w[i]);
  }
  for (int i = threadIdx.x; i < PIX * CIN * 9; i += blockDim.x) {
    const int pp = i / (CIN * 9), k = i % (CIN * 9);
    const int ci = k / 9, tap = k % 9;
    const int p = p0 + pp, yy = p / W + tap / 3 - 1, xx = p % W + tap % 3 - 1;
    float v = 0.f;
    if (p < HW && yy >= 0 && yy < H && xx >= 0 && xx < W) v = __half2float(x[(((size_t)b * CIN + ci) * H + yy) * W + xx]);
    sp[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PIX * Cout; i += blockDim.x) {
    const int pp = i / Cout, co = i % Cout;
    if (p0 + pp >= HW) continue;
    float acc = __half2float(bias[co]);
#pragma unroll
    for (int k = 0; k < CIN * 9; ++k) acc += sp[pp * CIN * 9 + k] * sw[k * Cout + co];
    if (silu) acc = silu_f(acc);
    y[((size_t)b * HW + p0 + pp) * Cout + co] = __float2half_rn(acc);
  }
}

template <int CIN>
static int conv_in_launch(const __half* x, const __half* w, const __half* bias, __half* y, int B, int H, int W, int Cout,
                          int silu, cudaStream_t st) {
  constexpr int PIX = 32;
  const size_t smem = (size_t)(CIN * 9 * Cout + PIX * CIN * 9) * sizeof(float);
  auto kern = conv_in_kernel<CIN, PIX>;
  TG_SMEM_ATTR_ONCE(kern, 96 * 1024);
  TG_CHECK(smem <= 96 * 1024, "conv_in: Cout=%d too large", Cout);
  TG_CUDA(launch_kernel(kern, dim3((H * W + PIX - 1) / PIX, B), dim3(256), smem, st, 1, x, w, bias, y, H, W, Cout, silu));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// Cin = 4: the UNet / ControlNet conv_in on latents; Cin = 3: the first convolution of the ControlNet conditioning
// embedder on the RGB hint image (diffusers ControlNetConditioningEmbedding.conv_in, followed by SiLU).
int conv_in_nchw_to_nhwc(const __half* x, const __half* w, const __half* bias, __half* y, int B, int Cin, int H, int W,
                         int Cout, int silu, cudaStream_t st) {
  if (Cin == 4) return conv_in_launch<4>(x, w, bias, y, B, H, W, Cout, silu, st);
  if (Cin == 3) return conv_in_launch<3>(x, w, bias, y, B, H, W, Cout, silu, st);
  TG_CHECK(false, "conv_in: Cin=%d unsupported (3 or 4)", Cin);
}

// ---------------------------------------------------------------------------------------------------------
// conv_out: NHWC [B,H,W,C] (already GroupNorm+SiLU-ed) -> NCHW [B,4,H,W], 3x3 pad 1
// (ip_adapter/unet_2d_condition.py:1120).  One warp per output pixel; weights [4][3][3][C] via L1.
// ---------------------------------------------------------------------------------------------------------
// Weights [COUT][3][3][C] are staged ONCE per CTA in shared memory (23 KB at C = 320) and reused for the CTA's 64
// pixels; a warp owns a pixel at a time, lanes split the channels (half2), warp-reduce at the end.  (The first version
// re-read the weights through L1 for every pixel: 128 us at B = 10; this one is bound by the 9 input reads.)
static constexpr int CONV_OUT_PIX = 64;
template <int COUT>
__global__ void __launch_bounds__(256)
conv_out_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                __half* __restrict__ y, int B, int H, int W, int C) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ __align__(16) uint8_t sm_co[];
  __half2* sw = reinterpret_cast<__half2*>(sm_co);  // [COUT*9][C/2]
  const int c2n = C / 2, HW = H * W;
  for (int i = threadIdx.x; i < COUT * 9 * c2n; i += blockDim.x) sw[i] = reinterpret_cast<const __half2*>(w)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = B * HW;
  for (int pi = warp; pi < CONV_OUT_PIX; pi += 8) {
    const int gp = blockIdx.x * CONV_OUT_PIX + pi;
    if (gp >= total) break;
    const int b = gp / HW, p = gp % HW, py = p / W, px = p % W;
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    // a filter row (3 taps) at a time: all of its loads are issued before any is used (out-of-image taps read a valid
    // address and are multiplied by zero), so a lane keeps 3 x ceil(C/64) independent loads in flight
    constexpr int MAXI = 5;  // C <= 320
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
      __half2 xv[3][MAXI];
      float ok[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        ok[kx] = (xx >= 0 && xx < W) ? 1.f : 0.f;
        const int xc = min(max(xx, 0), W - 1);
        const __half2* xp = reinterpret_cast<const __half2*>(x + (((size_t)b * H + yy) * W + xc) * C);
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
          const int c = lane + 32 * i;
          xv[kx][i] = c < c2n ? xp[c] : __half2{};
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tap = ky * 3 + kx;
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
          const int c = lane + 32 * i;
          if (c < c2n) {
            float2 xf = __half22float2(xv[kx][i]);
            xf.x *= ok[kx];
            xf.y *= ok[kx];
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
              const float2 wv = __half22float2(sw[(o * 9 + tap) * c2n + c]);
              acc[o] = fmaf(xf.x, wv.x, fmaf(xf.y, wv.y, acc[o]));
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      const float s_ = warp_sum(acc[o]);
      if (lane == 0) y[((size_t)b * COUT + o) * HW + p] = __float2half_rn(s_ + __half2float(bias[o]));
    }
  }
}

template <int COUT>
static int conv_out_launch(const __half* x, const __half* w, const __half* bias, __half* y, int B, int H, int W, int C,
                           cudaStream_t st) {
  const size_t smem = (size_t)COUT * 9 * C * 2;
  TG_SMEM_ATTR_ONCE(conv_out_kernel<COUT>, 96 * 1024);
  TG_CHECK(smem <= 96 * 1024, "conv_out: C=%d too large", C);
  const int total = B * H * W;
  TG_CUDA(launch_kernel(conv_out_kernel<COUT>, dim3((total + CONV_OUT_PIX - 1) / CONV_OUT_PIX), dim3(256), smem, st, 1, x, w, bias, y, B, H, W, C));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// Cout = 4: the UNet's conv_out (eps); Cout = 3: the VAE decoder's conv_out (RGB)
int conv_out_nhwc_to_nchw(const __half* x, const __half* w, const __half* bias, __half* y, int B, int H, int W, int C,
                          int Cout, cudaStream_t st) {
  TG_CHECK(C % 2 == 0 && C <= 320, "conv_out: C=%d (even, <= 320)", C);
  if (Cout == 4) return conv_out_launch<4>(x, w, bias, y, B, H, W, C, st);
  if (Cout == 3) return conv_out_launch<3>(x, w, bias, y, B, H, W, C, st);
  TG_CHECK(false, "conv_out: Cout=%d unsupported (3 or 4)", Cout);
}

// ---------------------------------------------------------------------------------------------------------
// 1x1 convolution between tiny channel counts on an NCHW tensor (the VAE's post_quant_conv, 4 -> 4, on the latents):
// y[b, o, p] = bias[o] + sum_c w[o, c] x[b, c, p]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pointwise_nchw_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                      __half* __restrict__ y, int Cin, int Cout, int HW, int total) {
  grid_dep_launch();
  grid_dep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i % HW, o = (i / HW) % Cout, b = i / (HW * Cout);
    float acc = __half2float(bias[o]);
    for (int c = 0; c < Cin; ++c) acc = fmaf(__half2float(w[o * Cin + c]), __half2float(x[((size_t)b * Cin + c) * HW + p]), acc);
    y[i] = __float2half_rn(acc);
  }
}
int pointwise_nchw(const __half* x, const __half* w, const __half* bias, __half* y, int B, int Cin, int Cout, int HW,
                   cudaStream_t st) {
  TG_CHECK(Cin >= 1 && Cin <= 16 && Cout >= 1 && Cout <= 16, "pointwise: Cin=%d Cout=%d (1..16)", Cin, Cout);
  const int total = B * Cout * HW;
  TG_CUDA(launch_kernel(pointwise_nchw_kernel, dim3((total + 255) / 256), dim3(256), 0, st, 1, x, w, bias, y, Cin, Cout, HW, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Token + position embedding of the CLIP text encoder: out[b][t][:] = table[ids[b][t]][:] + pos[t][:] (fp32 add, one
// rounding).  One warp per token row, 16-byte accesses; ids outside [0, vocab) are a caller error and read row 0 after the
// device-side flag is raised (checked by the host wrapper's launch check in debug runs).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ ids, const __half* __restrict__ table, const __half* __restrict__ pos,
                    __half* __restrict__ out, int rows, int T, int C, int vocab) {
  grid_dep_launch();
  grid_dep_wait();
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    long long id = ids[row];
    if (id < 0 || id >= vocab) id = 0;
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * C);
    const uint4* ps = reinterpret_cast<const uint4*>(pos + (size_t)(row % T) * C);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * C);
    for (int i = lane; i < C / 8; i += 32) {
      uint4 a = __ldg(src + i), p = __ldg(ps + i), o;
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* ph = reinterpret_cast<const __half2*>(&p);
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = __half22float2(ah[j]), y = __half22float2(ph[j]);
        oh[j] = __floats2half2_rn(x.x + y.x, x.y + y.y);
      }
      dst[i] = o;
    }
  }
}
int embed_tokens(const long long* ids, const __half* table, const __half* pos, __half* out, int B, int T, int C,
                 int vocab, cudaStream_t st) {
  TG_CHECK(C % 8 == 0 && B > 0 && T > 0 && vocab > 0, "embed_tokens: B=%d T=%d C=%d vocab=%d", B, T, C, vocab);
  const int rows = B * T;
  const int grid = (rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8;
  TG_CUDA(launch_kernel(embed_tokens_kernel, dim3(grid), dim3(256), 0, st, 1, ids, table, pos, out, rows, T, C, vocab));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Non-overlapping patch embedding as a GEMM (CLIP ViT: Conv2d(3, width, kernel = stride = P, bias = False)):
// out[(b, py, px)][k] = x[b][c][py*P + ky][px*P + kx] with k = c*P*P + ky*P + kx (a torch [width][3][P][P] weight
// flattened), zero for k >= Cin*P*P (K padded to a multiple of 64).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_patch_kernel(const __half* __restrict__ x, __half* __restrict__ out, int Cin, int H, int W, int P, int Kpad,
                    size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  const int gw = W / P, gh = H / P, KK = Cin * P * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const size_t row = i / Kpad;
    const int px = (int)(row % gw), py = (int)((row / gw) % gh);
    const size_t b = row / ((size_t)gw * gh);
    __half v = __float2half(0.f);
    if (k < KK) {
      const int c = k / (P * P), r = k - c * P * P, ky = r / P, kx = r - ky * P;
      v = x[((b * Cin + c) * H + py * P + ky) * W + px * P + kx];
    }
    out[i] = v;
  }
}
int im2col_patch(const __half* x, __half* out, int B, int Cin, int H, int W, int P, int Kpad, cudaStream_t st) {
  TG_CHECK(H % P == 0 && W % P == 0 && Kpad % 64 == 0 && Kpad >= Cin * P * P, "im2col_patch: H=%d W=%d P=%d Kpad=%d", H, W, P, Kpad);
  const size_t total = (size_t)B * (H / P) * (W / P) * Kpad;
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(im2col_patch_kernel, dim3(grid), dim3(256), 0, st, 1, x, out, Cin, H, W, P, Kpad, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// diffusers DiagonalGaussianDistribution.sample on the VAE encoder's moments (NCHW [B, 2L, HW]: L mean channels then L
// log-variance channels): z = (mean + exp(0.5 clamp(logvar, -30, 20)) * noise) * out_scale      (models/pipelines.py:624-626)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gaussian_sample_kernel(const __half* __restrict__ moments, const __half* __restrict__ noise, __half* __restrict__ z, int L,
                       int HW, int total, float out_scale) {
  grid_dep_launch();
  grid_dep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i % HW, c = (i / HW) % L, b = i / (HW * L);
    const float mean = __half2float(moments[((size_t)b * 2 * L + c) * HW + p]);
    const float logvar = fminf(fmaxf(__half2float(moments[((size_t)b * 2 * L + L + c) * HW + p]), -30.f), 20.f);
    // the reference computes in fp16 tensors: std and the product are rounded to fp16 like there
    const float std_ = __half2float(__float2half_rn(expf(0.5f * logvar)));
    const float n = noise ? __half2float(noise[i]) : 0.f;
    z[i] = __float2half_rn(__half2float(__float2half_rn(mean + __half2float(__float2half_rn(std_ * n)))) * out_scale);
  }
}
int gaussian_sample(const __half* moments, const __half* noise, __half* z, int B, int L, int HW, float out_scale,
                    cudaStream_t st) {
  const int total = B * L * HW;
  TG_CUDA(launch_kernel(gaussian_sample_kernel, dim3((total + 255) / 256), dim3(256), 0, st, 1, moments, noise, z, L, HW, total, out_scale));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Row softmax: out[r, :] = softmax(scale * x[r, :]) in fp32 (the VAE mid-block attention: one head of 512 channels over
// 4096 tokens is computed as GEMM -> this -> GEMM; diffusers upcasts that softmax).  One CTA per row, N <= 8192.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const __half* __restrict__ x, int ldx, __half* __restrict__ out, int ldo, int N, float scale_log2) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float s_red[8];
  const __half2* px = reinterpret_cast<const __half2*>(x + (size_t)blockIdx.x * ldx);
  __half2* po = reinterpret_cast<__half2*>(out + (size_t)blockIdx.x * ldo);
  const int n2 = N / 2, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 v[16];  // up to 8192 columns
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < n2) {
      v[i] = __half22float2(px[c]);
      mx = fmaxf(mx, fmaxf(v[i].x, v[i].y));
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < n2) {
      v[i].x = exp2f((v[i].x - mx) * scale_log2);
      v[i].y = exp2f((v[i].y - mx) * scale_log2);
      sum += v[i].x + v[i].y;
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += s_red[i];
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < n2) po[c] = __floats2half2_rn(v[i].x * inv, v[i].y * inv);
  }
}
int softmax_rows(const __half* x, int ldx, __half* out, int ldo, int M, int N, float scale, cudaStream_t st) {
  TG_CHECK(N % 2 == 0 && N <= 8192 && ldx % 2 == 0 && ldo % 2 == 0 && scale > 0.f, "softmax_rows: N=%d (even, <= 8192), scale > 0", N);
  TG_CUDA(launch_kernel(softmax_rows_kernel, dim3(M), dim3(256), 0, st, 1, x, ldx, out, ldo, N, scale * 1.4426950408889634f));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// im2col of a small-channel NCHW image for a 3x3 / pad 1 / stride 1 convolution done as a tensor-core GEMM:
// out[(b, y, x)][k] = x[b][ci][y + ky - 1][x + kx - 1] with k = ci * 9 + ky * 3 + kx (the flattening of a torch
// [Cout][Cin][3][3] weight), columns k >= Cin * 9 zero (K padded to 64).  Used for conv_in (Cin = 4 latents,
// ip_adapter/unet_2d_condition.py:999) and the first convolution of the ControlNet hint embedder (Cin = 3).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_small_kernel(const __half* __restrict__ x, uint4* __restrict__ out, int Cin, int H, int W, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  const int HW = H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);  // 8 columns of the 64
    const size_t pix = i >> 3;
    const int p = (int)(pix % HW), py = p / W, px = p % W;
    const size_t b = pix / HW;
    uint4 v;
    __half* hv = reinterpret_cast<__half*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = chunk * 8 + j;
      const int ci = k / 9, tap = k - ci * 9;
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      hv[j] = (ci < Cin && yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[((b * Cin + ci) * H + yy) * W + xx] : __float2half(0.f);
    }
    out[i] = v;
  }
}
int im2col_small_nchw(const __half* x, __half* out, int B, int Cin, int H, int W, cudaStream_t st) {
  TG_CHECK(Cin * 9 <= 64, "im2col_small: Cin=%d (Cin * 9 must fit 64 columns)", Cin);
  const size_t total = (size_t)B * H * W * 8;
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(im2col_small_kernel, dim3(grid), dim3(256), 0, st, 1, x, reinterpret_cast<uint4*>(out), Cin, H, W, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Timesteps(320, flip_sin_to_cos=True, shift=0): out[b] = [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i / half)
// computed in fp32 then cast to fp16 (ip_adapter/unet_2d_condition.py:897-902).  t is read from device memory so a
// captured CUDA graph can be replayed for every timestep.
// ---------------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, __half* __restrict__ out, int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x, half_dim = dim / 2;
  const float tv = t[b];
  for (int i = threadIdx.x; i < half_dim; i += blockDim.x) {
    const float f = expf(-9.210340371976184f * (float)i / (float)half_dim);
    const float e = tv * f;
    out[(size_t)b * dim + i] = __float2half_rn(cosf(e));
    out[(size_t)b * dim + half_dim + i] = __float2half_rn(sinf(e));
  }
}

int timestep_embedding(const float* t_dev, __half* out, int B, int dim, cudaStream_t st) {
  TG_CUDA(launch_kernel(timestep_embed_kernel, dim3(B), dim3(160), 0, st, 1, t_dev, out, dim));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// y += x (ControlNet residual adds, ip_adapter/unet_2d_condition.py:1044-1053, 1077-1078)
__global__ void __launch_bounds__(256) add_inplace_kernel(__half2* __restrict__ y, const __half2* __restrict__ x, size_t n2) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __hadd2(y[i], x[i]);
}
int add_inplace(__half* y, const __half* x, size_t n, cudaStream_t st) {
  TG_CHECK(n % 2 == 0, "add: odd length");
  const int grid = (int)(((n / 2 + 255) / 256) < (size_t)(148 * 8) ? ((n / 2 + 255) / 256) : (size_t)(148 * 8));
  TG_CUDA(launch_kernel(add_inplace_kernel, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<__half2*>(y), reinterpret_cast<const __half2*>(x), n / 2));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// NCHW <-> NHWC for the (tiny) ControlNet residual interface
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int C, int HW, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int hw = (int)(p % HW);
    const size_t b = p / HW;
    y[i] = x[(b * C + c) * HW + hw];
  }
}
int nchw_to_nhwc(const __half* x, __half* y, int B, int C, int HW, cudaStream_t st) {
  const size_t total = (size_t)B * C * HW;
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 8) ? ((total + 255) / 256) : (size_t)(148 * 8));
  TG_CUDA(launch_kernel(nchw_to_nhwc_kernel, dim3(grid), dim3(256), 0, st, 1, x, y, C, HW, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// First Cout (<= 8) channels of a channels-last tensor [B,HW,Cs] as NCHW planes [B,Cout,HW]: the tail of conv_out when it
// runs on the tensor-core convolution with its 4 output channels padded to a 64-column tile
// (ip_adapter/unet_2d_condition.py:1120).  One thread per pixel: one 16-byte read, Cout coalesced plane writes.
__global__ void __launch_bounds__(256)
nhwc_head_to_nchw_kernel(const __half* __restrict__ x, __half* __restrict__ y, int Cs, int Cout, int HW, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / HW;
    const int hw = (int)(i - b * HW);
    const uint4 v = *reinterpret_cast<const uint4*>(x + i * Cs);
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < Cout) y[(b * Cout + c) * HW + hw] = h[c];
  }
}
int nhwc_head_to_nchw(const __half* x, __half* y, int B, int Cs, int Cout, int HW, cudaStream_t st) {
  TG_CHECK(Cs % 8 == 0 && Cout >= 1 && Cout <= 8 && Cout <= Cs, "nhwc_head_to_nchw: Cs=%d Cout=%d", Cs, Cout);
  const size_t total = (size_t)B * HW;
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 8) ? ((total + 255) / 256) : (size_t)(148 * 8));
  TG_CUDA(launch_kernel(nhwc_head_to_nchw_kernel, dim3(grid), dim3(256), 0, st, 1, x, y, Cs, Cout, HW, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + DDIM step + masked composite, one launch per denoise step
//   eps    = eps_u + g (eps_c - eps_u)                                  (models/pipelines.py:441-442, 820-821)
//   x0     = (x - sqrt(1-a_t) eps) / sqrt(a_t);  x' = sqrt(a_prev) x0 + sqrt(1-a_prev) eps     (DDIM, eta = 0)
//   x'     = frozen * m + x' * (1 - m)     when frozen != null          (models/pipelines.py:833-834)
// eps is [2K,4,H,W] (first K rows unconditional), latents [K,4,H,W]; coef = {g, sqrt(a_t), sqrt(1-a_t),
// sqrt(a_prev), sqrt(1-a_prev)} lives in device memory (graph replay).  mask is [H,W] broadcast over K and C.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cfg_ddim_kernel(const __half* __restrict__ eps, __half* __restrict__ lat, const float* __restrict__ coef,
                const __half* __restrict__ frozen, const __half* __restrict__ mask, int K, int CHW, int HW,
                __half* __restrict__ lat_copy, int pair_stride) {
  grid_dep_launch();
  grid_dep_wait();
  const float g = coef[0], sa = coef[1], s1a = coef[2], sap = coef[3], s1ap = coef[4];
  const int total = K * CHW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / CHW, r = i % CHW;
    const float eu = __half2float(eps[(size_t)k * CHW + r]);
    const float ec = __half2float(eps[(size_t)(pair_stride + k) * CHW + r]);
    // every intermediate of the reference is an fp16 tensor (fp32 python/0-d scalars do not promote)
    const float e = __half2float(__float2half_rn(eu + g * (ec - eu)));
    const float x = __half2float(lat[i]);
    const float x0 = (x - s1a * e) / sa;
    float xn = sap * x0 + s1ap * e;
    if (frozen != nullptr) {
      const float m = __half2float(mask[r % HW]);
      xn = __half2float(frozen[i]) * m + __half2float(__float2half_rn(xn)) * (1.f - m);
    }
    const __half o = __float2half_rn(xn);
    lat[i] = o;
    if (lat_copy != nullptr) lat_copy[i] = o;
  }
}

int cfg_ddim_composite(const __half* eps, __half* latents, const float* coef_dev, const __half* frozen,
                       const __half* mask, int K, int C, int HW, __half* latents_copy, int pair_stride,
                       cudaStream_t st) {
  if (pair_stride <= 0) pair_stride = K;
  const int total = K * C * HW;
  TG_CUDA(launch_kernel(cfg_ddim_kernel, dim3((total + 255) / 256), dim3(256), 0, st, 1, eps, latents, coef_dev, frozen, mask, K, C * HW, HW,
                        latents_copy, pair_stride));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// duplicate latents for CFG: x_in[2K] = cat([latents]*2) (models/pipelines.py:409, 747)
__global__ void __launch_bounds__(256)
dup_kernel(const __half* __restrict__ lat, __half* __restrict__ out, int n) {
  grid_dep_launch();
  grid_dep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const __half v = lat[i];
    out[i] = v;
    out[n + i] = v;
  }
}
int duplicate_latents(const __half* lat, __half* out, int n, cudaStream_t st) {
  TG_CUDA(launch_kernel(dup_kernel, dim3((n + 255) / 256), dim3(256), 0, st, 1, lat, out, n));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

}  // namespace tg
