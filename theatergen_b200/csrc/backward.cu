// Input-gradient kernels of the guidance step: d(energy)/d(latents) through the conditional UNet forward
// (latent_backward_guidance, models/pipelines.py:62-128: `torch.autograd.grad(loss, [latents])`).
//
// Only DATA gradients exist on this path (the weights are frozen), so every GEMM / convolution backward is the forward
// tensor-core kernel on a transposed / rotated weight copy; this file holds what is left: the backward of the
// normalisations, of the row softmax, of GEGLU, of the text + image-token cross-attention (query side only: keys and
// values come from the constant context), of nearest-x2 upsampling and of the stride-2 im2col, and the latent update.
// All tensors are fp16 with fp32 arithmetic inside a kernel, like the reference's fp16 autograd.
#include "common.cuh"
#include "ops.h"

namespace tg {

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();  // s_red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s_red[i];
  return t;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward (input gradient): xhat = (x - mu) rstd, g = dy gamma, dx = rstd (g - mean(g) - xhat mean(g xhat)).
// One warp per row; the row is re-read from L1/L2 for each of the four passes.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __half* __restrict__ x, int ldx, const __half* __restrict__ gamma, const __half* __restrict__ dy,
              int lddy, __half* __restrict__ dx, int lddx, int M, int C, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  const int lane = threadIdx.x & 31, c2 = C / 2;
  const __half2* g2 = reinterpret_cast<const __half2*>(gamma);
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += gridDim.x * 8) {
    const __half2* px = reinterpret_cast<const __half2*>(x + (size_t)row * ldx);
    const __half2* pd = reinterpret_cast<const __half2*>(dy + (size_t)row * lddy);
    __half2* po = reinterpret_cast<__half2*>(dx + (size_t)row * lddx);
    float s = 0.f;
    for (int i = lane; i < c2; i += 32) {
      const float2 v = __half22float2(px[i]);
      s += v.x + v.y;
    }
    const float mu = warp_sum(s) / C;
    s = 0.f;
    for (int i = lane; i < c2; i += 32) {
      const float2 v = __half22float2(px[i]);
      s += (v.x - mu) * (v.x - mu) + (v.y - mu) * (v.y - mu);
    }
    const float rstd = rsqrtf(warp_sum(s) / C + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < c2; i += 32) {
      const float2 v = __half22float2(px[i]), d = __half22float2(pd[i]), gm = __half22float2(g2[i]);
      const float g0 = d.x * gm.x, g1 = d.y * gm.y;
      s1 += g0 + g1;
      s2 += g0 * (v.x - mu) * rstd + g1 * (v.y - mu) * rstd;
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int i = lane; i < c2; i += 32) {
      const float2 v = __half22float2(px[i]), d = __half22float2(pd[i]), gm = __half22float2(g2[i]);
      const float h0 = (v.x - mu) * rstd, h1 = (v.y - mu) * rstd;
      po[i] = __floats2half2_rn(rstd * (d.x * gm.x - s1 - h0 * s2), rstd * (d.y * gm.y - s1 - h1 * s2));
    }
  }
}
int layernorm_bwd(const __half* x, int ldx, const __half* gamma, const __half* dy, int lddy, __half* dx, int lddx, int M,
                  int C, float eps, cudaStream_t st) {
  TG_CHECK(C % 2 == 0 && M > 0, "layernorm_bwd: M=%d C=%d", M, C);
  const int grid = (M + 7) / 8 < 148 * 8 ? (M + 7) / 8 : 148 * 8;
  TG_CUDA(launch_kernel(ln_bwd_kernel, dim3(grid), dim3(256), 0, st, 1, x, ldx, gamma, dy, lddy, dx, lddx, M, C, eps));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(32) (+SiLU) backward over channels-last [B,HW,C]: y = act(z), z = xhat gamma + beta.
//   dz = dy act'(z);  g = dz gamma;  dx = rstd (g - mean_grp(g) - xhat mean_grp(g xhat))
// One CTA per (image, group); its slice (HW x C/32 values, <= 320 KB) is re-read from L2 by each of the four passes.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
              const __half* __restrict__ dy, __half* __restrict__ dx, int HW, int C, float eps, int silu) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float s_red[8];
  const int grp = blockIdx.x, b = blockIdx.y, cpg = C / 32, p2 = cpg / 2;
  const size_t base = (size_t)b * HW * C + (size_t)grp * cpg;
  const int total = HW * p2;
  const float n = (float)HW * cpg;
  auto at = [&](const __half* t, int i) {
    const int pix = i / p2, pr = i - pix * p2;
    return __half22float2(*reinterpret_cast<const __half2*>(t + base + (size_t)pix * C + 2 * pr));
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const float2 v = at(x, i);
    s += v.x + v.y;
  }
  const float mu = block_sum_256(s, s_red) / n;
  s = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const float2 v = at(x, i);
    s += (v.x - mu) * (v.x - mu) + (v.y - mu) * (v.y - mu);
  }
  const float rstd = rsqrtf(block_sum_256(s, s_red) / n + eps);
  auto gval = [&](float xv, float dyv, float gm, float bt, float& xhat) {
    xhat = (xv - mu) * rstd;
    float dz = dyv;
    if (silu) {
      const float z = xhat * gm + bt, sg = 1.f / (1.f + __expf(-z));
      dz *= sg * (1.f + z * (1.f - sg));
    }
    return dz * gm;
  };
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int pr = i % p2;
    const float2 v = at(x, i), d = at(dy, i);
    const float2 gm = __half22float2(*reinterpret_cast<const __half2*>(gamma + grp * cpg + 2 * pr));
    const float2 bt = __half22float2(*reinterpret_cast<const __half2*>(beta + grp * cpg + 2 * pr));
    float h0, h1;
    const float g0 = gval(v.x, d.x, gm.x, bt.x, h0), g1 = gval(v.y, d.y, gm.y, bt.y, h1);
    s1 += g0 + g1;
    s2 += g0 * h0 + g1 * h1;
  }
  s1 = block_sum_256(s1, s_red) / n;
  s2 = block_sum_256(s2, s_red) / n;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int pix = i / p2, pr = i - pix * p2;
    const float2 v = at(x, i), d = at(dy, i);
    const float2 gm = __half22float2(*reinterpret_cast<const __half2*>(gamma + grp * cpg + 2 * pr));
    const float2 bt = __half22float2(*reinterpret_cast<const __half2*>(beta + grp * cpg + 2 * pr));
    float h0, h1;
    const float g0 = gval(v.x, d.x, gm.x, bt.x, h0), g1 = gval(v.y, d.y, gm.y, bt.y, h1);
    *reinterpret_cast<__half2*>(dx + base + (size_t)pix * C + 2 * pr) =
        __floats2half2_rn(rstd * (g0 - s1 - h0 * s2), rstd * (g1 - s1 - h1 * s2));
  }
}
int groupnorm_bwd(const __half* x, const __half* gamma, const __half* beta, const __half* dy, __half* dx, int B, int HW,
                  int C, float eps, int silu, cudaStream_t st) {
  TG_CHECK(C % 64 == 0 && B > 0 && HW > 0, "groupnorm_bwd: B=%d HW=%d C=%d (C must be a multiple of 64)", B, HW, C);
  TG_CUDA(launch_kernel(gn_bwd_kernel, dim3(32, B), dim3(256), 0, st, 1, x, gamma, beta, dy, dx, HW, C, eps, silu));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Row-softmax backward: dS = scale * P o (dP - sum_j P dP) for P = softmax(scale S); dS may alias dP.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const __half* __restrict__ p, int ldp, const __half* dp, int lddp, __half* ds, int ldds, int N,
                        float scale) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float s_red[8];
  const __half2* pp = reinterpret_cast<const __half2*>(p + (size_t)blockIdx.x * ldp);
  const __half2* pd = reinterpret_cast<const __half2*>(dp + (size_t)blockIdx.x * lddp);
  __half2* po = reinterpret_cast<__half2*>(ds + (size_t)blockIdx.x * ldds);
  const int n2 = N / 2;
  float2 pv[16], dv[16];  // up to 8192 columns
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < n2) {
      pv[i] = __half22float2(pp[c]);
      dv[i] = __half22float2(pd[c]);
      dot += pv[i].x * dv[i].x + pv[i].y * dv[i].y;
    }
  }
  dot = block_sum_256(dot, s_red);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < n2) po[c] = __floats2half2_rn(scale * pv[i].x * (dv[i].x - dot), scale * pv[i].y * (dv[i].y - dot));
  }
}
int softmax_bwd_rows(const __half* p, int ldp, const __half* dp, int lddp, __half* ds, int ldds, int M, int N,
                     float scale, cudaStream_t st) {
  TG_CHECK(N % 2 == 0 && N <= 8192 && M > 0, "softmax_bwd_rows: M=%d N=%d", M, N);
  TG_CUDA(launch_kernel(softmax_bwd_rows_kernel, dim3(M), dim3(256), 0, st, 1, p, ldp, dp, lddp, ds, ldds, N, scale));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// GEGLU backward on the packed pre-activation layout (per 64 columns: [32 value | 32 gate], bias included):
//   y = v gelu(g):  dv = dy gelu(g),  dg = dy v (Phi(g) + g phi(g))      (ip_adapter/attention.py:359-367)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __half* __restrict__ pre, const __half* __restrict__ dy, __half* __restrict__ dpre, int F,
                 size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / F;
    const int o = (int)(i - m * F), blk = o >> 5, j = o & 31;
    const size_t pv = m * 2 * F + (size_t)blk * 64 + j;
    const float v = __half2float(pre[pv]), g = __half2float(pre[pv + 32]), d = __half2float(dy[i]);
    const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
    dpre[pv] = __float2half_rn(d * g * cdf);
    dpre[pv + 32] = __float2half_rn(d * v * (cdf + g * pdf));
  }
}
int geglu_bwd(const __half* pre, const __half* dy, __half* dpre, int M, int F, cudaStream_t st) {
  TG_CHECK(F % 32 == 0 && M > 0, "geglu_bwd: M=%d F=%d", M, F);
  const size_t total = (size_t)M * F;
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(geglu_bwd_kernel, dim3(grid), dim3(256), 0, st, 1, pre, dy, dpre, F, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Cross-attention backward, query side (IPAttnProcessor, ip_adapter/attention_processor.py:478-516):
//   out = softmax(a q Kt^T) Vt + ip_scale * softmax(a q Ki^T) Vi
// Keys / values are projections of the constant context, so only dq is needed.  Per (query, head), one warp:
//   p = softmax(a q.k_j);  dp_j = dout.v_j (+ dprobs_j, the energy's gradient on the captured text map);
//   ds_j = p_j (dp_j - sum p dp);  dq = a sum_j ds_j k_j         (both branches, the image one scaled by ip_scale)
// K/V rows are read from L1/L2 (81 x C x 2 B per tensor at most); lanes own head-dim elements lane, lane+32, ...
// ---------------------------------------------------------------------------------------------------------
template <int DPL>  // head-dim elements per lane: d <= 32 * DPL
__global__ void __launch_bounds__(256)
cross_ip_bwd_q_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ kt, const __half* __restrict__ vt,
                      const __half* __restrict__ ki, const __half* __restrict__ vi, int ldkv, int n_text, int n_ip,
                      float ip_scale, const __half* __restrict__ dout, int lddo, const __half* __restrict__ dprobs,
                      __half* __restrict__ dq, int lddq, int Nq, int heads, int d, float alpha) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float s_s[8][96];  // scores -> probabilities -> ds of this warp's (query, head)
  __shared__ float s_d[8][96];  // dp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + warp;
  if (item >= Nq * heads) return;  // whole warps leave together; only __syncwarp below
  const int n = item / heads, h = item - n * heads;
  float qv[DPL], dv[DPL], acc[DPL];
#pragma unroll
  for (int e = 0; e < DPL; ++e) {
    const int c = lane + 32 * e;
    qv[e] = c < d ? __half2float(q[(size_t)n * ldq + h * d + c]) : 0.f;
    dv[e] = c < d ? __half2float(dout[(size_t)n * lddo + h * d + c]) : 0.f;
    acc[e] = 0.f;
  }
  float* ss = s_s[warp];
  float* sd = s_d[warp];
  for (int branch = 0; branch < 2; ++branch) {
    const int nk = branch ? n_ip : n_text;
    if (nk == 0) continue;
    const __half* K = (branch ? ki : kt) + (size_t)(branch ? n_text : 0) * ldkv + h * d;  // image tokens follow the text
    const __half* V = (branch ? vi : vt) + (size_t)(branch ? n_text : 0) * ldkv + h * d;
    const float bscale = branch ? ip_scale : 1.f;
    float mx = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      float s = 0.f, t = 0.f;
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const int c = lane + 32 * e;
        if (c < d) {
          s += qv[e] * __half2float(K[(size_t)j * ldkv + c]);
          t += dv[e] * __half2float(V[(size_t)j * ldkv + c]);
        }
      }
      s = warp_sum(s) * alpha;
      t = warp_sum(t) * bscale;
      if (!branch && dprobs) t += __half2float(dprobs[((size_t)h * Nq + n) * n_text + j]);
      if (lane == 0) {
        ss[j] = s;
        sd[j] = t;
      }
      mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float e = __expf(ss[j] - mx);
      ss[j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    float dot = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float pj = ss[j] * inv;
      ss[j] = pj;
      dot += pj * sd[j];
    }
    dot = warp_sum(dot);
    for (int j = lane; j < nk; j += 32) ss[j] = ss[j] * (sd[j] - dot);
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const float dsj = ss[j];
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const int c = lane + 32 * e;
        if (c < d) acc[e] += dsj * __half2float(K[(size_t)j * ldkv + c]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int e = 0; e < DPL; ++e) {
    const int c = lane + 32 * e;
    if (c < d) dq[(size_t)n * lddq + h * d + c] = __float2half_rn(alpha * acc[e]);
  }
}
int cross_ip_bwd_q(const __half* q, int ldq, const __half* kt, const __half* vt, const __half* ki, const __half* vi,
                   int ldkv, int n_text, int n_ip, float ip_scale, const __half* dout, int lddo, const __half* dprobs,
                   __half* dq, int lddq, int Nq, int heads, int d, float alpha, cudaStream_t st) {
  TG_CHECK(n_text > 0 && n_text <= 96 && n_ip >= 0 && n_ip <= 96 && d > 0 && d <= 160 && Nq > 0,
           "cross_ip_bwd_q: n_text=%d n_ip=%d d=%d Nq=%d", n_text, n_ip, d, Nq);
  TG_CHECK(n_ip == 0 || (ki && vi), "cross_ip_bwd_q: image-token keys / values missing");
  const int grid = (Nq * heads + 7) / 8;
  if (d <= 64) {
    TG_CUDA(launch_kernel(cross_ip_bwd_q_kernel<2>, dim3(grid), dim3(256), 0, st, 1, q, ldq, kt, vt, ki, vi, ldkv, n_text,
                          n_ip, ip_scale, dout, lddo, dprobs, dq, lddq, Nq, heads, d, alpha));
  } else {
    TG_CUDA(launch_kernel(cross_ip_bwd_q_kernel<5>, dim3(grid), dim3(256), 0, st, 1, q, ldq, kt, vt, ki, vi, ldkv, n_text,
                          n_ip, ip_scale, dout, lddo, dprobs, dq, lddq, Nq, heads, d, alpha));
  }
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// nearest-x2 upsampling backward: dx[b][y][x][c] = sum of the 2x2 block of dy (fp32 sum, one rounding)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const __half2* __restrict__ dy, __half2* __restrict__ dx, int H, int W, int C2, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C2);
    size_t p = i / C2;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const size_t b = p / H;
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 v = __half22float2(dy[((b * 2 * H + 2 * y + (t >> 1)) * 2 * W + 2 * x + (t & 1)) * C2 + c]);
      a.x += v.x;
      a.y += v.y;
    }
    dx[i] = __floats2half2_rn(a.x, a.y);
  }
}
int upsample2x_bwd(const __half* dy, int B, int H, int W, int C, __half* dx, cudaStream_t st) {
  TG_CHECK(C % 2 == 0, "upsample2x_bwd: C %% 2 != 0");
  const size_t total = (size_t)B * H * W * (C / 2);
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(upsample2x_bwd_kernel, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const __half2*>(dy),
                        reinterpret_cast<__half2*>(dx), H, W, C / 2, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the stride-2 im2col (misc.cu, pad_lo = 1): dx[b][y][x][c] = sum over the taps (ky,kx) whose window covers
// (y,x): dcols[b*Ho*Wo + oy*Wo + ox][(ky*3+kx)*C + c] with 2oy + ky - 1 = y, 2ox + kx - 1 = x.  Gather form, no atomics.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
col2im_s2_kernel(const __half2* __restrict__ dcols, __half2* __restrict__ dx, int H, int W, int C2, size_t total) {
  grid_dep_launch();
  grid_dep_wait();
  const int Ho = H / 2, Wo = W / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C2);
    size_t p = i / C2;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const size_t b = p / H;
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
        const size_t row = (b * Ho + (ty >> 1)) * Wo + (tx >> 1);
        const float2 v = __half22float2(dcols[(row * 9 + ky * 3 + kx) * C2 + c]);
        a.x += v.x;
        a.y += v.y;
      }
    }
    dx[i] = __floats2half2_rn(a.x, a.y);
  }
}
int col2im_s2(const __half* dcols, int B, int H, int W, int C, __half* dx, cudaStream_t st) {
  TG_CHECK(C % 2 == 0 && H % 2 == 0 && W % 2 == 0, "col2im_s2: bad dims");
  const size_t total = (size_t)B * H * W * (C / 2);
  const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
  TG_CUDA(launch_kernel(col2im_s2_kernel, dim3(grid), dim3(256), 0, st, 1, reinterpret_cast<const __half2*>(dcols),
                        reinterpret_cast<__half2*>(dx), H, W, C / 2, total));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// The guidance update (models/pipelines.py:106-115, DDIM branch): latents -= (grad / grad_scale) * step, fp32 math;
// grad is the fp16 output of the backward pass, carried at grad_scale times its true value against underflow.
// Non-finite gradients (an fp16 overflow upstream) raise the flag instead of poisoning the latents.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
guidance_update_kernel(__half* __restrict__ latents, const __half* __restrict__ grad, float inv_scale, float step,
                       float* __restrict__ grad_out, int* __restrict__ bad, int n) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = __half2float(grad[i]) * inv_scale;
  if (!isfinite(g)) {
    if (bad) atomicExch(bad, 1);
    return;
  }
  if (grad_out) grad_out[i] = g;
  latents[i] = __float2half_rn(__half2float(latents[i]) - g * step);
}
int guidance_update(__half* latents, const __half* grad, float inv_scale, float step, float* grad_out, int* bad, int n,
                    cudaStream_t st) {
  TG_CHECK(n > 0, "guidance_update: n=%d", n);
  TG_CUDA(launch_kernel(guidance_update_kernel, dim3((n + 255) / 256), dim3(256), 0, st, 1, latents, grad, inv_scale, step,
                        grad_out, bad, n));
  TG_COUNT_LAUNCH();
  TG_LAUNCH_CHECK();
  return 0;
}

}  // namespace tg
