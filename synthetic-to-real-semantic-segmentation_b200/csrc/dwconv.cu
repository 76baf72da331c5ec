// Depthwise 3x3 convolution (forward, data gradient, weight gradient) on NHWC bf16.
// Replaces the groups=hidden_dim nn.Conv2d of InvertedResidual
// (modeling/backbone/mobilenet.py:40,54 of the reference): stride 1|2, dilation d,
// applied to the already padded hidden tensor (padding 0), or with implicit zero
// padding for the expand_ratio==1 block.
//
// HBM-bound: 0.53 GFLOP vs 146 MB per 512x1024 image.  One thread owns a vector of
// 8 channels (16 B) of one pixel, channel-group fastest, so a warp reads/writes
// contiguous NHWC memory; the 3x3 taps of neighbouring pixels hit L1/L2.  The
// [C][3][3] fp32 filter is staged once per CTA in shared memory as [9][C].
#include "common.cuh"
#include "../../include/s2r_b200.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void stage_filter(const float* __restrict__ w, float* sw, int C) {
  // sw[k][c] = w[c][k]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) {
    const int c = i / 9, k = i - c * 9;
    sw[k * C + c] = __ldg(w + i);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
dw_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
              __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo, int stride,
              int dil, int pad) {
  extern __shared__ float sw[];
  stage_filter(w, sw, C);
  const int cg = C / 8;
  const long long total = (long long)N * Ho * Wo * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const int g = (int)(t % cg);
    long long p = t / cg;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh * stride + ky * dil - pad;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow * stride + kx * dil - pad;
        if (iw < 0 || iw >= W) continue;
        float f[8];
        bf16x8_to_float(ldg16(x + (((long long)n * H + ih) * W + iw) * C + g * 8), f);
        const float* wk = sw + (ky * 3 + kx) * C + g * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wk[i], acc[i]);
      }
    }
    *reinterpret_cast<uint4*>(y + t * 8) = float_to_bf16x8(acc);
  }
}

// dx[n,ih,iw,c] = sum_{ky,kx} dy[n,oh,ow,c] w[c,ky,kx] with ih = oh*s + ky*d - pad
__global__ void __launch_bounds__(kThreads)
dw_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo, int stride,
                int dil, int pad) {
  extern __shared__ float sw[];
  stage_filter(w, sw, C);
  const int cg = C / 8;
  const long long total = (long long)N * H * W * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const int g = (int)(t % cg);
    long long p = t / cg;
    const int iw = (int)(p % W);
    p /= W;
    const int ih = (int)(p % H);
    const int n = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hn = ih + pad - ky * dil;
      if (hn < 0 || hn % stride) continue;
      const int oh = hn / stride;
      if (oh >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int wn = iw + pad - kx * dil;
        if (wn < 0 || wn % stride) continue;
        const int ow = wn / stride;
        if (ow >= Wo) continue;
        float f[8];
        bf16x8_to_float(ldg16(dy + (((long long)n * Ho + oh) * Wo + ow) * C + g * 8), f);
        const float* wk = sw + (ky * 3 + kx) * C + g * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wk[i], acc[i]);
      }
    }
    *reinterpret_cast<uint4*>(dx + t * 8) = float_to_bf16x8(acc);
  }
}

// dw[c][k] += sum_{n,oh,ow} dy[n,oh,ow,c] x[n,ih,iw,c]
__global__ void dw_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                float* __restrict__ dw, int N, int H, int W, int C, int Ho, int Wo,
                                int stride, int dil, int pad, int rows) {
  extern __shared__ float sm[];  // [rows][cg][24]
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  float acc[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  const long long P = (long long)N * Ho * Wo;
  for (long long p = (long long)blockIdx.x * rows + r; p < P; p += (long long)gridDim.x * rows) {
    long long q = p;
    const int ow = (int)(q % Wo);
    q /= Wo;
    const int oh = (int)(q % Ho);
    const int n = (int)(q / Ho);
    float gd[8];
    bf16x8_to_float(ldg16(dy + p * C + g * 8), gd);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh * stride + ky * dil - pad;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow * stride + kx * dil - pad;
        if (iw < 0 || iw >= W) continue;
        float f[8];
        bf16x8_to_float(ldg16(x + (((long long)n * H + ih) * W + iw) * C + g * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[ky * 3 + kx][i] = fmaf(f[i], gd[i], acc[ky * 3 + kx][i]);
      }
    }
  }
  // three rounds of 3 taps keep the staging buffer at 96 B per thread
  for (int round = 0; round < 3; ++round) {
    float* mine = sm + ((size_t)r * cg + g) * 24;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[k * 8 + i] = acc[round * 3 + k][i];
    __syncthreads();
    for (int t = threadIdx.x; t < cg * 24; t += blockDim.x) {
      const int gg = t / 24, j = t % 24;
      float s = 0.f;
      for (int rr = 0; rr < rows; ++rr) s += sm[((size_t)rr * cg + gg) * 24 + j];
      const int c = gg * 8 + (j & 7), k = round * 3 + (j >> 3);
      atomicAdd(&dw[c * 9 + k], s);
    }
    __syncthreads();
  }
}

inline int dw_check(const void* a, const void* b, int N, int H, int W, int C, int stride, int dil,
                    int pad, int* Ho, int* Wo) {
  S2R_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, S2R_ERR_SHAPE,
              "dwconv3x3: bad shape N=%d H=%d W=%d C=%d (C must be a multiple of 8)", N, H, W, C);
  S2R_REQUIRE(stride == 1 || stride == 2, S2R_ERR_SHAPE, "dwconv3x3: stride %d not in {1,2}", stride);  // mobilenet.py:30
  S2R_REQUIRE(dil >= 1 && pad >= 0, S2R_ERR_SHAPE, "dwconv3x3: bad dilation/padding");
  S2R_REQUIRE(C * 9 * sizeof(float) <= 46 * 1024, S2R_ERR_UNSUPPORTED, "dwconv3x3: C=%d too large for the staged filter", C);
  S2R_REQUIRE(((uintptr_t)a | (uintptr_t)b) % 16 == 0, S2R_ERR_SHAPE, "dwconv3x3: unaligned tensor");
  const int he = H + 2 * pad - 2 * dil - 1, we = W + 2 * pad - 2 * dil - 1;
  S2R_REQUIRE(he >= 0 && we >= 0, S2R_ERR_SHAPE, "dwconv3x3: input smaller than the dilated filter");
  *Ho = he / stride + 1;
  *Wo = we / stride + 1;
  return S2R_OK;
}

}  // namespace

extern "C" int s2r_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C,
                                 int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(x, y, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  dw_fwd_kernel<<<s2r_grid(total, kThreads * 4, 8), kThreads, C * 9 * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, stride, dil, pad);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int H, int W, int C,
                                   int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(dy, dx, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  const long long total = (long long)N * H * W * (C / 8);
  dw_dgrad_kernel<<<s2r_grid(total, kThreads * 4, 8), kThreads, C * 9 * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo, stride, dil, pad);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C,
                                   int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(x, dy, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  const int cg = C / 8;
  int rows = 256 / cg;
  if (rows < 1) rows = 1;
  const long long P = (long long)N * Ho * Wo;
  long long blocks = (P + (long long)rows * 64 - 1) / ((long long)rows * 64);
  const long long cap = (long long)s2r_sm_count() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)rows * cg * 24 * sizeof(float);
  dw_wgrad_kernel<<<(int)blocks, rows * cg, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, N, H, W, C, Ho, Wo, stride, dil, pad, rows);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
