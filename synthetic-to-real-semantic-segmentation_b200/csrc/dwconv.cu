// Depthwise 3x3 convolution (forward, data gradient, weight gradient) on NHWC bf16, fused with
// the BatchNorm + ReLU6 that precedes it and with the block's zero padding.
//
// Replaces, for one InvertedResidual (modeling/backbone/mobilenet.py:26-68 of the reference):
//   fixed_padding (:17-23,62)  ->  BN(hidden) + ReLU6 (:51-52)  ->  groups=hidden Conv2d (:40,54)
// The reference pads the block INPUT, so the 1x1 expand conv, its BN and ReLU6 run on the padded
// tensor: the depthwise conv (padding=0) sees relu6(shift[c]) in the halo, not zero, and the BN
// backward sums run over the padded domain.  Here the expand output stays unpadded in HBM; the
// BN+ReLU6 is a load prologue and the halo constant is synthesised per channel.
//
// HBM-bound (0.53 GFLOP vs 146 MB per 512x1024 image).  One thread owns 8 channels (16 B) of a
// pixel, channel-group fastest, so a warp touches contiguous NHWC memory; neighbouring taps hit
// L1/L2.  The [C][3][3] fp32 filter is staged once per CTA in shared memory as [9][C].  The
// per-channel BN statistics of the output (forward) and the BN-backward sums of the input
// (data gradient) are accumulated in registers and leave the CTA as one fp64 atomic per channel.
#include "common.cuh"

namespace {

struct DwCfg {
  int cg, rows, threads, grid;
};

inline DwCfg dw_cfg(long long P, int C, int min_pix_per_thread) {
  DwCfg c;
  c.cg = C / 8;
  c.rows = 256 / c.cg;
  if (c.rows < 1) c.rows = 1;
  c.threads = c.cg * c.rows;
  long long blocks = (P + (long long)c.rows * min_pix_per_thread - 1) / ((long long)c.rows * min_pix_per_thread);
  const long long cap = (long long)s2r_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  c.grid = (int)blocks;
  return c;
}

__device__ __forceinline__ void stage_filter(const float* __restrict__ w, float* sw, int C) {
  // sw[k][c] = w[c][k]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) {
    const int c = i / 9, k = i - c * 9;
    sw[k * C + c] = __ldg(w + i);
  }
  __syncthreads();
}

struct Prologue {
  float sc[8], sh[8], halo[8];
  bool on;
  int act;
};

__device__ __forceinline__ void prologue_init(Prologue& pr, const float* __restrict__ scale_shift, int C,
                                              int g, int act, int halo_const) {
  pr.on = scale_shift != nullptr;
  pr.act = act;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    pr.sc[i] = pr.on ? __ldg(scale_shift + g * 8 + i) : 1.f;
    pr.sh[i] = pr.on ? __ldg(scale_shift + C + g * 8 + i) : 0.f;
    pr.halo[i] = (pr.on && halo_const) ? apply_act(pr.sh[i], act, 0.f) : 0.f;
  }
}

// activated input at (n, ih, iw); outside the tensor: the halo constant
__device__ __forceinline__ void load_in(const Prologue& pr, const __nv_bfloat16* __restrict__ x, int n,
                                        int ih, int iw, int H, int W, int C, int g, float* f) {
  if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
    bf16x8_to_float(ldg16(x + (((long long)n * H + ih) * W + iw) * C + g * 8), f);
    if (pr.on) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = apply_act(fmaf(f[i], pr.sc[i], pr.sh[i]), pr.act, 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = pr.halo[i];
  }
}

// block-level reduction of two 8-vectors per thread into fp64 global sums [2][C]
__device__ __forceinline__ void block_channel_reduce(float* red, const float* s, const float* q, int cg,
                                                     int g, int r, int rows, int C,
                                                     double* __restrict__ sums) {
  float* mine = red + ((size_t)r * cg + g) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[i] = s[i];
    mine[8 + i] = q[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
    const int gg = t / 16, k = t % 16;
    double acc = 0;
    for (int rr = 0; rr < rows; ++rr) acc += (double)red[((size_t)rr * cg + gg) * 16 + k];
    atomicAdd(&sums[(k >> 3) * C + gg * 8 + (k & 7)], acc);
  }
}

__global__ void __launch_bounds__(256)
dw_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift, int in_act,
              int halo_const, const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
              double* __restrict__ stats, int N, int H, int W, int C, int Ho, int Wo, int stride,
              int dil, int pad, int rows) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;             // [9][C]
  float* red = sm + 9 * C;    // [rows][cg][16]
  stage_filter(w, sw, C);
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  Prologue pr;
  prologue_init(pr, scale_shift, C, g, in_act, halo_const);
  const float* wg = sw + g * 8;  // tap k at wg[k * C + i]
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const long long P = (long long)N * Ho * Wo;
  for (long long p = (long long)blockIdx.x * rows + r; p < P; p += (long long)gridDim.x * rows) {
    const int ow = (int)(p % Wo);
    const long long t = p / Wo;
    const int oh = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh * stride + ky * dil - pad;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow * stride + kx * dil - pad;
        float f[8];
        load_in(pr, x, n, ih, iw, H, W, C, g, f);
        float wv[8];
        *reinterpret_cast<float4*>(wv) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * C);
        *reinterpret_cast<float4*>(wv + 4) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * C + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wv[i], acc[i]);
      }
    }
    *reinterpret_cast<uint4*>(y + p * C + g * 8) = float_to_bf16x8(acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += acc[i]; q[i] += acc[i] * acc[i]; }
  }
  if (stats) block_channel_reduce(red, s, q, cg, g, r, rows, C, stats);
}

// g[n, ih+ext, iw+ext, c] = act'(pre) * sum_{ky,kx} dy[n,oh,ow,c] w[c,ky,kx],  ih = oh*s + ky*d - pad
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift,
                const float* __restrict__ mean_invstd, int in_act, int ext,
                __nv_bfloat16* __restrict__ gout, double* __restrict__ bsums, int N, int H, int W,
                int C, int Ho, int Wo, int stride, int dil, int pad, int rows) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;
  float* red = sm + 9 * C;
  stage_filter(w, sw, C);
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  Prologue pr;
  prologue_init(pr, scale_shift, C, g, in_act, 0);
  float mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean_invstd ? __ldg(mean_invstd + g * 8 + i) : 0.f;
    is[i] = mean_invstd ? __ldg(mean_invstd + C + g * 8 + i) : 0.f;
  }
  const float* wg = sw + g * 8;  // tap k at wg[k * C + i]
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const int He = H + 2 * ext, We = W + 2 * ext;
  const long long P = (long long)N * He * We;
  for (long long p = (long long)blockIdx.x * rows + r; p < P; p += (long long)gridDim.x * rows) {
    const int we = (int)(p % We);
    const long long t = p / We;
    const int he = (int)(t % He);
    const int n = (int)(t / He);
    const int ih = he - ext, iw = we - ext;
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
        float wv[8];
        *reinterpret_cast<float4*>(wv) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * C);
        *reinterpret_cast<float4*>(wv + 4) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * C + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wv[i], acc[i]);
      }
    }
    if (pr.on) {
      float xv[8];
      const bool inside = (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
      if (inside) {
        bf16x8_to_float(ldg16(x + (((long long)n * H + ih) * W + iw) * C + g * 8), xv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float pre = fmaf(xv[i], pr.sc[i], pr.sh[i]);
        acc[i] *= act_grad(pre, pr.act, 0.f);
        s[i] += acc[i];
        q[i] += acc[i] * (xv[i] - mu[i]) * is[i];
      }
    }
    *reinterpret_cast<uint4*>(gout + p * C + g * 8) = float_to_bf16x8(acc);
  }
  if (bsums) block_channel_reduce(red, s, q, cg, g, r, rows, C, bsums);
}

// dw[c][k] += sum_{n,oh,ow} dy[n,oh,ow,c] in(n, oh*s + ky*d - pad, ow*s + kx*d - pad, c)
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift, int in_act,
                int halo_const, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int N,
                int H, int W, int C, int Ho, int Wo, int stride, int dil, int pad, int rows) {
  extern __shared__ __align__(16) float sm[];  // [rows][cg][24]
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  Prologue pr;
  prologue_init(pr, scale_shift, C, g, in_act, halo_const);
  float acc[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  const long long P = (long long)N * Ho * Wo;
  for (long long p = (long long)blockIdx.x * rows + r; p < P; p += (long long)gridDim.x * rows) {
    const int ow = (int)(p % Wo);
    const long long t = p / Wo;
    const int oh = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float gd[8];
    bf16x8_to_float(ldg16(dy + p * C + g * 8), gd);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh * stride + ky * dil - pad;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow * stride + kx * dil - pad;
        float f[8];
        load_in(pr, x, n, ih, iw, H, W, C, g, f);
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
      float sacc = 0.f;
      for (int rr = 0; rr < rows; ++rr) sacc += sm[((size_t)rr * cg + gg) * 24 + j];
      const int c = gg * 8 + (j & 7), k = round * 3 + (j >> 3);
      atomicAdd(&dw[c * 9 + k], sacc);
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
  S2R_REQUIRE(C <= 1024, S2R_ERR_UNSUPPORTED, "dwconv3x3: C=%d too large for the staged filter", C);
  S2R_REQUIRE(a && b && ((uintptr_t)a | (uintptr_t)b) % 16 == 0, S2R_ERR_SHAPE, "dwconv3x3: null or unaligned tensor");
  const int he = H + 2 * pad - 2 * dil - 1, we = W + 2 * pad - 2 * dil - 1;
  S2R_REQUIRE(he >= 0 && we >= 0, S2R_ERR_SHAPE, "dwconv3x3: input smaller than the dilated filter");
  *Ho = he / stride + 1;
  *Wo = we / stride + 1;
  return S2R_OK;
}

template <typename K>
inline int dw_smem_attr(K kernel, size_t smem) {
  if (smem > 48 * 1024) S2R_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return S2R_OK;
}

}  // namespace

extern "C" int s2r_dwconv3x3_fwd(const void* x, const float* in_scale_shift, int in_act, int halo_const,
                                 const float* w, void* y, double* stats, int N, int H, int W, int C,
                                 int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(x, y, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  const DwCfg cfg = dw_cfg((long long)N * Ho * Wo, C, 8);
  const size_t smem = ((size_t)9 * C + (size_t)cfg.rows * cfg.cg * 16) * sizeof(float);
  rc = dw_smem_attr(dw_fwd_kernel, smem);
  if (rc) return rc;
  dw_fwd_kernel<<<cfg.grid, cfg.threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, in_scale_shift, in_act, halo_const, w, (__nv_bfloat16*)y, stats, N, H,
      W, C, Ho, Wo, stride, dil, pad, cfg.rows);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_dgrad(const void* dy, const float* w, const void* x,
                                   const float* in_scale_shift, const float* in_mean_invstd,
                                   int in_act, int ext, void* g, double* bwd_sums, int N, int H, int W,
                                   int C, int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(dy, g, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  S2R_REQUIRE(ext >= 0 && ext <= pad, S2R_ERR_SHAPE, "dwconv3x3_dgrad: ext=%d outside [0,pad]", ext);
  S2R_REQUIRE(!in_scale_shift || (x && (uintptr_t)x % 16 == 0 && in_mean_invstd), S2R_ERR_SHAPE,
              "dwconv3x3_dgrad: the masked variant needs x and mean/invstd");
  const DwCfg cfg = dw_cfg((long long)N * (H + 2 * ext) * (W + 2 * ext), C, 8);
  const size_t smem = ((size_t)9 * C + (size_t)cfg.rows * cfg.cg * 16) * sizeof(float);
  rc = dw_smem_attr(dw_dgrad_kernel, smem);
  if (rc) return rc;
  dw_dgrad_kernel<<<cfg.grid, cfg.threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, w, (const __nv_bfloat16*)x, in_scale_shift, in_mean_invstd, in_act, ext,
      (__nv_bfloat16*)g, bwd_sums, N, H, W, C, Ho, Wo, stride, dil, pad, cfg.rows);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_wgrad(const void* x, const float* in_scale_shift, int in_act,
                                   int halo_const, const void* dy, float* dw, int N, int H, int W,
                                   int C, int stride, int dil, int pad, s2r_stream_t stream) {
  int Ho, Wo;
  int rc = dw_check(x, dy, N, H, W, C, stride, dil, pad, &Ho, &Wo);
  if (rc) return rc;
  const DwCfg cfg = dw_cfg((long long)N * Ho * Wo, C, 64);
  const size_t smem = (size_t)cfg.rows * cfg.cg * 24 * sizeof(float);
  dw_wgrad_kernel<<<cfg.grid, cfg.threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, in_scale_shift, in_act, halo_const, (const __nv_bfloat16*)dy, dw, N, H,
      W, C, Ho, Wo, stride, dil, pad, cfg.rows);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
