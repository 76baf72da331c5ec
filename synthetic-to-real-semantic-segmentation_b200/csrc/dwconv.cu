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
// HBM-bound (0.53 GFLOP vs 146 MB per 512x1024 image).  Work decomposition, all three kernels:
//   CTA   = one image x a block of ROWS output rows x a block of TW output columns x all channels
//   thread= one column x one group of 8 channels (16 B); channel group fastest, so a warp reads and
//           writes contiguous NHWC runs of TW*C*2 bytes
//   the thread walks DOWN its column keeping the 3x3 input window in registers: for the common
//   stride-1 / dilation-1 case only one new input row (3 vectors) is fetched and activated per
//   output row; horizontal neighbours are served by L1.
// No per-element 64-bit division: coordinates are block indices plus loop counters.
// The [C][3][3] fp32 filter is staged once per CTA in shared memory as [9][C].  Per-channel BN
// statistics of the output (forward) and BN-backward sums of the input (data gradient) are kept in
// registers per thread (its channel group is fixed) and leave the CTA as one fp64 atomic per channel.
#include <stdlib.h>

#include "bn_tail.cuh"

namespace {

constexpr int ROWS = 8;  // output rows per CTA

struct DwGeom {
  int N, H, W, C, Ho, Wo, stride, dil, pad;
};

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

// activated input vector at (ih, iw) of image `img` (pointer to its first element, channel group added)
__device__ __forceinline__ void load_in(const Prologue& pr, const __nv_bfloat16* __restrict__ img, int ih, int iw,
                                        int H, int W, int C, float* f) {
  if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
    bf16x8_to_float(ldg16(img + (ih * W + iw) * C), f);
    if (pr.on) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = apply_act(fmaf(f[i], pr.sc[i], pr.sh[i]), pr.act, 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = pr.halo[i];
  }
}

__device__ __forceinline__ void load_row3(const Prologue& pr, const __nv_bfloat16* __restrict__ img, int ih, int iw0,
                                          int dil, int H, int W, int C, float (*r)[8]) {
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) load_in(pr, img, ih, iw0 + kx * dil, H, W, C, r[kx]);
}

__device__ __forceinline__ void fma_row3(float* acc, const float (*r)[8], const float* wrow, int C) {
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    float wv[8];
    *reinterpret_cast<float4*>(wv) = *reinterpret_cast<const float4*>(wrow + kx * C);
    *reinterpret_cast<float4*>(wv + 4) = *reinterpret_cast<const float4*>(wrow + kx * C + 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(r[kx][i], wv[i], acc[i]);
  }
}

// block-level reduction of two 8-vectors per thread into fp64 global sums [2][C]; TW threads share a channel group
__device__ __forceinline__ void block_channel_reduce(float* red, const float* s, const float* q, int cg, int g,
                                                     int col, int tw, int C, double* __restrict__ sums) {
  float* mine = red + ((size_t)col * cg + g) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[i] = s[i];
    mine[8 + i] = q[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
    const int gg = t / 16, k = t % 16;
    float acc = 0.f;
    for (int cc = 0; cc < tw; ++cc) acc += red[((size_t)cc * cg + gg) * 16 + k];
    atomicAdd(&sums[(k >> 3) * C + gg * 8 + (k & 7)], (double)acc);
  }
}

// ------------------------------------------------------------------------------------ forward
template <bool SLIDE>  // SLIDE: stride 1, dilation 1 -> rolling 3-row window
__global__ void __launch_bounds__(256)
dw_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift, int in_act,
              int halo_const, const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
              double* __restrict__ stats, DwGeom G, int tw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;               // [9][C]
  float* red = sm + 9 * G.C;    // [tw][cg][16]
  stage_filter(w, sw, G.C);
  const int cg = G.C / 8;
  const int g = threadIdx.x % cg, col = threadIdx.x / cg;
  const int ow = blockIdx.x * tw + col;
  const int oh0 = blockIdx.y * ROWS;
  const int n = blockIdx.z;
  const bool active = ow < G.Wo;
  Prologue pr;
  prologue_init(pr, scale_shift, G.C, g, in_act, halo_const);
  const __nv_bfloat16* img = x + (size_t)n * G.H * G.W * G.C + g * 8;
  __nv_bfloat16* out = y + (size_t)n * G.Ho * G.Wo * G.C + g * 8;
  const float* wg = sw + g * 8;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (active) {
    const int iw0 = ow * G.stride - G.pad;
    const int rows = min(ROWS, G.Ho - oh0);
    if (SLIDE) {
      float win[3][3][8];
      int ih = oh0 - G.pad;
      load_row3(pr, img, ih, iw0, 1, G.H, G.W, G.C, win[0]);
      load_row3(pr, img, ih + 1, iw0, 1, G.H, G.W, G.C, win[1]);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (r < rows) {
          load_row3(pr, img, ih + r + 2, iw0, 1, G.H, G.W, G.C, win[(r + 2) % 3]);
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) fma_row3(acc, win[(r + ky) % 3], wg + ky * 3 * G.C, G.C);
          *reinterpret_cast<uint4*>(out + ((oh0 + r) * G.Wo + ow) * G.C) = float_to_bf16x8(acc);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += acc[i]; q[i] += acc[i] * acc[i]; }
        }
      }
    } else {
      for (int r = 0; r < rows; ++r) {
        const int ih0 = (oh0 + r) * G.stride - G.pad;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float row[3][8];
          load_row3(pr, img, ih0 + ky * G.dil, iw0, G.dil, G.H, G.W, G.C, row);
          fma_row3(acc, row, wg + ky * 3 * G.C, G.C);
        }
        *reinterpret_cast<uint4*>(out + ((oh0 + r) * G.Wo + ow) * G.C) = float_to_bf16x8(acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += acc[i]; q[i] += acc[i] * acc[i]; }
      }
    }
  }
  if (stats) block_channel_reduce(red, s, q, cg, g, col, tw, G.C, stats);
}

// ------------------------------------------------------------------------------------ data gradient
// g[n, ih+ext, iw+ext, c] = act'(pre) * sum_{ky,kx} dy[n,oh,ow,c] w[c,ky,kx],  ih = oh*s + ky*d - pad
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift,
                const float* __restrict__ mean_invstd, int in_act, int ext, int interior, __nv_bfloat16* __restrict__ gout,
                double* __restrict__ bsums, DwGeom G, int tw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;
  float* red = sm + 9 * G.C;
  stage_filter(w, sw, G.C);
  const int cg = G.C / 8;
  const int g = threadIdx.x % cg, col = threadIdx.x / cg;
  const int He = G.H + 2 * ext, We = G.W + 2 * ext;
  const int we = blockIdx.x * tw + col;
  const int he0 = blockIdx.y * ROWS;
  const int n = blockIdx.z;
  const bool active = we < We;
  Prologue pr;
  prologue_init(pr, scale_shift, G.C, g, in_act, 0);
  float mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean_invstd ? __ldg(mean_invstd + g * 8 + i) : 0.f;
    is[i] = mean_invstd ? __ldg(mean_invstd + G.C + g * 8 + i) : 0.f;
  }
  const __nv_bfloat16* dimg = dy + (size_t)n * G.Ho * G.Wo * G.C + g * 8;
  const __nv_bfloat16* ximg = x ? x + (size_t)n * G.H * G.W * G.C + g * 8 : nullptr;
  // interior layout: g is [N][H][W][C] and the border positions only feed the sums
  __nv_bfloat16* out = gout + (interior ? (size_t)n * G.H * G.W * G.C : (size_t)n * He * We * G.C) + g * 8;
  const float* wg = sw + g * 8;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (active) {
    const int iw = we - ext;
    // column taps: ow = (iw + pad - kx*dil) / stride when divisible and in range
    int owk[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int wn = iw + G.pad - kx * G.dil;
      owk[kx] = (wn >= 0 && wn % G.stride == 0 && wn / G.stride < G.Wo) ? wn / G.stride : -1;
    }
    const int rows = min(ROWS, He - he0);
    for (int r = 0; r < rows; ++r) {
      const int ih = he0 + r - ext;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hn = ih + G.pad - ky * G.dil;
        if (hn < 0 || hn % G.stride) continue;
        const int oh = hn / G.stride;
        if (oh >= G.Ho) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (owk[kx] < 0) continue;
          float f[8], wv[8];
          bf16x8_to_float(ldg16(dimg + (oh * G.Wo + owk[kx]) * G.C), f);
          *reinterpret_cast<float4*>(wv) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * G.C);
          *reinterpret_cast<float4*>(wv + 4) = *reinterpret_cast<const float4*>(wg + (ky * 3 + kx) * G.C + 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wv[i], acc[i]);
        }
      }
      if (pr.on) {
        float xv[8];
        if ((unsigned)ih < (unsigned)G.H && (unsigned)iw < (unsigned)G.W) {
          bf16x8_to_float(ldg16(ximg + (ih * G.W + iw) * G.C), xv);
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
      if (!interior)
        *reinterpret_cast<uint4*>(out + ((he0 + r) * We + we) * G.C) = float_to_bf16x8(acc);
      else if ((unsigned)ih < (unsigned)G.H && (unsigned)iw < (unsigned)G.W)
        *reinterpret_cast<uint4*>(out + (ih * G.W + iw) * G.C) = float_to_bf16x8(acc);
    }
  }
  if (bsums) block_channel_reduce(red, s, q, cg, g, col, tw, G.C, bsums);
}

// ------------------------------------------------------------------------------------ weight gradient
// dw[c][ky][kx] += sum_{n,oh,ow} dy[n,oh,ow,c] in(n, oh*s + ky*d - pad, ow*s + kx*d - pad, c)
// thread = (column, kernel row ky, channel group): 24 accumulators; a CTA walks WROWS output rows.
constexpr int WROWS = 64;

__global__ void __launch_bounds__(384)
dw_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift, int in_act,
                int halo_const, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, DwGeom G, int tw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];  // [tw][3*cg][24]
  const int cg = G.C / 8;
  const int g = threadIdx.x % cg;
  const int ky = (threadIdx.x / cg) % 3;
  const int col = threadIdx.x / (3 * cg);
  const int ow = blockIdx.x * tw + col;
  const int oh0 = blockIdx.y * WROWS;
  const int n = blockIdx.z;
  Prologue pr;
  prologue_init(pr, scale_shift, G.C, g, in_act, halo_const);
  const __nv_bfloat16* img = x + (size_t)n * G.H * G.W * G.C + g * 8;
  const __nv_bfloat16* dimg = dy + (size_t)n * G.Ho * G.Wo * G.C + g * 8;
  float acc[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  if (ow < G.Wo) {
    const int iw0 = ow * G.stride - G.pad;
    const int rows = min(WROWS, G.Ho - oh0);
    for (int r = 0; r < rows; ++r) {
      const int oh = oh0 + r;
      float gd[8], row[3][8];
      bf16x8_to_float(ldg16(dimg + (oh * G.Wo + ow) * G.C), gd);
      load_row3(pr, img, oh * G.stride + ky * G.dil - G.pad, iw0, G.dil, G.H, G.W, G.C, row);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[kx][i] = fmaf(row[kx][i], gd[i], acc[kx][i]);
    }
  }
  float* mine = sm + (size_t)threadIdx.x * 24;  // layout [col][ky][g][24]
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) mine[k * 8 + i] = acc[k][i];
  __syncthreads();
  // (ky, g, kx, i) columns summed over the tw thread columns
  for (int t = threadIdx.x; t < 3 * cg * 24; t += blockDim.x) {
    float sacc = 0.f;
    for (int cc = 0; cc < tw; ++cc) sacc += sm[(size_t)cc * 3 * cg * 24 + t];
    const int j = t % 24, rest = t / 24;
    const int gg = rest % cg, kyy = rest / cg;
    const int c = gg * 8 + (j & 7), kx = j >> 3;
    atomicAdd(&dw[c * 9 + kyy * 3 + kx], sacc);
  }
}

inline int dw_check(const void* a, const void* b, int N, int H, int W, int C, int stride, int dil,
                    int pad, DwGeom* G) {
  S2R_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, S2R_ERR_SHAPE,
              "dwconv3x3: bad shape N=%d H=%d W=%d C=%d (C must be a multiple of 8)", N, H, W, C);
  S2R_REQUIRE(stride == 1 || stride == 2, S2R_ERR_SHAPE, "dwconv3x3: stride %d not in {1,2}", stride);  // mobilenet.py:30
  S2R_REQUIRE(dil >= 1 && pad >= 0, S2R_ERR_SHAPE, "dwconv3x3: bad dilation/padding");
  S2R_REQUIRE(C <= 2048, S2R_ERR_UNSUPPORTED, "dwconv3x3: C=%d too large (one thread per 8 channels per CTA)", C);
  S2R_REQUIRE(a && b && ((uintptr_t)a | (uintptr_t)b) % 16 == 0, S2R_ERR_SHAPE, "dwconv3x3: null or unaligned tensor");
  S2R_REQUIRE((long long)(H + 2 * pad) * (W + 2 * pad) * C < (1ll << 31), S2R_ERR_UNSUPPORTED,
              "dwconv3x3: one image plane exceeds 2^31 elements");
  S2R_REQUIRE(N <= 65535, S2R_ERR_UNSUPPORTED, "dwconv3x3: batch %d > 65535", N);
  const int he = H + 2 * pad - 2 * dil - 1, we = W + 2 * pad - 2 * dil - 1;
  S2R_REQUIRE(he >= 0 && we >= 0, S2R_ERR_SHAPE, "dwconv3x3: input smaller than the dilated filter");
  G->N = N; G->H = H; G->W = W; G->C = C;
  G->Ho = he / stride + 1;
  G->Wo = we / stride + 1;
  G->stride = stride; G->dil = dil; G->pad = pad;
  return S2R_OK;
}

// threads per CTA = tw columns x cg channel groups (x 3 kernel rows for the weight gradient)
inline int pick_tw(int cg, int width, int per_col) {
  int tw = 256 / (cg * per_col);
  if (tw < 1) tw = 1;
  if (tw > width) tw = width;
  return tw;
}

template <typename K>
inline int dw_smem_attr(K kernel, size_t smem) {
  if (smem > 48 * 1024) S2R_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return S2R_OK;
}

}  // namespace

// streaming stride-1 kernels (dwconv_s1.cu)
int s2r_dw_s1_fwd(const void* x, const float* ss, const s2r_bn_tail* in_bn, int halo_const, const float* w, void* y,
                  double* stats, const float* oss, int N, int H, int W, int C, int dil, cudaStream_t stream);
int s2r_dw_s1_bwd(const void* dy, const void* x, const float* ss, const float* mi, const float* w, int ext,
                  int interior, void* g, double* bsums, float* dw, int N, int H, int W, int C, int dil, cudaStream_t stream);

int s2r_dw_s2_fwd(const void* x, const float* ss, const s2r_bn_tail* in_bn, const float* w, void* y, double* stats,
                  const float* oss, int N, int H, int W, int C, cudaStream_t stream);
int s2r_dw_s2_bwd(const void* dy, const void* x, const float* ss, const float* mi, const float* w, int interior,
                  void* g, double* bsums, float* dw, int N, int H, int W, int C, cudaStream_t stream);

static bool s2_eligible(const float* ss, int in_act, int halo_const, int stride, int dil, int pad, int C) {
  return ss != nullptr && in_act == S2R_ACT_RELU6 && halo_const && stride == 2 && dil == 1 && pad == 1 && C % 16 == 0 &&
         getenv("S2R_DW_GENERIC") == nullptr;
}

static bool s1_eligible(const float* ss, int in_act, int stride, int dil, int pad, int C) {
  return ss != nullptr && in_act == S2R_ACT_RELU6 && stride == 1 && dil >= 1 && dil <= 4 && pad == dil && C % 16 == 0 &&
         getenv("S2R_DW_GENERIC") == nullptr;
}

// in_bn: the input's BatchNorm is still pending (s2r_bn_tail): the streaming kernels derive scale / shift from its
// sums in their prologue; with a cross-rank exchange, or on the generic kernel, it is finalised by one small launch first.
// out_scale_shift (inference, [2][C] or NULL): y = relu6(conv*scale + shift), the BatchNorm + ReLU6 that follows the
// convolution with its running statistics, applied in the kernel's store (the generic kernel: one bn_apply pass after it).
int s2r_bn_apply_inplace_relu6(void* y, int64_t P, int C, const float* scale_shift, cudaStream_t st);

extern "C" int s2r_dwconv3x3_fwd_bn(const void* x, const s2r_bn_tail* in_bn, const float* in_scale_shift, int in_act,
                                    int halo_const, const float* w, void* y, double* stats, const float* out_scale_shift,
                                    int N, int H, int W, int C, int stride, int dil, int pad, s2r_stream_t stream) {
  DwGeom G;
  int rc = dw_check(x, y, N, H, W, C, stride, dil, pad, &G);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  S2R_REQUIRE(!(out_scale_shift && stats), S2R_ERR_SHAPE, "dwconv3x3_fwd: an output BatchNorm (inference) and statistics (training) exclude each other");
  if (in_bn) {
    S2R_REQUIRE(in_bn->count > 1, S2R_ERR_SHAPE, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
    S2R_REQUIRE(in_bn->sums && in_bn->scale_shift && in_bn->mean_invstd && (uintptr_t)in_bn->sums % 16 == 0, S2R_ERR_SHAPE,
                "dwconv3x3_fwd: pending BatchNorm without sums / outputs");
    in_scale_shift = in_bn->scale_shift;
  }
  const bool vec = in_scale_shift && ((uintptr_t)in_scale_shift % 16 == 0);
  const bool s1 = s1_eligible(in_scale_shift, in_act, stride, dil, pad, C) && vec;
  const bool s2 = s2_eligible(in_scale_shift, in_act, halo_const, stride, dil, pad, C) && vec;
  const s2r_bn_tail* fused = ((s1 || s2) && bn_tail_fusable(in_bn)) ? in_bn : nullptr;
  if (in_bn && !fused) {
    rc = s2r_bn_tail_launch(in_bn, C, st);
    if (rc) return rc;
  }
  if (s1) {
    rc = s2r_dw_s1_fwd(x, in_scale_shift, fused, halo_const, w, y, stats, out_scale_shift, N, H, W, C, dil, st);
    if (rc != S2R_ERR_UNSUPPORTED) return rc;
  }
  if (s2) {
    rc = s2r_dw_s2_fwd(x, in_scale_shift, fused, w, y, stats, out_scale_shift, N, H, W, C, st);
    if (rc != S2R_ERR_UNSUPPORTED) return rc;
  }
  if (fused) {   // the streaming kernels declined the shape after all
    rc = s2r_bn_tail_launch(in_bn, C, st);
    if (rc) return rc;
  }
  const int cg = C / 8;
  const int tw = pick_tw(cg, G.Wo, 1);
  const size_t smem = ((size_t)9 * C + (size_t)tw * cg * 16) * sizeof(float);
  dim3 grid(s2r_div_up(G.Wo, tw), s2r_div_up(G.Ho, ROWS), N);
  const bool slide = stride == 1 && dil == 1;
  if (slide) {
    rc = dw_smem_attr(dw_fwd_kernel<true>, smem);
    if (rc) return rc;
    S2R_CUDA_OK(s2r_launch(dw_fwd_kernel<true>, dim3(grid), dim3(tw * cg), (size_t)(smem), st,
        (const __nv_bfloat16*)x, in_scale_shift, in_act, halo_const, w, (__nv_bfloat16*)y, stats, G, tw));
  } else {
    rc = dw_smem_attr(dw_fwd_kernel<false>, smem);
    if (rc) return rc;
    S2R_CUDA_OK(s2r_launch(dw_fwd_kernel<false>, dim3(grid), dim3(tw * cg), (size_t)(smem), st,
        (const __nv_bfloat16*)x, in_scale_shift, in_act, halo_const, w, (__nv_bfloat16*)y, stats, G, tw));
  }
  S2R_LAUNCH_OK();
  if (out_scale_shift) return s2r_bn_apply_inplace_relu6(y, (int64_t)N * G.Ho * G.Wo, C, out_scale_shift, st);
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_fwd(const void* x, const float* in_scale_shift, int in_act, int halo_const,
                                 const float* w, void* y, double* stats, int N, int H, int W, int C,
                                 int stride, int dil, int pad, s2r_stream_t stream) {
  return s2r_dwconv3x3_fwd_bn(x, nullptr, in_scale_shift, in_act, halo_const, w, y, stats, nullptr, N, H, W, C, stride, dil,
                              pad, stream);
}

static int dw_dgrad_generic(const void* dy, const float* w, const void* x,
                            const float* in_scale_shift, const float* in_mean_invstd,
                            int in_act, int ext, int interior, void* g, double* bwd_sums, int N, int H, int W,
                            int C, int stride, int dil, int pad, s2r_stream_t stream) {
  DwGeom G;
  int rc = dw_check(dy, g, N, H, W, C, stride, dil, pad, &G);
  if (rc) return rc;
  S2R_REQUIRE(ext >= 0 && ext <= pad, S2R_ERR_SHAPE, "dwconv3x3_dgrad: ext=%d outside [0,pad]", ext);
  S2R_REQUIRE(!in_scale_shift || (x && (uintptr_t)x % 16 == 0 && in_mean_invstd), S2R_ERR_SHAPE,
              "dwconv3x3_dgrad: the masked variant needs x and mean/invstd");
  const int cg = C / 8;
  const int He = H + 2 * ext, We = W + 2 * ext;
  const int tw = pick_tw(cg, We, 1);
  const size_t smem = ((size_t)9 * C + (size_t)tw * cg * 16) * sizeof(float);
  rc = dw_smem_attr(dw_dgrad_kernel, smem);
  if (rc) return rc;
  dim3 grid(s2r_div_up(We, tw), s2r_div_up(He, ROWS), N);
  S2R_CUDA_OK(s2r_launch(dw_dgrad_kernel, dim3(grid), dim3(tw * cg), (size_t)(smem), (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, w, (const __nv_bfloat16*)x, in_scale_shift, in_mean_invstd, in_act, ext, interior,
      (__nv_bfloat16*)g, bwd_sums, G, tw));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_dwconv3x3_dgrad(const void* dy, const float* w, const void* x,
                                   const float* in_scale_shift, const float* in_mean_invstd,
                                   int in_act, int ext, void* g, double* bwd_sums, int N, int H, int W,
                                   int C, int stride, int dil, int pad, s2r_stream_t stream) {
  return dw_dgrad_generic(dy, w, x, in_scale_shift, in_mean_invstd, in_act, ext, 0, g, bwd_sums, N, H, W, C, stride, dil,
                          pad, stream);
}

extern "C" int s2r_dwconv3x3_wgrad(const void* x, const float* in_scale_shift, int in_act,
                                   int halo_const, const void* dy, float* dw, int N, int H, int W,
                                   int C, int stride, int dil, int pad, s2r_stream_t stream) {
  DwGeom G;
  int rc = dw_check(x, dy, N, H, W, C, stride, dil, pad, &G);
  if (rc) return rc;
  const int cg = C / 8;
  S2R_REQUIRE(3 * cg <= 384, S2R_ERR_UNSUPPORTED, "dwconv3x3_wgrad: C=%d > 1024", C);
  const int tw = pick_tw(cg, G.Wo, 3);
  const int threads = tw * 3 * cg;
  const size_t smem = (size_t)threads * 24 * sizeof(float);
  rc = dw_smem_attr(dw_wgrad_kernel, smem);
  if (rc) return rc;
  dim3 grid(s2r_div_up(G.Wo, tw), s2r_div_up(G.Ho, WROWS), N);
  S2R_CUDA_OK(s2r_launch(dw_wgrad_kernel, dim3(grid), dim3(threads), (size_t)(smem), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, in_scale_shift, in_act, halo_const, (const __nv_bfloat16*)dy, dw, G, tw));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

// Fused data + weight gradient (one pass over dy and x where the streaming kernel applies, otherwise the
// two generic kernels).  halo_const selects the reference's padded-border semantics (ext = pad).
extern "C" int s2r_dwconv3x3_bwd(const void* dy, const float* w, const void* x, const float* in_scale_shift,
                                 const float* in_mean_invstd, int in_act, int halo_const, int g_interior, void* g,
                                 double* bwd_sums, float* dw, int N, int H, int W, int C, int stride, int dil,
                                 int pad, s2r_stream_t stream) {
  const int ext = halo_const ? pad : 0;
  if (ext == 0) g_interior = 0;   // same layout
  if (s1_eligible(in_scale_shift, in_act, stride, dil, pad, C) && x && (bwd_sums == nullptr || in_mean_invstd) &&
      ((uintptr_t)in_scale_shift % 16 == 0) && (in_mean_invstd == nullptr || (uintptr_t)in_mean_invstd % 16 == 0)) {
    DwGeom G;
    int rc = dw_check(dy, g, N, H, W, C, stride, dil, pad, &G);
    if (rc) return rc;
    rc = s2r_dw_s1_bwd(dy, x, in_scale_shift, in_mean_invstd, w, ext, g_interior, g, bwd_sums, dw, N, H, W, C, dil,
                       (cudaStream_t)stream);
    if (rc != S2R_ERR_UNSUPPORTED) return rc;
  }
  if (s2_eligible(in_scale_shift, in_act, halo_const, stride, dil, pad, C) && x && (bwd_sums == nullptr || in_mean_invstd) &&
      ((uintptr_t)in_scale_shift % 16 == 0) && (in_mean_invstd == nullptr || (uintptr_t)in_mean_invstd % 16 == 0)) {
    DwGeom G;
    int rc = dw_check(dy, g, N, H, W, C, stride, dil, pad, &G);
    if (rc) return rc;
    rc = s2r_dw_s2_bwd(dy, x, in_scale_shift, in_mean_invstd, w, g_interior, g, bwd_sums, dw, N, H, W, C, (cudaStream_t)stream);
    if (rc != S2R_ERR_UNSUPPORTED) return rc;
  }
  if (dw) {
    int rc = s2r_dwconv3x3_wgrad(x, in_scale_shift, in_act, halo_const, dy, dw, N, H, W, C, stride, dil, pad, stream);
    if (rc) return rc;
  }
  return dw_dgrad_generic(dy, w, in_scale_shift ? x : nullptr, in_scale_shift, in_mean_invstd, in_act, ext, g_interior, g,
                             bwd_sums, N, H, W, C, stride, dil, pad, stream);
}
