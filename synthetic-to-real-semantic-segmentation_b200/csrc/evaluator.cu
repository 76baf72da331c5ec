// Evaluator confusion matrix on device.
// Replaces utils/metrics.py:34-43 (Evaluator._generate_matrix / add_batch) and the
// host argmax at val_adapt.py:131-135 of the reference.  Integer counts, bit-exact.
//
// HBM-bound: 12 B/pixel (float gt + int64 pred) or (4*C + 4) B/pixel for the fused
// argmax variant.  One shared-memory histogram per CTA, warp-aggregated with
// match.any so spatially coherent label maps do not serialise on one bank, then
// one 64-bit global atomic per non-empty bin per CTA.
#include "common.cuh"
#include "lerp.cuh"

namespace {

constexpr int kMaxClass = 64;
constexpr int kThreads = 256;

__device__ __forceinline__ void hist_add(unsigned int* hist, int bin) {
  // bin < 0 means "no contribution"; all 32 lanes must call this together
  unsigned peers = __match_any_sync(0xffffffffu, bin);
  int leader = __ffs(peers) - 1;
  if (bin >= 0 && (int)(threadIdx.x & 31) == leader) atomicAdd(&hist[bin], (unsigned)__popc(peers));
}

template <typename GT>
__device__ __forceinline__ bool gt_valid(GT g, int nc, int* cls) {
  // mask = (gt >= 0) & (gt < nc); label uses astype('int') == truncation
  if (g >= (GT)0 && g < (GT)nc) {
    *cls = (int)g;
    return true;
  }
  return false;
}

template <typename GT>
__global__ void __launch_bounds__(kThreads)
confusion_kernel(const GT* __restrict__ gt, const long long* __restrict__ pred, long long n, int nc,
                 unsigned long long* __restrict__ counts, unsigned long long* __restrict__ bad) {
  __shared__ unsigned int hist[kMaxClass * kMaxClass];
  __shared__ unsigned int nbad;
  const int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += kThreads) hist[i] = 0;
  if (threadIdx.x == 0) nbad = 0;
  __syncthreads();

  const long long stride = (long long)gridDim.x * kThreads;
  // block-uniform trip count so that match.any sees all 32 lanes
  const long long iters = (n + stride - 1) / stride;
  const long long first = (long long)blockIdx.x * kThreads + threadIdx.x;
  for (long long it = 0; it < iters; it += 4) {
    GT g[4];
    long long p[4];
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      long long i = first + (it + u) * stride;
      in[u] = (it + u) < iters && i < n;
      g[u] = in[u] ? gt[i] : (GT)-1;
      p[u] = in[u] ? pred[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int cls = 0;
      int bin = -1;
      if (in[u] && gt_valid<GT>(g[u], nc, &cls)) {
        if (p[u] >= 0 && p[u] < nc) bin = cls * nc + (int)p[u];
        else atomicAdd(&nbad, 1u);
      }
      hist_add(hist, bin);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kThreads) {
    unsigned int c = hist[i];
    if (c) atomicAdd(&counts[i], (unsigned long long)c);
  }
  if (threadIdx.x == 0 && nbad && bad) atomicAdd(bad, (unsigned long long)nbad);
}

// Fused argmax over the class planes of NCHW fp32 logits + confusion histogram.
// One thread per pixel, class planes walked with stride HW (each plane read is a
// coalesced 128 B line per warp).  First maximum wins, NaN counts as maximum
// (numpy.argmax semantics).
__global__ void __launch_bounds__(kThreads)
argmax_confusion_kernel(const float* __restrict__ logits, const float* __restrict__ gt, int C,
                        long long HW, long long npix, int nc, unsigned long long* __restrict__ counts,
                        long long* __restrict__ pred_out) {
  __shared__ unsigned int hist[kMaxClass * kMaxClass];
  const int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += kThreads) hist[i] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * kThreads;
  const long long iters = (npix + stride - 1) / stride;
  const long long first = (long long)blockIdx.x * kThreads + threadIdx.x;
  for (long long it = 0; it < iters; ++it) {
    long long i = first + it * stride;
    int bin = -1;
    if (i < npix) {
      long long img = i / HW, px = i - img * HW;
      const float* base = logits + img * (long long)C * HW + px;
      float best = __ldg(base);
      int arg = 0;
      for (int c = 1; c < C; ++c) {
        float v = __ldg(base + (long long)c * HW);
        // strict > keeps the first maximum; a NaN beats any non-NaN
        if (v > best || (v != v && best == best)) {
          best = v;
          arg = c;
        }
      }
      if (pred_out) pred_out[i] = arg;
      if (gt) {
        int cls = 0;
        if (gt_valid<float>(__ldg(gt + i), nc, &cls) && arg < nc) bin = cls * nc + arg;
      }
    }
    if (gt) hist_add(hist, bin);
  }
  __syncthreads();
  if (gt) {
    for (int i = threadIdx.x; i < bins; i += kThreads) {
      unsigned int c = hist[i];
      if (c) atomicAdd(&counts[i], (unsigned long long)c);
    }
  }
}


// The same on the decoder's LOW-RESOLUTION logits (NHWC bf16 [N][Hi][Wi][pitch], C classes): the final
// F.interpolate(x, size, mode='bilinear', align_corners=True) of deeplab.py:31, the argmax of val_adapt.py:133 and the
// histogram of metrics.py:34-43 in one pass -- the fp32 [N,C,Ho,Wo] logits (159 MB per 1024x2048 image, written by
// the up-sampling and read back by the argmax) never exist.  Every thread evaluates the C interpolated values of its
// pixel with resize.cu's own arithmetic (lerp.cuh), so the predictions are those of the two-step path bit for bit.
template <int CG>
__global__ void __launch_bounds__(kThreads)
up_argmax_confusion_kernel(const __nv_bfloat16* __restrict__ x, int xpitch, int Hi, int Wi, int C,
                           const float* __restrict__ gt, int Ho, int Wo, long long npix, float sh, float sw, int nc,
                           unsigned long long* __restrict__ counts) {
  __shared__ unsigned int hist[kMaxClass * kMaxClass];
  const int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += kThreads) hist[i] = 0;
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const long long stride = (long long)gridDim.x * kThreads;
  const long long iters = (npix + stride - 1) / stride;   // block-uniform: match.any needs all 32 lanes
  const long long first = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long plane = (long long)Ho * Wo;
  for (long long it = 0; it < iters; ++it) {
    const long long i = first + it * stride;
    int bin = -1;
    if (i < npix) {
      const int n = (int)(i / plane);
      const int r = (int)(i - (long long)n * plane);
      const int oh = r / Wo, ow = r - oh * Wo;
      const Lerp ly = lerp_src(oh, sh, Hi), lx = lerp_src(ow, sw, Wi);
      const __nv_bfloat16* b = x + (long long)n * Hi * Wi * xpitch;
      const __nv_bfloat16* p00 = b + ((long long)ly.i0 * Wi + lx.i0) * xpitch;
      const __nv_bfloat16* p01 = b + ((long long)ly.i0 * Wi + lx.i1) * xpitch;
      const __nv_bfloat16* p10 = b + ((long long)ly.i1 * Wi + lx.i0) * xpitch;
      const __nv_bfloat16* p11 = b + ((long long)ly.i1 * Wi + lx.i1) * xpitch;
      float best = 0.f;
      int arg = 0;
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        float v00[8], v01[8], v10[8], v11[8];
        bf16x8_to_float(ldg16(p00 + g * 8), v00);
        bf16x8_to_float(ldg16(p01 + g * 8), v01);
        bf16x8_to_float(ldg16(p10 + g * 8), v10);
        bf16x8_to_float(ldg16(p11 + g * 8), v11);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = g * 8 + k;
          if (c < C) {
            const float v = bilerp(ly, lx, v00[k], v01[k], v10[k], v11[k]);
            // strict > keeps the first maximum; a NaN beats any non-NaN (numpy.argmax)
            if (c == 0 || v > best || (v != v && best == best)) {
              best = v;
              arg = c;
            }
          }
        }
      }
      int cls = 0;
      if (gt_valid<float>(__ldg(gt + i), nc, &cls) && arg < nc) bin = cls * nc + arg;
    }
    hist_add(hist, bin);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kThreads) {
    const unsigned int c = hist[i];
    if (c) atomicAdd(&counts[i], (unsigned long long)c);
  }
}


// Prediction export (test_adapt.py:118-157 / val_adapt.py:179-218 imgsaver + the host argmax at test_adapt.py:170-171):
// argmax over the class planes at the source pixel PIL's NEAREST resize picks for every output position, then the two
// byte tables (trainId -> labelId, trainId -> RGB).  ids u8 [N][OH][OW], rgb u8 [N][OH][OW][3]; ties -> lowest class
// (np.argmax); values without a table entry stay 0 like the reference's zero-initialised images.
__global__ void __launch_bounds__(kThreads)
export_prediction_kernel(const float* __restrict__ logits, int N, int C, int H, int W, const int* __restrict__ xtab,
                         const int* __restrict__ ytab, int OH, int OW, const unsigned char* __restrict__ id_table,
                         const unsigned char* __restrict__ rgb_table, int ntab, unsigned char* __restrict__ ids,
                         unsigned char* __restrict__ rgb) {
  const long long total = (long long)N * OH * OW;
  const long long plane = (long long)H * W;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % OW);
    long long q = i / OW;
    const int y = (int)(q % OH);
    const int n = (int)(q / OH);
    const int sx = __ldg(xtab + x), sy = __ldg(ytab + y);
    unsigned char id = 0, r = 0, g = 0, b = 0;
    if (sx >= 0 && sy >= 0) {
      const float* p = logits + (long long)n * C * plane + (long long)sy * W + sx;
      float best = __ldg(p);
      int arg = 0;
      for (int c = 1; c < C; ++c) {
        const float v = __ldg(p + c * plane);
        if (v > best) { best = v; arg = c; }   // strict: the first maximum wins, as in np.argmax
      }
      if (arg < ntab) {
        id = id_table[arg];
        r = rgb_table[3 * arg]; g = rgb_table[3 * arg + 1]; b = rgb_table[3 * arg + 2];
      }
    }
    if (ids) ids[i] = id;
    if (rgb) { rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b; }
  }
}

}  // namespace

extern "C" int s2r_confusion_matrix(const void* gt, int gt_is_i64, const int64_t* pred, int64_t n,
                                    int num_class, int64_t* counts, int64_t* bad_pred,
                                    s2r_stream_t stream) {
  S2R_REQUIRE(num_class >= 1 && num_class <= kMaxClass, S2R_ERR_UNSUPPORTED,
              "confusion_matrix: num_class %d outside [1,%d]", num_class, kMaxClass);
  S2R_REQUIRE(n >= 0, S2R_ERR_SHAPE, "confusion_matrix: negative element count");
  if (n == 0) return S2R_OK;
  S2R_REQUIRE(gt && pred && counts, S2R_ERR_SHAPE, "confusion_matrix: null pointer");
  int grid = s2r_grid(n, kThreads * 16, 4);
  if (gt_is_i64)
    confusion_kernel<long long><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const long long*)gt, (const long long*)pred, n, num_class, (unsigned long long*)counts,
        (unsigned long long*)bad_pred);
  else
    confusion_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const float*)gt, (const long long*)pred, n, num_class, (unsigned long long*)counts,
        (unsigned long long*)bad_pred);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_argmax_confusion_nchw(const float* logits, const float* gt, int N, int C,
                                         int64_t HW, int num_class, int64_t* counts,
                                         int64_t* pred_out, s2r_stream_t stream) {
  S2R_REQUIRE(N >= 0 && C >= 1 && HW >= 0, S2R_ERR_SHAPE, "argmax_confusion: bad shape");
  S2R_REQUIRE(!gt || (num_class >= 1 && num_class <= kMaxClass), S2R_ERR_UNSUPPORTED,
              "argmax_confusion: num_class %d outside [1,%d]", num_class, kMaxClass);
  S2R_REQUIRE(!gt || counts, S2R_ERR_SHAPE, "argmax_confusion: counts is null");
  long long npix = (long long)N * HW;
  if (npix == 0) return S2R_OK;
  int grid = s2r_grid(npix, kThreads * 4, 8);
  argmax_confusion_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      logits, gt, C, HW, npix, num_class > 0 ? num_class : 1, (unsigned long long*)counts,
      (long long*)pred_out);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_upsample_argmax_confusion_nhwc(const void* x, int xpitch, int N, int Hi, int Wi, int C,
                                                  const float* gt, int Ho, int Wo, int num_class, int64_t* counts,
                                                  s2r_stream_t stream) {
  S2R_REQUIRE(N >= 0 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1, S2R_ERR_SHAPE, "upsample_argmax_confusion: bad shape");
  const int cgs = (C + 7) / 8;
  S2R_REQUIRE(x && xpitch % 8 == 0 && xpitch >= cgs * 8 && (uintptr_t)x % 16 == 0, S2R_ERR_SHAPE,
              "upsample_argmax_confusion: pitch %d must be a multiple of 8 covering C=%d", xpitch, C);
  S2R_REQUIRE(cgs <= 4, S2R_ERR_UNSUPPORTED, "upsample_argmax_confusion: C=%d > 32 not supported", C);
  S2R_REQUIRE(num_class >= 1 && num_class <= kMaxClass, S2R_ERR_UNSUPPORTED,
              "upsample_argmax_confusion: num_class %d outside [1,%d]", num_class, kMaxClass);
  S2R_REQUIRE(gt && counts, S2R_ERR_SHAPE, "upsample_argmax_confusion: gt / counts is null");
  S2R_REQUIRE((long long)Ho * Wo < (1ll << 31), S2R_ERR_UNSUPPORTED, "upsample_argmax_confusion: image too large");
  const long long npix = (long long)N * Ho * Wo;
  if (npix == 0) return S2R_OK;
  const int grid = s2r_grid(npix, kThreads * 4, 8);
  const float sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  cudaStream_t st = (cudaStream_t)stream;
#define S2R_UAC(CG_) S2R_CUDA_OK(s2r_launch(up_argmax_confusion_kernel<CG_>, dim3(grid), dim3(kThreads), (size_t)0, st, xb, xpitch, Hi, Wi, C, gt, Ho, Wo, npix, sh, sw, num_class, (unsigned long long*)counts))
  switch (cgs) {
    case 1: S2R_UAC(1); break;
    case 2: S2R_UAC(2); break;
    case 3: S2R_UAC(3); break;
    default: S2R_UAC(4); break;
  }
#undef S2R_UAC
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_export_prediction_nchw(const float* logits, int N, int C, int H, int W, const int32_t* xtab,
                                          const int32_t* ytab, int OH, int OW, const uint8_t* id_table,
                                          const uint8_t* rgb_table, int ntab, uint8_t* ids, uint8_t* rgb,
                                          s2r_stream_t stream) {
  S2R_REQUIRE(logits && xtab && ytab && id_table && rgb_table && (ids || rgb), S2R_ERR_SHAPE, "export_prediction: null pointer");
  S2R_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1 && OH >= 1 && OW >= 1 && ntab >= 0, S2R_ERR_SHAPE, "export_prediction: bad shape");
  const long long total = (long long)N * OH * OW;
  export_prediction_kernel<<<s2r_grid(total, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(
      logits, N, C, H, W, xtab, ytab, OH, OW, id_table, rgb_table, ntab, ids, rgb);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
