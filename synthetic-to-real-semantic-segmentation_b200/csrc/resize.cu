// Bilinear resampling with align_corners=True, global average pooling and the
// NCHW<->NHWC boundary conversions.
// Replaces F.interpolate(..., mode='bilinear', align_corners=True) at
// modeling/deeplab.py:31, modeling/decoder.py:39, modeling/assp.py:71 and
// nn.AdaptiveAvgPool2d((1,1)) at modeling/assp.py:55 of the reference.
// Source index arithmetic follows PyTorch's upsample_bilinear2d: scale =
// (in-1)/(out-1) in fp32, src = scale*dst, i0 = int(src), i1 = i0 + (i0 < in-1).
// Backward kernels are gather-formulated (no atomics, deterministic).
#include "common.cuh"
#include "lerp.cuh"

namespace {

constexpr int kThreads = 256;

// candidate output range that can reference input index i
__device__ __forceinline__ void cand_range(int i, float scale, int out, int* lo, int* hi) {
  if (scale <= 0.f) {
    *lo = 0;
    *hi = out - 1;
    return;
  }
  int a = (int)floorf((float)(i - 1) / scale) - 1;
  int b = (int)ceilf((float)(i + 1) / scale) + 1;
  *lo = a < 0 ? 0 : a;
  *hi = b > out - 1 ? out - 1 : b;
}

__device__ __forceinline__ float lerp_weight(const Lerp& l, int i) {
  return (l.i0 == i ? l.w0 : 0.f) + (l.i1 == i ? l.w1 : 0.f);
}

// ------------------------------------------------------------ NHWC bf16 -> NHWC bf16
// IT: index type of the flattened element counter.  unsigned (32-bit) whenever the tensor allows it: the three
// div/mod pairs per element are ~10 instructions in 32 bits and ~100 in 64 bits, more than the interpolation itself.
template <typename IT>
__global__ void __launch_bounds__(kThreads)
up_nhwc_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int Hi, int Wi, int C,
                   __nv_bfloat16* __restrict__ y, int Ho, int Wo, int ypitch, int yoff, float sh,
                   float sw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = C / 8;
  const long long total = (long long)N * Ho * Wo * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const IT tt = (IT)t;
    const int g = (int)(tt % (IT)cg);
    const IT p = tt / (IT)cg;
    const int ow = (int)(p % (IT)Wo);
    const IT q = p / (IT)Wo;
    const int oh = (int)(q % (IT)Ho);
    const int n = (int)(q / (IT)Ho);
    const Lerp ly = lerp_src(oh, sh, Hi), lx = lerp_src(ow, sw, Wi);
    const __nv_bfloat16* b = x + (long long)n * Hi * Wi * C + g * 8;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    bf16x8_to_float(ldg16(b + ((long long)ly.i0 * Wi + lx.i0) * C), v00);
    bf16x8_to_float(ldg16(b + ((long long)ly.i0 * Wi + lx.i1) * C), v01);
    bf16x8_to_float(ldg16(b + ((long long)ly.i1 * Wi + lx.i0) * C), v10);
    bf16x8_to_float(ldg16(b + ((long long)ly.i1 * Wi + lx.i1) * C), v11);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = ly.w0 * (lx.w0 * v00[i] + lx.w1 * v01[i]) + ly.w1 * (lx.w0 * v10[i] + lx.w1 * v11[i]);
    *reinterpret_cast<uint4*>(y + (long long)p * ypitch + yoff + g * 8) = float_to_bf16x8(o);
  }
}

template <typename IT>
__global__ void __launch_bounds__(kThreads)
up_nhwc_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dypitch, int dyoff, int N, int Hi, int Wi,
                   int C, int Ho, int Wo, __nv_bfloat16* __restrict__ dx, float sh, float sw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = C / 8;
  const long long total = (long long)N * Hi * Wi * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const IT tt = (IT)t;
    const int g = (int)(tt % (IT)cg);
    const IT p = tt / (IT)cg;
    const int iw = (int)(p % (IT)Wi);
    const IT q = p / (IT)Wi;
    const int ih = (int)(q % (IT)Hi);
    const int n = (int)(q / (IT)Hi);
    int hlo, hhi, wlo, whi;
    cand_range(ih, sh, Ho, &hlo, &hhi);
    cand_range(iw, sw, Wo, &wlo, &whi);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int oh = hlo; oh <= hhi; ++oh) {
      const float wy = lerp_weight(lerp_src(oh, sh, Hi), ih);
      if (wy == 0.f) continue;
      for (int ow = wlo; ow <= whi; ++ow) {
        const float wx = lerp_weight(lerp_src(ow, sw, Wi), iw);
        if (wx == 0.f) continue;
        float f[8];
        bf16x8_to_float(ldg16(dy + (((long long)n * Ho + oh) * Wo + ow) * dypitch + dyoff + g * 8), f);
        const float wgt = wy * wx;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
      }
    }
    *reinterpret_cast<uint4*>(dx + t * 8) = float_to_bf16x8(acc);
  }
}

// ------------------------------------------------------------ NHWC bf16 -> NCHW fp32
template <int CG, typename IT>  // ceil(C/8) vectors per pixel
__global__ void __launch_bounds__(kThreads)
up_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int xpitch, int N, int Hi, int Wi, int C,
                  float* __restrict__ y, int Ho, int Wo, float sh, float sw) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const long long total = (long long)N * Ho * Wo;
  const long long plane = (long long)Ho * Wo;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const IT tt = (IT)t;
    const int ow = (int)(tt % (IT)Wo);
    const IT q = tt / (IT)Wo;
    const int oh = (int)(q % (IT)Ho);
    const int n = (int)(q / (IT)Ho);
    const Lerp ly = lerp_src(oh, sh, Hi), lx = lerp_src(ow, sw, Wi);
    const __nv_bfloat16* b = x + (long long)n * Hi * Wi * xpitch;
    const __nv_bfloat16* p00 = b + ((long long)ly.i0 * Wi + lx.i0) * xpitch;
    const __nv_bfloat16* p01 = b + ((long long)ly.i0 * Wi + lx.i1) * xpitch;
    const __nv_bfloat16* p10 = b + ((long long)ly.i1 * Wi + lx.i0) * xpitch;
    const __nv_bfloat16* p11 = b + ((long long)ly.i1 * Wi + lx.i1) * xpitch;
    float* out = y + (long long)n * C * plane + (long long)oh * Wo + ow;
#pragma unroll
    for (int g = 0; g < CG; ++g) {
      float v00[8], v01[8], v10[8], v11[8];
      bf16x8_to_float(ldg16(p00 + g * 8), v00);
      bf16x8_to_float(ldg16(p01 + g * 8), v01);
      bf16x8_to_float(ldg16(p10 + g * 8), v10);
      bf16x8_to_float(ldg16(p11 + g * 8), v11);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        if (c < C)
          out[(long long)c * plane] = bilerp(ly, lx, v00[i], v01[i], v10[i], v11[i]);
      }
    }
  }
}

// dy NCHW fp32 [N,C,Ho,Wo] -> dx NHWC bf16 [N,Hi,Wi,dxpitch]; thread per (n, c, ih, iw)
__global__ void __launch_bounds__(kThreads)
up_from_nchw_bwd_kernel(const float* __restrict__ dy, int N, int C, int Ho, int Wo,
                        __nv_bfloat16* __restrict__ dx, int dxpitch, int Hi, int Wi, float sh, float sw,
                        const float* __restrict__ gscale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const long long total = (long long)N * dxpitch * Hi * Wi;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const int iw = (int)(t % Wi);
    long long q = t / Wi;
    const int ih = (int)(q % Hi);
    q /= Hi;
    const int c = (int)(q % dxpitch);
    const int n = (int)(q / dxpitch);
    float acc = 0.f;
    if (c < C) {
      int hlo, hhi, wlo, whi;
      cand_range(ih, sh, Ho, &hlo, &hhi);
      cand_range(iw, sw, Wo, &wlo, &whi);
      const float* src = dy + ((long long)n * C + c) * Ho * Wo;
      for (int oh = hlo; oh <= hhi; ++oh) {
        const float wy = lerp_weight(lerp_src(oh, sh, Hi), ih);
        if (wy == 0.f) continue;
        for (int ow = wlo; ow <= whi; ++ow) {
          const float wx = lerp_weight(lerp_src(ow, sw, Wi), iw);
          if (wx != 0.f) acc = fmaf(wy * wx, __ldg(src + (long long)oh * Wo + ow), acc);
        }
      }
    }
    dx[(((long long)n * Hi + ih) * Wi + iw) * dxpitch + c] = __float2bfloat16(gscale ? acc * __ldg(gscale) : acc);
  }
}

// Separable, row-staged variant of the kernel above for the final x4 up-sampling (deeplab.py:31), whose
// gradient is the largest activation of the step (N x 19 x 512 x 1024 fp32): a CTA owns (image, channel,
// UR coarse rows, 256 coarse columns).  The fine rows are walked one coarse INTERVAL at a time (the ~1/scale
// rows whose source coordinate lies in [ih, ih+1)): they are staged in shared memory with coalesced float4 loads
// (each fine element read once per CTA), every thread reduces its column neighbourhood with weights held in
// registers (exactly the <= UKX fine columns with a non-zero weight), and the two row weights of an interval are
// applied to two accumulators with static indices.  Same arithmetic as up_from_nchw_bwd_kernel up to summation order.
constexpr int UR = 8;      // coarse rows per CTA
constexpr int UKX = 10;    // max fine columns with a non-zero weight for one coarse column (floor(2/scale) + 1)
constexpr int URB = 6;     // fine rows staged per barrier pair (an interval has ceil(1/scale) <= URB rows)

// smallest fine row oh >= 0 whose (clamped) source row index is >= i
__device__ __forceinline__ int first_fine_row(int i, float sh, int Hi, int Ho) {
  if (i <= 0) return 0;
  int o = (int)ceilf((float)i / sh);
  if (o > Ho) o = Ho;
  while (o > 0 && lerp_src(o - 1, sh, Hi).i0 >= i) --o;
  while (o < Ho && lerp_src(o, sh, Hi).i0 < i) ++o;
  return o;
}

__global__ void __launch_bounds__(kThreads)
up_from_nchw_bwd_rows_kernel(const float* __restrict__ dy, int C, int Ho, int Wo, __nv_bfloat16* __restrict__ dx,
                             int dxpitch, int Hi, int Wi, float sh, float sw, int col_chunks, int seg_max,
                             const float* __restrict__ gscale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float rowbuf[];   // [URB][seg_max]
  const int ihb = blockIdx.x / col_chunks, cb = blockIdx.x - ihb * col_chunks;
  const int c = blockIdx.y, n = blockIdx.z;
  const int ih0 = ihb * UR;
  const int iw = cb * kThreads + threadIdx.x;
  const bool col_ok = iw < Wi;
  __nv_bfloat16* out = dx + (((long long)n * Hi + ih0) * Wi + (col_ok ? iw : 0)) * dxpitch + c;
  if (c >= C) {   // padding channels of the NHWC buffer stay zero
    if (col_ok)
      for (int r = 0; r < UR && ih0 + r < Hi; ++r) out[(long long)r * Wi * dxpitch] = __float2bfloat16(0.f);
    return;
  }
  // fine column segment needed by this CTA's coarse columns
  int seg_lo, seg_hi, t0, t1;
  cand_range(cb * kThreads, sw, Wo, &seg_lo, &t0);
  cand_range(min(cb * kThreads + kThreads - 1, Wi - 1), sw, Wo, &t1, &seg_hi);
  seg_lo &= ~3;                                     // float4-aligned start
  const int seg_len = seg_hi - seg_lo + 1;
  // this thread's column weights: the first fine column with a non-zero weight and the UKX columns from there
  int wlo = 0;
  float wx[UKX];
#pragma unroll
  for (int k = 0; k < UKX; ++k) wx[k] = 0.f;
  if (col_ok) {
    int whi;
    cand_range(iw, sw, Wo, &wlo, &whi);
    while (wlo < whi && lerp_weight(lerp_src(wlo, sw, Wi), iw) == 0.f) ++wlo;
#pragma unroll
    for (int k = 0; k < UKX; ++k)
      if (wlo + k <= whi) wx[k] = lerp_weight(lerp_src(wlo + k, sw, Wi), iw);
  }
  const int koff = wlo - seg_lo;
  float acc[UR];
#pragma unroll
  for (int r = 0; r < UR; ++r) acc[r] = 0.f;
  const float* plane = dy + ((long long)n * C + c) * Ho * Wo;
  const bool vec = (Wo % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
#pragma unroll
  for (int k = 0; k <= UR; ++k) {
    const int ih = ih0 - 1 + k;                      // interval [ih, ih+1): weight w0 to row ih, w1 to row ih+1
    if (ih < 0 || ih > Hi - 1) continue;             // uniform over the CTA
    const int a = first_fine_row(ih, sh, Hi, Ho);
    const int b = ih == Hi - 1 ? Ho : first_fine_row(ih + 1, sh, Hi, Ho);
    float lo = 0.f, hi = 0.f;
    for (int o0 = a; o0 < b; o0 += URB) {
      const int nr = min(URB, b - o0);
      for (int rr = 0; rr < nr; ++rr) {
        const float* src = plane + (long long)(o0 + rr) * Wo + seg_lo;
        float* dstrow = rowbuf + rr * seg_max;
        // element j is stored at j + (j >> 5): the readers walk the row with a stride of ~1/scale (4) words per
        // lane, which would be a 4-way bank conflict on a dense row and is conflict-free on the skewed one
        if (vec) {
          // seg_lo and Wo are multiples of 4, so a float4 never straddles the row end
          for (int i = threadIdx.x * 4; i < seg_len; i += kThreads * 4) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(src + i));
            float* d = dstrow + i + (i >> 5);
            d[0] = v4.x; d[1] = v4.y; d[2] = v4.z; d[3] = v4.w;
          }
        } else {
          for (int i = threadIdx.x; i < seg_len; i += kThreads) dstrow[i + (i >> 5)] = __ldg(src + i);
        }
      }
      __syncthreads();
      for (int rr = 0; rr < nr; ++rr) {
        const float* row = rowbuf + rr * seg_max;
        float v = 0.f;
#pragma unroll
        for (int kk = 0; kk < UKX; ++kk) {
          const int j = koff + kk;
          v = fmaf(wx[kk], (j < seg_len && j >= 0) ? row[j + (j >> 5)] : 0.f, v);
        }
        const Lerp ly = lerp_src(o0 + rr, sh, Hi);  // ly.i0 == ih
        lo = fmaf(ly.w0, v, lo);
        hi = fmaf(ly.w1, v, hi);
      }
      __syncthreads();
    }
    if (k >= 1) acc[k - 1] += lo;
    if (ih + 1 <= Hi - 1) {
      if (k <= UR - 1) acc[k] += hi;
    } else if (k >= 1) {
      acc[k - 1] += hi;                              // i1 is clamped to the last row
    }
  }
  const float gs = gscale ? __ldg(gscale) : 1.f;
  if (col_ok)
    for (int r = 0; r < UR && ih0 + r < Hi; ++r) out[(long long)r * Wi * dxpitch] = __float2bfloat16(acc[r] * gs);
}

// Vertical-first variant of the row-staged kernel (used when the rows can be read as float4): the kernel above reduces
// every FINE row horizontally (10 bounds-checked shared-memory taps per fine row and coarse column: 179 M warp
// instructions for the final x4 gradient, issue-bound at 1.3 TB/s).  Here a thread first reduces its four fine columns
// VERTICALLY in registers while the rows stream in (one LDG.128 + 8 FMA per fine row, all loads of an interval in
// flight together), the UR vertically reduced rows go to shared memory once, and the horizontal 10-tap reduction runs
// once per COARSE row.  Same sums, different order.
__global__ void __launch_bounds__(kThreads, 3)   // <= 85 registers: without the bound ptxas hoists all 54 row loads (255 registers, spills)
up_from_nchw_bwd_sep_kernel(const float* __restrict__ dy, int C, int Ho, int Wo, __nv_bfloat16* __restrict__ dx,
                            int dxpitch, int Hi, int Wi, float sh, float sw, int col_chunks, int seg_max,
                            const float* __restrict__ gscale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float rowbuf[];   // [UR][seg_max], skewed like above
  const int ihb = blockIdx.x / col_chunks, cb = blockIdx.x - ihb * col_chunks;
  const int c = blockIdx.y, n = blockIdx.z;
  const int ih0 = ihb * UR;
  const int iw = cb * kThreads + threadIdx.x;
  const bool col_ok = iw < Wi;
  __nv_bfloat16* out = dx + (((long long)n * Hi + ih0) * Wi + (col_ok ? iw : 0)) * dxpitch + c;
  if (c >= C) {   // padding channels of the NHWC buffer stay zero
    if (col_ok)
      for (int r = 0; r < UR && ih0 + r < Hi; ++r) out[(long long)r * Wi * dxpitch] = __float2bfloat16(0.f);
    return;
  }
  int seg_lo, seg_hi, t0, t1;
  cand_range(cb * kThreads, sw, Wo, &seg_lo, &t0);
  cand_range(min(cb * kThreads + kThreads - 1, Wi - 1), sw, Wo, &t1, &seg_hi);
  seg_lo &= ~3;
  const int seg_len = seg_hi - seg_lo + 1;
  const int ngroups = (seg_len + 3) >> 2;
  // interval k = 0 .. UR belongs to coarse row ih = ih0 - 1 + k: its fine rows [fa(ih), fa(ih + 1)) weigh w0 on row ih and
  // w1 on row ih + 1, so output row k - 1 = hi(interval k - 1) + lo(interval k) and is complete -- and written to shared
  // memory -- as soon as interval k has been reduced: three float4 accumulators instead of UR of them
  const float* plane = dy + ((long long)n * C + c) * Ho * Wo + seg_lo;
  for (int g0 = 0; g0 < ngroups; g0 += kThreads) {
    const int g = g0 + threadIdx.x;
    const bool gv = g < ngroups;
    float ph[4] = {0.f, 0.f, 0.f, 0.f};
    int ra = ih0 - 1 <= 0 ? 0 : first_fine_row(ih0 - 1, sh, Hi, Ho);
#pragma unroll 1
    for (int k = 0; k <= UR; ++k) {
      const int ih = ih0 - 1 + k;
      const int rb = ih + 1 <= 0 ? 0 : (ih + 1 > Hi - 1 ? Ho : first_fine_row(ih + 1, sh, Hi, Ho));
      float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
      for (int r0 = ra; r0 < rb; r0 += URB) {
        float4 vv[URB];
#pragma unroll
        for (int rr = 0; rr < URB; ++rr) {
          vv[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gv && r0 + rr < rb) vv[rr] = __ldg(reinterpret_cast<const float4*>(plane + (long long)(r0 + rr) * Wo) + g);
        }
#pragma unroll
        for (int rr = 0; rr < URB; ++rr) {
          const float4 v = vv[rr];
          const Lerp ly = lerp_src(r0 + rr, sh, Hi);
          lo[0] = fmaf(ly.w0, v.x, lo[0]); lo[1] = fmaf(ly.w0, v.y, lo[1]); lo[2] = fmaf(ly.w0, v.z, lo[2]); lo[3] = fmaf(ly.w0, v.w, lo[3]);
          hi[0] = fmaf(ly.w1, v.x, hi[0]); hi[1] = fmaf(ly.w1, v.y, hi[1]); hi[2] = fmaf(ly.w1, v.z, hi[2]); hi[3] = fmaf(ly.w1, v.w, hi[3]);
        }
      }
      const bool clamped = ih + 1 > Hi - 1;          // i1 is clamped to the last row: w1 stays on row ih
      if (k >= 1 && gv) {
        float* d = rowbuf + (k - 1) * seg_max + 4 * g + ((4 * g) >> 5);   // the four columns share one skew offset
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = ph[e] + lo[e] + (clamped ? hi[e] : 0.f);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) ph[e] = clamped ? 0.f : hi[e];
      ra = rb;
    }
  }
  __syncthreads();
  if (!col_ok) return;
  const float gs = gscale ? __ldg(gscale) : 1.f;   // deferred scale of the loss (mean cross entropy): see functional.py
  int wlo, whi;
  cand_range(iw, sw, Wo, &wlo, &whi);
  while (wlo < whi && lerp_weight(lerp_src(wlo, sw, Wi), iw) == 0.f) ++wlo;
  float wx[UKX];
  int jj[UKX];
#pragma unroll
  for (int k = 0; k < UKX; ++k) {
    int j = wlo + k - seg_lo;
    const bool ok = wlo + k <= whi && j >= 0 && j < seg_len;
    wx[k] = ok ? lerp_weight(lerp_src(wlo + k, sw, Wi), iw) : 0.f;
    if (!ok) j = 0;
    jj[k] = j + (j >> 5);
  }
#pragma unroll
  for (int r = 0; r < UR; ++r) {
    if (ih0 + r >= Hi) break;
    const float* row = rowbuf + r * seg_max;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < UKX; ++k) v = fmaf(wx[k], row[jj[k]], v);
    out[(long long)r * Wi * dxpitch] = __float2bfloat16(v * gs);
  }
}

// NHWC bf16 gradient of the align-corners up-sampling, column-stationary: the gather kernel up_nhwc_bwd_kernel walks
// the ~12 x 12 candidate fine pixels of every coarse pixel and recomputes both interpolation weights for each of them
// (83 M warp instructions for the ASPP -> decoder x4 gradient: issue-bound at 1.4 TB/s).  Here a thread owns (coarse
// column, 8-channel group) and walks the fine rows of UCR coarse rows once: per fine row it reduces its <= UKX fine
// columns with weights held in registers (16-byte loads, 512 contiguous bytes per pixel across the warp), applies the
// two row weights to rolling accumulators (row k - 1 = hi(interval k - 1) + lo(interval k)) and writes every coarse
// row as soon as it is complete.  No shared memory, no atomics, deterministic.
constexpr int UCR = 4;     // coarse rows per CTA of the column-stationary kernel (more CTAs; 1.25x vertical halo reads)
__global__ void __launch_bounds__(kThreads, 2)
up_nhwc_bwd_cols_kernel(const __nv_bfloat16* __restrict__ dy, int dypitch, int dyoff, int Hi, int Wi, int C, int Ho, int Wo,
                        __nv_bfloat16* __restrict__ dx, float sh, float sw, int col_chunks) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = C >> 3, cols = kThreads / cg;
  const int g = threadIdx.x % cg, col = threadIdx.x / cg;
  const int ihb = blockIdx.x / col_chunks, cb = blockIdx.x - ihb * col_chunks;
  const int n = blockIdx.y, ih0 = ihb * UCR;
  const int iw = cb * cols + col;
  if (col >= cols || iw >= Wi) return;
  int wlo, whi;
  cand_range(iw, sw, Wo, &wlo, &whi);
  while (wlo < whi && lerp_weight(lerp_src(wlo, sw, Wi), iw) == 0.f) ++wlo;
  float wx[UKX];
#pragma unroll
  for (int k = 0; k < UKX; ++k) wx[k] = wlo + k <= whi ? lerp_weight(lerp_src(wlo + k, sw, Wi), iw) : 0.f;
  const __nv_bfloat16* base = dy + (long long)n * Ho * Wo * dypitch + dyoff + g * 8 + (long long)wlo * dypitch;
  __nv_bfloat16* out = dx + (((long long)n * Hi + ih0) * Wi + iw) * C + g * 8;
  float ph[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ph[i] = 0.f;
  int ra = ih0 - 1 <= 0 ? 0 : first_fine_row(ih0 - 1, sh, Hi, Ho);
#pragma unroll 1
  for (int k = 0; k <= UCR; ++k) {
    const int ih = ih0 - 1 + k;
    const int rb = ih + 1 <= 0 ? 0 : (ih + 1 > Hi - 1 ? Ho : first_fine_row(ih + 1, sh, Hi, Ho));
    float lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { lo[i] = 0.f; hi[i] = 0.f; }
#pragma unroll 2
    for (int r = ra; r < rb; ++r) {
      const __nv_bfloat16* rowp = base + (long long)r * Wo * dypitch;
      uint4 v[UKX];
#pragma unroll
      for (int kk = 0; kk < UKX; ++kk) {
        v[kk] = make_uint4(0, 0, 0, 0);
        if (wx[kk] != 0.f) v[kk] = ldg16(rowp + (long long)kk * dypitch);
      }
      float h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < UKX; ++kk) {
        if (wx[kk] == 0.f) continue;                 // uniform over the warp (one coarse column per warp when C >= 256)
        float f[8];
        bf16x8_to_float(v[kk], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = fmaf(wx[kk], f[i], h[i]);
      }
      const Lerp ly = lerp_src(r, sh, Hi);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lo[i] = fmaf(ly.w0, h[i], lo[i]);
        hi[i] = fmaf(ly.w1, h[i], hi[i]);
      }
    }
    const bool clamped = ih + 1 > Hi - 1;            // i1 is clamped to the last row: w1 stays on row ih
    if (k >= 1 && ih0 + k - 1 < Hi) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = ph[i] + lo[i] + (clamped ? hi[i] : 0.f);
      *reinterpret_cast<uint4*>(out + (long long)(k - 1) * Wi * C) = float_to_bf16x8(o);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) ph[i] = clamped ? 0.f : hi[i];
    ra = rb;
  }
}

// ------------------------------------------------------------ global average pool
// one CTA per (image, 64-channel slab); y[n][c] = mean_p x[n][p][c]
__global__ void __launch_bounds__(kThreads)
avgpool_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int pitch, int coff,
               __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, float scale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  __shared__ float sm[kThreads / 8][8][8];
  const int n = blockIdx.y;
  const int g = blockIdx.x * 8 + (threadIdx.x & 7);  // channel group of 8
  const int r = threadIdx.x >> 3;                    // 32 row slots
  const int cg = C / 8;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (g < cg) {
    const __nv_bfloat16* b = x + (long long)n * HW * pitch + coff + g * 8;
    for (int p = r; p < HW; p += kThreads / 8) {
      float f[8];
      bf16x8_to_float(ldg16(b + (long long)p * pitch), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[r][threadIdx.x & 7][i] = s[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int gg = threadIdx.x >> 3, i = threadIdx.x & 7;
    float t = 0.f;
    for (int rr = 0; rr < kThreads / 8; ++rr) t += sm[rr][gg][i];
    const int c = (blockIdx.x * 8 + gg) * 8 + i;
    if (c < C) {
      if (y_bf16) y_bf16[(long long)n * C + c] = __float2bfloat16(t * scale);
      if (y_f32) y_f32[(long long)n * C + c] = t * scale;
    }
  }
}

// dx[n][p][c] = dy[n][c] * scale   (backward of the mean; also a plain broadcast with scale = 1)
__global__ void __launch_bounds__(kThreads)
broadcast_kernel(const __nv_bfloat16* __restrict__ v, int N, int HW, int C, float scale, int accumulate,
                 __nv_bfloat16* __restrict__ y, int ypitch, int yoff) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = C / 8;
  const long long total = (long long)N * HW * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const int g = (int)(t % cg);
    const long long p = t / cg;
    const int n = (int)(p / HW);
    float f[8];
    bf16x8_to_float(ldg16(v + (long long)n * C + g * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] *= scale;
    if (accumulate) {
      float o[8];
      bf16x8_to_float(*reinterpret_cast<const uint4*>(y + p * ypitch + yoff + g * 8), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += o[i];
    }
    *reinterpret_cast<uint4*>(y + p * ypitch + yoff + g * 8) = float_to_bf16x8(f);
  }
}

// ------------------------------------------------------------ layout conversions
// x NCHW fp32 -> y NHWC bf16 with channels [C, ypitch) zero filled; pixel index fastest
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, long long HW,
                    __nv_bfloat16* __restrict__ y, int ypitch) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = ypitch / 8;
  const long long npix = (long long)N * HW;
  const long long total = npix * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const long long p = t % npix;
    const int g = (int)(t / npix);
    const long long n = p / HW, px = p - n * HW;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      f[i] = c < C ? __ldg(x + (n * C + c) * HW + px) : 0.f;
    }
    *reinterpret_cast<uint4*>(y + p * ypitch + g * 8) = float_to_bf16x8(f);
  }
}

// x NHWC bf16 (pitch xpitch) -> y NCHW fp32, first C channels
__global__ void __launch_bounds__(kThreads)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int xpitch, int N, int C, long long HW,
                    float* __restrict__ y) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int cg = (C + 7) / 8;
  const long long npix = (long long)N * HW;
  const long long total = npix * cg;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (long long)gridDim.x * kThreads) {
    const long long p = t % npix;
    const int g = (int)(t / npix);
    const long long n = p / HW, px = p - n * HW;
    float f[8];
    bf16x8_to_float(ldg16(x + p * xpitch + g * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      if (c < C) y[(n * C + c) * HW + px] = f[i];
    }
  }
}

// dx = dy * (y > 0 ? 1 : slope); y is the POST-activation tensor (sign preserved by leaky relu)
__global__ void __launch_bounds__(kThreads)
leaky_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                 __nv_bfloat16* __restrict__ dx, long long nvec, float slope) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < nvec;
       t += (long long)gridDim.x * kThreads) {
    float g[8], a[8];
    bf16x8_to_float(ldg16(dy + t * 8), g);
    bf16x8_to_float(ldg16(y + t * 8), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] *= (a[i] > 0.f ? 1.f : slope);
    *reinterpret_cast<uint4*>(dx + t * 8) = float_to_bf16x8(g);
  }
}

inline bool al16(const void* p) { return (uintptr_t)p % 16 == 0; }

}  // namespace

extern "C" int s2r_upsample_bilinear_nhwc(const void* x, int N, int Hi, int Wi, int C, void* y, int Ho,
                                          int Wo, int ypitch, int yoff, s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, S2R_ERR_SHAPE, "upsample: bad shape");
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && ypitch % 8 == 0 && yoff % 8 == 0 && ypitch >= yoff + C && al16(x) && al16(y),
              S2R_ERR_SHAPE, "upsample: channels/pitch must be multiples of 8 and buffers 16B aligned");
  const long long total = (long long)N * Ho * Wo * (C / 8);
  if (total < (1ll << 32))
    S2R_CUDA_OK(s2r_launch(up_nhwc_fwd_kernel<unsigned>, dim3(s2r_grid(total, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, N, Hi, Wi, C, (__nv_bfloat16*)y, Ho, Wo, ypitch, yoff, ac_scale(Hi, Ho),
        ac_scale(Wi, Wo)));
  else
    S2R_CUDA_OK(s2r_launch(up_nhwc_fwd_kernel<long long>, dim3(s2r_grid(total, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, N, Hi, Wi, C, (__nv_bfloat16*)y, Ho, Wo, ypitch, yoff, ac_scale(Hi, Ho),
        ac_scale(Wi, Wo)));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_upsample_bilinear_nhwc_bwd(const void* dy, int dypitch, int dyoff, int N, int Hi,
                                              int Wi, int C, int Ho, int Wo, void* dx,
                                              s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, S2R_ERR_SHAPE, "upsample_bwd: bad shape");
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && dypitch % 8 == 0 && dyoff % 8 == 0 && dypitch >= dyoff + C && al16(dy) && al16(dx),
              S2R_ERR_SHAPE, "upsample_bwd: channels/pitch must be multiples of 8 and buffers 16B aligned");
  const long long total = (long long)N * Hi * Wi * (C / 8);
  {
    // column-stationary kernel for up-sampling factors up to ~4.4 (at most UKX fine columns per coarse column)
    const float sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
    const int cg = C / 8;
    static int use_cols = -1;
    if (use_cols < 0) {
      const char* e = getenv("S2R_UPBWD_NHWC");     // S2R_UPBWD_NHWC=gather keeps the gather kernel (A/B testing)
      use_cols = (e && e[0] == 'g') ? 0 : 1;
    }
    if (use_cols && sw > 0.f && sh > 0.f && 2.f / sw + 1.f <= (float)UKX && cg <= kThreads && N <= 65535) {
      const int cols = kThreads / cg, col_chunks = s2r_div_up(Wi, cols);
      dim3 grid(s2r_div_up(Hi, UCR) * col_chunks, N);
      S2R_CUDA_OK(s2r_launch(up_nhwc_bwd_cols_kernel, grid, dim3(kThreads), (size_t)0, (cudaStream_t)stream,
                             (const __nv_bfloat16*)dy, dypitch, dyoff, Hi, Wi, C, Ho, Wo, (__nv_bfloat16*)dx, sh, sw, col_chunks));
      S2R_LAUNCH_OK();
      return S2R_OK;
    }
  }
  if (total < (1ll << 32))
    S2R_CUDA_OK(s2r_launch(up_nhwc_bwd_kernel<unsigned>, dim3(s2r_grid(total, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)dy, dypitch, dyoff, N, Hi, Wi, C, Ho, Wo, (__nv_bfloat16*)dx,
        ac_scale(Hi, Ho), ac_scale(Wi, Wo)));
  else
    S2R_CUDA_OK(s2r_launch(up_nhwc_bwd_kernel<long long>, dim3(s2r_grid(total, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)dy, dypitch, dyoff, N, Hi, Wi, C, Ho, Wo, (__nv_bfloat16*)dx,
        ac_scale(Hi, Ho), ac_scale(Wi, Wo)));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_upsample_bilinear_nhwc_to_nchw(const void* x, int xpitch, int N, int Hi, int Wi,
                                                  int C, float* y, int Ho, int Wo,
                                                  s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1, S2R_ERR_SHAPE, "upsample_to_nchw: bad shape");
  const int cgs = (C + 7) / 8;
  S2R_REQUIRE(xpitch % 8 == 0 && xpitch >= cgs * 8 && al16(x), S2R_ERR_SHAPE,
              "upsample_to_nchw: pitch %d must be a multiple of 8 covering C=%d", xpitch, C);
  S2R_REQUIRE(cgs <= 4, S2R_ERR_UNSUPPORTED, "upsample_to_nchw: C=%d > 32 not supported", C);
  const long long total = (long long)N * Ho * Wo;
  const int grid = s2r_grid(total, kThreads, 16);
  const float sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  cudaStream_t st = (cudaStream_t)stream;
#define S2R_UP(CG_, IT_) S2R_CUDA_OK(s2r_launch(up_to_nchw_kernel<CG_, IT_>, dim3(grid), dim3(kThreads), (size_t)0, st, xb, xpitch, N, Hi, Wi, C, y, Ho, Wo, sh, sw))
  const bool small = total < (1ll << 32);
  switch (cgs) {
    case 1: if (small) S2R_UP(1, unsigned); else S2R_UP(1, long long); break;
    case 2: if (small) S2R_UP(2, unsigned); else S2R_UP(2, long long); break;
    case 3: if (small) S2R_UP(3, unsigned); else S2R_UP(3, long long); break;
    default: if (small) S2R_UP(4, unsigned); else S2R_UP(4, long long); break;
  }
#undef S2R_UP
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled(const float* dy, int N, int C, int Ho, int Wo, void* dx,
                                                             int dxpitch, int Hi, int Wi, const float* scale,
                                                             s2r_stream_t stream);

extern "C" int s2r_upsample_bilinear_nchw_bwd_to_nhwc(const float* dy, int N, int C, int Ho, int Wo,
                                                      void* dx, int dxpitch, int Hi, int Wi,
                                                      s2r_stream_t stream) {
  return s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled(dy, N, C, Ho, Wo, dx, dxpitch, Hi, Wi, nullptr, stream);
}

/* scale: device scalar the result is multiplied by (NULL = 1): lets a loss hand over its UNSCALED gradient and the
 * 1 / sum-of-weights factor separately instead of making one more pass over the largest tensor of the step. */
extern "C" int s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled(const float* dy, int N, int C, int Ho, int Wo, void* dx,
                                                             int dxpitch, int Hi, int Wi, const float* scale,
                                                             s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1 && dxpitch >= C, S2R_ERR_SHAPE,
              "upsample_from_nchw_bwd: bad shape");
  const float sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
  // row-staged kernel when at most UKX fine columns weigh on a coarse column and an interval has at most URB rows
  // (up-sampling factors up to ~4.4 in both directions)
  if (sw > 0.f && sh > 0.f && 2.f / sw + 1.f <= (float)UKX && 1.f / sh + 1.f <= (float)URB && N <= 65535 && dxpitch <= 65535) {
    const int col_chunks = s2r_div_up(Wi, kThreads);
    int seg_max = (int)((kThreads + 2) / sw) + 16;
    seg_max = (seg_max + (seg_max >> 5) + 4) & ~3;   // skewed row length
    dim3 grid(s2r_div_up(Hi, UR) * col_chunks, dxpitch, N);
    static int use_sep = -1;
    if (use_sep < 0) {
      const char* e = getenv("S2R_UPBWD");        // S2R_UPBWD=rows keeps the horizontal-first kernel (A/B testing)
      use_sep = (e && e[0] == 'r') ? 0 : 1;
    }
    const size_t sep_smem = (size_t)UR * seg_max * sizeof(float);
    if (use_sep && Wo % 4 == 0 && ((uintptr_t)dy & 15) == 0 && sep_smem <= 48 * 1024)
      S2R_CUDA_OK(s2r_launch(up_from_nchw_bwd_sep_kernel, dim3(grid), dim3(kThreads), (size_t)(sep_smem), (cudaStream_t)stream, 
          dy, C, Ho, Wo, (__nv_bfloat16*)dx, dxpitch, Hi, Wi, sh, sw, col_chunks, seg_max, scale));
    else
      S2R_CUDA_OK(s2r_launch(up_from_nchw_bwd_rows_kernel, dim3(grid), dim3(kThreads), (size_t)((size_t)URB * seg_max * sizeof(float)), (cudaStream_t)stream, 
          dy, C, Ho, Wo, (__nv_bfloat16*)dx, dxpitch, Hi, Wi, sh, sw, col_chunks, seg_max, scale));
    S2R_LAUNCH_OK();
    return S2R_OK;
  }
  const long long total = (long long)N * dxpitch * Hi * Wi;
  S2R_CUDA_OK(s2r_launch(up_from_nchw_bwd_kernel, dim3(s2r_grid(total, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      dy, N, C, Ho, Wo, (__nv_bfloat16*)dx, dxpitch, Hi, Wi, sh, sw, scale));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_avgpool_nhwc(const void* x, int N, int HW, int C, int pitch, int coff, float scale,
                                void* y_bf16, float* y_f32, s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0 && pitch % 8 == 0 && coff % 8 == 0 && al16(x),
              S2R_ERR_SHAPE, "avgpool: bad shape/alignment");
  dim3 grid(s2r_div_up(C / 8, 8), N);
  S2R_CUDA_OK(s2r_launch(avgpool_kernel, dim3(grid), dim3(kThreads), (size_t)0, (cudaStream_t)stream, (const __nv_bfloat16*)x, HW, C, pitch, coff,
                                                             (__nv_bfloat16*)y_bf16, y_f32, scale));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_broadcast_nhwc(const void* v, int N, int HW, int C, float scale, int accumulate,
                                  void* y, int ypitch, int yoff, s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0 && ypitch % 8 == 0 && yoff % 8 == 0 &&
                  ypitch >= yoff + C && al16(v) && al16(y),
              S2R_ERR_SHAPE, "broadcast: bad shape/alignment");
  const long long total = (long long)N * HW * (C / 8);
  S2R_CUDA_OK(s2r_launch(broadcast_kernel, dim3(s2r_grid(total, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)v, N, HW, C, scale, accumulate, (__nv_bfloat16*)y, ypitch, yoff));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int64_t HW, void* y, int ypitch,
                                         s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && C >= 1 && HW >= 1 && ypitch % 8 == 0 && ypitch >= C && al16(y), S2R_ERR_SHAPE,
              "nchw_to_nhwc: bad shape (pitch must be a multiple of 8 >= C)");
  const long long total = (long long)N * HW * (ypitch / 8);
  S2R_CUDA_OK(s2r_launch(nchw_to_nhwc_kernel, dim3(s2r_grid(total, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      x, N, C, HW, (__nv_bfloat16*)y, ypitch));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_nhwc_bf16_to_nchw_f32(const void* x, int xpitch, int N, int C, int64_t HW, float* y,
                                         s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && C >= 1 && HW >= 1 && xpitch % 8 == 0 && xpitch >= ((C + 7) / 8) * 8 && al16(x),
              S2R_ERR_SHAPE, "nhwc_to_nchw: bad shape");
  const long long total = (long long)N * HW * ((C + 7) / 8);
  S2R_CUDA_OK(s2r_launch(nhwc_to_nchw_kernel, dim3(s2r_grid(total, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, xpitch, N, C, HW, y));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_leaky_relu_bwd_bf16(const void* dy, const void* y, void* dx, int64_t n, float slope,
                                       s2r_stream_t stream) {
  S2R_REQUIRE(n % 8 == 0 && al16(dy) && al16(y) && al16(dx), S2R_ERR_SHAPE, "leaky_relu_bwd: n must be a multiple of 8");
  if (n == 0) return S2R_OK;
  S2R_CUDA_OK(s2r_launch(leaky_bwd_kernel, dim3(s2r_grid(n / 8, kThreads * 2, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, n / 8, slope));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
