// Source index / weight arithmetic of bilinear resampling with align_corners=True (PyTorch's upsample_bilinear2d:
// scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = int(src), i1 = i0 + (i0 < in-1)), shared by resize.cu and by the
// fused up-sampling + argmax + confusion-matrix kernel of evaluator.cu -- ONE definition of the interpolated value, with
// explicit roundings, so that the fused kernel sees bit for bit the logits the two-step path writes.
#pragma once
#include "common.cuh"

struct Lerp {
  int i0, i1;
  float w0, w1;
};

__device__ __forceinline__ Lerp lerp_src(int o, float scale, int in) {
  Lerp l;
  const float t = scale * (float)o;
  int i0 = (int)t;
  if (i0 > in - 1) i0 = in - 1;
  l.i0 = i0;
  l.i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l.w1 = t - (float)i0;
  l.w0 = 1.f - l.w1;
  return l;
}

__host__ __device__ __forceinline__ float ac_scale(int in, int out) {
  return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
}

// ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11), every rounding spelled out
__device__ __forceinline__ float bilerp(const Lerp& ly, const Lerp& lx, float v00, float v01, float v10, float v11) {
  const float top = __fmaf_rn(lx.w1, v01, __fmul_rn(lx.w0, v00));
  const float bot = __fmaf_rn(lx.w1, v11, __fmul_rn(lx.w0, v10));
  return __fmaf_rn(ly.w1, bot, __fmul_rn(ly.w0, top));
}
