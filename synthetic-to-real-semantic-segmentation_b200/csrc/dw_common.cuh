// Device helpers shared by the streaming depthwise kernels (dwconv_s1.cu, dwconv_s2.cu): packed fp32 math
// (FFMA2 / FMUL2 / FADD2 on register pairs), bf16x4 <-> fp32 conversion, the one-instruction ReLU6 prologue,
// per-thread filter load and the end-of-CTA channel reduction.
#pragma once
#include "tma_util.cuh"

namespace s2r_dw {

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

// 4 packed bf16 -> two float2 (channels 0,1 and 2,3)
__device__ __forceinline__ void unpack4(uint2 u, float2& a, float2& b) {
  a.x = __uint_as_float(u.x << 16);
  a.y = __uint_as_float(u.x & 0xffff0000u);
  b.x = __uint_as_float(u.y << 16);
  b.y = __uint_as_float(u.y & 0xffff0000u);
}
__device__ __forceinline__ uint2 pack4(float2 a, float2 b) {
  uint2 u;
  __nv_bfloat162 p = __floats2bfloat162_rn(a.x, a.y), q = __floats2bfloat162_rn(b.x, b.y);
  u.x = *reinterpret_cast<uint32_t*>(&p);
  u.y = *reinterpret_cast<uint32_t*>(&q);
  return u;
}
// 8-byte global store under a predicate, without control flow (a C++ `if (p) *dst = v` inside the row loop makes
// the compiler branch around the store and costs 30 % of the stride-2 backward kernel)
__device__ __forceinline__ void st8_if(void* dst, uint2 v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.v2.u32 [%0], {%1, %2};\n\t}"
               ::"l"(dst), "r"(v.x), "r"(v.y), "r"((unsigned)p)
               : "memory");
}

// a6 = sat(x*sc6 + sh6) for 4 packed bf16
__device__ __forceinline__ void act4(uint2 raw, float2 scA, float2 scB, float2 shA, float2 shB, float2& aA, float2& aB) {
  float2 xa, xb;
  unpack4(raw, xa, xb);
  aA.x = __saturatef(fmaf(xa.x, scA.x, shA.x));
  aA.y = __saturatef(fmaf(xa.y, scA.y, shA.y));
  aB.x = __saturatef(fmaf(xb.x, scB.x, shB.x));
  aB.y = __saturatef(fmaf(xb.y, scB.y, shB.y));
}

// sum NV per-thread floats over the thread columns of a CTA (consumer threads tid = g + CG*j share g);
// thread t = g*NV + k < CG*NV returns the total for (g, k).  red: [NV][NCONS+1] floats.  All threads call.
template <int NV, int NCONS>
__device__ __forceinline__ float column_reduce(float* red, const float (&v)[NV], int CG, int ncol, bool consumer) {
  if (consumer) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * (NCONS + 1) + threadIdx.x] = v[k];
  }
  __syncthreads();
  float acc = 0.f;
  const int t = threadIdx.x;
  if (t < CG * NV) {
    const int g = t / NV, k = t - g * NV;
    const float* p = red + k * (NCONS + 1) + g;
    for (int j = 0; j < ncol; ++j) acc += p[j * CG];
  }
  return acc;
}

// per-thread filter: w[c..c+3][9] is 36 contiguous floats (144 B, 16-byte aligned since c % 4 == 0)
__device__ __forceinline__ void load_filter(const float* __restrict__ w, int c, float scale, float2* wA, float2* wB) {
  float f[36];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + c * 9) + i);
    f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wA[k] = make_float2(scale * f[k], scale * f[9 + k]);
    wB[k] = make_float2(scale * f[18 + k], scale * f[27 + k]);
  }
}

}  // namespace s2r_dw
