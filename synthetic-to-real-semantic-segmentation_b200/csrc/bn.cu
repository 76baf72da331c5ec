// Batch-norm family on NHWC bf16 activations (fp32/fp64 statistics).
// Replaces modeling/sync_batchnorm/batchnorm.py:48-125 (_SynchronizedBatchNorm.forward,
// _compute_mean_std) and the F.batch_norm fallback at :50-53 of the reference, fused
// with the ReLU/ReLU6 that always follows (mobilenet.py:12-13,41-42,51-56; assp.py:19-21;
// decoder.py:35-37), the residual add (mobilenet.py:65) and Dropout (assp.py:78,
// decoder.py:25,29; domian.py:18,22).
//
// Two-phase structure: channel sums (this rank) -> [host: all-reduce across ranks] ->
// finalize (mean / inv-std / running stats / scale+shift) -> apply.  All kernels are
// HBM-bound: 16-byte vector of 8 channels per thread, channel-group fastest so a warp
// touches contiguous memory.
#include <stdlib.h>

#include "dw_common.cuh"
#include "bn_tail.cuh"

namespace {

// Launch geometry for kernels that reduce over the pixel axis of a [P][C] tensor.
struct RowReduceCfg {
  int cg;       // channel groups of 8
  int rows;     // rows handled concurrently by one CTA
  int threads;  // cg * rows
  int grid;
};

inline RowReduceCfg row_reduce_cfg(long long P, int C) {
  RowReduceCfg c;
  c.cg = C / 8;
  c.rows = 256 / c.cg;
  if (c.rows < 1) c.rows = 1;
  c.threads = c.cg * c.rows;
  // >= 8 rows per thread; one wave of CTAs: three per SM (the kernels' launch bound) on the large tensors, ONE per SM
  // on the small ones, where the fp64 atomics of the final reduction -- every CTA adds into the same 2*C addresses --
  // are the tail (measured, tests/tools/bn_bench.py: 16384x384 10.5 -> 7.4 us, 65536x192 15.0 -> 10.3 us,
  // 1048576x96 80.4 -> 65.4 us = 0.94 of the copy peak)
  long long blocks = (P + (long long)c.rows * 8 - 1) / ((long long)c.rows * 8);
  long long cap = (long long)s2r_sm_count() * (P * (long long)C >= 24LL * 1024 * 1024 ? 3 : 1);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  c.grid = (int)blocks;
  return c;
}

constexpr int UNR = 4;   // independent 16-byte loads in flight per thread and tensor (latency hiding at low occupancy)

// Keeps the batch of loads issued above this point from being sunk to their uses (ptxas otherwise re-serialises
// load -> use -> load to save registers, which costs the memory-level parallelism the unrolling is for).
#define S2R_ISSUE_LOADS_FIRST() asm volatile("" ::: "memory")

// 16-byte read-only load that stays where it is written: LLVM sinks a plain __ldg into the (conditional) block
// that consumes it, which turns a batch of independent loads back into load -> use -> load.
__device__ __forceinline__ uint4 ldg16_pinned(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// sums[0][c] += sum_p x[p][c], sums[1][c] += sum_p x[p][c]^2
__global__ void channel_sums_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, int pitch,
                                    int coff, int rows, double* __restrict__ sums) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ float sm[];  // [rows][cg][16]
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const long long step = (long long)gridDim.x * rows;
  long long p0 = (long long)blockIdx.x * rows + r;
  const __nv_bfloat16* xp = x + p0 * pitch + coff + g * 8;
  const long long sx = step * pitch;
  auto one = [&](const uint4& vv) {
    float f[8];
    bf16x8_to_float(vv, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
  };
  constexpr int UNS = 2 * UNR;   // single input stream: twice the batch
  for (; p0 + (UNS - 1) * step < P; p0 += UNS * step, xp += UNS * sx) {
    uint4 v[UNS];
#pragma unroll
    for (int u = 0; u < UNS; ++u) v[u] = ldg16(xp + u * sx);
#pragma unroll
    for (int u = 0; u < UNS; ++u) one(v[u]);
  }
  for (; p0 < P; p0 += step, xp += sx) one(ldg16(xp));
  float* mine = sm + ((size_t)r * cg + g) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
  __syncthreads();
  // one thread per (channel-group, value) column sums over the rows
  for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
    const int gg = t / 16, k = t % 16;
    double acc = 0;
    for (int rr = 0; rr < rows; ++rr) acc += (double)sm[((size_t)rr * cg + gg) * 16 + k];
    const int ch = gg * 8 + (k & 7);
    atomicAdd(&sums[(k >> 3) * C + ch], acc);
  }
}

// mean/inv-std from the (already cross-rank reduced) sums; running-stat update;
// scale = gamma * invstd, shift = beta - mean * scale.
// clamp_mode 0: invstd = (var + eps)^-1/2   (F.batch_norm, batchnorm.py:50-53)
// clamp_mode 1: invstd = max(var, eps)^-1/2 (sync branch, batchnorm.py:125)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, int clamp_mode, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_invstd, float* __restrict__ scale_shift,
                                   int C) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double sumvar = sums[C + c] - sums[c] * mean;  // batchnorm.py:117-118
  if (sumvar < 0) sumvar = 0;
  const double var = sumvar / count;
  const double invstd = clamp_mode ? 1.0 / sqrt(var > (double)eps ? var : (double)eps)
                                   : 1.0 / sqrt(var + (double)eps);
  if (running_mean) {
    const double unbiased = count > 1 ? sumvar / (count - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = (float)(g * invstd);
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = (float)invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = (float)(b - mean * (double)sc);
}

// eval mode: statistics are the running buffers
__global__ void bn_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                               float* __restrict__ mean_invstd, float* __restrict__ scale_shift, int C) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  mean_invstd[c] = rm[c];
  mean_invstd[C + c] = invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = b - rm[c] * sc;
}

// eval mode, every BatchNorm layer of a model in ONE launch (blockIdx.y = layer): the inference path folds these
// scale / shift pairs into the producing kernels' epilogues, so this is the only BatchNorm launch of a forward pass
__global__ void bn_eval_multi_kernel(const s2r_bn_eval_job* __restrict__ jobs) {
  pdl_wait();
  pdl_trigger();
  const s2r_bn_eval_job j = jobs[blockIdx.y];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.C; c += gridDim.x * blockDim.x) {
    const float invstd = rsqrtf(j.running_var[c] + j.eps);
    const float g = j.gamma ? j.gamma[c] : 1.f, b = j.beta ? j.beta[c] : 0.f;
    const float sc = g * invstd;
    j.mean_invstd[c] = j.running_mean[c];
    j.mean_invstd[j.C + c] = invstd;
    j.scale_shift[c] = sc;
    j.scale_shift[j.C + c] = b - j.running_mean[c] * sc;
  }
}

// Per-thread channel constants: a thread owns one group of 8 channels for its whole lifetime, so the
// per-channel parameters are loaded once into registers (no per-element division or parameter loads).
struct ChanConst {
  float sc[8], sh[8], mu[8], is[8];
};

__device__ __forceinline__ void load_f8(const float* __restrict__ p, float* f) {
  *reinterpret_cast<float4*>(f) = __ldg(reinterpret_cast<const float4*>(p));
  *reinterpret_cast<float4*>(f + 4) = __ldg(reinterpret_cast<const float4*>(p) + 1);
}

__device__ __forceinline__ void chan_init(ChanConst& k, const float* __restrict__ scale_shift,
                                          const float* __restrict__ mean_invstd, int C, int g) {
  load_f8(scale_shift + g * 8, k.sc);
  load_f8(scale_shift + C + g * 8, k.sh);
  if (mean_invstd) {
    load_f8(mean_invstd + g * 8, k.mu);
    load_f8(mean_invstd + C + g * 8, k.is);
  }
}

// y = dropout(act(x * scale + shift)) + residual
// thread = (row slot r, channel group g); rows advance by gridDim.x * rows
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, int xpitch, int xoff,
                const float* __restrict__ scale_shift, int act,
                const __nv_bfloat16* __restrict__ residual, float drop_p, unsigned long long seed,
                const unsigned long long* __restrict__ seed_dev, __nv_bfloat16* __restrict__ y, int ypitch,
                int yoff, int rows) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  if (seed_dev) seed += *seed_dev;
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  ChanConst k;
  chan_init(k, scale_shift, nullptr, C, g);
  const long long step = (long long)gridDim.x * rows;
  for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += UNR * step) {
   uint4 xv[UNR], rv[UNR];
#pragma unroll
   for (int u = 0; u < UNR; ++u) {
     const long long p = min(p0 + u * step, P - 1);   // clamped: unconditional loads are issued back to back
     xv[u] = ldg16_pinned(x + p * xpitch + xoff + g * 8);
     if (residual) rv[u] = ldg16_pinned(residual + p * C + g * 8);
   }
   S2R_ISSUE_LOADS_FIRST();
#pragma unroll
   for (int u = 0; u < UNR; ++u) {
    const long long p = p0 + u * step;
    if (p >= P) break;
    float f[8];
    bf16x8_to_float(xv[u], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = apply_act(fmaf(f[i], k.sc[i], k.sh[i]), act, 0.f);
    if (drop_p > 0.f) {
      float m[8];
      dropout_scale8(seed, (unsigned long long)(p * cg + g), drop_p, m);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= m[i];
    }
    if (residual) {
      float rr[8];
      bf16x8_to_float(rv[u], rr);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += rr[i];
    }
    *reinterpret_cast<uint4*>(y + p * ypitch + yoff + g * 8) = float_to_bf16x8(f);
   }
  }
}

// effective upstream gradient through dropout and the activation:
// dy' = dY * dropout_mask * act'(x * scale + shift)
__device__ __forceinline__ void bn_bwd_load(const uint4& dyv, const uint4& xv, const ChanConst& k,
                                            int act, float drop_p, unsigned long long seed, long long t,
                                            float* gdy, float* xhat) {
  float f[8];
  bf16x8_to_float(dyv, gdy);
  bf16x8_to_float(xv, f);
  float m[8];
  if (drop_p > 0.f) dropout_scale8(seed, (unsigned long long)t, drop_p, m);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float pre = fmaf(f[i], k.sc[i], k.sh[i]);
    float gd = gdy[i] * act_grad(pre, act, 0.f);
    if (drop_p > 0.f) gd *= m[i];
    gdy[i] = gd;
    xhat[i] = (f[i] - k.mu[i]) * k.is[i];
  }
}

// dsums[0][c] += sum_p dy', dsums[1][c] += sum_p dy' * xhat
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int dypitch, int dyoff,
                                     const __nv_bfloat16* __restrict__ x, int xpitch, int xoff,
                                     const float* __restrict__ mean_invstd,
                                     const float* __restrict__ scale_shift, int act, float drop_p,
                                     unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                                     long long P, int C, int rows, double* __restrict__ dsums) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  if (seed_dev) seed += *seed_dev;
  extern __shared__ float sm[];
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  ChanConst k;
  chan_init(k, scale_shift, mean_invstd, C, g);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const long long step = (long long)gridDim.x * rows;
  for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += UNR * step) {
    uint4 dv[UNR], xv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long p = min(p0 + u * step, P - 1);   // clamped: unconditional loads are issued back to back
      dv[u] = ldg16_pinned(dy + p * dypitch + dyoff + g * 8);
      xv[u] = ldg16_pinned(x + p * xpitch + xoff + g * 8);
    }
    S2R_ISSUE_LOADS_FIRST();
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long p = p0 + u * step;
      if (p >= P) break;
      float gdy[8], xh[8];
      bn_bwd_load(dv[u], xv[u], k, act, drop_p, seed, p * cg + g, gdy, xh);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += gdy[i]; q[i] += gdy[i] * xh[i]; }
    }
  }
  float* mine = sm + ((size_t)r * cg + g) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = q[i]; }
  __syncthreads();
  for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
    const int gg = t / 16, kk = t % 16;
    double acc = 0;
    for (int rr = 0; rr < rows; ++rr) acc += (double)sm[((size_t)rr * cg + gg) * 16 + kk];
    atomicAdd(&dsums[(kk >> 3) * C + gg * 8 + (kk & 7)], acc);
  }
}

// dx = scale * (dy' - mean(dy') - xhat * mean(dy' xhat));  count <= 0: frozen stats, dx = scale * dy'
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dypitch, int dyoff,
                    const __nv_bfloat16* __restrict__ x, int xpitch, int xoff,
                    const float* __restrict__ mean_invstd, const float* __restrict__ scale_shift,
                    int act, float drop_p, unsigned long long seed,
                    const unsigned long long* __restrict__ seed_dev,
                    const double* __restrict__ dsums, double count, long long P, int C,
                    __nv_bfloat16* __restrict__ dx, int dxpitch, int dxoff, int win_H, int win_W,
                    int win_pad, int rows) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  if (seed_dev) seed += *seed_dev;
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  ChanConst k;
  chan_init(k, scale_shift, mean_invstd, C, g);
  float m1[8], m2[8];
  const double inv_n = count > 0 ? 1.0 / count : 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m1[i] = count > 0 ? (float)(dsums[g * 8 + i] * inv_n) : 0.f;
    m2[i] = count > 0 ? (float)(dsums[C + g * 8 + i] * inv_n) : 0.f;
  }
  const long long step = (long long)gridDim.x * rows;
  const int Wp = win_W + 2 * win_pad;
  const long long plane = (long long)win_H * win_W;
  for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += UNR * step) {
    uint4 dv[UNR], xv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long p = min(p0 + u * step, P - 1);   // clamped: unconditional loads are issued back to back
      long long pd = p;  // pixel index inside dy (interior window of a padded grid when win_pad > 0)
      if (win_pad > 0) {
        const long long n_ = p / plane;
        const int rem = (int)(p - n_ * plane);
        const int h_ = rem / win_W, w_ = rem - h_ * win_W;
        pd = (n_ * (win_H + 2 * win_pad) + h_ + win_pad) * Wp + w_ + win_pad;
      }
      dv[u] = ldg16_pinned(dy + pd * dypitch + dyoff + g * 8);
      xv[u] = ldg16_pinned(x + p * xpitch + xoff + g * 8);
    }
    S2R_ISSUE_LOADS_FIRST();
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long p = p0 + u * step;
      if (p >= P) break;
      float gdy[8], xh[8], o[8];
      bn_bwd_load(dv[u], xv[u], k, act, drop_p, seed, p * cg + g, gdy, xh);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = k.sc[i] * (gdy[i] - m1[i] - xh[i] * m2[i]);
      *reinterpret_cast<uint4*>(dx + p * dxpitch + dxoff + g * 8) = float_to_bf16x8(o);
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Lean variants of the three streaming kernels above for the common case (no dropout, no gradient window).
// The generic kernels spend ~125-210 instructions per 16-byte vector (64-bit index arithmetic per load, generic
// activation code) and are ISSUE-bound at half the HBM roofline (ncu: profiles/r1_bn_ncu.txt).  Here:
//   * pointer-increment addressing, one fast path for full batches of UNR rows;
//   * packed FFMA2 arithmetic with the per-channel affine maps folded on entry:
//       apply:      y  = act(x*sc + sh) (+ residual)        relu6(v) = 6*sat(v/6): one FFMA.SAT
//       bwd reduce: s += gd, q += gd*x  with gd = dy*act'(pre); the caller's (sum gd, sum gd*xhat) follow as
//                   q_hat = invstd*(q - mean*s)
//       bwd apply:  dx = A*gd + B*x + D,  A = sc, B = -sc*m2*invstd, D = sc*(m2*invstd*mean - m1)
template <int ACT>
__device__ __forceinline__ float2 act_mask2(float2 pre6, float2 v) {
  // pre6: pre-activation (ACT_RELU) or sat(pre/6) (ACT_RELU6); returns v where act' = 1, else 0
  if (ACT == S2R_ACT_RELU) {
    v.x = pre6.x > 0.f ? v.x : 0.f;
    v.y = pre6.y > 0.f ? v.y : 0.f;
  } else if (ACT == S2R_ACT_RELU6) {
    v.x = (pre6.x > 0.f && pre6.x < 1.f) ? v.x : 0.f;
    v.y = (pre6.y > 0.f && pre6.y < 1.f) ? v.y : 0.f;
  }
  return v;
}

struct Lean8 {   // 8 channels as four float2 pairs
  float2 v[4];
};
__device__ __forceinline__ Lean8 lean_unpack(const uint4& u) {
  Lean8 r;
  r.v[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
  r.v[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  r.v[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
  r.v[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
  return r;
}
__device__ __forceinline__ uint4 lean_pack(const Lean8& r) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(r.v[0].x, r.v[0].y); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r.v[1].x, r.v[1].y); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r.v[2].x, r.v[2].y); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r.v[3].x, r.v[3].y); u.w = *reinterpret_cast<uint32_t*>(&t);
  return u;
}
__device__ __forceinline__ void load_pairs(const float* __restrict__ p, float scale, float2* o) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = make_float2(a.x * scale, a.y * scale);
  o[1] = make_float2(a.z * scale, a.w * scale);
  o[2] = make_float2(b.x * scale, b.y * scale);
  o[3] = make_float2(b.z * scale, b.w * scale);
}

template <int ACT, bool RES, bool DROP>
__global__ void __launch_bounds__(256)
bn_apply_lean_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, int xpitch, int xoff,
                     const float* __restrict__ scale_shift, const __nv_bfloat16* __restrict__ residual,
                     __nv_bfloat16* __restrict__ y, int ypitch, int yoff, int rows, float drop_p,
                     unsigned long long seed, const unsigned long long* __restrict__ seed_dev, const BnTail bn) {
  pdl_wait();
  pdl_trigger();
  if (DROP && seed_dev) seed += *seed_dev;
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  const float k = ACT == S2R_ACT_RELU6 ? (1.f / 6.f) : 1.f;
  float2 sc[4], sh[4];
  if (bn.enabled) {
    // pending BatchNorm: scale / shift from the producer's sums (bn_tail.cuh), computed ONCE per CTA by the threads
    // of row slot 0 and handed to the other row slots through shared memory; block 0 publishes them
    extern __shared__ float s_fin[];   // [2][C], pre-multiplied by k
    if (r == 0) {
      float fsc[8], fsh[8];
      const bool pub = blockIdx.x == 0;
      bn_fin4(bn, C, g * 8, pub, fsc, fsh);
      bn_fin4(bn, C, g * 8 + 4, pub, fsc + 4, fsh + 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s_fin[g * 8 + i] = fsc[i] * k;
        s_fin[C + g * 8 + i] = fsh[i] * k;
      }
    }
    __syncthreads();
    const float4 a0 = *reinterpret_cast<const float4*>(s_fin + g * 8), a1 = *reinterpret_cast<const float4*>(s_fin + g * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(s_fin + C + g * 8), b1 = *reinterpret_cast<const float4*>(s_fin + C + g * 8 + 4);
    sc[0] = make_float2(a0.x, a0.y); sc[1] = make_float2(a0.z, a0.w); sc[2] = make_float2(a1.x, a1.y); sc[3] = make_float2(a1.z, a1.w);
    sh[0] = make_float2(b0.x, b0.y); sh[1] = make_float2(b0.z, b0.w); sh[2] = make_float2(b1.x, b1.y); sh[3] = make_float2(b1.z, b1.w);
  } else {
    load_pairs(scale_shift + g * 8, k, sc);
    load_pairs(scale_shift + C + g * 8, k, sh);
  }
  const long long step = (long long)gridDim.x * rows;
  long long p0 = (long long)blockIdx.x * rows + r;
  const __nv_bfloat16* xp = x + p0 * xpitch + xoff + g * 8;
  const __nv_bfloat16* rp = RES ? residual + p0 * C + g * 8 : nullptr;
  __nv_bfloat16* yp = y + p0 * ypitch + yoff + g * 8;
  const long long sx = step * xpitch, sr = step * C, sy = step * ypitch;
  auto one = [&](const uint4& xvv, const uint4& rvv, __nv_bfloat16* dst, long long prow) {
    Lean8 v = lean_unpack(xvv);
    float m[8];
    if (DROP) dropout_scale8(seed, (unsigned long long)(prow * cg + g), drop_p, m);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (ACT == S2R_ACT_RELU6) {
        v.v[i].x = 6.f * __saturatef(fmaf(v.v[i].x, sc[i].x, sh[i].x));
        v.v[i].y = 6.f * __saturatef(fmaf(v.v[i].y, sc[i].y, sh[i].y));
      } else {
        v.v[i] = s2r_dw::ffma2(v.v[i], sc[i], sh[i]);
        if (ACT == S2R_ACT_RELU) {
          v.v[i].x = fmaxf(v.v[i].x, 0.f);
          v.v[i].y = fmaxf(v.v[i].y, 0.f);
        }
      }
      if (DROP) {
        v.v[i].x *= m[2 * i];
        v.v[i].y *= m[2 * i + 1];
      }
    }
    if (RES) {
      const Lean8 rr = lean_unpack(rvv);
#pragma unroll
      for (int i = 0; i < 4; ++i) v.v[i] = s2r_dw::fadd2(v.v[i], rr.v[i]);
    }
    *reinterpret_cast<uint4*>(dst) = lean_pack(v);
  };
  // full batches: straight-line code (no per-row branches), so all UNR loads issue before the first use
  for (; p0 + (UNR - 1) * step < P; p0 += UNR * step, xp += UNR * sx, yp += UNR * sy, rp += RES ? UNR * sr : 0) {
    uint4 xv[UNR], rv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      xv[u] = ldg16(xp + u * sx);
      rv[u] = RES ? ldg16(rp + u * sr) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) one(xv[u], rv[u], yp + u * sy, p0 + u * step);
  }
  for (; p0 < P; p0 += step, xp += sx, yp += sy, rp += RES ? sr : 0)
    one(ldg16(xp), RES ? ldg16(rp) : make_uint4(0u, 0u, 0u, 0u), yp, p0);
}

// APPLY = false: dsums += [sum gd, sum gd*xhat];  APPLY = true: dx = scale*(gd - m1 - xhat*m2)
template <int ACT, bool APPLY, bool DROP>
__global__ void __launch_bounds__(256, 3)
bn_bwd_lean_kernel(const __nv_bfloat16* __restrict__ dy, int dypitch, int dyoff, const __nv_bfloat16* __restrict__ x,
                   int xpitch, int xoff, const float* __restrict__ mean_invstd, const float* __restrict__ scale_shift,
                   double* __restrict__ dsums, double count, long long P, int C, __nv_bfloat16* __restrict__ dx,
                   int dxpitch, int dxoff, int rows, float* __restrict__ dgamma, float* __restrict__ dbeta, float drop_p,
                   unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ float sm[];
  pdl_wait();
  pdl_trigger();
  if (DROP && seed_dev) seed += *seed_dev;
  const int cg = C / 8;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  if (APPLY && blockIdx.x == 0 && r == 0) {   // parameter gradients from the (final) sums: dgamma += sum gd*xhat
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // atomic: two backward passes (source / target) may add into the same parameter gradient from two streams
      if (dbeta) atomicAdd(dbeta + g * 8 + i, (float)dsums[g * 8 + i]);
      if (dgamma) atomicAdd(dgamma + g * 8 + i, (float)dsums[C + g * 8 + i]);
    }
  }
  const float k = ACT == S2R_ACT_RELU6 ? (1.f / 6.f) : 1.f;
  float2 sc[4], sh[4];   // activation mask: pre (or pre/6)
  if (ACT != S2R_ACT_NONE) {
    load_pairs(scale_shift + g * 8, k, sc);
    load_pairs(scale_shift + C + g * 8, k, sh);
  }
  float2 A[4], B[4], D[4];   // APPLY: dx = A*gd + B*x + D
  float2 s[4], q[4];         // reduce
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = q[i] = make_float2(0.f, 0.f);
  if (APPLY) {
    float scl[8], mu[8], is[8];
    load_f8(scale_shift + g * 8, scl);
    load_f8(mean_invstd + g * 8, mu);
    load_f8(mean_invstd + C + g * 8, is);
    const double inv_n = count > 0 ? 1.0 / count : 0.0;
    float a[8], b[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m1 = count > 0 ? (float)(dsums[g * 8 + i] * inv_n) : 0.f;
      const float m2 = count > 0 ? (float)(dsums[C + g * 8 + i] * inv_n) : 0.f;
      a[i] = scl[i];
      b[i] = -scl[i] * m2 * is[i];
      d[i] = scl[i] * (m2 * is[i] * mu[i] - m1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      A[i] = make_float2(a[2 * i], a[2 * i + 1]);
      B[i] = make_float2(b[2 * i], b[2 * i + 1]);
      D[i] = make_float2(d[2 * i], d[2 * i + 1]);
    }
  }
  const long long step = (long long)gridDim.x * rows;
  long long p0 = (long long)blockIdx.x * rows + r;
  const __nv_bfloat16* dp = dy + p0 * dypitch + dyoff + g * 8;
  const __nv_bfloat16* xp = x + p0 * xpitch + xoff + g * 8;
  __nv_bfloat16* op = APPLY ? dx + p0 * dxpitch + dxoff + g * 8 : nullptr;
  const long long sd = step * dypitch, sx = step * xpitch, so = step * dxpitch;
  auto one = [&](const uint4& dvv, const uint4& xvv, __nv_bfloat16* dst, long long prow) {
    Lean8 gd = lean_unpack(dvv);
    const Lean8 xx = lean_unpack(xvv);
    Lean8 o;
    float m[8];
    if (DROP) dropout_scale8(seed, (unsigned long long)(prow * cg + g), drop_p, m);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (DROP) {
        gd.v[i].x *= m[2 * i];
        gd.v[i].y *= m[2 * i + 1];
      }
      if (ACT == S2R_ACT_RELU6) {
        float2 a6;
        a6.x = __saturatef(fmaf(xx.v[i].x, sc[i].x, sh[i].x));
        a6.y = __saturatef(fmaf(xx.v[i].y, sc[i].y, sh[i].y));
        gd.v[i] = act_mask2<ACT>(a6, gd.v[i]);
      } else if (ACT == S2R_ACT_RELU) {
        gd.v[i] = act_mask2<ACT>(s2r_dw::ffma2(xx.v[i], sc[i], sh[i]), gd.v[i]);
      }
      if (APPLY) {
        o.v[i] = s2r_dw::ffma2(A[i], gd.v[i], s2r_dw::ffma2(B[i], xx.v[i], D[i]));
      } else {
        s[i] = s2r_dw::fadd2(s[i], gd.v[i]);
        q[i] = s2r_dw::ffma2(gd.v[i], xx.v[i], q[i]);
      }
    }
    if (APPLY) *reinterpret_cast<uint4*>(dst) = lean_pack(o);
  };
  // full batches: straight-line code (no per-row branches), so all 2*UNR loads issue before the first use
  for (; p0 + (UNR - 1) * step < P; p0 += UNR * step, dp += UNR * sd, xp += UNR * sx, op += APPLY ? UNR * so : 0) {
    uint4 dv[UNR], xv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      dv[u] = ldg16(dp + u * sd);
      xv[u] = ldg16(xp + u * sx);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) one(dv[u], xv[u], APPLY ? op + u * so : nullptr, p0 + u * step);
  }
  for (; p0 < P; p0 += step, dp += sd, xp += sx, op += APPLY ? so : 0) one(ldg16(dp), ldg16(xp), op, p0);
  if (!APPLY) {
    // xhat form: sum gd*xhat = invstd * (sum gd*x - mean * sum gd), per thread before the block reduction
    float mu[8], is[8];
    load_f8(mean_invstd + g * 8, mu);
    load_f8(mean_invstd + C + g * 8, is);
    float* mine = sm + ((size_t)r * cg + g) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mine[2 * i] = s[i].x;
      mine[2 * i + 1] = s[i].y;
      mine[8 + 2 * i] = is[2 * i] * (q[i].x - mu[2 * i] * s[i].x);
      mine[8 + 2 * i + 1] = is[2 * i + 1] * (q[i].y - mu[2 * i + 1] * s[i].y);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
      const int gg = t / 16, kk = t % 16;
      double acc = 0;
      for (int rr = 0; rr < rows; ++rr) acc += (double)sm[((size_t)rr * cg + gg) * 16 + kk];
      atomicAdd(&dsums[(kk >> 3) * C + gg * 8 + (kk & 7)], acc);
    }
  }
}

// dgamma += sum dy' xhat, dbeta += sum dy'
__global__ void bn_param_grad_kernel(const double* __restrict__ dsums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) atomicAdd(dbeta + c, (float)dsums[c]);
  if (dgamma) atomicAdd(dgamma + c, (float)dsums[C + c]);
}

// launch geometry of the element-wise kernels: threads = rows x channel groups, enough CTAs for ~16 per SM
struct ElemCfg {
  int rows, threads, grid;
};
inline ElemCfg elem_cfg(long long P, int C) {
  ElemCfg c;
  const int cg = C / 8;
  c.rows = 256 / cg;
  if (c.rows < 1) c.rows = 1;
  c.threads = c.rows * cg;
  // >= 8 rows per thread (two full batches); 16 CTAs per SM on the large tensors, 8 on the small ones, where every
  // CTA's prologue (channel constants, a pending BatchNorm's finalize) weighs more (tests/tools/bn_bench.py)
  long long blocks = (P + (long long)c.rows * 8 - 1) / ((long long)c.rows * 8);
  const long long cap = (long long)s2r_sm_count() * (P * (long long)C >= 24LL * 1024 * 1024 ? 16 : 8);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  c.grid = (int)blocks;
  return c;
}

inline bool vec_ok(const void* p, int pitch, int off) {
  return ((uintptr_t)p % 16 == 0) && (pitch % 8 == 0) && (off % 8 == 0);
}

}  // namespace

extern "C" int s2r_channel_sums_bf16(const void* x, int64_t P, int C, int pitch, int coff,
                                     double* sums, s2r_stream_t stream) {
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && C <= 4096, S2R_ERR_SHAPE, "channel_sums: C=%d must be a multiple of 8 in [8,4096]", C);
  S2R_REQUIRE(vec_ok(x, pitch, coff) && pitch >= C + coff, S2R_ERR_SHAPE, "channel_sums: bad pitch/offset/alignment");
  if (P == 0) return S2R_OK;
  RowReduceCfg cfg = row_reduce_cfg(P, C);
  size_t smem = (size_t)cfg.rows * cfg.cg * 16 * sizeof(float);
  S2R_CUDA_OK(s2r_launch(channel_sums_kernel, dim3(cfg.grid), dim3(cfg.threads), (size_t)(smem), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, P, C, pitch, coff, cfg.rows, sums));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                               float eps, int clamp_mode, float momentum, float* running_mean,
                               float* running_var, float* mean_invstd, float* scale_shift, int C,
                               s2r_stream_t stream) {
  S2R_REQUIRE(C >= 1, S2R_ERR_SHAPE, "bn_finalize: C=%d", C);
  // batchnorm.py:116 -- the statistics need more than one value per channel
  S2R_REQUIRE(count > 1, S2R_ERR_SHAPE, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
  S2R_CUDA_OK(s2r_launch(bn_finalize_kernel, dim3(s2r_div_up(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums, count,
                         gamma, beta, eps, clamp_mode, momentum, running_mean, running_var, mean_invstd, scale_shift, C));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_eval_scale_shift(const float* gamma, const float* beta, const float* running_mean,
                                       const float* running_var, float eps, float* mean_invstd,
                                       float* scale_shift, int C, s2r_stream_t stream) {
  S2R_REQUIRE(C >= 1, S2R_ERR_SHAPE, "bn_eval: C=%d", C);
  S2R_CUDA_OK(s2r_launch(bn_eval_kernel, dim3(s2r_div_up(C, 128)), dim3(128), (size_t)0, (cudaStream_t)stream, 
      gamma, beta, running_mean, running_var, eps, mean_invstd, scale_shift, C));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_eval_multi(const s2r_bn_eval_job* jobs, int njobs, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535, S2R_ERR_SHAPE, "bn_eval_multi: %d jobs", njobs);
  if (njobs == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "bn_eval_multi: null table");
  S2R_CUDA_OK(s2r_launch(bn_eval_multi_kernel, dim3(2, njobs), dim3(256), (size_t)0, (cudaStream_t)stream, jobs));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_apply_act_bn(const void* x, int64_t P, int C, int xpitch, int xoff, const s2r_bn_tail* pending,
                                   const float* scale_shift, int act, const void* residual,
                                   float drop_p, uint64_t seed, const uint64_t* seed_dev, void* y, int ypitch,
                                   int yoff, s2r_stream_t stream) {
  if (pending) {
    S2R_REQUIRE(pending->count > 1, S2R_ERR_SHAPE, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
    S2R_REQUIRE(pending->sums && pending->scale_shift && pending->mean_invstd && (uintptr_t)pending->sums % 16 == 0,
                S2R_ERR_SHAPE, "bn_apply: pending BatchNorm without sums / outputs");
    scale_shift = pending->scale_shift;
  }
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && C <= 2048, S2R_ERR_SHAPE, "bn_apply: C=%d must be a multiple of 8 in [8,2048]", C);
  S2R_REQUIRE(vec_ok(x, xpitch, xoff) && vec_ok(y, ypitch, yoff) && vec_ok(residual, 8, 0) &&
                  ((uintptr_t)scale_shift % 16 == 0),
              S2R_ERR_SHAPE, "bn_apply: bad pitch/offset/alignment");
  S2R_REQUIRE(drop_p >= 0.f && drop_p < 1.f, S2R_ERR_SHAPE, "bn_apply: dropout p=%f", drop_p);
  const bool lean = getenv("S2R_BN_GENERIC") == nullptr && (drop_p == 0.f || (act == S2R_ACT_RELU && !residual)) &&
                    (act == S2R_ACT_NONE || act == S2R_ACT_RELU || act == S2R_ACT_RELU6);
  // pending BatchNorm: the lean kernels finalise it in their prologue; otherwise (cross-rank exchange, generic kernel,
  // empty tensor) one small launch does
  const bool fused = pending && lean && P > 0 && bn_tail_fusable(pending);
  if (pending && !fused) {
    const int rt = s2r_bn_tail_launch(pending, C, (cudaStream_t)stream);
    if (rt) return rt;
  }
  const BnTail bt = bn_tail_from(fused ? pending : nullptr);
  if (P == 0) return S2R_OK;
  const ElemCfg cfg = elem_cfg(P, C);
  if (lean) {
    const __nv_bfloat16 *xb = (const __nv_bfloat16*)x, *rb = (const __nv_bfloat16*)residual;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* sd = (const unsigned long long*)seed_dev;
#define S2R_APPLY(ACT_, RES_, DROP_)                                                                                     \
  S2R_CUDA_OK(s2r_launch(bn_apply_lean_kernel<ACT_, RES_, DROP_>, dim3(cfg.grid), dim3(cfg.threads),                     \
                         fused ? (size_t)2 * C * sizeof(float) : (size_t)0, st, xb,                                        \
                         (long long)P, C, xpitch, xoff, scale_shift, rb, yb, ypitch, yoff, cfg.rows, drop_p,               \
                         (unsigned long long)seed, sd, bt))
    bool done = true;
    if (drop_p > 0.f) S2R_APPLY(S2R_ACT_RELU, false, true);
    else if (act == S2R_ACT_NONE) { if (residual) S2R_APPLY(S2R_ACT_NONE, true, false); else S2R_APPLY(S2R_ACT_NONE, false, false); }
    else if (act == S2R_ACT_RELU) { if (residual) S2R_APPLY(S2R_ACT_RELU, true, false); else S2R_APPLY(S2R_ACT_RELU, false, false); }
    else if (act == S2R_ACT_RELU6) { if (residual) S2R_APPLY(S2R_ACT_RELU6, true, false); else S2R_APPLY(S2R_ACT_RELU6, false, false); }
    else done = false;
#undef S2R_APPLY
    if (done) {
      S2R_LAUNCH_OK();
      return S2R_OK;
    }
  }
  S2R_CUDA_OK(s2r_launch(bn_apply_kernel, dim3(cfg.grid), dim3(cfg.threads), (size_t)0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, P, C, xpitch, xoff, scale_shift, act,
      (const __nv_bfloat16*)residual, drop_p, seed, (const unsigned long long*)seed_dev, (__nv_bfloat16*)y, ypitch,
      yoff, cfg.rows));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_apply_act(const void* x, int64_t P, int C, int xpitch, int xoff,
                                const float* scale_shift, int act, const void* residual,
                                float drop_p, uint64_t seed, const uint64_t* seed_dev, void* y, int ypitch,
                                int yoff, s2r_stream_t stream) {
  return s2r_bn_apply_act_bn(x, P, C, xpitch, xoff, nullptr, scale_shift, act, residual, drop_p, seed, seed_dev, y, ypitch,
                             yoff, stream);
}

// y = relu6(y*scale + shift) in place (the generic depthwise kernel's inference epilogue, dwconv.cu)
int s2r_bn_apply_inplace_relu6(void* y, int64_t P, int C, const float* scale_shift, cudaStream_t st) {
  return s2r_bn_apply_act_bn(y, P, C, C, 0, nullptr, scale_shift, S2R_ACT_RELU6, nullptr, 0.f, 0, nullptr, y, C, 0, (s2r_stream_t)st);
}

extern "C" int s2r_bn_bwd_reduce(const void* dy, int dypitch, int dyoff, const void* x, int xpitch,
                                 int xoff, const float* mean_invstd, const float* scale_shift, int act,
                                 float drop_p, uint64_t seed, const uint64_t* seed_dev, int64_t P, int C,
                                 double* dsums, s2r_stream_t stream) {
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && C <= 4096, S2R_ERR_SHAPE, "bn_bwd_reduce: C=%d", C);
  S2R_REQUIRE(vec_ok(dy, dypitch, dyoff) && vec_ok(x, xpitch, xoff) &&
                  ((uintptr_t)scale_shift % 16 == 0) && ((uintptr_t)mean_invstd % 16 == 0),
              S2R_ERR_SHAPE, "bn_bwd_reduce: bad pitch/offset/alignment");
  if (P == 0) return S2R_OK;
  RowReduceCfg cfg = row_reduce_cfg(P, C);
  size_t smem = (size_t)cfg.rows * cfg.cg * 16 * sizeof(float);
  if (getenv("S2R_BN_GENERIC") == nullptr &&
      ((drop_p == 0.f && (act == S2R_ACT_NONE || act == S2R_ACT_RELU || act == S2R_ACT_RELU6)) || (drop_p > 0.f && act == S2R_ACT_RELU))) {
    const __nv_bfloat16 *db = (const __nv_bfloat16*)dy, *xb = (const __nv_bfloat16*)x;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* sd = (const unsigned long long*)seed_dev;
#define S2R_RED(ACT_, DROP_)                                                                                          \
  S2R_CUDA_OK(s2r_launch(bn_bwd_lean_kernel<ACT_, false, DROP_>, dim3(cfg.grid), dim3(cfg.threads), smem, st, db,      \
      dypitch, dyoff, xb, xpitch, xoff, mean_invstd, scale_shift, dsums, 0.0, (long long)P, C, (__nv_bfloat16*)nullptr,  \
      0, 0, cfg.rows, (float*)nullptr, (float*)nullptr, drop_p, (unsigned long long)seed, sd))
    if (drop_p > 0.f) S2R_RED(S2R_ACT_RELU, true);
    else if (act == S2R_ACT_NONE) S2R_RED(S2R_ACT_NONE, false);
    else if (act == S2R_ACT_RELU) S2R_RED(S2R_ACT_RELU, false);
    else S2R_RED(S2R_ACT_RELU6, false);
#undef S2R_RED
    S2R_LAUNCH_OK();
    return S2R_OK;
  }
  S2R_CUDA_OK(s2r_launch(bn_bwd_reduce_kernel, dim3(cfg.grid), dim3(cfg.threads), (size_t)(smem), (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, dypitch, dyoff, (const __nv_bfloat16*)x, xpitch, xoff, mean_invstd,
      scale_shift, act, drop_p, seed, (const unsigned long long*)seed_dev, P, C, cfg.rows, dsums));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_bwd_apply(const void* dy, int dypitch, int dyoff, const void* x, int xpitch,
                                int xoff, const float* mean_invstd, const float* scale_shift, int act,
                                float drop_p, uint64_t seed, const uint64_t* seed_dev, const double* dsums,
                                double count, int64_t P, int C, void* dx, int dxpitch, int dxoff, float* dgamma,
                                float* dbeta, int win_H, int win_W, int win_pad, s2r_stream_t stream) {
  S2R_REQUIRE(C >= 8 && C % 8 == 0 && C <= 2048, S2R_ERR_SHAPE, "bn_bwd_apply: C=%d must be a multiple of 8 in [8,2048]", C);
  S2R_REQUIRE(vec_ok(dy, dypitch, dyoff) && vec_ok(x, xpitch, xoff) && vec_ok(dx, dxpitch, dxoff) &&
                  ((uintptr_t)scale_shift % 16 == 0) && ((uintptr_t)mean_invstd % 16 == 0),
              S2R_ERR_SHAPE, "bn_bwd_apply: bad pitch/offset/alignment");
  const bool lean = dx && P > 0 && win_pad == 0 && getenv("S2R_BN_GENERIC") == nullptr &&
                    ((drop_p == 0.f && (act == S2R_ACT_NONE || act == S2R_ACT_RELU || act == S2R_ACT_RELU6)) ||
                     (drop_p > 0.f && act == S2R_ACT_RELU));
  if (lean) {
    const ElemCfg cfg = elem_cfg(P, C);
    const __nv_bfloat16 *db = (const __nv_bfloat16*)dy, *xb = (const __nv_bfloat16*)x;
    __nv_bfloat16* ob = (__nv_bfloat16*)dx;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long* sd = (const unsigned long long*)seed_dev;
    // the parameter gradients (dgamma, dbeta) are added by block 0 of the same launch
#define S2R_APP(ACT_, DROP_)                                                                                          \
  S2R_CUDA_OK(s2r_launch(bn_bwd_lean_kernel<ACT_, true, DROP_>, dim3(cfg.grid), dim3(cfg.threads), (size_t)0, st, db,  \
      dypitch, dyoff, xb, xpitch, xoff, mean_invstd, scale_shift, const_cast<double*>(dsums), count, (long long)P, C,   \
      ob, dxpitch, dxoff, cfg.rows, dgamma, dbeta, drop_p, (unsigned long long)seed, sd))
    if (drop_p > 0.f) S2R_APP(S2R_ACT_RELU, true);
    else if (act == S2R_ACT_NONE) S2R_APP(S2R_ACT_NONE, false);
    else if (act == S2R_ACT_RELU) S2R_APP(S2R_ACT_RELU, false);
    else S2R_APP(S2R_ACT_RELU6, false);
#undef S2R_APP
    S2R_LAUNCH_OK();
    return S2R_OK;
  }
  if (dgamma || dbeta) {
    S2R_CUDA_OK(s2r_launch(bn_param_grad_kernel, dim3(s2r_div_up(C, 128)), dim3(128), (size_t)0, (cudaStream_t)stream, dsums, dgamma, dbeta, C));
    S2R_LAUNCH_OK();
  }
  if (P == 0 || !dx) return S2R_OK;
  S2R_REQUIRE(win_pad == 0 || (win_H >= 1 && win_W >= 1 && P % ((int64_t)win_H * win_W) == 0), S2R_ERR_SHAPE,
              "bn_bwd_apply: window %dx%d does not tile P", win_H, win_W);
  const ElemCfg cfg = elem_cfg(P, C);
  S2R_CUDA_OK(s2r_launch(bn_bwd_apply_kernel, dim3(cfg.grid), dim3(cfg.threads), (size_t)0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, dypitch, dyoff, (const __nv_bfloat16*)x, xpitch, xoff, mean_invstd,
      scale_shift, act, drop_p, seed, (const unsigned long long*)seed_dev, dsums, count, P, C, (__nv_bfloat16*)dx,
      dxpitch, dxoff, win_H,
      win_W, win_pad, cfg.rows));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
