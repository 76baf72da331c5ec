// Shared helpers for the sm_100a kernels of the segmentation hot path.
// Everything here is device-side utility code or host-side error plumbing; the
// C-ABI entry points live in the individual .cu files and are declared in
// include/s2r_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/s2r_b200.h"

void s2r_set_error(const char* fmt, ...);

#define S2R_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      s2r_set_error(__VA_ARGS__);      \
      return (code);                   \
    }                                  \
  } while (0)

#define S2R_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      s2r_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                    __FILE__, __LINE__);                                         \
      return S2R_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define S2R_LAUNCH_OK()                                                          \
  do {                                                                           \
    cudaError_t _e = cudaPeekAtLastError();                                      \
    if (_e != cudaSuccess) {                                                     \
      s2r_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                    __FILE__, __LINE__);                                         \
      return S2R_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)


static inline int s2r_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int s2r_div_up(long a, long b) { return (int)((a + b - 1) / b); }

// grid size for a grid-stride kernel: enough CTAs for `work` items at `per_block`
// items per CTA pass, capped at `waves` resident waves over all SMs.
static inline int s2r_grid(long work, int per_block, int waves = 8) {
  long need = (work + per_block - 1) / per_block;
  long cap = (long)s2r_sm_count() * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Programmatic dependent launch: kernels launched through s2r_launch may start (and run their prologue: barrier /
// TMEM / shared-memory set-up, tensor-map prefetch) while the previous kernel in the stream is still draining.
// They call pdl_wait() before touching any global memory -- it returns once every earlier kernel has completed and
// its writes are visible -- and pdl_trigger() right after, which lets the NEXT kernel do the same.  Every kernel of the
// training step goes through s2r_launch (the peer-exchange kernel of comm.cu, which spins on remote flags, does not).
// Measured on the captured step (B=8, 512x1024): 19.84-19.98 ms with, 20.03 ms without when only the GEMM / depthwise
// kernels took part; on by default, S2R_PDL=0 disables it.
static inline bool s2r_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("S2R_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

#ifdef __CUDACC__

template <typename... KArgs, typename... Args>
static inline cudaError_t s2r_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = s2r_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == S2R_ACT_RELU) return fmaxf(v, 0.f);
  if (act == S2R_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  if (act == S2R_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}

// derivative mask of the activation evaluated at pre-activation value v
__device__ __forceinline__ float act_grad(float v, int act, float slope) {
  if (act == S2R_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (act == S2R_ACT_RELU6) return (v > 0.f && v < 6.f) ? 1.f : 0.f;
  if (act == S2R_ACT_LEAKY) return v > 0.f ? 1.f : slope;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Counter-based RNG for dropout masks: Philox-4x32 with 7 rounds keyed by
// (seed_lo, seed_hi), counter = (idx, stream). The mask for element i is a pure
// function of (seed, i) so backward regenerates it instead of storing it.
__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t x0 = c0, x1 = c1, x2 = 0x9E3779B9u, x3 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    uint32_t n0 = hi1 ^ x1 ^ k0, n1 = lo1, n2 = hi0 ^ x3 ^ k1, n3 = lo0;
    x0 = n0; x1 = n1; x2 = n2; x3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(x0, x1, x2, x3);
}

// keep-scale for 8 consecutive elements starting at element index `idx8*8`:
// returns per-element multiplier (0 or 1/(1-p)).
__device__ __forceinline__ void dropout_scale8(unsigned long long seed, unsigned long long idx8,
                                               float p, float* m) {
  uint4 r = philox4x32_7((uint32_t)idx8, (uint32_t)(idx8 >> 32), (uint32_t)seed,
                         (uint32_t)(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const float inv = 1.f / (1.f - p);
  const uint32_t thr = (uint32_t)(p * 65536.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = ((w[i] & 0xFFFFu) >= thr) ? inv : 0.f;
    m[2 * i + 1] = ((w[i] >> 16) >= thr) ? inv : 0.f;
  }
}

#endif  // __CUDACC__
