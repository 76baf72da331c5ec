// Error plumbing, version and small utility entry points of the C ABI (include/s2r_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void s2r_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* s2r_last_error(void) { return g_err; }

extern "C" int s2r_version(void) { return 100; }

extern "C" int s2r_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

namespace {

__global__ void __launch_bounds__(256)
add_bf16_kernel(__nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long nvec) {
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < nvec;
       t += (long long)gridDim.x * 256) {
    float x[8], y[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(a + t * 8), x);
    bf16x8_to_float(ldg16(b + t * 8), y);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += y[i];
    *reinterpret_cast<uint4*>(a + t * 8) = float_to_bf16x8(x);
  }
}

__global__ void add_f64_to_f32_kernel(const double* __restrict__ s, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(o + i, (float)s[i]);   // atomic: gradient buffers may be accumulated from two streams
}

}  // namespace

extern "C" int s2r_add_bf16(void* a, const void* b, int64_t n, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 0 && n % 8 == 0 && ((uintptr_t)a | (uintptr_t)b) % 16 == 0, S2R_ERR_SHAPE,
              "add_bf16: n must be a multiple of 8 and buffers 16B aligned");
  if (n == 0) return S2R_OK;
  add_bf16_kernel<<<s2r_grid(n / 8, 256 * 2, 16), 256, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)a, (const __nv_bfloat16*)b, n / 8);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_add_f64_to_f32(const double* sums, float* out, int n, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 0, S2R_ERR_SHAPE, "add_f64_to_f32: n < 0");
  if (n == 0) return S2R_OK;
  add_f64_to_f32_kernel<<<s2r_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(sums, out, n);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
