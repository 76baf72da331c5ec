// Multi-tensor optimizer steps.
// Replaces torch.optim.SGD.step / torch.optim.Adam.step at train_adapt.py:58-60,180-181 and
// train.py:63-82,202-204 of the reference: one launch walks a device table of
// (param, grad, state) slots instead of one launch chain per parameter tensor (DeepLab has 182
// parameter tensors).  Learning rate and Adam bias corrections are read from device memory so a
// captured CUDA graph of the whole step replays with the schedule's current values.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerSlot = 16;

// hyper[0] = lr
__global__ void __launch_bounds__(kThreads)
sgd_kernel(const s2r_param_slot* __restrict__ slots, const float* __restrict__ hyper, float momentum,
           float dampening, float wd, int nesterov, float gscale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const s2r_param_slot s = slots[blockIdx.y];
  if (s.g == nullptr) return;
  const float lr = hyper[0] * s.lr_mult;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < s.n;
       i += (long long)gridDim.x * kThreads) {
    const float p = s.p[i];
    float g = s.g[i] * gscale + wd * p;
    if (momentum != 0.f) {
      const float b = momentum * s.s0[i] + (1.f - dampening) * g;
      s.s0[i] = b;
      g = nesterov ? g + momentum * b : b;
    }
    s.p[i] = p - lr * g;
  }
}

// hyper[0] = lr, hyper[1] = 1 - beta1^t, hyper[2] = 1 - beta2^t
__global__ void __launch_bounds__(kThreads)
adam_kernel(const s2r_param_slot* __restrict__ slots, const float* __restrict__ hyper, float beta1,
            float beta2, float eps, float wd, float gscale) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const s2r_param_slot s = slots[blockIdx.y];
  if (s.g == nullptr) return;
  const float lr = hyper[0] * s.lr_mult;
  const float step_size = lr / hyper[1];
  const float inv_sqrt_bc2 = rsqrtf(hyper[2]);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < s.n;
       i += (long long)gridDim.x * kThreads) {
    const float p = s.p[i];
    const float g = s.g[i] * gscale + wd * p;
    const float m = beta1 * s.s0[i] + (1.f - beta1) * g;
    const float v = beta2 * s.s1[i] + (1.f - beta2) * g * g;
    s.s0[i] = m;
    s.s1[i] = v;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    s.p[i] = p - step_size * (m / denom);
  }
}

// Up to 8 floats passed BY VALUE as kernel arguments and stored to device memory: the values are snapshotted when
// the launch is enqueued, so the host may overwrite its copy immediately (a pinned-host -> device copy node would
// read the host buffer when the GPU executes it, i.e. possibly after the host has advanced by several steps).
struct F32x8 {
  float v[8];
};

__global__ void store_f32_kernel(float* __restrict__ dst, int n, F32x8 vals) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}

}  // namespace

extern "C" int s2r_store_f32(float* dst, int n, const float* host_vals, s2r_stream_t stream) {
  S2R_REQUIRE(dst && host_vals && n >= 1 && n <= 8, S2R_ERR_SHAPE, "store_f32: n=%d (1..8)", n);
  F32x8 v = {};
  for (int i = 0; i < n; ++i) v.v[i] = host_vals[i];
  store_f32_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst, n, v);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_sgd_step(const s2r_param_slot* slots, int nslots, const float* hyper, float momentum,
                            float dampening, float weight_decay, int nesterov, float gscale,
                            s2r_stream_t stream) {
  S2R_REQUIRE(nslots >= 0 && nslots <= 65535, S2R_ERR_SHAPE, "sgd_step: %d slots", nslots);
  if (nslots == 0) return S2R_OK;
  S2R_REQUIRE(slots && hyper, S2R_ERR_SHAPE, "sgd_step: null table");
  dim3 grid(kBlocksPerSlot, nslots);
  S2R_CUDA_OK(s2r_launch(sgd_kernel, dim3(grid), dim3(kThreads), (size_t)0, (cudaStream_t)stream, slots, hyper, momentum, dampening, weight_decay,
                                                         nesterov, gscale));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_adam_step(const s2r_param_slot* slots, int nslots, const float* hyper, float beta1,
                             float beta2, float eps, float weight_decay, float gscale,
                             s2r_stream_t stream) {
  S2R_REQUIRE(nslots >= 0 && nslots <= 65535, S2R_ERR_SHAPE, "adam_step: %d slots", nslots);
  if (nslots == 0) return S2R_OK;
  S2R_REQUIRE(slots && hyper, S2R_ERR_SHAPE, "adam_step: null table");
  dim3 grid(kBlocksPerSlot, nslots);
  S2R_CUDA_OK(s2r_launch(adam_kernel, dim3(grid), dim3(kThreads), (size_t)0, (cudaStream_t)stream, slots, hyper, beta1, beta2, eps, weight_decay,
                                                          gscale));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
