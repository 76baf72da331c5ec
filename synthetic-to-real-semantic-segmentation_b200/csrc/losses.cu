// Loss-side kernels of the adaptation step (all HBM-bound, fp32 arithmetic):
//   * softmax over the BATCH axis (train_adapt.py:151,166,174 uses F.softmax(x, dim=0))
//   * ignore-index / class-weighted cross entropy (utils/loss.py:21-30, 57-69)
//   * BCE-with-logits against a constant or tensor target (train_adapt.py:75,153,168,176)
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- softmax(dim=0)
// x viewed as [B][M]; every column j is normalised independently over b.
template <int V>
__global__ void __launch_bounds__(kThreads)
softmax0_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, long long M) {
  const long long nvec = M / V;
  for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < nvec;
       j += (long long)gridDim.x * kThreads) {
    float mx[V], s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { mx[v] = -INFINITY; s[v] = 0.f; }
    for (int b = 0; b < B; ++b) {
      float t[V];
      if (V == 4) *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(x + b * M) + j);
      else t[0] = __ldg(x + b * M + j);
#pragma unroll
      for (int v = 0; v < V; ++v) mx[v] = fmaxf(mx[v], t[v]);
    }
    for (int b = 0; b < B; ++b) {
      float t[V];
      if (V == 4) *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(x + b * M) + j);
      else t[0] = __ldg(x + b * M + j);
#pragma unroll
      for (int v = 0; v < V; ++v) s[v] += __expf(t[v] - mx[v]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 1.f / s[v];
    for (int b = 0; b < B; ++b) {
      float t[V];
      if (V == 4) *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(x + b * M) + j);
      else t[0] = __ldg(x + b * M + j);
#pragma unroll
      for (int v = 0; v < V; ++v) t[v] = __expf(t[v] - mx[v]) * s[v];
      if (V == 4) reinterpret_cast<float4*>(y + b * M)[j] = *reinterpret_cast<float4*>(t);
      else y[b * M + j] = t[0];
    }
  }
}

// dx_b = y_b * (dy_b - sum_b' dy_b' y_b')
template <int V>
__global__ void __launch_bounds__(kThreads)
softmax0_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                    float* __restrict__ dx, int B, long long M) {
  const long long nvec = M / V;
  for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < nvec;
       j += (long long)gridDim.x * kThreads) {
    float dot[V];
#pragma unroll
    for (int v = 0; v < V; ++v) dot[v] = 0.f;
    for (int b = 0; b < B; ++b) {
      float a[V], g[V];
      if (V == 4) {
        *reinterpret_cast<float4*>(a) = __ldg(reinterpret_cast<const float4*>(y + b * M) + j);
        *reinterpret_cast<float4*>(g) = __ldg(reinterpret_cast<const float4*>(dy + b * M) + j);
      } else {
        a[0] = __ldg(y + b * M + j);
        g[0] = __ldg(dy + b * M + j);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) dot[v] += a[v] * g[v];
    }
    for (int b = 0; b < B; ++b) {
      float a[V], g[V];
      if (V == 4) {
        *reinterpret_cast<float4*>(a) = __ldg(reinterpret_cast<const float4*>(y + b * M) + j);
        *reinterpret_cast<float4*>(g) = __ldg(reinterpret_cast<const float4*>(dy + b * M) + j);
      } else {
        a[0] = __ldg(y + b * M + j);
        g[0] = __ldg(dy + b * M + j);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) a[v] = a[v] * (g[v] - dot[v]);
      if (V == 4) reinterpret_cast<float4*>(dx + b * M)[j] = *reinterpret_cast<float4*>(a);
      else dx[b * M + j] = a[0];
    }
  }
}

// ---------------------------------------------------------------- cross entropy
// logits NCHW fp32, target float (cast with truncation like .long()) or a constant.
// sums[0] += sum_valid w_t (lse - x_t), sums[1] += sum_valid w_t, sums[2] += #(argmax == t)
// grad (optional) receives the UNSCALED gradient w_t (softmax - onehot); the
// 1/sum(w) factor needs the global reduction and is applied by s2r_scale_by_ratio.
template <int V>
__global__ void __launch_bounds__(kThreads)
ce_kernel(const float* __restrict__ logits, const float* __restrict__ target, int const_target,
          const float* __restrict__ weight, int C, long long HW, long long npix, int ignore_index,
          double* __restrict__ sums, float* __restrict__ grad) {
  float loss_acc = 0.f, w_acc = 0.f, hit_acc = 0.f;
  const long long nvec = npix / V;  // HW % V == 0 guaranteed by the launcher
  for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < nvec;
       q += (long long)gridDim.x * kThreads) {
    const long long i = q * V;
    const long long img = i / HW, px = i - img * HW;
    const float* base = logits + img * (long long)C * HW + px;
    int tcls[V];
    bool valid[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      long long t = target ? (long long)__ldg(target + i + v) : (long long)const_target;
      valid[v] = (t != ignore_index) && t >= 0 && t < C;
      tcls[v] = valid[v] ? (int)t : 0;
    }
    float mx[V], s[V], xt[V];
    int arg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { mx[v] = -INFINITY; s[v] = 0.f; xt[v] = 0.f; arg[v] = 0; }
    for (int c = 0; c < C; ++c) {
      float t[V];
      if (V == 4) *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(base + (long long)c * HW));
      else t[0] = __ldg(base + (long long)c * HW);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if (t[v] > mx[v]) {
          s[v] = s[v] * __expf(mx[v] - t[v]) + 1.f;
          mx[v] = t[v];
          arg[v] = c;
        } else {
          s[v] += __expf(t[v] - mx[v]);
        }
        if (c == tcls[v]) xt[v] = t[v];
      }
    }
    float lse[V], wt[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      lse[v] = mx[v] + __logf(s[v]);
      wt[v] = valid[v] ? (weight ? __ldg(weight + tcls[v]) : 1.f) : 0.f;
      loss_acc += wt[v] * (lse[v] - xt[v]);
      w_acc += wt[v];
      hit_acc += (valid[v] && arg[v] == tcls[v]) ? 1.f : 0.f;
    }
    if (grad) {
      float* gbase = grad + img * (long long)C * HW + px;
      for (int c = 0; c < C; ++c) {
        float t[V];
        if (V == 4) *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(base + (long long)c * HW));
        else t[0] = __ldg(base + (long long)c * HW);
#pragma unroll
        for (int v = 0; v < V; ++v)
          t[v] = wt[v] * (__expf(t[v] - lse[v]) - (c == tcls[v] ? 1.f : 0.f));
        if (V == 4) *reinterpret_cast<float4*>(gbase + (long long)c * HW) = *reinterpret_cast<float4*>(t);
        else gbase[(long long)c * HW] = t[0];
      }
    }
  }
  __shared__ double red[3][kThreads / 32];
  double a = warp_sum_d((double)loss_acc), b = warp_sum_d((double)w_acc),
         h = warp_sum_d((double)hit_acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = h; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

__global__ void ratio_kernel(const double* __restrict__ sums, float* __restrict__ out, double denom_override) {
  // out[0] = sums[0] / (denom_override > 0 ? denom_override : sums[1])
  double d = denom_override > 0 ? denom_override : sums[1];
  out[0] = (float)(sums[0] / d);
}

// g[i] *= gout[0] / (denom_override > 0 ? denom_override : sums[1])
__global__ void __launch_bounds__(kThreads)
scale_by_ratio_kernel(float* __restrict__ g, long long n, const float* __restrict__ gout,
                      const double* __restrict__ sums, double denom_override) {
  const double d = denom_override > 0 ? denom_override : sums[1];
  const float sc = (float)((double)gout[0] / d);
  const long long n4 = n / 4;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4;
       i += (long long)gridDim.x * kThreads) {
    float4 t = g4[i];
    t.x *= sc; t.y *= sc; t.z *= sc; t.w *= sc;
    g4[i] = t;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    g[i] *= sc;
}

// ---------------------------------------------------------------- BCE with logits
// loss_i = max(x,0) - x t + log(1 + exp(-|x|)); sums[0] += sum_i loss_i
__global__ void __launch_bounds__(kThreads)
bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ target, float const_target,
               long long n, double* __restrict__ sums) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    float v = x[i], t = target ? target[i] : const_target;
    acc += fmaxf(v, 0.f) - v * t + log1pf(__expf(-fabsf(v)));
  }
  __shared__ double red[kThreads / 32];
  double a = warp_sum_d((double)acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    atomicAdd(&sums[0], t);
  }
}

// dx_i = (sigmoid(x_i) - t_i) * gout / n
__global__ void __launch_bounds__(kThreads)
bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ target, float const_target,
               long long n, const float* __restrict__ gout, float* __restrict__ dx) {
  const float sc = gout[0] / (float)n;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    float v = x[i], t = target ? target[i] : const_target;
    float sg = 1.f / (1.f + __expf(-v));
    dx[i] = (sg - t) * sc;
  }
}

}  // namespace

extern "C" int s2r_softmax_dim0_fwd(const float* x, float* y, int B, int64_t M, s2r_stream_t stream) {
  S2R_REQUIRE(B >= 1 && M >= 0, S2R_ERR_SHAPE, "softmax_dim0: bad shape B=%d M=%lld", B, (long long)M);
  if (M == 0) return S2R_OK;
  const bool vec = (M % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
  if (vec)
    softmax0_fwd_kernel<4><<<s2r_grid(M / 4, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(x, y, B, M);
  else
    softmax0_fwd_kernel<1><<<s2r_grid(M, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(x, y, B, M);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_softmax_dim0_bwd(const float* y, const float* dy, float* dx, int B, int64_t M,
                                    s2r_stream_t stream) {
  S2R_REQUIRE(B >= 1 && M >= 0, S2R_ERR_SHAPE, "softmax_dim0_bwd: bad shape");
  if (M == 0) return S2R_OK;
  const bool vec = (M % 4 == 0) && (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)dx) % 16 == 0);
  if (vec)
    softmax0_bwd_kernel<4><<<s2r_grid(M / 4, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(y, dy, dx, B, M);
  else
    softmax0_bwd_kernel<1><<<s2r_grid(M, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(y, dy, dx, B, M);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_cross_entropy_nchw(const float* logits, const float* target, int const_target,
                                      const float* weight, int N, int C, int64_t HW,
                                      int ignore_index, double* sums, float* grad_unscaled,
                                      s2r_stream_t stream) {
  S2R_REQUIRE(N >= 0 && C >= 1 && HW >= 0, S2R_ERR_SHAPE, "cross_entropy: bad shape");
  S2R_REQUIRE(logits && sums, S2R_ERR_SHAPE, "cross_entropy: null pointer");
  const long long npix = (long long)N * HW;
  if (npix == 0) return S2R_OK;
  const bool vec = (HW % 4 == 0) &&
                   (((uintptr_t)logits | (uintptr_t)grad_unscaled) % 16 == 0);
  if (vec)
    ce_kernel<4><<<s2r_grid(npix / 4, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        logits, target, const_target, weight, C, HW, npix, ignore_index, sums, grad_unscaled);
  else
    ce_kernel<1><<<s2r_grid(npix, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        logits, target, const_target, weight, C, HW, npix, ignore_index, sums, grad_unscaled);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_ratio(const double* sums, double denom_override, float* out, s2r_stream_t stream) {
  ratio_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, out, denom_override);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_scale_by_ratio(float* g, int64_t n, const float* gout, const double* sums,
                                  double denom_override, s2r_stream_t stream) {
  if (n == 0) return S2R_OK;
  S2R_REQUIRE(((uintptr_t)g) % 16 == 0, S2R_ERR_SHAPE, "scale_by_ratio: unaligned buffer");
  scale_by_ratio_kernel<<<s2r_grid(n / 4 + 1, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(
      g, n, gout, sums, denom_override);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bce_logits_fwd(const float* x, const float* target, float const_target, int64_t n,
                                  double* sums, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 1, S2R_ERR_SHAPE, "bce_logits: empty input");
  bce_fwd_kernel<<<s2r_grid(n, kThreads, 2), kThreads, 0, (cudaStream_t)stream>>>(x, target, const_target, n, sums);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bce_logits_bwd(const float* x, const float* target, float const_target, int64_t n,
                                  const float* gout, float* dx, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 1, S2R_ERR_SHAPE, "bce_logits_bwd: empty input");
  bce_bwd_kernel<<<s2r_grid(n, kThreads, 2), kThreads, 0, (cudaStream_t)stream>>>(x, target, const_target, n, gout, dx);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
