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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
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
      // a class id outside [0, C) that is not ignore_index: nn.CrossEntropyLoss raises a device assert for it
      // (utils/loss.py:27-28 of the reference would stop); here the LOSS becomes NaN -- a mislabelled data set (raw
      // GTA ids 19..254) must not train quietly on the remaining pixels
      if (!valid[v] && t != ignore_index) loss_acc = __int_as_float(0x7fc00000);
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

// The same for C <= CMAX classes (19 on this path) with the 4-pixel logit vectors of ALL classes held in registers:
// the C loads of a pixel group are issued back to back (the generic kernel's runtime class loop serialises load ->
// exp -> load), the maximum is taken first (no rescaling branch in the exp chain), and the gradient is computed from the
// registers instead of a second read of the logits.
template <int CMAX>
__global__ void __launch_bounds__(kThreads)
ce_regs_kernel(const float* __restrict__ logits, const float* __restrict__ target, int const_target,
               const float* __restrict__ weight, int C, long long HW, long long npix, int ignore_index,
               double* __restrict__ sums, float* __restrict__ grad) {
  pdl_wait();
  pdl_trigger();
  float loss_acc = 0.f, w_acc = 0.f, hit_acc = 0.f;
  const long long nvec = npix / 4;
  for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < nvec; q += (long long)gridDim.x * kThreads) {
    const long long i = q * 4;
    const long long img = i / HW, px = i - img * HW;
    const float* base = logits + img * (long long)C * HW + px;
    float4 x[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      x[c] = c < C ? __ldg(reinterpret_cast<const float4*>(base + (long long)c * HW)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    float tv[4];
    if (target) *reinterpret_cast<float4*>(tv) = __ldg(reinterpret_cast<const float4*>(target + i));
    int tcls[4];
    bool valid[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long t = target ? (long long)tv[v] : (long long)const_target;
      valid[v] = (t != ignore_index) && t >= 0 && t < C;
      tcls[v] = valid[v] ? (int)t : 0;
      if (!valid[v] && t != ignore_index) loss_acc = __int_as_float(0x7fc00000);   // see ce_kernel
    }
    float mx[4], s[4] = {0.f, 0.f, 0.f, 0.f}, xt[4] = {0.f, 0.f, 0.f, 0.f};
    int arg[4] = {0, 0, 0, 0};
    mx[0] = x[0].x; mx[1] = x[0].y; mx[2] = x[0].z; mx[3] = x[0].w;
#pragma unroll
    for (int c = 1; c < CMAX; ++c) {
      const float t[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (t[v] > mx[v]) { mx[v] = t[v]; arg[v] = c; }     // first maximum, like the generic kernel
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      float t[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (c == tcls[v]) xt[v] = t[v];
        t[v] = __expf(t[v] - mx[v]);     // exp(-inf) = 0 for the classes past C
        s[v] += t[v];
      }
      x[c] = make_float4(t[0], t[1], t[2], t[3]);
    }
    float wt[4], inv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float lse = mx[v] + __logf(s[v]);
      wt[v] = valid[v] ? (weight ? __ldg(weight + tcls[v]) : 1.f) : 0.f;
      loss_acc += wt[v] * (lse - xt[v]);
      w_acc += wt[v];
      hit_acc += (valid[v] && arg[v] == tcls[v]) ? 1.f : 0.f;
      inv[v] = wt[v] / s[v];
    }
    if (grad) {
      float* gbase = grad + img * (long long)C * HW + px;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          float4 g;
          g.x = x[c].x * inv[0] - (c == tcls[0] ? wt[0] : 0.f);
          g.y = x[c].y * inv[1] - (c == tcls[1] ? wt[1] : 0.f);
          g.z = x[c].z * inv[2] - (c == tcls[2] ? wt[2] : 0.f);
          g.w = x[c].w * inv[3] - (c == tcls[3] ? wt[3] : 0.f);
          *reinterpret_cast<float4*>(gbase + (long long)c * HW) = g;
        }
    }
  }
  __shared__ double red[3][kThreads / 32];
  double a = warp_sum_d((double)loss_acc), b = warp_sum_d((double)w_acc), h = warp_sum_d((double)hit_acc);
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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  // out[0] = sums[0] / (denom_override > 0 ? denom_override : sums[1])
  double d = denom_override > 0 ? denom_override : sums[1];
  out[0] = (float)(sums[0] / d);
}

// g[i] *= gout[0] / (denom_override > 0 ? denom_override : sums[1])
__global__ void __launch_bounds__(kThreads)
scale_by_ratio_kernel(float* __restrict__ g, long long n, const float* __restrict__ gout,
                      const double* __restrict__ sums, double denom_override) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
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
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const float sc = gout[0] / (float)n;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads) {
    float v = x[i], t = target ? target[i] : const_target;
    float sg = 1.f / (1.f + __expf(-v));
    dx[i] = (sg - t) * sc;
  }
}


// ---------------------------------------------------------------- softmax(dim=0) -> zero-padded NHWC bf16
// The discriminator consumes F.softmax(logits, dim=0) (train_adapt.py:151,166,174) through a 4x4 stride-2 pad-1
// convolution.  Instead of an fp32 NCHW softmax tensor followed by a patch matrix (16x the input in bf16), the batch
// softmax is written ONCE as bf16 into a zero-padded NHWC buffer [B][H+2][W+2][Cp]: in that buffer the four kw taps of
// a filter row are one contiguous run of 4*Cp channels, so conv1 becomes a 4-tap GEMM over an overlapping strided view
// (engine.rowtap_*).  A CTA handles a strip of 128 pixels of one image row for up to 8 batch entries: coalesced fp32
// reads along w, transpose through shared memory, coalesced 16-byte stores of whole pixels.
constexpr int SP_TW = 128;
constexpr int SP_THREADS = 256;
constexpr int SP_MAXB = 8;

// gmax / gsum (both or neither; [C][H][W] fp32): maximum and sum of exp(x - gmax) over the GLOBAL batch, i.e. over the
// images of all ranks (s2r_softmax0_batch_stats + two all-reduces): F.softmax(x, dim=0) of the gathered batch, as the
// reference's single-process DataParallel evaluates it (train_adapt.py:87-88,151), on a rank that holds B of the images.
template <bool SOFTMAX>
__global__ void __launch_bounds__(SP_THREADS)
softmax0_to_nhwc_pad_kernel(const float* __restrict__ x, int B, int C, int H, int W, __nv_bfloat16* __restrict__ y, int Cp,
                            const float* __restrict__ gmax, const float* __restrict__ gsum) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t sp_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(sp_smem);   // [bc][SP_TW][Cp]
  const int w0 = blockIdx.x * SP_TW, h = blockIdx.y, b0 = blockIdx.z * SP_MAXB;
  const int bc = min(SP_MAXB, B - b0), npx = min(SP_TW, W - w0);
  const long long plane = (long long)H * W;
  const int Wp = W + 2, Hp = H + 2;
  // every channel pair of the padded pixel is written by the loop below (pairs past C as zeros), so the tile needs no
  // clearing pass
  const int cpairs = (C + 1) >> 1, ppairs = Cp >> 1;
  for (int it = threadIdx.x; it < ppairs * SP_TW; it += SP_THREADS) {
    const int w = it & (SP_TW - 1), cp = it / SP_TW;
    if (w >= npx) continue;
    const int c0 = 2 * cp;
    if (cp >= cpairs) {                     // pad channels
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b)
        if (b < bc) *reinterpret_cast<uint32_t*>(tile + ((long long)b * SP_TW + w) * Cp + c0) = 0u;
      continue;
    }
    const bool has1 = c0 + 1 < C;
    float t0[SP_MAXB], t1[SP_MAXB];
    const float* src = x + ((long long)b0 * C + c0) * plane + (long long)h * W + w0 + w;
#pragma unroll
    for (int b = 0; b < SP_MAXB; ++b) {
      t0[b] = b < bc ? __ldg(src + (long long)b * C * plane) : -INFINITY;
      t1[b] = (b < bc && has1) ? __ldg(src + (long long)b * C * plane + plane) : -INFINITY;
    }
    if (SOFTMAX) {
      float m0 = -INFINITY, m1 = -INFINITY, s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) { m0 = fmaxf(m0, t0[b]); m1 = fmaxf(m1, t1[b]); }
      const long long gi = (long long)c0 * plane + (long long)h * W + w0 + w;
      if (gmax) {
        m0 = __ldg(gmax + gi);
        if (has1) m1 = __ldg(gmax + gi + plane);
      }
      if (!has1) m1 = 0.f;
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) {
        t0[b] = b < bc ? __expf(t0[b] - m0) : 0.f;
        t1[b] = (b < bc && has1) ? __expf(t1[b] - m1) : 0.f;
        s0 += t0[b]; s1 += t1[b];
      }
      if (gsum) {
        s0 = __ldg(gsum + gi);
        if (has1) s1 = __ldg(gsum + gi + plane);
      }
      s0 = 1.f / s0; s1 = has1 ? 1.f / s1 : 0.f;
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) { t0[b] *= s0; t1[b] *= s1; }
    }
#pragma unroll
    for (int b = 0; b < SP_MAXB; ++b)
      if (b < bc) {
        const __nv_bfloat162 v = __floats2bfloat162_rn(t0[b], has1 ? t1[b] : 0.f);
        *reinterpret_cast<__nv_bfloat162*>(tile + ((long long)b * SP_TW + w) * Cp + c0) = v;
      }
  }
  __syncthreads();
  const int vpp = Cp / 8;   // 16-byte vectors per pixel
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  for (int b = 0; b < bc; ++b) {
    __nv_bfloat16* img = y + (long long)(b0 + b) * Hp * Wp * Cp;
    uint4* dst = reinterpret_cast<uint4*>(img + ((long long)(h + 1) * Wp + w0 + 1) * Cp);
    const uint4* srcv = reinterpret_cast<const uint4*>(tile + (long long)b * SP_TW * Cp);
    for (int v = threadIdx.x; v < npx * vpp; v += SP_THREADS) dst[v] = srcv[v];
    // zero border: left / right pixel of this row, and the rows above / below the image for this strip (+ corners)
    if (w0 == 0 && threadIdx.x < vpp) reinterpret_cast<uint4*>(img + (long long)(h + 1) * Wp * Cp)[threadIdx.x] = z4;
    if (w0 + npx == W && threadIdx.x < vpp)
      reinterpret_cast<uint4*>(img + ((long long)(h + 1) * Wp + W + 1) * Cp)[threadIdx.x] = z4;
    if (h == 0 || h == H - 1) {
      const int lo = w0 == 0 ? 0 : w0 + 1, hi = w0 + npx == W ? W + 2 : w0 + npx + 1;   // padded columns [lo, hi)
      for (int r = 0; r < 2; ++r) {
        if ((r == 0 && h != 0) || (r == 1 && h != H - 1)) continue;
        uint4* row = reinterpret_cast<uint4*>(img + ((long long)(r == 0 ? 0 : H + 1) * Wp + lo) * Cp);
        for (int v = threadIdx.x; v < (hi - lo) * vpp; v += SP_THREADS) row[v] = z4;
      }
    }
  }
}

// dx[b][c][h][w] = y_b * (g_b - sum_b' g_b' y_b') with y = softmax over b of x (recomputed) and g read from the
// interior of the padded NHWC bf16 gradient; SOFTMAX = false: plain layout conversion dx = g.
template <bool SOFTMAX>
__global__ void __launch_bounds__(SP_THREADS)
softmax0_nhwc_pad_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ gp, int B, int C, int H,
                             int W, int Cp, float* __restrict__ dx, const float* __restrict__ gmax,
                             const float* __restrict__ gsum, float* __restrict__ tpart, const float* __restrict__ tglob) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t sp_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(sp_smem);
  const int w0 = blockIdx.x * SP_TW, h = blockIdx.y, b0 = blockIdx.z * SP_MAXB;
  const int bc = min(SP_MAXB, B - b0), npx = min(SP_TW, W - w0);
  const long long plane = (long long)H * W;
  const int Wp = W + 2, Hp = H + 2, vpp = Cp / 8;
  for (int b = 0; b < bc; ++b) {
    const uint4* src = reinterpret_cast<const uint4*>(gp + (((long long)(b0 + b) * Hp + h + 1) * Wp + w0 + 1) * Cp);
    uint4* dstv = reinterpret_cast<uint4*>(tile + (long long)b * SP_TW * Cp);
    for (int v = threadIdx.x; v < npx * vpp; v += SP_THREADS) dstv[v] = __ldg(src + v);
  }
  __syncthreads();
  for (int it = threadIdx.x; it < C * SP_TW; it += SP_THREADS) {
    const int w = it & (SP_TW - 1), c = it / SP_TW;
    if (w >= npx) continue;
    const long long off = ((long long)b0 * C + c) * plane + (long long)h * W + w0 + w;
    float g[SP_MAXB], t[SP_MAXB];
#pragma unroll
    for (int b = 0; b < SP_MAXB; ++b) {
      g[b] = b < bc ? __bfloat162float(tile[((long long)b * SP_TW + w) * Cp + c]) : 0.f;
      t[b] = (SOFTMAX && b < bc) ? __ldg(x + off + (long long)b * C * plane) : -INFINITY;
    }
    if (SOFTMAX) {
      float m = -INFINITY, s = 0.f, dot = 0.f;
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) m = fmaxf(m, t[b]);
      // global batch (see the forward kernel): statistics of all ranks' images; the sum over b' of g y then also runs
      // over all ranks -- this launch leaves its part in tpart, a second one takes the all-reduced sum from tglob
      const long long gi = (long long)c * plane + (long long)h * W + w0 + w;
      if (gmax) m = __ldg(gmax + gi);
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) { t[b] = b < bc ? __expf(t[b] - m) : 0.f; s += t[b]; }
      if (gsum) s = __ldg(gsum + gi);
      s = 1.f / s;
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) { t[b] *= s; dot = fmaf(t[b], g[b], dot); }
      if (tpart) {
        tpart[gi] = dot;
        continue;
      }
      if (tglob) dot = __ldg(tglob + gi);
#pragma unroll
      for (int b = 0; b < SP_MAXB; ++b) g[b] = t[b] * (g[b] - dot);
    }
#pragma unroll
    for (int b = 0; b < SP_MAXB; ++b)
      if (b < bc) dx[off + (long long)b * C * plane] = g[b];
  }
}

// out[i] = max_b x[b][i] (gmax == NULL) or sum_b exp(x[b][i] - gmax[i]), i over the M = C*H*W positions of one image
__global__ void __launch_bounds__(kThreads)
softmax0_batch_stats_kernel(const float* __restrict__ x, int B, long long M, const float* __restrict__ gmax,
                            float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < M; i += (long long)gridDim.x * kThreads) {
    float a = gmax ? 0.f : -INFINITY;
    const float m = gmax ? __ldg(gmax + i) : 0.f;
    for (int b = 0; b < B; ++b) {
      const float v = __ldg(x + (long long)b * M + i);
      a = gmax ? a + __expf(v - m) : fmaxf(a, v);
    }
    out[i] = a;
  }
}

}  // namespace

extern "C" int s2r_softmax0_batch_stats(const float* x, int B, int64_t M, const float* gmax, float* out,
                                        s2r_stream_t stream) {
  S2R_REQUIRE(x && out && B >= 1 && M >= 0, S2R_ERR_SHAPE, "softmax0_batch_stats: bad arguments");
  if (M == 0) return S2R_OK;
  S2R_CUDA_OK(s2r_launch(softmax0_batch_stats_kernel, dim3(s2r_grid(M, kThreads, 16)), dim3(kThreads), (size_t)0,
                         (cudaStream_t)stream, x, B, (long long)M, gmax, out));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_softmax_dim0_fwd(const float* x, float* y, int B, int64_t M, s2r_stream_t stream) {
  S2R_REQUIRE(B >= 1 && M >= 0, S2R_ERR_SHAPE, "softmax_dim0: bad shape B=%d M=%lld", B, (long long)M);
  if (M == 0) return S2R_OK;
  const bool vec = (M % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
  if (vec)
    S2R_CUDA_OK(s2r_launch(softmax0_fwd_kernel<4>, dim3(s2r_grid(M / 4, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, x, y, B, M));
  else
    S2R_CUDA_OK(s2r_launch(softmax0_fwd_kernel<1>, dim3(s2r_grid(M, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, x, y, B, M));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_softmax_dim0_bwd(const float* y, const float* dy, float* dx, int B, int64_t M,
                                    s2r_stream_t stream) {
  S2R_REQUIRE(B >= 1 && M >= 0, S2R_ERR_SHAPE, "softmax_dim0_bwd: bad shape");
  if (M == 0) return S2R_OK;
  const bool vec = (M % 4 == 0) && (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)dx) % 16 == 0);
  if (vec)
    S2R_CUDA_OK(s2r_launch(softmax0_bwd_kernel<4>, dim3(s2r_grid(M / 4, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, y, dy, dx, B, M));
  else
    S2R_CUDA_OK(s2r_launch(softmax0_bwd_kernel<1>, dim3(s2r_grid(M, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, y, dy, dx, B, M));
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
  const bool tvec = !target || (uintptr_t)target % 16 == 0;
  if (vec && tvec && C >= 2 && C <= 19 && getenv("S2R_CE_GENERIC") == nullptr)
    S2R_CUDA_OK(s2r_launch(ce_regs_kernel<19>, dim3(s2r_grid(npix / 4, kThreads, 8)), dim3(kThreads), (size_t)0, (cudaStream_t)stream,
        logits, target, const_target, weight, C, HW, npix, ignore_index, sums, grad_unscaled));
  else if (vec)
    S2R_CUDA_OK(s2r_launch(ce_kernel<4>, dim3(s2r_grid(npix / 4, kThreads, 8)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        logits, target, const_target, weight, C, HW, npix, ignore_index, sums, grad_unscaled));
  else
    S2R_CUDA_OK(s2r_launch(ce_kernel<1>, dim3(s2r_grid(npix, kThreads, 8)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
        logits, target, const_target, weight, C, HW, npix, ignore_index, sums, grad_unscaled));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_ratio(const double* sums, double denom_override, float* out, s2r_stream_t stream) {
  S2R_CUDA_OK(s2r_launch(ratio_kernel, dim3(1), dim3(1), (size_t)0, (cudaStream_t)stream, sums, out, denom_override));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_scale_by_ratio(float* g, int64_t n, const float* gout, const double* sums,
                                  double denom_override, s2r_stream_t stream) {
  if (n == 0) return S2R_OK;
  S2R_REQUIRE(((uintptr_t)g) % 16 == 0, S2R_ERR_SHAPE, "scale_by_ratio: unaligned buffer");
  S2R_CUDA_OK(s2r_launch(scale_by_ratio_kernel, dim3(s2r_grid(n / 4 + 1, kThreads, 16)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, 
      g, n, gout, sums, denom_override));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bce_logits_fwd(const float* x, const float* target, float const_target, int64_t n,
                                  double* sums, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 1, S2R_ERR_SHAPE, "bce_logits: empty input");
  S2R_CUDA_OK(s2r_launch(bce_fwd_kernel, dim3(s2r_grid(n, kThreads, 2)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, x, target, const_target, n, sums));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bce_logits_bwd(const float* x, const float* target, float const_target, int64_t n,
                                  const float* gout, float* dx, s2r_stream_t stream) {
  S2R_REQUIRE(n >= 1, S2R_ERR_SHAPE, "bce_logits_bwd: empty input");
  S2R_CUDA_OK(s2r_launch(bce_bwd_kernel, dim3(s2r_grid(n, kThreads, 2)), dim3(kThreads), (size_t)0, (cudaStream_t)stream, x, target, const_target, n, gout, dx));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

template <bool SOFTMAX>
static int launch_sp_fwd(const float* x, int B, int C, int H, int W, void* yp, int Cp, cudaStream_t st,
                         const float* gmax = nullptr, const float* gsum = nullptr) {
  const int smem = min(B, SP_MAXB) * SP_TW * Cp * 2;
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(softmax0_to_nhwc_pad_kernel<SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_MAXB * SP_TW * 64 * 2));
    attr = true;
  }
  dim3 grid(s2r_div_up(W, SP_TW), H, s2r_div_up(B, SP_MAXB));
  S2R_CUDA_OK(s2r_launch(softmax0_to_nhwc_pad_kernel<SOFTMAX>, dim3(grid), dim3(SP_THREADS), (size_t)(smem), st, x, B, C, H, W, (__nv_bfloat16*)yp, Cp, gmax, gsum));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

template <bool SOFTMAX>
static int launch_sp_bwd(const float* x, const void* gp, int B, int C, int H, int W, int Cp, float* dx, cudaStream_t st,
                         const float* gmax = nullptr, const float* gsum = nullptr, float* tpart = nullptr,
                         const float* tglob = nullptr) {
  const int smem = min(B, SP_MAXB) * SP_TW * Cp * 2;
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(softmax0_nhwc_pad_bwd_kernel<SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_MAXB * SP_TW * 64 * 2));
    attr = true;
  }
  dim3 grid(s2r_div_up(W, SP_TW), H, s2r_div_up(B, SP_MAXB));
  S2R_CUDA_OK(s2r_launch(softmax0_nhwc_pad_bwd_kernel<SOFTMAX>, dim3(grid), dim3(SP_THREADS), (size_t)(smem), st, x, (const __nv_bfloat16*)gp, B, C, H, W, Cp, dx, gmax, gsum, tpart, tglob));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_softmax0_nchw_to_nhwc_pad(const float* x, int B, int C, int H, int W, int softmax, void* yp, int Cp,
                                             s2r_stream_t stream) {
  S2R_REQUIRE(x && yp && B >= 1 && C >= 1 && H >= 1 && W >= 1 && H <= 65535, S2R_ERR_SHAPE, "softmax0_to_nhwc_pad: bad shape");
  S2R_REQUIRE(Cp % 8 == 0 && Cp >= C && Cp <= 64 && (uintptr_t)yp % 16 == 0, S2R_ERR_SHAPE,
              "softmax0_to_nhwc_pad: channel pitch %d (need a multiple of 8 in [C, 64]) / alignment", Cp);
  S2R_REQUIRE(!softmax || B <= SP_MAXB, S2R_ERR_UNSUPPORTED, "softmax0_to_nhwc_pad: batch %d > %d", B, SP_MAXB);
  return softmax ? launch_sp_fwd<true>(x, B, C, H, W, yp, Cp, (cudaStream_t)stream)
                 : launch_sp_fwd<false>(x, B, C, H, W, yp, Cp, (cudaStream_t)stream);
}

extern "C" int s2r_softmax0_nhwc_pad_bwd(const float* x, const void* gp, int B, int C, int H, int W, int Cp, int softmax,
                                         float* dx, s2r_stream_t stream) {
  S2R_REQUIRE(gp && dx && (x || !softmax) && B >= 1 && C >= 1 && H >= 1 && W >= 1 && H <= 65535, S2R_ERR_SHAPE,
              "softmax0_nhwc_pad_bwd: bad shape");
  S2R_REQUIRE(Cp % 8 == 0 && Cp >= C && Cp <= 64 && (uintptr_t)gp % 16 == 0, S2R_ERR_SHAPE,
              "softmax0_nhwc_pad_bwd: channel pitch %d (need a multiple of 8 in [C, 64]) / alignment", Cp);
  S2R_REQUIRE(!softmax || B <= SP_MAXB, S2R_ERR_UNSUPPORTED, "softmax0_nhwc_pad_bwd: batch %d > %d", B, SP_MAXB);
  return softmax ? launch_sp_bwd<true>(x, gp, B, C, H, W, Cp, dx, (cudaStream_t)stream)
                 : launch_sp_bwd<false>(x, gp, B, C, H, W, Cp, dx, (cudaStream_t)stream);
}

// F.softmax(x, dim=0) over the GLOBAL batch of a data-parallel run (the reference's DataParallel gathers the logits on one
// device before train_adapt.py:151,166,174): gmax / gsum = s2r_softmax0_batch_stats all-reduced (MAX, then SUM) over the ranks.
extern "C" int s2r_softmax0_nchw_to_nhwc_pad_global(const float* x, int B, int C, int H, int W, const float* gmax,
                                                    const float* gsum, void* yp, int Cp, s2r_stream_t stream) {
  S2R_REQUIRE(x && yp && gmax && gsum && B >= 1 && C >= 1 && H >= 1 && W >= 1 && H <= 65535, S2R_ERR_SHAPE,
              "softmax0_to_nhwc_pad_global: bad arguments");
  S2R_REQUIRE(Cp % 8 == 0 && Cp >= C && Cp <= 64 && (uintptr_t)yp % 16 == 0, S2R_ERR_SHAPE,
              "softmax0_to_nhwc_pad_global: channel pitch %d (need a multiple of 8 in [C, 64]) / alignment", Cp);
  S2R_REQUIRE(B <= SP_MAXB, S2R_ERR_UNSUPPORTED, "softmax0_to_nhwc_pad_global: batch %d > %d per rank", B, SP_MAXB);
  return launch_sp_fwd<true>(x, B, C, H, W, yp, Cp, (cudaStream_t)stream, gmax, gsum);
}

// Backward of the above in two launches around one all-reduce(SUM): tpart != NULL: tpart[c][h][w] = this rank's part of
// sum_b' g_b' y_b' (dx untouched); tpart == NULL: dx_b = y_b * (g_b - tglob) with the all-reduced sum in tglob.
extern "C" int s2r_softmax0_nhwc_pad_bwd_global(const float* x, const void* gp, int B, int C, int H, int W, int Cp,
                                                const float* gmax, const float* gsum, float* tpart, const float* tglob,
                                                float* dx, s2r_stream_t stream) {
  S2R_REQUIRE(x && gp && gmax && gsum && (tpart || (tglob && dx)) && B >= 1 && C >= 1 && H >= 1 && W >= 1 && H <= 65535,
              S2R_ERR_SHAPE, "softmax0_nhwc_pad_bwd_global: bad arguments");
  S2R_REQUIRE(Cp % 8 == 0 && Cp >= C && Cp <= 64 && (uintptr_t)gp % 16 == 0, S2R_ERR_SHAPE,
              "softmax0_nhwc_pad_bwd_global: channel pitch %d (need a multiple of 8 in [C, 64]) / alignment", Cp);
  S2R_REQUIRE(B <= SP_MAXB, S2R_ERR_UNSUPPORTED, "softmax0_nhwc_pad_bwd_global: batch %d > %d per rank", B, SP_MAXB);
  return launch_sp_bwd<true>(x, gp, B, C, H, W, Cp, dx, (cudaStream_t)stream, gmax, gsum, tpart, tglob);
}
