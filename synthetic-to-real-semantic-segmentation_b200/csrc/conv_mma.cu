// Generic tap-GEMM convolution on warp-level tensor-core MMAs (mma.sync m16n8k16 bf16, fp32
// accumulate).  This is the shape-agnostic path of s2r_conv_fwd / s2r_conv_wgrad: any tap
// table, any Cin multiple of 8, any Cout.  The tcgen05/TMA kernels in conv_tc.cu take over the
// shapes they support; this file also serves as their on-device cross-check.
//
// Forward/data-gradient:  out[m][co] = epi(sum_t sum_ci src_t[pix(m)+d_t][ci] * W[t][co][ci])
//   CTA tile 128 pixels x BN couts x 64 channels, 3-stage cp.async pipeline, A gathered per
//   16-byte channel chunk with zero fill outside the tap's view, 128B-row XOR-swizzled smem,
//   ldmatrix fragments.  Epilogue: bias, BN statistics (sum / sum of squares per channel, fp32
//   accumulators reduced by shuffles + shared atomics, one fp64 atomic per channel per CTA),
//   activation, residual add or leaky mask, bf16 store.
// Weight gradient:  dW[t][co][ci] += sum_m dy[m][co] * src_t[pix(m)+d_t][ci]
//   pixels are the contraction axis: both operands are read pixel-major and fed through
//   ldmatrix.trans; the pixel range is split across CTAs and reduced with fp32 atomics.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int THREADS = 256;

struct TapDev {
  const __nv_bfloat16* base;
  long long sn, sh, sw;
  int H, W, dh, dw;
  int wslice;
  int pad_;
  long long wofs;
};

struct FwdParams {
  TapDev taps[S2R_MAX_TAPS];
  int ntaps;
  int N, OH, OW, Cin, Cout;
  int M;
  const __nv_bfloat16* w;
  int Cout_pad, Kpad;
  __nv_bfloat16* out;
  long long on, oh, ow;
  const float* bias;
  const float* oscale;
  int act;
  float slope;
  int aux_mode;
  const __nv_bfloat16* aux;
  long long an, ah, aw;
  double* stats;
};

struct WgParams {
  TapDev taps[S2R_MAX_TAPS];
  int ntaps;
  int N, OH, OW, Cin, Cout;
  int P;
  const __nv_bfloat16* dy;
  long long dn, dh, dw;
  float* dwt;
  long long s_co, s_ci;
  int pix_per_split;
  int ci_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------ forward
template <int BN>
__global__ void __launch_bounds__(THREADS)
tap_fwd_kernel(const __grid_constant__ FwdParams p) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int WN = BN / 32, WM = 8 / WN, WTM = BM / WM, MI = WTM / 16, NI = 4;
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_sum[BN], s_sq[BN];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp / WN, warp_n = warp % WN;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const uint32_t sbase = smem_u32(smem);

  if (p.stats) {
    for (int i = tid; i < BN; i += THREADS) {
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
    }
  }

  // loader: this thread serves chunk (tid&7) of rows (tid>>3) + 32*j
  const int chunk = tid & 7;
  int r_n[4], r_oh[4], r_ow[4];
  bool r_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + (tid >> 3) + 32 * j;
    r_ok[j] = m < p.M;
    const int mm = r_ok[j] ? m : 0;
    r_ow[j] = mm % p.OW;
    const int t = mm / p.OW;
    r_oh[j] = t % p.OH;
    r_n[j] = t / p.OH;
  }
  const int kchunks = (p.Cin + BK - 1) / BK;
  const int kiters = p.ntaps * kchunks;

  auto load_stage = [&](int it, int stage) {
    const int tap = it / kchunks, kc = it - tap * kchunks;
    const TapDev& T = p.taps[tap];
    const int c = kc * BK + chunk * 8;
    const bool cok = c < p.Cin;
    const uint32_t a_base = sbase + stage * STAGE_BYTES;
    const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = (tid >> 3) + 32 * j;
      const int y = r_oh[j] + T.dh, x = r_ow[j] + T.dw;
      const bool ok = r_ok[j] && cok && (unsigned)y < (unsigned)T.H && (unsigned)x < (unsigned)T.W;
      const __nv_bfloat16* src = ok ? T.base + r_n[j] * T.sn + y * T.sh + x * T.sw + c : T.base;
      cp_async16(a_base + row * 128 + ((chunk ^ (row & 7)) << 4), src, ok);
    }
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) {
      const int row = (tid >> 3) + 32 * j;
      const int co = n0 + row;
      const bool ok = co < p.Cout_pad;
      const __nv_bfloat16* src =
          p.w + ((long long)T.wslice * p.Cout_pad + (ok ? co : 0)) * p.Kpad + kc * BK + chunk * 8;
      cp_async16(b_base + row * 128 + ((chunk ^ (row & 7)) << 4), src, ok);
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < kiters) load_stage(s, s);
    cp_async_commit();
  }

  for (int it = 0; it < kiters; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < kiters) load_stage(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const uint32_t a_base = sbase + (it % STAGES) * STAGE_BYTES;
    const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
    for (int kk = 0; kk < BK / 16; ++kk) {
      uint32_t af[MI][4], bf[NI][2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int row = warp_m * WTM + mi * 16 + (lane & 15);
        const int ch = kk * 2 + (lane >> 4);
        ldmatrix_x4(af[mi], a_base + row * 128 + ((ch ^ (row & 7)) << 4));
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {
        const int row = warp_n * 32 + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int ch = kk * 2 + ((lane >> 3) & 1);
        uint32_t r[4];
        ldmatrix_x4(r, b_base + row * 128 + ((ch ^ (row & 7)) << 4));
        bf[nj * 2][0] = r[0];
        bf[nj * 2][1] = r[1];
        bf[nj * 2 + 1][0] = r[2];
        bf[nj * 2 + 1][1] = r[3];
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) mma_bf16(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue
  const int gid = lane >> 2, tq = lane & 3;
  float csum[NI][2], csq[NI][2];
#pragma unroll
  for (int ni = 0; ni < NI; ++ni) {
    csum[ni][0] = csum[ni][1] = 0.f;
    csq[ni][0] = csq[ni][1] = 0.f;
  }
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + warp_m * WTM + mi * 16 + gid + h * 8;
      const bool mok = m < p.M;
      const int mm = mok ? m : 0;
      const int ow_ = mm % p.OW;
      const int t = mm / p.OW;
      const int oh_ = t % p.OH;
      const int n_ = t / p.OH;
      __nv_bfloat16* orow = p.out + n_ * p.on + oh_ * p.oh + ow_ * p.ow;
      const __nv_bfloat16* arow = p.aux ? p.aux + n_ * p.an + oh_ * p.ah + ow_ * p.aw : nullptr;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int co = n0 + warp_n * 32 + ni * 8 + tq * 2;
        float v0 = acc[mi][ni][h * 2], v1 = acc[mi][ni][h * 2 + 1];
        const bool c0 = co < p.Cout, c1 = co + 1 < p.Cout;
        if (p.oscale) {
          if (c0) v0 *= __ldg(p.oscale + co);
          if (c1) v1 *= __ldg(p.oscale + co + 1);
        }
        if (p.bias) {
          if (c0) v0 += __ldg(p.bias + co);
          if (c1) v1 += __ldg(p.bias + co + 1);
        }
        if (!mok || !c0) continue;
        if (p.aux_mode == S2R_AUX_LEAKY_MASK) {
          const float a0 = __bfloat162float(arow[co]);
          v0 *= (a0 > 0.f ? 1.f : p.slope);
          if (c1) v1 *= (__bfloat162float(arow[co + 1]) > 0.f ? 1.f : p.slope);
        } else {
          v0 = apply_act(v0, p.act, p.slope);
          v1 = apply_act(v1, p.act, p.slope);
          if (p.aux_mode == S2R_AUX_ADD) {
            v0 += __bfloat162float(arow[co]);
            if (c1) v1 += __bfloat162float(arow[co + 1]);
          }
        }
        // statistics of the STORED values (after activation / auxiliary operand, rounded to bf16), like conv_tc.cu
        v0 = __bfloat162float(__float2bfloat16(v0));
        v1 = c1 ? __bfloat162float(__float2bfloat16(v1)) : 0.f;
        csum[ni][0] += v0;
        csum[ni][1] += v1;
        csq[ni][0] += v0 * v0;
        csq[ni][1] += v1 * v1;
        if (c1 && ((reinterpret_cast<uintptr_t>(orow + co) & 3) == 0)) {
          *reinterpret_cast<__nv_bfloat162*>(orow + co) = __floats2bfloat162_rn(v0, v1);
        } else {
          orow[co] = __float2bfloat16(v0);
          if (c1) orow[co + 1] = __float2bfloat16(v1);
        }
      }
    }
  }
  if (p.stats) {
    // reduce over the 8 row groups of the warp (lane bits 2..4), then shared atomics
#pragma unroll
    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s = csum[ni][j], q = csq[ni][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (gid == 0) {
          const int cl = warp_n * 32 + ni * 8 + tq * 2 + j;
          atomicAdd(&s_sum[cl], s);
          atomicAdd(&s_sq[cl], q);
        }
      }
    __syncthreads();
    for (int i = tid; i < BN; i += THREADS) {
      const int co = n0 + i;
      if (co < p.Cout) {
        atomicAdd(&p.stats[co], (double)s_sum[i]);
        atomicAdd(&p.stats[p.Cout + co], (double)s_sq[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------ weight gradient
constexpr int WBK = 32;  // pixels per pipeline stage

template <int BN>
__global__ void __launch_bounds__(THREADS)
tap_wgrad_kernel(const __grid_constant__ WgParams p) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  constexpr int WN = BN / 32, WM = 8 / WN, WTM = BM / WM, MI = WTM / 16, NI = 4;
  constexpr int A_ROW = BM * 2, B_ROW = BN * 2;  // bytes per pixel row
  constexpr int A_BYTES = WBK * A_ROW, B_BYTES = WBK * B_ROW, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int A_CH = BM / 8, B_CH = BN / 8;  // 16-byte chunks per row
  extern __shared__ __align__(128) uint8_t smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp / WN, warp_n = warp % WN;
  const int co0 = blockIdx.x * BM;
  const int tap = blockIdx.y / p.ci_tiles;
  const int ci0 = (blockIdx.y - tap * p.ci_tiles) * BN;
  const TapDev& T = p.taps[tap];
  const int pbeg = blockIdx.z * p.pix_per_split;
  const int pend = min(p.P, pbeg + p.pix_per_split);
  if (pbeg >= pend) return;
  const int kiters = (pend - pbeg + WBK - 1) / WBK;
  const uint32_t sbase = smem_u32(smem);
  const int cout8 = (p.Cout + 7) & ~7, cin8 = (p.Cin + 7) & ~7;

  auto load_stage = [&](int it, int stage) {
    const uint32_t a_base = sbase + stage * STAGE_BYTES;
    const uint32_t b_base = a_base + A_BYTES;
    const int k0 = pbeg + it * WBK;
#pragma unroll
    for (int j = 0; j < (WBK * A_CH) / THREADS; ++j) {
      const int idx = tid + THREADS * j;
      const int row = idx / A_CH, ch = idx % A_CH;
      const int pix = k0 + row;
      const bool pok = pix < pend;
      const int pp = pok ? pix : 0;
      const int ow_ = pp % p.OW;
      const int t = pp / p.OW;
      const int oh_ = t % p.OH;
      const int n_ = t / p.OH;
      const int co = co0 + ch * 8;
      const bool ok = pok && co < cout8;
      const __nv_bfloat16* src = ok ? p.dy + n_ * p.dn + oh_ * p.dh + ow_ * p.dw + co : p.dy;
      cp_async16(a_base + row * A_ROW + ((ch ^ (row & 7)) << 4), src, ok);
    }
#pragma unroll
    for (int j = 0; j < (WBK * B_CH) / THREADS; ++j) {
      const int idx = tid + THREADS * j;
      const int row = idx / B_CH, ch = idx % B_CH;
      const int pix = k0 + row;
      const bool pok = pix < pend;
      const int pp = pok ? pix : 0;
      const int ow_ = pp % p.OW;
      const int t = pp / p.OW;
      const int oh_ = t % p.OH;
      const int n_ = t / p.OH;
      const int y = oh_ + T.dh, x = ow_ + T.dw;
      const int ci = ci0 + ch * 8;
      const bool ok = pok && ci < cin8 && (unsigned)y < (unsigned)T.H && (unsigned)x < (unsigned)T.W;
      const __nv_bfloat16* src = ok ? T.base + n_ * T.sn + y * T.sh + x * T.sw + ci : T.base;
      cp_async16(b_base + row * B_ROW + ((ch ^ (row & 7)) << 4), src, ok);
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < kiters) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < kiters; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < kiters) load_stage(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const uint32_t a_base = sbase + (it % STAGES) * STAGE_BYTES;
    const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
    for (int kk = 0; kk < WBK / 16; ++kk) {
      uint32_t af[MI][4], bf[NI][2];
      const int j = lane >> 3;
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int k = kk * 16 + (lane & 7) + (j >> 1) * 8;
        const int mch = (warp_m * WTM + mi * 16) / 8 + (j & 1);
        ldmatrix_x4_trans(af[mi], a_base + k * A_ROW + ((mch ^ (k & 7)) << 4));
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {
        const int k = kk * 16 + (lane & 7) + (j & 1) * 8;
        const int nch = (warp_n * 32 + nj * 16) / 8 + (j >> 1);
        uint32_t r[4];
        ldmatrix_x4_trans(r, b_base + k * B_ROW + ((nch ^ (k & 7)) << 4));
        bf[nj * 2][0] = r[0];
        bf[nj * 2][1] = r[1];
        bf[nj * 2 + 1][0] = r[2];
        bf[nj * 2 + 1][1] = r[3];
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) mma_bf16(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();

  const int gid = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int co = co0 + warp_m * WTM + mi * 16 + gid + h * 8;
      if (co >= p.Cout) continue;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = ci0 + warp_n * 32 + ni * 8 + tq * 2 + e;
          if (ci < p.Cin) atomicAdd(p.dwt + co * p.s_co + ci * p.s_ci + T.wofs, acc[mi][ni][h * 2 + e]);
        }
    }
}

// ------------------------------------------------------------------------------ weight packing
// Source element of packed[t][a][b] in the OIHW fp32 filter, or -1 for zero padding.
//   mode 0: forward, (a, b) = (co, ci), t = filter tap           mode 1: data gradient, (a, b) = (ci, co)
// Row-tap forms of a 4x4 stride-2 pad-1 filter over a zero-padded NHWC buffer with Cp = round8(Cin) channels per pixel
// (FCDiscriminator conv1; see engine.rowtap_*):
//   mode 2: forward, slice t = kh, a = co, b = kw*Cp + c   (the four kw taps of a filter row are one contiguous run
//           of 4*Cp channels in memory)
//   mode 3/4: data gradient of the padded rows of parity ph = mode - 3, slice t = 2*ta + tb (dy offset (-ta, -tb)),
//           a = pw*Cp + c (two adjacent padded pixels), b = co, filter element (kh, kw) = (ph + 2*ta, pw + 2*tb)
__device__ __forceinline__ long long pack_src(int mode, int Cout, int Cin, int RS, int t, int a, int b) {
  if (mode <= 1) {
    const int co = mode ? b : a, ci = mode ? a : b;
    return (co < Cout && ci < Cin) ? ((long long)co * Cin + ci) * RS + t : -1;
  }
  const int Cp = (Cin + 7) & ~7;
  if (mode == 2) {
    const int kw = b / Cp, c = b - kw * Cp;
    return (a < Cout && kw < 4 && c < Cin) ? ((long long)a * Cin + c) * 16 + t * 4 + kw : -1;
  }
  const int ph = mode - 3, pw = a / Cp, c = a - pw * Cp;
  const int kh = ph + 2 * (t >> 1), kw = pw + 2 * (t & 1);
  return (b < Cout && pw < 2 && c < Cin) ? ((long long)b * Cin + c) * 16 + kh * 4 + kw : -1;
}

__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int RS, int nslices, int mode,
                                   __nv_bfloat16* __restrict__ out, int A_pad, int B_pad) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)nslices * A_pad * B_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % B_pad);
    const long long q = i / B_pad;
    const int a = (int)(q % A_pad);
    const int t = (int)(q / A_pad);
    const long long src = pack_src(mode, Cout, Cin, RS, t, a, b);
    out[i] = __float2bfloat16(src >= 0 ? __ldg(w + src) : 0.f);
  }
}

// One launch for many filters: blockIdx.y walks a device table of jobs (each a range of at most a few thousand
// output elements of one packed filter), so re-packing all ~190 filter copies after an optimizer step is one
// launch instead of ~100 (each of which was mostly launch latency).
__global__ void pack_weights_multi_kernel(const s2r_pack_job* __restrict__ jobs) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const s2r_pack_job j = jobs[blockIdx.y];
  const float* __restrict__ w = j.w;
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(j.packed);
  if (j.transpose & 16) {
    // modes 0 / 1, several taps: [begin, end) counts (a, b) PAIRS and a thread writes all RS taps of its pair -- the RS
    // source floats of a pair are contiguous in the OIHW filter (36 bytes of a 3x3 filter), so a warp of consecutive b
    // (mode 0: consecutive input channels) reads one contiguous run once instead of RS strided passes over it, and the
    // index arithmetic is paid once per pair
    const int mode = j.transpose & 15;
    const unsigned B = (unsigned)j.B_pad, end = (unsigned)j.end;
    const size_t slice = (size_t)j.A_pad * j.B_pad;
    for (unsigned i = (unsigned)j.begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
      const unsigned a = i / B, b = i - a * B;
      const int co = mode ? (int)b : (int)a, ci = mode ? (int)a : (int)b;
      const bool in = co < j.Cout && ci < j.Cin;
      const float* src = w + ((size_t)co * j.Cin + ci) * j.RS;
      for (int t = 0; t < j.RS; ++t) out[(size_t)t * slice + i] = __float2bfloat16(in ? __ldg(src + t) : 0.f);
    }
    return;
  }
  if (j.end <= 0x7fffffffLL) {
    // 32-bit index arithmetic: the three div / mod pairs per element are ~10 instructions each in 32 bits and ~40 in
    // 64 bits, and they were most of this kernel (149 us per feature-adaptation step for 19 M elements)
    const unsigned B = (unsigned)j.B_pad, A = (unsigned)j.A_pad, end = (unsigned)j.end;
    for (unsigned i = (unsigned)j.begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
      const unsigned q = i / B, b = i - q * B;
      const unsigned t = q / A, a = q - t * A;
      const long long src = pack_src(j.transpose, j.Cout, j.Cin, j.RS, (int)t, (int)a, (int)b);
      out[i] = __float2bfloat16(src >= 0 ? __ldg(w + src) : 0.f);
    }
    return;
  }
  for (long long i = j.begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.end;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % j.B_pad);
    const long long q = i / j.B_pad;
    const int a = (int)(q % j.A_pad);
    const int t = (int)(q / j.A_pad);
    const long long src = pack_src(j.transpose, j.Cout, j.Cin, j.RS, t, a, b);
    out[i] = __float2bfloat16(src >= 0 ? __ldg(w + src) : 0.f);
  }
}

// w.grad[co][ci][t] += G[t][co][ci] (row pitch Cp): multi-tap weight gradients are accumulated tap-major, so that the
// atomics of the gradient kernels cover 32 consecutive input channels (one 128-byte line) per instruction -- in the
// OIHW tensor itself those 32 elements are R*S floats apart -- and are moved into the OIHW gradient here
__global__ void wgrad_scatter_taps_kernel(const float* __restrict__ G, float* __restrict__ dw, int Cout, int Cin, int RS,
                                          int Cp) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const long long total = (long long)Cout * Cin * RS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % RS);
    const long long q = i / RS;
    const int ci = (int)(q % Cin), co = (int)(q / Cin);
    atomicAdd(dw + i, G[((long long)t * Cout + co) * Cp + ci]);   // atomic: see s2r_add_f64_to_f32
  }
}

// w.grad[co][c][kh][kw] += G[co][kh][kw*Cp + c]: the row-tap weight gradient (mode 2 layout, fp32) back to OIHW
__global__ void rowtap_wgrad_scatter_kernel(const float* __restrict__ G, float* __restrict__ dw, int Cout, int Cin, int Cp) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  const int total = Cout * Cin * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kw = i & 3, kh = (i >> 2) & 3, q = i >> 4;
    const int c = q % Cin, co = q / Cin;
    atomicAdd(dw + i, G[((long long)co * 4 + kh) * (4 * Cp) + kw * Cp + c]);
  }
}

void copy_taps(TapDev* dst, const s2r_tap* src, int n) {
  for (int i = 0; i < n; ++i) {
    dst[i].base = (const __nv_bfloat16*)src[i].base;
    dst[i].sn = src[i].sn;
    dst[i].sh = src[i].sh;
    dst[i].sw = src[i].sw;
    dst[i].H = src[i].H;
    dst[i].W = src[i].W;
    dst[i].dh = src[i].dh;
    dst[i].dw = src[i].dw;
    dst[i].wslice = src[i].wslice;
    dst[i].pad_ = 0;
    dst[i].wofs = src[i].wofs;
  }
}

template <int BN>
int launch_fwd(const FwdParams& p, cudaStream_t st) {
  constexpr int smem = STAGES * (BM * 128 + BN * 128);
  static bool attr_set = false;
  if (!attr_set) {
    S2R_CUDA_OK(cudaFuncSetAttribute(tap_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(s2r_div_up(p.M, BM), s2r_div_up(p.Cout, BN));
  S2R_CUDA_OK(s2r_launch(tap_fwd_kernel<BN>, dim3(grid), dim3(THREADS), (size_t)(smem), st, p));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

template <int BN>
int launch_wgrad(WgParams& p, cudaStream_t st) {
  constexpr int smem = STAGES * (WBK * BM * 2 + WBK * BN * 2);
  static bool attr_set = false;
  if (!attr_set) {
    S2R_CUDA_OK(cudaFuncSetAttribute(tap_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  p.ci_tiles = s2r_div_up(p.Cin, BN);
  const int tiles = s2r_div_up(p.Cout, BM) * p.ci_tiles * p.ntaps;
  // split the pixel axis so that roughly 4 CTAs per SM are in flight, at least 8 stages each
  int splits = s2r_div_up((long)s2r_sm_count() * 4, tiles);
  const int max_splits = s2r_div_up(p.P, WBK * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int pps = s2r_div_up(p.P, splits);
  pps = (pps + WBK - 1) / WBK * WBK;
  splits = s2r_div_up(p.P, pps);
  p.pix_per_split = pps;
  dim3 grid(s2r_div_up(p.Cout, BM), p.ci_tiles * p.ntaps, splits);
  S2R_CUDA_OK(s2r_launch(tap_wgrad_kernel<BN>, dim3(grid), dim3(THREADS), (size_t)(smem), st, p));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

int check_taps(const s2r_tap* taps, int ntaps, int Cin, const char* who) {
  S2R_REQUIRE(ntaps >= 1 && ntaps <= S2R_MAX_TAPS, S2R_ERR_UNSUPPORTED, "%s: %d taps outside [1,%d]", who, ntaps, S2R_MAX_TAPS);
  for (int i = 0; i < ntaps; ++i) {
    S2R_REQUIRE(taps[i].base != nullptr && (uintptr_t)taps[i].base % 16 == 0, S2R_ERR_SHAPE, "%s: tap %d base null or not 16B aligned", who, i);
    S2R_REQUIRE(taps[i].sn % 8 == 0 && taps[i].sh % 8 == 0 && taps[i].sw % 8 == 0, S2R_ERR_SHAPE, "%s: tap %d strides must be multiples of 8 elements", who, i);
    S2R_REQUIRE(taps[i].H >= 1 && taps[i].W >= 1, S2R_ERR_SHAPE, "%s: tap %d empty view", who, i);
  }
  S2R_REQUIRE(Cin >= 8 && Cin % 8 == 0, S2R_ERR_SHAPE, "%s: Cin=%d must be a positive multiple of 8", who, Cin);
  return S2R_OK;
}

}  // namespace

// implemented in conv_tc.cu: returns 1 when it handled the problem, 0 to fall through, <0 on error
int s2r_conv_fwd_tc(const s2r_conv_args* a, cudaStream_t st);

extern "C" int s2r_conv_fwd_mma(const s2r_conv_args* a, s2r_stream_t stream) {
  S2R_REQUIRE(a && a->struct_size == sizeof(s2r_conv_args), S2R_ERR_SHAPE, "conv_fwd: struct size mismatch (%u vs %zu)", a ? a->struct_size : 0u, sizeof(s2r_conv_args));
  int rc = check_taps(a->taps, a->ntaps, a->Cin, "conv_fwd");
  if (rc) return rc;
  S2R_REQUIRE(a->N >= 1 && a->OH >= 1 && a->OW >= 1 && a->Cout >= 1, S2R_ERR_SHAPE, "conv_fwd: bad output shape");
  S2R_REQUIRE((long long)a->N * a->OH * a->OW < (1ll << 31), S2R_ERR_UNSUPPORTED, "conv_fwd: more than 2^31 output pixels");
  S2R_REQUIRE(a->w && (uintptr_t)a->w % 16 == 0 && a->Kpad % BK == 0 && a->Kpad >= a->Cin && a->Cout_pad >= a->Cout,
              S2R_ERR_SHAPE, "conv_fwd: packed weight must be 16B aligned with Kpad %% 64 == 0, Kpad >= Cin, Cout_pad >= Cout");
  S2R_REQUIRE(a->out != nullptr, S2R_ERR_SHAPE, "conv_fwd: null output");
  S2R_REQUIRE(a->aux_mode == S2R_AUX_NONE || a->aux != nullptr, S2R_ERR_SHAPE, "conv_fwd: aux_mode set but aux is null");
  FwdParams p;
  copy_taps(p.taps, a->taps, a->ntaps);
  p.ntaps = a->ntaps;
  p.N = a->N; p.OH = a->OH; p.OW = a->OW; p.Cin = a->Cin; p.Cout = a->Cout;
  p.M = a->N * a->OH * a->OW;
  p.w = (const __nv_bfloat16*)a->w;
  p.Cout_pad = a->Cout_pad; p.Kpad = a->Kpad;
  p.out = (__nv_bfloat16*)a->out;
  p.on = a->on; p.oh = a->oh; p.ow = a->ow;
  p.bias = a->bias; p.act = a->act; p.slope = a->slope;
  p.oscale = a->oscale;
  p.aux_mode = a->aux_mode;
  p.aux = a->aux_mode == S2R_AUX_NONE ? nullptr : (const __nv_bfloat16*)a->aux;
  p.an = a->an; p.ah = a->ah; p.aw = a->aw;
  p.stats = a->stats;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->Cout > 64) return launch_fwd<128>(p, st);
  if (a->Cout > 32) return launch_fwd<64>(p, st);
  return launch_fwd<32>(p, st);
}

extern "C" int s2r_conv_fwd(const s2r_conv_args* a, s2r_stream_t stream) {
  S2R_REQUIRE(a && a->struct_size == sizeof(s2r_conv_args), S2R_ERR_SHAPE, "conv_fwd: struct size mismatch");
  int rc = s2r_conv_fwd_tc(a, (cudaStream_t)stream);
  if (rc != 0) return rc < 0 ? rc : S2R_OK;
  return s2r_conv_fwd_mma(a, stream);
}

int s2r_conv_wgrad_tc(const s2r_wgrad_args* a, cudaStream_t st);

extern "C" int s2r_conv_wgrad(const s2r_wgrad_args* a, s2r_stream_t stream) {
  S2R_REQUIRE(a && a->struct_size == sizeof(s2r_wgrad_args), S2R_ERR_SHAPE, "conv_wgrad: struct size mismatch (%u vs %zu)", a ? a->struct_size : 0u, sizeof(s2r_wgrad_args));
  S2R_REQUIRE(a->dweight != nullptr, S2R_ERR_SHAPE, "conv_wgrad: null dweight");
  {
    const int rc_tc = s2r_conv_wgrad_tc(a, (cudaStream_t)stream);
    if (rc_tc != 0) return rc_tc < 0 ? rc_tc : S2R_OK;
  }
  int rc = check_taps(a->taps, a->ntaps, (a->Cin + 7) & ~7, "conv_wgrad");
  if (rc) return rc;
  S2R_REQUIRE(a->N >= 1 && a->OH >= 1 && a->OW >= 1 && a->Cout >= 1, S2R_ERR_SHAPE, "conv_wgrad: bad shape");
  S2R_REQUIRE((long long)a->N * a->OH * a->OW < (1ll << 31), S2R_ERR_UNSUPPORTED, "conv_wgrad: more than 2^31 pixels");
  S2R_REQUIRE(a->dy && (uintptr_t)a->dy % 16 == 0 && a->dn % 8 == 0 && a->dh % 8 == 0 && a->dw % 8 == 0,
              S2R_ERR_SHAPE, "conv_wgrad: dy must be 16B aligned with strides multiple of 8");
  S2R_REQUIRE(a->dweight != nullptr, S2R_ERR_SHAPE, "conv_wgrad: null dweight");
  WgParams p;
  copy_taps(p.taps, a->taps, a->ntaps);
  p.ntaps = a->ntaps;
  p.N = a->N; p.OH = a->OH; p.OW = a->OW; p.Cin = a->Cin; p.Cout = a->Cout;
  p.P = a->N * a->OH * a->OW;
  p.dy = (const __nv_bfloat16*)a->dy;
  p.dn = a->dn; p.dh = a->dh; p.dw = a->dw;
  p.dwt = a->dweight;
  p.s_co = a->s_co; p.s_ci = a->s_ci;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->Cin > 64) return launch_wgrad<128>(p, st);
  return launch_wgrad<64>(p, st);
}

extern "C" int s2r_pack_weight(const float* w, int Cout, int Cin, int R, int S, int mode,
                               void* packed, int A_pad, int B_pad, s2r_stream_t stream) {
  S2R_REQUIRE(w && packed && Cout >= 1 && Cin >= 1 && R >= 1 && S >= 1 && mode >= 0 && mode <= 4, S2R_ERR_SHAPE,
              "pack_weight: bad arguments");
  S2R_REQUIRE(mode <= 1 || (R == 4 && S == 4), S2R_ERR_UNSUPPORTED, "pack_weight: row-tap modes take 4x4 filters");
  const int Cp = (Cin + 7) & ~7;
  const int A = mode == 0 ? Cout : mode == 1 ? Cin : mode == 2 ? Cout : 2 * Cp;
  const int B = mode == 0 ? Cin : mode == 1 ? Cout : mode == 2 ? 4 * Cp : Cout;
  const int nslices = mode <= 1 ? R * S : 4;
  S2R_REQUIRE(A_pad >= A && B_pad >= B, S2R_ERR_SHAPE, "pack_weight: padded dims (%d,%d) smaller than (%d,%d)", A_pad, B_pad, A, B);
  const long long total = (long long)nslices * A_pad * B_pad;
  S2R_CUDA_OK(s2r_launch(pack_weight_kernel, dim3(s2r_grid(total, 256, 8)), dim3(256), (size_t)0, (cudaStream_t)stream, w,
                         Cout, Cin, R * S, nslices, mode, (__nv_bfloat16*)packed, A_pad, B_pad));
  return S2R_OK;
}

extern "C" int s2r_wgrad_scatter_taps(const float* G, float* dw, int Cout, int Cin, int RS, int Cp, s2r_stream_t stream) {
  S2R_REQUIRE(G && dw && Cout >= 1 && Cin >= 1 && RS >= 1 && Cp >= Cin, S2R_ERR_SHAPE, "wgrad_scatter_taps: bad arguments");
  const long long total = (long long)Cout * Cin * RS;
  S2R_CUDA_OK(s2r_launch(wgrad_scatter_taps_kernel, dim3(s2r_grid(total, 256, 4)), dim3(256), (size_t)0, (cudaStream_t)stream, G, dw, Cout, Cin, RS, Cp));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_rowtap_wgrad_scatter(const float* G, float* dw, int Cout, int Cin, s2r_stream_t stream) {
  S2R_REQUIRE(G && dw && Cout >= 1 && Cin >= 1, S2R_ERR_SHAPE, "rowtap_wgrad_scatter: bad arguments");
  const int total = Cout * Cin * 16;
  S2R_CUDA_OK(s2r_launch(rowtap_wgrad_scatter_kernel, dim3(s2r_grid(total, 256, 1)), dim3(256), (size_t)0, (cudaStream_t)stream, G, dw, Cout, Cin, (Cin + 7) & ~7));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_pack_weights_multi(const s2r_pack_job* jobs, int njobs, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535, S2R_ERR_SHAPE, "pack_weights_multi: %d jobs", njobs);
  if (njobs == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "pack_weights_multi: null table");
  S2R_CUDA_OK(s2r_launch(pack_weights_multi_kernel, dim3(dim3(4, njobs)), dim3(256), (size_t)0, (cudaStream_t)stream, jobs));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
