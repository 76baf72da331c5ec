// Streaming depthwise 3x3 (stride 1, dilation 1, padding 1) with the BN + ReLU6 prologue: the
// forward kernel and a fused data-gradient + weight-gradient kernel.  Covers 14 of the 17
// InvertedResidual blocks of the reference (modeling/backbone/mobilenet.py:26-68); dwconv.cu keeps
// the generic (stride 2 / dilated) kernels and dispatches here.
//
// Why a second design: at the HBM roofline (2 B in + 2 B out per element, 23 B/clk/SM) an SM has
// ~0.17 clk per element, i.e. ~22 issue slots per warp-row of 128 elements -- a 3x3 depthwise conv
// with its BN prologue and its statistics is almost ALU-bound on B200.  So:
//   * every input element is loaded from HBM once, converted once and activated once
//     (relu6(x*sc+sh) = 6*sat(x*sc/6+sh/6): ONE FFMA.SAT, the 6 folded into the filter);
//   * horizontal neighbours are exchanged through a 2-slot shared-memory ring as packed bf16
//     (one STS.64 + two LDS.64 per thread and row) instead of being re-loaded and re-activated;
//   * vertical reuse is a rolling 3x3 register window: a CTA walks DOWN a strip of TW columns x
//     one channel chunk for `rs` rows (one barrier per row);
//   * all multiply-adds are packed FFMA2 (fma.rn.f32x2, two channels per instruction);
//   * global loads are register-prefetched PF rows ahead (27 KB in flight per SM).
//   * fused backward: g = act'(pre) * sum_k dy(pos+1-k) w[k] and dw[k] += a(pos) dy(pos+1-k) use
//     the SAME dy window, so dy and x are read once (6 B/element instead of 10 for two kernels).
// thread = (column j, 4 consecutive channels); lane order is channel-fastest, so a warp touches
// contiguous CG*8-byte pixel segments.  Column 0 and TW+1 of a CTA are load-only halo columns.
#include "common.cuh"

namespace {

constexpr int PF = 6;          // rows of loads in flight per thread; multiple of 6 (ring parity x window)
constexpr int NT_MAX = 288;    // threads per CTA

struct S1Geom {
  int N, H, W, C;   // tensor
  int CG;           // 4-channel groups per CTA chunk (chunk = CG*4 channels)
  int TW;           // output columns per CTA
  int rs;           // rows per CTA
  int nseg;         // row segments per image
  int ext;          // backward: gradient domain extension (0 or 1)
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
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
__device__ __forceinline__ uint2 ldg8(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// sum `nv` per-thread floats over the thread columns of a CTA (threads tid = g + CG*j share g);
// result for (g, k) is returned to thread t = g*nv + k < CG*nv.  red: [nv][NT_MAX+1] floats.
template <int NV>
__device__ __forceinline__ float column_reduce(float* red, const float* v, int CG, int ncol) {
#pragma unroll
  for (int k = 0; k < NV; ++k) red[k * (NT_MAX + 1) + threadIdx.x] = v[k];
  __syncthreads();
  float acc = 0.f;
  const int t = threadIdx.x;
  if (t < CG * NV) {
    const int g = t / NV, k = t - g * NV;
    const float* p = red + k * (NT_MAX + 1) + g;
    for (int j = 0; j < ncol; ++j) acc += p[j * CG];
  }
  return acc;
}

// ------------------------------------------------------------------------------------ forward
// y[oh][ow] = sum_{ky,kx} a(oh+ky-1, ow+kx-1) w[ky][kx],  a = relu6(x*sc+sh) inside the image and
// HALO ? relu6(sh) : 0 outside.  stats += per-channel sum / sum of squares of y.
template <bool HALO>
__global__ void __launch_bounds__(NT_MAX, 2)
dw_s1_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ ss, const float* __restrict__ w,
                 __nv_bfloat16* __restrict__ y, double* __restrict__ stats, S1Geom G) {
  __shared__ uint2 ring[2][NT_MAX];
  __shared__ float red[8 * (NT_MAX + 1)];
  const int CG = G.CG, TWL = G.TW + 2;
  const int g = threadIdx.x % CG, j = threadIdx.x / CG;
  const int c = (blockIdx.x * CG + g) * 4;
  const int ow0 = blockIdx.y * G.TW;
  const int n = blockIdx.z / G.nseg, seg = blockIdx.z - n * G.nseg;
  const int oh0 = seg * G.rs, oh1 = min(oh0 + G.rs, G.H);
  const int iw = ow0 - 1 + j;
  const bool col_ok = j < TWL && (unsigned)iw < (unsigned)G.W;
  const bool compute = j >= 1 && j <= G.TW && iw < G.W;

  float4 t4 = __ldg(reinterpret_cast<const float4*>(ss + c));
  const float2 scA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), scB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
  t4 = __ldg(reinterpret_cast<const float4*>(ss + G.C + c));
  const float2 shA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), shB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
  float2 wA[9], wB[9];  // 6 * filter, channels (0,1) and (2,3)
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wA[k] = make_float2(6.f * __ldg(w + (c + 0) * 9 + k), 6.f * __ldg(w + (c + 1) * 9 + k));
    wB[k] = make_float2(6.f * __ldg(w + (c + 2) * 9 + k), 6.f * __ldg(w + (c + 3) * 9 + k));
  }

  const size_t rowp = (size_t)G.W * G.C;
  const __nv_bfloat16* xp = x + ((size_t)n * G.H * G.W + iw) * G.C + c;   // + ih*rowp
  __nv_bfloat16* yp = y + ((size_t)n * G.H * G.W + iw) * G.C + c;          // output column == input column
  const int ih_first = oh0 - 1;
  const int total = oh1 - oh0 + 2;

  uint2 pf[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    const int ih = ih_first + u;
    pf[u] = (col_ok && u < total && (unsigned)ih < (unsigned)G.H) ? ldg8(xp + (size_t)ih * rowp) : make_uint2(0u, 0u);
  }
  float2 winA[3][3], winB[3][3];
  float2 sA = make_float2(0.f, 0.f), sB = sA, qA = sA, qB = sA;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) winA[a][b] = winB[a][b] = make_float2(0.f, 0.f);

  for (int r0 = 0; r0 < total; r0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int r = r0 + u;
      if (r < total) {   // uniform over the CTA
        const int ih = ih_first + r;
        const uint2 raw = pf[u];
        {
          const int ihn = ih + PF;
          pf[u] = (col_ok && r + PF < total && (unsigned)ihn < (unsigned)G.H) ? ldg8(xp + (size_t)ihn * rowp) : make_uint2(0u, 0u);
        }
        float2 xa, xb, aA, aB;
        unpack4(raw, xa, xb);
        aA.x = __saturatef(fmaf(xa.x, scA.x, shA.x));
        aA.y = __saturatef(fmaf(xa.y, scA.y, shA.y));
        aB.x = __saturatef(fmaf(xb.x, scB.x, shB.x));
        aB.y = __saturatef(fmaf(xb.y, scB.y, shB.y));
        if (!HALO) {
          if (!(col_ok && (unsigned)ih < (unsigned)G.H)) aA = aB = make_float2(0.f, 0.f);
        }
        ring[u & 1][threadIdx.x] = pack4(aA, aB);
        __syncthreads();
        const int sl = u % 3;
        winA[sl][1] = aA;
        winB[sl][1] = aB;
        if (compute) {
          unpack4(ring[u & 1][threadIdx.x - CG], winA[sl][0], winB[sl][0]);
          unpack4(ring[u & 1][threadIdx.x + CG], winA[sl][2], winB[sl][2]);
          if (r >= 2) {
            float2 accA = make_float2(0.f, 0.f), accB = accA;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int s = (u + 1 + ky) % 3;   // row r-2+ky
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                accA = ffma2(winA[s][kx], wA[ky * 3 + kx], accA);
                accB = ffma2(winB[s][kx], wB[ky * 3 + kx], accB);
              }
            }
            *reinterpret_cast<uint2*>(yp + (size_t)(ih - 1) * rowp) = pack4(accA, accB);
            sA = fadd2(sA, accA);
            sB = fadd2(sB, accB);
            qA = ffma2(accA, accA, qA);
            qB = ffma2(accB, accB, qB);
          }
        }
      }
    }
  }
  if (stats) {
    const float v[8] = {sA.x, sA.y, sB.x, sB.y, qA.x, qA.y, qB.x, qB.y};
    const float tot = column_reduce<8>(red, v, CG, TWL);
    const int t = threadIdx.x;
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      atomicAdd(&stats[(k >> 2) * G.C + (blockIdx.x * CG + gg) * 4 + (k & 3)], (double)tot);
    }
  }
}

// ------------------------------------------------------------------------------------ fused backward
// On the domain extended by `ext` (0 or 1) pixels per side, position (ih, iw) = (he - ext, we - ext):
//   a6   = sat(x*sc/6 + sh/6)              (x = 0 outside the image: the reference's padded border)
//   acc  = sum_{ky,kx} dy(ih+1-ky, iw+1-kx) w[ky][kx]
//   g    = (0 < a6 < 1) ? acc : 0                                   -> stored, bf16
//   bsums += [sum g, sum g*(x-mean)*invstd]                          (BN backward of the producer)
//   dw[ky][kx] += 6*a6 * dy(ih+1-ky, iw+1-kx)
__global__ void __launch_bounds__(NT_MAX, 1)
dw_s1_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                 const float* __restrict__ ss, const float* __restrict__ mi, const float* __restrict__ w,
                 __nv_bfloat16* __restrict__ gout, double* __restrict__ bsums, float* __restrict__ dw, S1Geom G) {
  __shared__ uint2 ring[2][NT_MAX];
  __shared__ float red[12 * (NT_MAX + 1)];
  const int CG = G.CG, TWL = G.TW + 2, ext = G.ext;
  const int He = G.H + 2 * ext, We = G.W + 2 * ext;
  const int g = threadIdx.x % CG, j = threadIdx.x / CG;
  const int c = (blockIdx.x * CG + g) * 4;
  const int e0 = blockIdx.y * G.TW;
  const int n = blockIdx.z / G.nseg, seg = blockIdx.z - n * G.nseg;
  const int he0 = seg * G.rs, he1 = min(he0 + G.rs, He);
  const int we = e0 - 1 + j, iw = we - ext;
  const bool col_ok = j < TWL && (unsigned)iw < (unsigned)G.W;
  const bool compute = j >= 1 && j <= G.TW && we < We;

  float4 t4 = __ldg(reinterpret_cast<const float4*>(ss + c));
  const float2 scA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), scB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
  t4 = __ldg(reinterpret_cast<const float4*>(ss + G.C + c));
  const float2 shA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), shB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
  float2 nmuA = make_float2(0.f, 0.f), nmuB = nmuA;
  if (mi) {
    t4 = __ldg(reinterpret_cast<const float4*>(mi + c));
    nmuA = make_float2(-t4.x, -t4.y);
    nmuB = make_float2(-t4.z, -t4.w);
  }
  float2 wA[9], wB[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wA[k] = make_float2(__ldg(w + (c + 0) * 9 + k), __ldg(w + (c + 1) * 9 + k));
    wB[k] = make_float2(__ldg(w + (c + 2) * 9 + k), __ldg(w + (c + 3) * 9 + k));
  }
  float2 dA[9], dB[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) dA[k] = dB[k] = make_float2(0.f, 0.f);

  const size_t rowp = (size_t)G.W * G.C;
  const __nv_bfloat16* dyp = dy + ((size_t)n * G.H * G.W + iw) * G.C + c;
  const __nv_bfloat16* xp = x + ((size_t)n * G.H * G.W + iw) * G.C + c;
  __nv_bfloat16* gp = gout + ((size_t)n * He * We + we) * G.C + c;
  const size_t growp = (size_t)We * G.C;
  const int ih_start = he0 - ext;          // image row of the first output row
  const int total = he1 - he0 + 2;         // step r loads dy row ih_start-1+r and x row ih_start-2+r

  uint2 pfd[PF], pfx[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    const int dr = ih_start - 1 + u, xr = ih_start - 2 + u;
    pfd[u] = (col_ok && u < total && (unsigned)dr < (unsigned)G.H) ? ldg8(dyp + (size_t)dr * rowp) : make_uint2(0u, 0u);
    pfx[u] = (col_ok && u < total && u >= 2 && (unsigned)xr < (unsigned)G.H) ? ldg8(xp + (size_t)xr * rowp) : make_uint2(0u, 0u);
  }
  float2 winA[3][3], winB[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) winA[a][b] = winB[a][b] = make_float2(0.f, 0.f);
  float2 sA = make_float2(0.f, 0.f), sB = sA, qA = sA, qB = sA;

  for (int r0 = 0; r0 < total; r0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int r = r0 + u;
      if (r < total) {
        const uint2 rawd = pfd[u], rawx = pfx[u];
        {
          const int dr = ih_start - 1 + r + PF, xr = dr - 1;
          const bool more = col_ok && r + PF < total;
          pfd[u] = (more && (unsigned)dr < (unsigned)G.H) ? ldg8(dyp + (size_t)dr * rowp) : make_uint2(0u, 0u);
          pfx[u] = (more && (unsigned)xr < (unsigned)G.H) ? ldg8(xp + (size_t)xr * rowp) : make_uint2(0u, 0u);
        }
        ring[u & 1][threadIdx.x] = rawd;
        __syncthreads();
        const int sl = u % 3;
        unpack4(rawd, winA[sl][1], winB[sl][1]);
        if (compute) {
          unpack4(ring[u & 1][threadIdx.x - CG], winA[sl][0], winB[sl][0]);   // column iw-1
          unpack4(ring[u & 1][threadIdx.x + CG], winA[sl][2], winB[sl][2]);   // column iw+1
          if (r >= 2) {
            float2 xa, xb, aA, aB;
            unpack4(rawx, xa, xb);
            aA.x = __saturatef(fmaf(xa.x, scA.x, shA.x));
            aA.y = __saturatef(fmaf(xa.y, scA.y, shA.y));
            aB.x = __saturatef(fmaf(xb.x, scB.x, shB.x));
            aB.y = __saturatef(fmaf(xb.y, scB.y, shB.y));
            float2 accA = make_float2(0.f, 0.f), accB = accA;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int s = (u + 3 - ky) % 3;   // dy row ih+1-ky was loaded at step r-ky
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float2 vA = winA[s][2 - kx], vB = winB[s][2 - kx];   // dy column iw+1-kx
                accA = ffma2(vA, wA[ky * 3 + kx], accA);
                accB = ffma2(vB, wB[ky * 3 + kx], accB);
                dA[ky * 3 + kx] = ffma2(vA, aA, dA[ky * 3 + kx]);
                dB[ky * 3 + kx] = ffma2(vB, aB, dB[ky * 3 + kx]);
              }
            }
            accA.x = (aA.x > 0.f && aA.x < 1.f) ? accA.x : 0.f;
            accA.y = (aA.y > 0.f && aA.y < 1.f) ? accA.y : 0.f;
            accB.x = (aB.x > 0.f && aB.x < 1.f) ? accB.x : 0.f;
            accB.y = (aB.y > 0.f && aB.y < 1.f) ? accB.y : 0.f;
            *reinterpret_cast<uint2*>(gp + (size_t)(he0 + r - 2) * growp) = pack4(accA, accB);
            sA = fadd2(sA, accA);
            sB = fadd2(sB, accB);
            qA = ffma2(accA, fadd2(xa, nmuA), qA);
            qB = ffma2(accB, fadd2(xb, nmuB), qB);
          }
        }
      }
    }
  }
  const int t = threadIdx.x;
  if (bsums) {
    const float v[8] = {sA.x, sA.y, sB.x, sB.y, qA.x, qA.y, qB.x, qB.y};
    const float tot = column_reduce<8>(red, v, CG, TWL);
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      const int ch = (blockIdx.x * CG + gg) * 4 + (k & 3);
      const float f = (k >> 2) ? __ldg(mi + G.C + ch) : 1.f;
      atomicAdd(&bsums[(k >> 2) * G.C + ch], (double)(tot * f));
    }
    __syncthreads();
  }
  if (dw) {
    // 36 values per thread, reduced in three rounds of 12
#pragma unroll
    for (int part = 0; part < 3; ++part) {
      float v[12];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        v[k * 4 + 0] = dA[part * 3 + k].x;
        v[k * 4 + 1] = dA[part * 3 + k].y;
        v[k * 4 + 2] = dB[part * 3 + k].x;
        v[k * 4 + 3] = dB[part * 3 + k].y;
      }
      const float tot = column_reduce<12>(red, v, CG, TWL);
      if (t < CG * 12) {
        const int gg = t / 12, k = t % 12;
        const int ch = (blockIdx.x * CG + gg) * 4 + (k & 3);
        atomicAdd(&dw[ch * 9 + part * 3 + (k >> 2)], 6.f * tot);
      }
      __syncthreads();
    }
  }
}

// chunking: CG 4-channel groups per CTA, TW output columns, so that (TW+2)*CG <= NT_MAX
inline bool s1_plan(int N, int H, int W, int C, int ext, S1Geom* G, dim3* grid, int* threads) {
  int CG;
  if (C % 32 == 0) CG = 8;
  else if (C % 48 == 0) CG = 12;
  else if (C % 16 == 0) CG = 4;
  else return false;
  const int He = H + 2 * ext, We = W + 2 * ext;
  int TW = NT_MAX / CG - 2;
  const int tiles = s2r_div_up(We, TW);
  TW = s2r_div_up(We, tiles);              // balance the column tiles
  const int chunks = C / (CG * 4);
  // rows per CTA: as long as possible (vertical halo = 2 rows per segment) while filling the GPU ~4x
  int rs = 64;
  while (rs > 8 && (long)chunks * tiles * N * s2r_div_up(He, rs) < 4L * s2r_sm_count()) rs >>= 1;
  const int nseg = s2r_div_up(He, rs);
  if ((long)N * nseg > 65535 || tiles > 65535) return false;
  G->N = N; G->H = H; G->W = W; G->C = C; G->CG = CG; G->TW = TW; G->rs = rs; G->nseg = nseg; G->ext = ext;
  *grid = dim3(chunks, tiles, N * nseg);
  *threads = ((TW + 2) * CG + 31) / 32 * 32;
  return true;
}

}  // namespace

// Internal entry points (called from dwconv.cu); return S2R_ERR_UNSUPPORTED when the shape is not covered.
int s2r_dw_s1_fwd(const void* x, const float* ss, int halo_const, const float* w, void* y, double* stats,
                  int N, int H, int W, int C, cudaStream_t stream) {
  S1Geom G;
  dim3 grid;
  int threads;
  if (!s1_plan(N, H, W, C, 0, &G, &grid, &threads)) return S2R_ERR_UNSUPPORTED;
  if (halo_const)
    dw_s1_fwd_kernel<true><<<grid, threads, 0, stream>>>((const __nv_bfloat16*)x, ss, w, (__nv_bfloat16*)y, stats, G);
  else
    dw_s1_fwd_kernel<false><<<grid, threads, 0, stream>>>((const __nv_bfloat16*)x, ss, w, (__nv_bfloat16*)y, stats, G);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

int s2r_dw_s1_bwd(const void* dy, const void* x, const float* ss, const float* mi, const float* w, int ext,
                  void* g, double* bsums, float* dw, int N, int H, int W, int C, cudaStream_t stream) {
  S1Geom G;
  dim3 grid;
  int threads;
  if (!s1_plan(N, H, W, C, ext, &G, &grid, &threads)) return S2R_ERR_UNSUPPORTED;
  dw_s1_bwd_kernel<<<grid, threads, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, ss, mi, w,
                                                 (__nv_bfloat16*)g, bsums, dw, G);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
