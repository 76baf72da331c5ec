// Streaming depthwise 3x3, stride 2 (padding 1, dilation 1), with the BN + ReLU6 prologue and the
// reference's padded-border semantics (modeling/backbone/mobilenet.py:26-68, the three stride-2
// InvertedResidual blocks).  Same machinery as dwconv_s1.cu (TMA row ring, one producer warp, packed
// FFMA2, FFMA.SAT prologue), on the four PARITY PLANES of the input
//     X_pq(i, j) = x(2i + p, 2j + q),   p, q in {0 = even, 1 = odd}
// each addressed through its own TMA map with doubled strides, so every tile is dense in shared memory:
//   out(o, j) = ee(o,j) w11 + eo(o,j-1) w10 + eo(o,j) w12 + oe(o-1,j) w01 + oe(o,j) w21
//             + oo(o-1,j-1) w00 + oo(o-1,j) w02 + oo(o,j-1) w20 + oo(o,j) w22
// Forward is input-stationary in the odd rows: plane row i finishes output row i (filter rows 1, 2) and
// starts output row i+1 (filter row 0), so every input element is loaded and activated once.
// Backward (fused data + weight gradient): a thread owns one dy position (i, j) and the four input positions
// (2i+p, 2j+q); it needs dy(i..i+1, j..j+1) -- dy row i is carried in registers.
#include "dw_common.cuh"
#include "bn_tail.cuh"

using namespace s2r_tma;
using namespace s2r_dw;

namespace {

constexpr int RB = 3;            // plane rows per TMA stage
constexpr int F_CONS = 256, F_STAGES = 4;
constexpr int B_CONS = 256, B_STAGES = 3;

struct S2Geom {
  int N, H, W, C;       // input tensor
  int Ho, Wo;           // output / dy
  int CG, TW, rs, nseg;
  int stage_bytes;
  int off[5];           // byte offsets of the tiles inside a stage
  int interior;         // backward: g is [N][H][W][C], border positions are not stored
};

struct S2Maps {
  CUtensorMap ee, eo, oe, oo, dy;
};

// ------------------------------------------------------------------------------------ forward
// OAFF (inference): y = relu6(acc*oss[c] + oss[C+c]), the BatchNorm + ReLU6 that follows the convolution with its
// running statistics (mobilenet.py:55-56), applied before the store; no statistics in that mode.
template <bool OAFF>
__global__ void __launch_bounds__(F_CONS + 32, 2)
dw_s2_fwd_kernel(const __grid_constant__ S2Maps M, const float* __restrict__ ss, const float* __restrict__ w,
                 __nv_bfloat16* __restrict__ y, double* __restrict__ stats, const S2Geom G, const BnTail in_bn,
                 const float* __restrict__ oss) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t bar_full[F_STAGES], bar_empty[F_STAGES];
  __shared__ float red[8 * (F_CONS + 1)];
  const int CG = G.CG, TW = G.TW;
  const int ncons = TW * CG;
  const int ncw = (ncons + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.y * TW;
  const int n = blockIdx.z / G.nseg, seg = blockIdx.z - n * G.nseg;
  const int o0 = seg * G.rs, rows = min(G.rs, G.Ho - o0);
  const int nst = (rows + 1 + RB - 1) / RB;   // step r handles plane row o0 - 1 + r

  if (threadIdx.x == 0) {
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (warp == ncw) {
    if (lane == 0) {
      prefetch_tmap(&M.ee); prefetch_tmap(&M.eo); prefetch_tmap(&M.oe); prefetch_tmap(&M.oo);
      const int c0 = blockIdx.x * CG * 4;
      for (int k = 0; k < nst; ++k) {
        const int s = k % F_STAGES;
        if (k >= F_STAGES) mbar_wait(smem_addr(&bar_empty[s]), ((k / F_STAGES) - 1) & 1);
        const uint32_t full = smem_addr(&bar_full[s]);
        mbar_expect_tx(full, (uint32_t)(RB * (4 * TW + 2) * CG * 8));
        unsigned char* st = smem + (size_t)s * G.stage_bytes;
        const int i0 = o0 - 1 + k * RB;
        tma_load_4d(smem_addr(st + G.off[0]), &M.ee, full, c0, j0, i0, n);
        tma_load_4d(smem_addr(st + G.off[1]), &M.eo, full, c0, j0 - 1, i0, n);
        tma_load_4d(smem_addr(st + G.off[2]), &M.oe, full, c0, j0, i0, n);
        tma_load_4d(smem_addr(st + G.off[3]), &M.oo, full, c0, j0 - 1, i0, n);
      }
    }
  } else if (warp < ncw) {
    const bool live = threadIdx.x < ncons;
    const int tid = live ? threadIdx.x : 0;
    const int g = tid % CG, j = tid / CG;
    const int c = (blockIdx.x * CG + g) * 4;
    const int ow = j0 + j;
    const bool active = live && ow < G.Wo;

    float4 t4, u4;
    if (in_bn.enabled) {
      // pending input BatchNorm: scale / shift from the producer's sums; the first CTA of every channel chunk
      // publishes them (bn_tail.cuh)
      float fsc[4], fsh[4];
      bn_fin4(in_bn, G.C, c, blockIdx.y == 0 && blockIdx.z == 0 && live && j == 0, fsc, fsh);
      t4 = make_float4(fsc[0], fsc[1], fsc[2], fsc[3]);
      u4 = make_float4(fsh[0], fsh[1], fsh[2], fsh[3]);
    } else {
      t4 = __ldg(reinterpret_cast<const float4*>(ss + c));
      u4 = __ldg(reinterpret_cast<const float4*>(ss + G.C + c));
    }
    const float2 scA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), scB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
    const float2 shA = make_float2(u4.x * (1.f / 6.f), u4.y * (1.f / 6.f)), shB = make_float2(u4.z * (1.f / 6.f), u4.w * (1.f / 6.f));
    float2 wA[9], wB[9];
    load_filter(w, c, 6.f, wA, wB);
    float2 oscA = make_float2(0.f, 0.f), oscB = oscA, oshA = oscA, oshB = oscA;   // OAFF: scale / 6, shift / 6
    if (OAFF) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(oss + c)), p4 = __ldg(reinterpret_cast<const float4*>(oss + G.C + c));
      oscA = make_float2(o4.x * (1.f / 6.f), o4.y * (1.f / 6.f)); oscB = make_float2(o4.z * (1.f / 6.f), o4.w * (1.f / 6.f));
      oshA = make_float2(p4.x * (1.f / 6.f), p4.y * (1.f / 6.f)); oshB = make_float2(p4.z * (1.f / 6.f), p4.w * (1.f / 6.f));
    }

    const size_t rowp = (size_t)G.Wo * G.C;
    __nv_bfloat16* yp = y + (((size_t)n * G.Ho + o0) * G.Wo + min(ow, G.Wo - 1)) * G.C + c - rowp;   // row o0-1+r at + r*rowp
    float2 curA = make_float2(0.f, 0.f), curB = curA;   // output row being finished (has its filter row 0 part)
    float2 sA = curA, sB = curA, qA = curA, qB = curA;

    for (int k = 0; k < nst; ++k) {
      const int s = k % F_STAGES;
      mbar_wait(smem_addr(&bar_full[s]), (k / F_STAGES) & 1);
      const unsigned char* st = smem + (size_t)s * G.stage_bytes;
      const uint2* tee = reinterpret_cast<const uint2*>(st + G.off[0]) + j * CG + g;
      const uint2* teo = reinterpret_cast<const uint2*>(st + G.off[1]) + j * CG + g;
      const uint2* toe = reinterpret_cast<const uint2*>(st + G.off[2]) + j * CG + g;
      const uint2* too = reinterpret_cast<const uint2*>(st + G.off[3]) + j * CG + g;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = k * RB + u;   // plane row i = o0 - 1 + r
        float2 eA, eB, lA, lB, rA, rB;
        // even row i: finishes output row i with filter row 1
        act4(tee[u * TW * CG], scA, scB, shA, shB, eA, eB);
        act4(teo[u * (TW + 1) * CG], scA, scB, shA, shB, lA, lB);
        act4(teo[u * (TW + 1) * CG + CG], scA, scB, shA, shB, rA, rB);
        curA = ffma2(rA, wA[5], ffma2(eA, wA[4], ffma2(lA, wA[3], curA)));
        curB = ffma2(rB, wB[5], ffma2(eB, wB[4], ffma2(lB, wB[3], curB)));
        // odd row i: filter row 2 of output row i, filter row 0 of output row i + 1
        act4(toe[u * TW * CG], scA, scB, shA, shB, eA, eB);
        act4(too[u * (TW + 1) * CG], scA, scB, shA, shB, lA, lB);
        act4(too[u * (TW + 1) * CG + CG], scA, scB, shA, shB, rA, rB);
        curA = ffma2(rA, wA[8], ffma2(eA, wA[7], ffma2(lA, wA[6], curA)));
        curB = ffma2(rB, wB[8], ffma2(eB, wB[7], ffma2(lB, wB[6], curB)));
        if (r >= 1 && r <= rows && active) {
          if (OAFF) {
            float2 oA, oB;
            oA.x = 6.f * __saturatef(fmaf(curA.x, oscA.x, oshA.x)); oA.y = 6.f * __saturatef(fmaf(curA.y, oscA.y, oshA.y));
            oB.x = 6.f * __saturatef(fmaf(curB.x, oscB.x, oshB.x)); oB.y = 6.f * __saturatef(fmaf(curB.y, oscB.y, oshB.y));
            *reinterpret_cast<uint2*>(yp + (size_t)r * rowp) = pack4(oA, oB);
          } else {
            *reinterpret_cast<uint2*>(yp + (size_t)r * rowp) = pack4(curA, curB);
            sA = fadd2(sA, curA);
            sB = fadd2(sB, curB);
            qA = ffma2(curA, curA, qA);
            qB = ffma2(curB, curB, qB);
          }
        }
        curA = ffma2(rA, wA[2], ffma2(eA, wA[1], fmul2(lA, wA[0])));
        curB = ffma2(rB, wB[2], ffma2(eB, wB[1], fmul2(lB, wB[0])));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&bar_empty[s]));
    }
    if (stats) {
      const float v[8] = {sA.x, sA.y, sB.x, sB.y, qA.x, qA.y, qB.x, qB.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) red[k * (F_CONS + 1) + threadIdx.x] = live ? v[k] : 0.f;
    }
  }
  if (stats) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      const float* p = red + k * (F_CONS + 1) + gg;
      float tot = 0.f;
      for (int jj = 0; jj < TW; ++jj) tot += p[jj * CG];
      atomicAdd(&stats[(k >> 2) * G.C + (blockIdx.x * CG + gg) * 4 + (k & 3)], (double)tot);
    }
  }
}

// ------------------------------------------------------------------------------------ fused backward
// Gradient domain extended by one pixel per side (the reference's padded border): ih, iw in [-1, H] x [-1, W],
// g is [N][H+2][W+2][C].  Thread (i, j), i in [-1, ..], j in [-1, ..] (dy = 0 outside [0,Ho) x [0,Wo)):
//   g_ee(2i,   2j  ) = dy(i,j) w11
//   g_eo(2i,   2j+1) = dy(i,j) w12 + dy(i,j+1) w10
//   g_oe(2i+1, 2j  ) = dy(i,j) w21 + dy(i+1,j) w01
//   g_oo(2i+1, 2j+1) = dy(i,j) w22 + dy(i,j+1) w20 + dy(i+1,j) w02 + dy(i+1,j+1) w00
// each masked by act'(pre) of its own position, and the nine weight-gradient taps pair the same operands.
__device__ __forceinline__ void s2_emit(float2 vA, float2 vB, float2 aA, float2 aB, uint2 xraw, bool ok, bool st, __nv_bfloat16* dst,
                                        float2 nmuA, float2 nmuB, float2& sA, float2& sB, float2& qA, float2& qB) {
  if (ok) {
    vA.x = (aA.x > 0.f && aA.x < 1.f) ? vA.x : 0.f;
    vA.y = (aA.y > 0.f && aA.y < 1.f) ? vA.y : 0.f;
    vB.x = (aB.x > 0.f && aB.x < 1.f) ? vB.x : 0.f;
    vB.y = (aB.y > 0.f && aB.y < 1.f) ? vB.y : 0.f;
    st8_if(dst, pack4(vA, vB), st);
    float2 xa, xb;
    unpack4(xraw, xa, xb);
    sA = fadd2(sA, vA);
    sB = fadd2(sB, vB);
    qA = ffma2(vA, fadd2(xa, nmuA), qA);
    qB = ffma2(vB, fadd2(xb, nmuB), qB);
  }
}

__global__ void __launch_bounds__(B_CONS + 32, 1)
dw_s2_bwd_kernel(const __grid_constant__ S2Maps M, const float* __restrict__ ss, const float* __restrict__ mi,
                 const float* __restrict__ w, __nv_bfloat16* __restrict__ gout, double* __restrict__ bsums,
                 float* __restrict__ dw, const S2Geom G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t bar_full[B_STAGES], bar_empty[B_STAGES];
  __shared__ float red[12 * (B_CONS + 1)];
  const int CG = G.CG, TW = G.TW;
  const int gsh = G.interior ? 0 : 1;               // g element of (ih, iw) sits at (ih + gsh, iw + gsh)
  const int He = G.H + 2 * gsh, We = G.W + 2 * gsh;
  const int ncons = TW * CG;
  const int ncw = (ncons + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.y * TW - 1;                // first dy column of this CTA (starts at -1)
  const int n = blockIdx.z / G.nseg, seg = blockIdx.z - n * G.nseg;
  const int nI = G.H / 2 + 2;                        // i = -1 .. floor(H/2): covers ih = -1 .. H
  const int i0 = seg * G.rs - 1, rows = min(G.rs, nI - seg * G.rs);
  const int nst = (rows + 1 + RB - 1) / RB;          // step r: new dy row i0 + r, x rows i0 + r - 1

  if (threadIdx.x == 0) {
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  float2 dA[9], dB[9];
  float2 sA = make_float2(0.f, 0.f), sB = sA, qA = sA, qB = sA;
#pragma unroll
  for (int k = 0; k < 9; ++k) dA[k] = dB[k] = make_float2(0.f, 0.f);
  bool live = false;

  if (warp == ncw) {
    if (lane == 0) {
      prefetch_tmap(&M.ee); prefetch_tmap(&M.eo); prefetch_tmap(&M.oe); prefetch_tmap(&M.oo); prefetch_tmap(&M.dy);
      const int c0 = blockIdx.x * CG * 4;
      for (int k = 0; k < nst; ++k) {
        const int s = k % B_STAGES;
        if (k >= B_STAGES) mbar_wait(smem_addr(&bar_empty[s]), ((k / B_STAGES) - 1) & 1);
        const uint32_t full = smem_addr(&bar_full[s]);
        mbar_expect_tx(full, (uint32_t)(RB * (5 * TW + 1) * CG * 8));
        unsigned char* st = smem + (size_t)s * G.stage_bytes;
        const int d0 = i0 + k * RB;
        tma_load_4d(smem_addr(st + G.off[4]), &M.dy, full, c0, j0, d0, n);
        tma_load_4d(smem_addr(st + G.off[0]), &M.ee, full, c0, j0, d0 - 1, n);
        tma_load_4d(smem_addr(st + G.off[1]), &M.eo, full, c0, j0, d0 - 1, n);
        tma_load_4d(smem_addr(st + G.off[2]), &M.oe, full, c0, j0, d0 - 1, n);
        tma_load_4d(smem_addr(st + G.off[3]), &M.oo, full, c0, j0, d0 - 1, n);
      }
    }
  } else if (warp < ncw) {
    live = threadIdx.x < ncons;
    const int tid = live ? threadIdx.x : 0;
    const int g = tid % CG, j = tid / CG;
    const int c = (blockIdx.x * CG + g) * 4;
    const int jj = j0 + j;                                  // dy column
    // column validity of the two input columns 2jj, 2jj+1 inside [-1, W]; cs_*: the position is stored
    const bool ce_ok = live && 2 * jj >= -1 && 2 * jj <= G.W;
    const bool co_ok = live && 2 * jj + 1 >= -1 && 2 * jj + 1 <= G.W;
    const bool cse = !G.interior || (unsigned)(2 * jj) < (unsigned)G.W;
    const bool cso = !G.interior || (unsigned)(2 * jj + 1) < (unsigned)G.W;

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
    load_filter(w, c, 1.f, wA, wB);

    const long long growp = (long long)We * G.C;
    // g element of (ih, iw) is at ((n*He + ih + 1)*We + iw + 1)*C; this thread's even column is iw = 2jj
    __nv_bfloat16* gp = gout + ((long long)n * He * We + (long long)(2 * jj + gsh)) * G.C + c;
    float2 p0A = make_float2(0.f, 0.f), p0B = p0A, p1A = p0A, p1B = p0A;   // dy(i, jj), dy(i, jj+1)

    for (int k = 0; k < nst; ++k) {
      const int s = k % B_STAGES;
      mbar_wait(smem_addr(&bar_full[s]), (k / B_STAGES) & 1);
      const unsigned char* st = smem + (size_t)s * G.stage_bytes;
      const uint2* tdy = reinterpret_cast<const uint2*>(st + G.off[4]) + j * CG + g;
      const uint2* tee = reinterpret_cast<const uint2*>(st + G.off[0]) + j * CG + g;
      const uint2* teo = reinterpret_cast<const uint2*>(st + G.off[1]) + j * CG + g;
      const uint2* toe = reinterpret_cast<const uint2*>(st + G.off[2]) + j * CG + g;
      const uint2* too = reinterpret_cast<const uint2*>(st + G.off[3]) + j * CG + g;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = k * RB + u;
        float2 n0A, n0B, n1A, n1B;             // dy(i+1, jj), dy(i+1, jj+1)
        unpack4(tdy[u * (TW + 1) * CG], n0A, n0B);
        unpack4(tdy[u * (TW + 1) * CG + CG], n1A, n1B);
        const int i = i0 + r - 1;              // x rows / g rows of this step: ih = 2i, 2i+1
        const bool row_step = live && r >= 1 && r <= rows;
        const bool re_ok = row_step && 2 * i >= -1 && 2 * i <= G.H;
        const bool ro_ok = row_step && 2 * i + 1 >= -1 && 2 * i + 1 <= G.H;
        __nv_bfloat16* grow = gp + (long long)(2 * i + gsh) * growp;   // row ih = 2i
        const bool rse = !G.interior || (unsigned)(2 * i) < (unsigned)G.H;
        const bool rso = !G.interior || (unsigned)(2 * i + 1) < (unsigned)G.H;
        uint2 xr;
        float2 aA, aB, vA, vB;
        // ---- ee: (2i, 2jj)
        xr = tee[u * TW * CG];
        act4(xr, scA, scB, shA, shB, aA, aB);
        if (!row_step) aA = aB = make_float2(0.f, 0.f);
        vA = fmul2(p0A, wA[4]); vB = fmul2(p0B, wB[4]);
        dA[4] = ffma2(p0A, aA, dA[4]); dB[4] = ffma2(p0B, aB, dB[4]);
        s2_emit(vA, vB, aA, aB, xr, re_ok && ce_ok, rse && cse, grow, nmuA, nmuB, sA, sB, qA, qB);
        // ---- eo: (2i, 2jj+1)
        xr = teo[u * TW * CG];
        act4(xr, scA, scB, shA, shB, aA, aB);
        if (!row_step) aA = aB = make_float2(0.f, 0.f);
        vA = ffma2(p1A, wA[3], fmul2(p0A, wA[5])); vB = ffma2(p1B, wB[3], fmul2(p0B, wB[5]));
        dA[5] = ffma2(p0A, aA, dA[5]); dB[5] = ffma2(p0B, aB, dB[5]);
        dA[3] = ffma2(p1A, aA, dA[3]); dB[3] = ffma2(p1B, aB, dB[3]);
        s2_emit(vA, vB, aA, aB, xr, re_ok && co_ok, rse && cso, grow + G.C, nmuA, nmuB, sA, sB, qA, qB);
        // ---- oe: (2i+1, 2jj)
        xr = toe[u * TW * CG];
        act4(xr, scA, scB, shA, shB, aA, aB);
        if (!row_step) aA = aB = make_float2(0.f, 0.f);
        vA = ffma2(n0A, wA[1], fmul2(p0A, wA[7])); vB = ffma2(n0B, wB[1], fmul2(p0B, wB[7]));
        dA[7] = ffma2(p0A, aA, dA[7]); dB[7] = ffma2(p0B, aB, dB[7]);
        dA[1] = ffma2(n0A, aA, dA[1]); dB[1] = ffma2(n0B, aB, dB[1]);
        s2_emit(vA, vB, aA, aB, xr, ro_ok && ce_ok, rso && cse, grow + growp, nmuA, nmuB, sA, sB, qA, qB);
        // ---- oo: (2i+1, 2jj+1)
        xr = too[u * TW * CG];
        act4(xr, scA, scB, shA, shB, aA, aB);
        if (!row_step) aA = aB = make_float2(0.f, 0.f);
        vA = ffma2(n1A, wA[0], ffma2(n0A, wA[2], ffma2(p1A, wA[6], fmul2(p0A, wA[8]))));
        vB = ffma2(n1B, wB[0], ffma2(n0B, wB[2], ffma2(p1B, wB[6], fmul2(p0B, wB[8]))));
        dA[8] = ffma2(p0A, aA, dA[8]); dB[8] = ffma2(p0B, aB, dB[8]);
        dA[6] = ffma2(p1A, aA, dA[6]); dB[6] = ffma2(p1B, aB, dB[6]);
        dA[2] = ffma2(n0A, aA, dA[2]); dB[2] = ffma2(n0B, aB, dB[2]);
        dA[0] = ffma2(n1A, aA, dA[0]); dB[0] = ffma2(n1B, aB, dB[0]);
        s2_emit(vA, vB, aA, aB, xr, ro_ok && co_ok, rso && cso, grow + growp + G.C, nmuA, nmuB, sA, sB, qA, qB);
        p0A = n0A; p0B = n0B; p1A = n1A; p1B = n1B;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&bar_empty[s]));
    }
  }
  const int t = threadIdx.x;
  const bool consumer = t < B_CONS;
  if (bsums) {
    const float v[8] = {live ? sA.x : 0.f, live ? sA.y : 0.f, live ? sB.x : 0.f, live ? sB.y : 0.f,
                        live ? qA.x : 0.f, live ? qA.y : 0.f, live ? qB.x : 0.f, live ? qB.y : 0.f};
    const float tot = column_reduce<8, B_CONS>(red, v, CG, TW, consumer);
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      const int ch = (blockIdx.x * CG + gg) * 4 + (k & 3);
      const float f = (k >> 2) ? __ldg(mi + G.C + ch) : 1.f;
      atomicAdd(&bsums[(k >> 2) * G.C + ch], (double)(tot * f));
    }
    __syncthreads();
  }
  if (dw) {
#pragma unroll
    for (int part = 0; part < 3; ++part) {
      float v[12];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        v[k * 4 + 0] = live ? dA[part * 3 + k].x : 0.f;
        v[k * 4 + 1] = live ? dA[part * 3 + k].y : 0.f;
        v[k * 4 + 2] = live ? dB[part * 3 + k].x : 0.f;
        v[k * 4 + 3] = live ? dB[part * 3 + k].y : 0.f;
      }
      const float tot = column_reduce<12, B_CONS>(red, v, CG, TW, consumer);
      if (t < CG * 12) {
        const int gg = t / 12, k = t % 12;
        const int ch = (blockIdx.x * CG + gg) * 4 + (k & 3);
        atomicAdd(&dw[ch * 9 + part * 3 + (k >> 2)], 6.f * tot);
      }
      __syncthreads();
    }
  }
}

constexpr int S2_SMEM_CAP = 200 * 1024;

inline int pick_cg(int C) {
  if (C % 32 == 0) return 8;
  if (C % 48 == 0) return 12;
  if (C % 16 == 0) return 4;
  return 0;
}

// the four parity planes of x [N][H][W][C]; plane (p, q) has ceil((H-p)/2) x ceil((W-q)/2) pixels
inline bool s2_maps(S2Maps* M, const void* x, int N, int H, int W, int C, int CG, const int bw[4]) {
  const long long sw = 2LL * C, sh = 2LL * W * C, sn = (long long)H * W * C;
  CUtensorMap* m[4] = {&M->ee, &M->eo, &M->oe, &M->oo};
  for (int p = 0; p < 2; ++p)
    for (int q = 0; q < 2; ++q) {
      const int Hp = (H - p + 1) / 2, Wq = (W - q + 1) / 2;
      if (Hp <= 0 || Wq <= 0) return false;
      if (!encode_nhwc_view(m[p * 2 + q], (const __nv_bfloat16*)x + ((long long)p * W + q) * C, N, Hp, Wq, C, sw, sh, sn,
                            CG * 4, bw[p * 2 + q], RB))
        return false;
    }
  return true;
}

inline void s2_offsets(S2Geom* G, const int bw[5], int ntiles) {
  int o = 0;
  for (int t = 0; t < 5; ++t) {
    G->off[t] = o;
    if (t < ntiles) o += (RB * bw[t] * G->CG * 8 + 127) / 128 * 128;
  }
  G->stage_bytes = o;
}

}  // namespace

// Rows per CTA for a grid of `per_row_seg` CTAs per row segment on `slots` resident CTA slots: the kernel runs in
// ceil(CTAs / slots) waves of (rs + 1 halo + ~3 rows of prologue / reduction) row steps each; rs + 1 is a multiple of
// the stage depth RB so that no loaded row is wasted.  (The first version doubled the segment count until the grid had
// four CTAs per SM, which leaves e.g. 1.3 waves of work on 2 waves of time.)
static int s2_pick_rows(int rows_total, long per_row_seg, long slots) {
  long best = -1;
  int best_rs = RB - 1;
  for (int rs = RB + RB - 1; ; rs += RB) {       // 5, 8, 11, ...
    const long nsg = (rows_total + rs - 1) / rs;
    const long waves = (per_row_seg * nsg + slots - 1) / slots;
    const long cost = waves * (rs + 4);
    if (best < 0 || cost < best || (cost == best && rs > best_rs)) { best = cost; best_rs = rs; }
    if (rs >= rows_total) break;
  }
  return best_rs;
}

int s2r_dw_s2_fwd(const void* x, const float* ss, const s2r_bn_tail* in_bn, const float* w, void* y, double* stats,
                  const float* oss, int N, int H, int W, int C, cudaStream_t stream) {
  if (oss && (stats || (uintptr_t)oss % 16)) return S2R_ERR_UNSUPPORTED;   // output affine: inference only
  S2Geom G;
  G.N = N; G.H = H; G.W = W; G.C = C;
  G.Ho = (H - 1) / 2 + 1; G.Wo = (W - 1) / 2 + 1;
  G.CG = pick_cg(C);
  if (!G.CG || H < 2 || W < 2) return S2R_ERR_UNSUPPORTED;
  int TW = F_CONS / G.CG;
  const int tiles = s2r_div_up(G.Wo, TW);
  TW = s2r_div_up(G.Wo, tiles);
  G.TW = TW;
  const int chunks = C / (G.CG * 4);
  const int rs = s2_pick_rows(G.Ho, (long)chunks * tiles * N, 2L * s2r_sm_count());
  G.rs = rs; G.nseg = s2r_div_up(G.Ho, rs);
  if ((long)N * G.nseg > 65535) return S2R_ERR_UNSUPPORTED;
  const int bw[5] = {TW, TW + 1, TW, TW + 1, 0};
  s2_offsets(&G, bw, 4);
  S2Maps M;
  if (!s2_maps(&M, x, N, H, W, C, G.CG, bw)) return S2R_ERR_UNSUPPORTED;
  M.dy = M.ee;
  const size_t smem = (size_t)F_STAGES * G.stage_bytes + 128;
  if (smem > (size_t)S2_SMEM_CAP) return S2R_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(dw_s2_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM_CAP));
    S2R_CUDA_OK(cudaFuncSetAttribute(dw_s2_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM_CAP));
    attr = true;
  }
  int threads = (TW * G.CG + 31) / 32 * 32 + 32;
  if (threads < (G.CG * 12 + 31) / 32 * 32) threads = (G.CG * 12 + 31) / 32 * 32;
  const BnTail bt = bn_tail_from(in_bn);
  if (oss)
    S2R_CUDA_OK(s2r_launch(dw_s2_fwd_kernel<true>, dim3(chunks, tiles, N * G.nseg), dim3(threads), smem, stream, M, ss, w,
                           (__nv_bfloat16*)y, stats, G, bt, oss));
  else
    S2R_CUDA_OK(s2r_launch(dw_s2_fwd_kernel<false>, dim3(chunks, tiles, N * G.nseg), dim3(threads), smem, stream, M, ss, w,
                           (__nv_bfloat16*)y, stats, G, bt, oss));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

int s2r_dw_s2_bwd(const void* dy, const void* x, const float* ss, const float* mi, const float* w, int interior,
                  void* g, double* bsums, float* dw, int N, int H, int W, int C, cudaStream_t stream) {
  S2Geom G;
  G.interior = interior;
  G.N = N; G.H = H; G.W = W; G.C = C;
  G.Ho = (H - 1) / 2 + 1; G.Wo = (W - 1) / 2 + 1;
  G.CG = pick_cg(C);
  if (!G.CG || H < 2 || W < 2) return S2R_ERR_UNSUPPORTED;
  const int nJ = W / 2 + 2, nI = H / 2 + 2;   // dy positions -1 .. floor(W/2)
  int TW = B_CONS / G.CG;
  const int tiles = s2r_div_up(nJ, TW);
  TW = s2r_div_up(nJ, tiles);
  G.TW = TW;
  const int chunks = C / (G.CG * 4);
  const int rs = s2_pick_rows(nI, (long)chunks * tiles * N, 1L * s2r_sm_count());
  G.rs = rs; G.nseg = s2r_div_up(nI, rs);
  if ((long)N * G.nseg > 65535) return S2R_ERR_UNSUPPORTED;
  const int bw[5] = {TW, TW, TW, TW, TW + 1};
  s2_offsets(&G, bw, 5);
  S2Maps M;
  if (!s2_maps(&M, x, N, H, W, C, G.CG, bw)) return S2R_ERR_UNSUPPORTED;
  if (!encode_nhwc(&M.dy, dy, N, G.Ho, G.Wo, C, G.CG * 4, TW + 1, RB)) return S2R_ERR_UNSUPPORTED;
  const size_t smem = (size_t)B_STAGES * G.stage_bytes + 128;
  if (smem > (size_t)S2_SMEM_CAP) return S2R_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(dw_s2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM_CAP));
    attr = true;
  }
  int threads = (TW * G.CG + 31) / 32 * 32 + 32;
  if (threads < (G.CG * 12 + 31) / 32 * 32) threads = (G.CG * 12 + 31) / 32 * 32;
  S2R_CUDA_OK(s2r_launch(dw_s2_bwd_kernel, dim3(chunks, tiles, N * G.nseg), dim3(threads), smem, stream, M, ss, mi, w, (__nv_bfloat16*)g, bsums, dw, G));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
