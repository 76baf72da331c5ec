// Streaming depthwise 3x3 (stride 1, dilation 1, padding 1) with the BN + ReLU6 prologue: the
// forward kernel and a fused data-gradient + weight-gradient kernel.  Covers 14 of the 17
// InvertedResidual blocks of the reference (modeling/backbone/mobilenet.py:26-68); dwconv.cu keeps
// the generic (stride 2 / dilated) kernels and dispatches here.
//
// Why a second design: at the HBM roofline (2 B in + 2 B out per element, 23 B/clk/SM) an SM has
// ~0.17 clk per element, i.e. ~22 issue slots per warp-row of 128 elements -- a 3x3 depthwise conv
// with its BN prologue and its statistics is almost issue-bound on B200.  So:
//   * a CTA owns (one chunk of CG*4 channels) x (TW columns) and walks DOWN `rs` rows;
//   * input rows arrive through a TMA ring (3 rows x (TW+2) columns x chunk per stage, one producer
//     warp, full/empty mbarriers): no address arithmetic, no load instructions, no __syncthreads in
//     the row loop, and TMA's zero fill IS the reference's zero padding of the block input
//     (relu6(0*sc+sh) = relu6(sh) is exactly what the reference's BN+ReLU6 produce on its padded border);
//   * relu6(x*sc+sh) = 6*sat(x*sc/6+sh/6): ONE FFMA.SAT per element, the 6 folded into the filter;
//   * vertical reuse is input-stationary: each activated row (left, centre, right vector) is used for the
//     three output rows it touches, whose accumulators live in registers -- no 3x3 window copy;
//   * all multiply-adds are packed FFMA2 (fma.rn.f32x2, two channels per instruction);
//   * fused backward: g = act'(pre) * sum_k dy(pos+1-k) w[k] and dw[k] += a(pos) dy(pos+1-k) share
//     every dy operand, so dy and x are read once (6 B/element instead of 10 for two kernels).
// thread = (column j, 4 consecutive channels); lane order is channel-fastest, so a warp stores
// contiguous CG*8-byte pixel segments.
#include "dw_common.cuh"
#include "bn_tail.cuh"

using namespace s2r_tma;
using namespace s2r_dw;

namespace {

constexpr int RB = 3;            // rows per TMA stage (= accumulator rotation period)
constexpr int FWD_CONS = 256;    // forward: consumer threads (+ one producer warp)
constexpr int FWD_STAGES = 4;
constexpr int FWD2_CONS = 224;   // two-column forward: 7 consumer warps + the producer = 256 threads -> 128 registers at 2 CTAs / SM
constexpr int BWD_CONS = 352;    // backward: consumer threads (+ one producer warp)
constexpr int BWD_STAGES = 4;

struct S1Geom {
  int N, H, W, C;   // tensor
  int CG;           // 4-channel groups per CTA chunk (chunk = CG*4 channels)
  int TW;           // columns per CTA
  int rs;           // rows per CTA
  int nseg;         // row segments per image
  int tiles;        // column tiles
  int nunits;       // N * nseg * tiles work units; a CTA (blockIdx.y) takes units blockIdx.y, +gridDim.y, ...
  int ext;          // backward: gradient domain extension (0 or 1)
  int stage_bytes;  // bytes of one ring stage (padded to 128)
  int xoff;         // backward: byte offset of the x tile inside a stage
  // output (y / g) addressing in elements: the tensor may be a strided view (dilation = parity planes)
  long long os_pix, os_row, os_img;
  int interior;     // backward: g has the unextended layout, border positions are not stored
  int publish;      // forward with a pending input BatchNorm: this launch publishes it (one launch of a dilated set)
};

// 8-byte shared-memory load at a 32-bit shared address plus a compile-time byte offset: the three neighbours of
// a row are one address register and three immediates (CG is a template parameter)
template <int OFF>
__device__ __forceinline__ uint2 lds8(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
  return v;
}

// relu6(v*sc + sh) with sc6 = sc/6, sh6 = sh/6: 6 * sat(v*sc6 + sh6)
__device__ __forceinline__ float2 oaff2(float2 v, float2 sc6, float2 sh6) {
  return make_float2(6.f * __saturatef(fmaf(v.x, sc6.x, sh6.x)), 6.f * __saturatef(fmaf(v.y, sc6.y, sh6.y)));
}

// ------------------------------------------------------------------------------------ forward
// y[oh][ow] = sum_{ky,kx} a(oh+ky-1, ow+kx-1) w[ky][kx],  a = relu6(x*sc+sh) inside the image and
// HALO ? relu6(sh) : 0 outside.  stats += per-channel sum / sum of squares of y.
// OAFF (inference): the BatchNorm + ReLU6 that FOLLOWS the convolution (mobilenet.py:41-42,55-56) with its running
// statistics is applied to the accumulators before the store, y = relu6(acc*oss[c] + oss[C+c]) -- no bn_apply pass
// over the output; no statistics in that mode.
template <bool HALO, int CG, bool OAFF>
__global__ void __launch_bounds__(FWD_CONS + 32, 2)
dw_s1_fwd_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ ss, const float* __restrict__ w,
                 __nv_bfloat16* __restrict__ y, double* __restrict__ stats, const S1Geom G, const BnTail in_bn,
                 const float* __restrict__ oss) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t bar_full[FWD_STAGES], bar_empty[FWD_STAGES];
  __shared__ float red[8 * (FWD_CONS + 1)];
  const int TWL = G.TW + 2;
  const int ncons = G.TW * CG;                       // active consumer threads
  const int ncw = (ncons + 31) / 32;                 // consumer warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent over work units (image, row segment, column tile) of one channel chunk: the filter, the channel
  // constants and the statistics accumulators stay in registers, the TMA ring keeps running across units, and the
  // prologue / block reduction are paid once per CTA instead of once per tile.
#define S1_UNIT(unit_)                                                   \
  const int tile_ = (unit_) % G.tiles, rest_ = (unit_) / G.tiles;        \
  const int seg_ = rest_ % G.nseg, n = rest_ / G.nseg;                   \
  const int ow0 = tile_ * G.TW, oh0 = seg_ * G.rs;                       \
  const int rows = min(G.rs, G.H - oh0);                                 \
  const int nst = (rows + 2 + RB - 1) / RB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < FWD_STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (warp == ncw) {
    // ---------------- producer warp
    if (lane == 0) {
      prefetch_tmap(&xmap);
      int it = 0;
      for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
        S1_UNIT(unit)
        for (int k = 0; k < nst; ++k, ++it) {
          const int s = it % FWD_STAGES;
          if (it >= FWD_STAGES) mbar_wait(smem_addr(&bar_empty[s]), ((it / FWD_STAGES) - 1) & 1);
          const uint32_t full = smem_addr(&bar_full[s]);
          mbar_expect_tx(full, (uint32_t)(RB * TWL * CG * 8));
          tma_load_4d(smem_addr(smem + (size_t)s * G.stage_bytes), &xmap, full, blockIdx.x * CG * 4, ow0 - 1,
                      oh0 - 1 + k * RB, n);
        }
      }
    }
  } else if (warp < ncw) {
    // ---------------- consumers
    const bool live = threadIdx.x < ncons;
    const int tid = live ? threadIdx.x : 0;
    const int g = tid % CG, j = tid / CG;
    const int c = (blockIdx.x * CG + g) * 4;

    float4 t4, u4;
    if (in_bn.enabled) {
      // the input's BatchNorm is still pending: scale / shift from the producer's sums; the first CTA of every
      // channel chunk publishes them (and mean / inv-std, running statistics) for the backward pass
      float fsc[4], fsh[4];
      bn_fin4(in_bn, G.C, c, blockIdx.y == 0 && G.publish && live && j == 0, fsc, fsh);
      t4 = make_float4(fsc[0], fsc[1], fsc[2], fsc[3]);
      u4 = make_float4(fsh[0], fsh[1], fsh[2], fsh[3]);
    } else {
      t4 = __ldg(reinterpret_cast<const float4*>(ss + c));
      u4 = __ldg(reinterpret_cast<const float4*>(ss + G.C + c));
    }
    const float2 scA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), scB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
    const float2 shA = make_float2(u4.x * (1.f / 6.f), u4.y * (1.f / 6.f)), shB = make_float2(u4.z * (1.f / 6.f), u4.w * (1.f / 6.f));
    float2 wA[9], wB[9];   // 6 * filter, channels (0,1) and (2,3)
    load_filter(w, c, 6.f, wA, wB);
    float2 oscA = make_float2(0.f, 0.f), oscB = oscA, oshA = oscA, oshB = oscA;   // OAFF: scale / 6, shift / 6
    if (OAFF) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(oss + c)), p4 = __ldg(reinterpret_cast<const float4*>(oss + G.C + c));
      oscA = make_float2(o4.x * (1.f / 6.f), o4.y * (1.f / 6.f)); oscB = make_float2(o4.z * (1.f / 6.f), o4.w * (1.f / 6.f));
      oshA = make_float2(p4.x * (1.f / 6.f), p4.y * (1.f / 6.f)); oshB = make_float2(p4.z * (1.f / 6.f), p4.w * (1.f / 6.f));
    }

    const long long rowp = G.os_row;
    const uint32_t tile0 = smem_addr(smem) + (uint32_t)(j * CG + g) * 8u;   // this thread's left neighbour in row 0 of stage 0
    const uint32_t rowstride = (uint32_t)(TWL * CG) * 8u;
    float2 sA = make_float2(0.f, 0.f), sB = sA, qA = sA, qB = sA;
    int it = 0;
    for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
    S1_UNIT(unit)
    const int ow = ow0 + j;
    const bool active = live && ow < G.W;
    // running output pointer: row o = r - 2 of step r
    __nv_bfloat16* yrow = y + (long long)n * G.os_img + (long long)(oh0 - 2) * G.os_row + (long long)min(ow, G.W - 1) * G.os_pix + c;
    const bool lok = ow - 1 >= 0, rok = ow + 1 < G.W;   // !HALO: zero (not relu6(shift)) outside the image
    float2 accA[3], accB[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) accA[i] = accB[i] = make_float2(0.f, 0.f);

    for (int k = 0; k < nst; ++k, ++it) {
      const int s = it % FWD_STAGES;
      mbar_wait(smem_addr(&bar_full[s]), (it / FWD_STAGES) & 1);
      const uint32_t tile = tile0 + (uint32_t)s * (uint32_t)G.stage_bytes;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = k * RB + u;            // input row oh0 - 1 + r
        const uint32_t rowt = tile + (uint32_t)u * rowstride;
        float2 lA, lB, cA, cB, rA, rB;
        act4(lds8<0>(rowt), scA, scB, shA, shB, lA, lB);
        act4(lds8<CG * 8>(rowt), scA, scB, shA, shB, cA, cB);
        act4(lds8<CG * 16>(rowt), scA, scB, shA, shB, rA, rB);
        if (!HALO) {
          const bool row_ok = (unsigned)(oh0 - 1 + r) < (unsigned)G.H;
          if (!(row_ok && lok)) lA = lB = make_float2(0.f, 0.f);
          if (!row_ok) cA = cB = make_float2(0.f, 0.f);
          if (!(row_ok && rok)) rA = rB = make_float2(0.f, 0.f);
        }
        // output row o = r - ky gets ky's filter row; slot of o is o % 3 (u == r % 3)
        const int s0 = u, s1 = (u + 2) % 3, s2 = (u + 1) % 3;
        accA[s0] = ffma2(rA, wA[2], ffma2(cA, wA[1], fmul2(lA, wA[0])));
        accB[s0] = ffma2(rB, wB[2], ffma2(cB, wB[1], fmul2(lB, wB[0])));
        accA[s1] = ffma2(rA, wA[5], ffma2(cA, wA[4], ffma2(lA, wA[3], accA[s1])));
        accB[s1] = ffma2(rB, wB[5], ffma2(cB, wB[4], ffma2(lB, wB[3], accB[s1])));
        accA[s2] = ffma2(rA, wA[8], ffma2(cA, wA[7], ffma2(lA, wA[6], accA[s2])));
        accB[s2] = ffma2(rB, wB[8], ffma2(cB, wB[7], ffma2(lB, wB[6], accB[s2])));
        if ((unsigned)(r - 2) < (unsigned)rows && active) {
          if (OAFF) {
            *reinterpret_cast<uint2*>(yrow) = pack4(oaff2(accA[s2], oscA, oshA), oaff2(accB[s2], oscB, oshB));
          } else {
            *reinterpret_cast<uint2*>(yrow) = pack4(accA[s2], accB[s2]);
            sA = fadd2(sA, accA[s2]);
            sB = fadd2(sB, accB[s2]);
            qA = ffma2(accA[s2], accA[s2], qA);
            qB = ffma2(accB[s2], accB[s2], qB);
          }
        }
        yrow += rowp;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&bar_empty[s]));
    }
    }   // units
    if (stats) {
      const float v[8] = {sA.x, sA.y, sB.x, sB.y, qA.x, qA.y, qB.x, qB.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) red[k * (FWD_CONS + 1) + threadIdx.x] = live ? v[k] : 0.f;
    }
  }
  if (stats) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      const float* p = red + k * (FWD_CONS + 1) + gg;
      float tot = 0.f;
      for (int jj = 0; jj < G.TW; ++jj) tot += p[jj * CG];
      atomicAdd(&stats[(k >> 2) * G.C + (blockIdx.x * CG + gg) * 4 + (k & 3)], (double)tot);
    }
  }
}

// ------------------------------------------------------------------------------------ forward, two columns per thread
// The same computation with thread = (column pair 2j, 2j+1; 4 channels).  The activation a = relu6(x*sc+sh) is what the
// one-column kernel spends most of its non-FMA issue slots on -- every element is unpacked and activated three times, as
// the left, centre and right neighbour of three threads (3 x 8 of ~58 instructions per output vector).  A column pair
// needs four activated vectors for two outputs (2 x 8 per output), one shared-memory load per output less, and the row
// loop's fixed cost (barrier wait, pointer updates) is spread over twice the work: ~43 instead of ~58 instructions per
// output vector, with twelve independent accumulator chains per thread instead of six.
template <bool HALO, int CG, bool OAFF>
__global__ void __launch_bounds__(FWD2_CONS + 32, 2)
dw_s1_fwd2_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ ss, const float* __restrict__ w,
                  __nv_bfloat16* __restrict__ y, double* __restrict__ stats, const S1Geom G, const BnTail in_bn,
                  const float* __restrict__ oss) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t bar_full[FWD_STAGES], bar_empty[FWD_STAGES];
  __shared__ float red[8 * (FWD2_CONS + 1)];
  const int TWL = G.TW + 2;
  const int npair = G.TW >> 1;                       // G.TW is even
  const int ncons = npair * CG;                      // active consumer threads
  const int ncw = (ncons + 31) / 32;                 // consumer warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < FWD_STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (warp == ncw) {
    // ---------------- producer warp
    if (lane == 0) {
      prefetch_tmap(&xmap);
      int it = 0;
      for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
        S1_UNIT(unit)
        for (int k = 0; k < nst; ++k, ++it) {
          const int s = it % FWD_STAGES;
          if (it >= FWD_STAGES) mbar_wait(smem_addr(&bar_empty[s]), ((it / FWD_STAGES) - 1) & 1);
          const uint32_t full = smem_addr(&bar_full[s]);
          mbar_expect_tx(full, (uint32_t)(RB * TWL * CG * 8));
          tma_load_4d(smem_addr(smem + (size_t)s * G.stage_bytes), &xmap, full, blockIdx.x * CG * 4, ow0 - 1,
                      oh0 - 1 + k * RB, n);
        }
      }
    }
  } else if (warp < ncw) {
    // ---------------- consumers
    const bool live = threadIdx.x < ncons;
    const int tid = live ? threadIdx.x : 0;
    const int g = tid % CG, j = tid / CG;
    const int c = (blockIdx.x * CG + g) * 4;

    float4 t4, u4;
    if (in_bn.enabled) {
      float fsc[4], fsh[4];
      bn_fin4(in_bn, G.C, c, blockIdx.y == 0 && G.publish && live && j == 0, fsc, fsh);
      t4 = make_float4(fsc[0], fsc[1], fsc[2], fsc[3]);
      u4 = make_float4(fsh[0], fsh[1], fsh[2], fsh[3]);
    } else {
      t4 = __ldg(reinterpret_cast<const float4*>(ss + c));
      u4 = __ldg(reinterpret_cast<const float4*>(ss + G.C + c));
    }
    const float2 scA = make_float2(t4.x * (1.f / 6.f), t4.y * (1.f / 6.f)), scB = make_float2(t4.z * (1.f / 6.f), t4.w * (1.f / 6.f));
    const float2 shA = make_float2(u4.x * (1.f / 6.f), u4.y * (1.f / 6.f)), shB = make_float2(u4.z * (1.f / 6.f), u4.w * (1.f / 6.f));
    float2 wA[9], wB[9];   // 6 * filter, channels (0,1) and (2,3)
    load_filter(w, c, 6.f, wA, wB);
    float2 oscA = make_float2(0.f, 0.f), oscB = oscA, oshA = oscA, oshB = oscA;   // OAFF: scale / 6, shift / 6
    if (OAFF) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(oss + c)), p4 = __ldg(reinterpret_cast<const float4*>(oss + G.C + c));
      oscA = make_float2(o4.x * (1.f / 6.f), o4.y * (1.f / 6.f)); oscB = make_float2(o4.z * (1.f / 6.f), o4.w * (1.f / 6.f));
      oshA = make_float2(p4.x * (1.f / 6.f), p4.y * (1.f / 6.f)); oshB = make_float2(p4.z * (1.f / 6.f), p4.w * (1.f / 6.f));
    }

    const long long rowp = G.os_row;
    const uint32_t tile0 = smem_addr(smem) + (uint32_t)(2 * j * CG + g) * 8u;   // left neighbour of column 2j, row 0, stage 0
    const uint32_t rowstride = (uint32_t)(TWL * CG) * 8u;
    float2 sA = make_float2(0.f, 0.f), sB = sA, qA = sA, qB = sA;
    int it = 0;
    for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
    S1_UNIT(unit)
    const int ow = ow0 + 2 * j;
    const bool act0 = live && ow < G.W, act1 = live && ow + 1 < G.W;
    // running output pointers: row o = r - 2 of step r
    __nv_bfloat16* yrow0 = y + (long long)n * G.os_img + (long long)(oh0 - 2) * G.os_row + (long long)min(ow, G.W - 1) * G.os_pix + c;
    __nv_bfloat16* yrow1 = y + (long long)n * G.os_img + (long long)(oh0 - 2) * G.os_row + (long long)min(ow + 1, G.W - 1) * G.os_pix + c;
    // !HALO: zero (not relu6(shift)) outside the image
    const bool ok0 = ow - 1 >= 0 && ow - 1 < G.W, ok1 = ow < G.W, ok2 = ow + 1 < G.W, ok3 = ow + 2 < G.W;
    float2 accA[2][3], accB[2][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) accA[0][i] = accB[0][i] = accA[1][i] = accB[1][i] = make_float2(0.f, 0.f);

    for (int k = 0; k < nst; ++k, ++it) {
      const int s = it % FWD_STAGES;
      mbar_wait(smem_addr(&bar_full[s]), (it / FWD_STAGES) & 1);
      const uint32_t tile = tile0 + (uint32_t)s * (uint32_t)G.stage_bytes;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = k * RB + u;            // input row oh0 - 1 + r
        const uint32_t rowt = tile + (uint32_t)u * rowstride;
        float2 aA[4], aB[4];                 // activated columns 2j-1 .. 2j+2
        act4(lds8<0>(rowt), scA, scB, shA, shB, aA[0], aB[0]);
        act4(lds8<CG * 8>(rowt), scA, scB, shA, shB, aA[1], aB[1]);
        act4(lds8<CG * 16>(rowt), scA, scB, shA, shB, aA[2], aB[2]);
        act4(lds8<CG * 24>(rowt), scA, scB, shA, shB, aA[3], aB[3]);
        if (!HALO) {
          const bool row_ok = (unsigned)(oh0 - 1 + r) < (unsigned)G.H;
          if (!(row_ok && ok0)) aA[0] = aB[0] = make_float2(0.f, 0.f);
          if (!(row_ok && ok1)) aA[1] = aB[1] = make_float2(0.f, 0.f);
          if (!(row_ok && ok2)) aA[2] = aB[2] = make_float2(0.f, 0.f);
          if (!(row_ok && ok3)) aA[3] = aB[3] = make_float2(0.f, 0.f);
        }
        // output row o = r - ky gets ky's filter row; slot of o is o % 3 (u == r % 3)
        const int s0 = u, s1 = (u + 2) % 3, s2 = (u + 1) % 3;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          accA[q][s0] = ffma2(aA[q + 2], wA[2], ffma2(aA[q + 1], wA[1], fmul2(aA[q], wA[0])));
          accB[q][s0] = ffma2(aB[q + 2], wB[2], ffma2(aB[q + 1], wB[1], fmul2(aB[q], wB[0])));
          accA[q][s1] = ffma2(aA[q + 2], wA[5], ffma2(aA[q + 1], wA[4], ffma2(aA[q], wA[3], accA[q][s1])));
          accB[q][s1] = ffma2(aB[q + 2], wB[5], ffma2(aB[q + 1], wB[4], ffma2(aB[q], wB[3], accB[q][s1])));
          accA[q][s2] = ffma2(aA[q + 2], wA[8], ffma2(aA[q + 1], wA[7], ffma2(aA[q], wA[6], accA[q][s2])));
          accB[q][s2] = ffma2(aB[q + 2], wB[8], ffma2(aB[q + 1], wB[7], ffma2(aB[q], wB[6], accB[q][s2])));
        }
        if ((unsigned)(r - 2) < (unsigned)rows && act0) {
          if (OAFF) {
            *reinterpret_cast<uint2*>(yrow0) = pack4(oaff2(accA[0][s2], oscA, oshA), oaff2(accB[0][s2], oscB, oshB));
            if (act1) *reinterpret_cast<uint2*>(yrow1) = pack4(oaff2(accA[1][s2], oscA, oshA), oaff2(accB[1][s2], oscB, oshB));
          } else {
            *reinterpret_cast<uint2*>(yrow0) = pack4(accA[0][s2], accB[0][s2]);
            sA = fadd2(sA, accA[0][s2]);
            sB = fadd2(sB, accB[0][s2]);
            qA = ffma2(accA[0][s2], accA[0][s2], qA);
            qB = ffma2(accB[0][s2], accB[0][s2], qB);
            if (act1) {
              *reinterpret_cast<uint2*>(yrow1) = pack4(accA[1][s2], accB[1][s2]);
              sA = fadd2(sA, accA[1][s2]);
              sB = fadd2(sB, accB[1][s2]);
              qA = ffma2(accA[1][s2], accA[1][s2], qA);
              qB = ffma2(accB[1][s2], accB[1][s2], qB);
            }
          }
        }
        yrow0 += rowp;
        yrow1 += rowp;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&bar_empty[s]));
    }
    }   // units
    if (stats) {
      const float v[8] = {sA.x, sA.y, sB.x, sB.y, qA.x, qA.y, qB.x, qB.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) red[k * (FWD2_CONS + 1) + threadIdx.x] = live ? v[k] : 0.f;
    }
  }
  if (stats) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < CG * 8) {
      const int gg = t / 8, k = t % 8;
      const float* p = red + k * (FWD2_CONS + 1) + gg;
      float tot = 0.f;
      for (int jj = 0; jj < npair; ++jj) tot += p[jj * CG];
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
// Input-stationary in dy: the dy row of step r feeds the g rows r, r-1, r-2 (filter rows 2, 1, 0) and, with
// the a6 rows r-2, r-1, r, all nine weight-gradient taps.
template <int CG>
__global__ void __launch_bounds__(BWD_CONS + 32, 1)
dw_s1_bwd_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                 const float* __restrict__ ss, const float* __restrict__ mi, const float* __restrict__ w,
                 __nv_bfloat16* __restrict__ gout, double* __restrict__ bsums, float* __restrict__ dw, const S1Geom G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t bar_full[BWD_STAGES], bar_empty[BWD_STAGES];
  __shared__ float red[12 * (BWD_CONS + 1)];
  const int TWL = G.TW + 2, ext = G.ext;
  const int He = G.H + 2 * ext, We = G.W + 2 * ext;
  const int ncons = G.TW * CG;
  const int ncw = (ncons + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define S1_BUNIT(unit_)                                                  \
  const int tile_ = (unit_) % G.tiles, rest_ = (unit_) / G.tiles;        \
  const int seg_ = rest_ % G.nseg, n = rest_ / G.nseg;                   \
  const int e0 = tile_ * G.TW, he0 = seg_ * G.rs;                        \
  const int rows = min(G.rs, He - he0);                                  \
  const int ih_start = he0 - ext; /* image row of the first g row */     \
  const int nst = (rows + 2 + RB - 1) / RB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < BWD_STAGES; ++s) {
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
      prefetch_tmap(&dymap);
      prefetch_tmap(&xmap);
      int it = 0;
      for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
        S1_BUNIT(unit)
        for (int k = 0; k < nst; ++k, ++it) {
          const int s = it % BWD_STAGES;
          if (it >= BWD_STAGES) mbar_wait(smem_addr(&bar_empty[s]), ((it / BWD_STAGES) - 1) & 1);
          const uint32_t full = smem_addr(&bar_full[s]);
          mbar_expect_tx(full, (uint32_t)(RB * (TWL + G.TW) * CG * 8));
          unsigned char* st = smem + (size_t)s * G.stage_bytes;
          // step r uses dy row ih_start-1+r (with one halo column per side) and x row ih_start+r
          tma_load_4d(smem_addr(st), &dymap, full, blockIdx.x * CG * 4, e0 - ext - 1, ih_start - 1 + k * RB, n);
          tma_load_4d(smem_addr(st + G.xoff), &xmap, full, blockIdx.x * CG * 4, e0 - ext, ih_start + k * RB, n);
        }
      }
    }
  } else if (warp < ncw) {
    live = threadIdx.x < ncons;
    const int tid = live ? threadIdx.x : 0;
    const int g = tid % CG, j = tid / CG;
    const int c = (blockIdx.x * CG + g) * 4;

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

    const long long growp = G.os_row;
    // interior layout: position (he, we) of the extended domain lives at (he - ext, we - ext) and exists only inside
    const int gsh = G.interior ? ext : 0;
    const uint32_t dtile0 = smem_addr(smem) + (uint32_t)(j * CG + g) * 8u;
    const uint32_t xtile0 = dtile0 + (uint32_t)G.xoff;
    const uint32_t drowstride = (uint32_t)(TWL * CG) * 8u, xrowstride = (uint32_t)(G.TW * CG) * 8u;
    int it = 0;
    for (int unit = blockIdx.y; unit < G.nunits; unit += gridDim.y) {
    S1_BUNIT(unit)
    const int we = e0 + j;
    const bool active = live && we < We;
    // running output pointer: row o = r - 2 of step r
    __nv_bfloat16* grow = gout + (long long)n * G.os_img + (long long)(he0 - gsh - 2) * G.os_row +
                          (long long)(min(we, We - 1) - gsh) * G.os_pix + c;
    const bool col_in = !G.interior || (unsigned)(we - ext) < (unsigned)G.W;
    float2 gA[3], gB[3], aA[3], aB[3];
    uint2 xr[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gA[i] = gB[i] = aA[i] = aB[i] = make_float2(0.f, 0.f);
      xr[i] = make_uint2(0u, 0u);
    }

    for (int k = 0; k < nst; ++k, ++it) {
      const int s = it % BWD_STAGES;
      mbar_wait(smem_addr(&bar_full[s]), (it / BWD_STAGES) & 1);
      const uint32_t soff = (uint32_t)s * (uint32_t)G.stage_bytes;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int r = k * RB + u;            // dy row ih_start-1+r; g / x row o = r (relative to ih_start)
        const uint32_t drow = dtile0 + soff + (uint32_t)u * drowstride;
        float2 lA, lB, cA, cB, rA, rB;       // dy at columns iw-1, iw, iw+1
        unpack4(lds8<0>(drow), lA, lB);
        unpack4(lds8<CG * 8>(drow), cA, cB);
        unpack4(lds8<CG * 16>(drow), rA, rB);
        const int s0 = u, s1 = (u + 2) % 3, s2 = (u + 1) % 3;   // slots of g rows r, r-1, r-2
        // x / a6 of row o = r (zero contribution outside this CTA's rows or columns)
        xr[s0] = lds8<0>(xtile0 + soff + (uint32_t)u * xrowstride);
        act4(xr[s0], scA, scB, shA, shB, aA[s0], aB[s0]);
        if (!(r < rows && active)) aA[s0] = aB[s0] = make_float2(0.f, 0.f);
        // data gradient: g row o uses dy row r = o + 2 - ky; dy column iw + 1 - kx  (kx=0: right, 2: left)
        gA[s0] = ffma2(lA, wA[8], ffma2(cA, wA[7], fmul2(rA, wA[6])));
        gB[s0] = ffma2(lB, wB[8], ffma2(cB, wB[7], fmul2(rB, wB[6])));
        gA[s1] = ffma2(lA, wA[5], ffma2(cA, wA[4], ffma2(rA, wA[3], gA[s1])));
        gB[s1] = ffma2(lB, wB[5], ffma2(cB, wB[4], ffma2(rB, wB[3], gB[s1])));
        gA[s2] = ffma2(lA, wA[2], ffma2(cA, wA[1], ffma2(rA, wA[0], gA[s2])));
        gB[s2] = ffma2(lB, wB[2], ffma2(cB, wB[1], ffma2(rB, wB[0], gB[s2])));
        // weight gradient: tap (ky, kx) pairs a6 of row o = r - 2 + ky with this dy row
        dA[0] = ffma2(rA, aA[s2], dA[0]); dB[0] = ffma2(rB, aB[s2], dB[0]);
        dA[1] = ffma2(cA, aA[s2], dA[1]); dB[1] = ffma2(cB, aB[s2], dB[1]);
        dA[2] = ffma2(lA, aA[s2], dA[2]); dB[2] = ffma2(lB, aB[s2], dB[2]);
        dA[3] = ffma2(rA, aA[s1], dA[3]); dB[3] = ffma2(rB, aB[s1], dB[3]);
        dA[4] = ffma2(cA, aA[s1], dA[4]); dB[4] = ffma2(cB, aB[s1], dB[4]);
        dA[5] = ffma2(lA, aA[s1], dA[5]); dB[5] = ffma2(lB, aB[s1], dB[5]);
        dA[6] = ffma2(rA, aA[s0], dA[6]); dB[6] = ffma2(rB, aB[s0], dB[6]);
        dA[7] = ffma2(cA, aA[s0], dA[7]); dB[7] = ffma2(cB, aB[s0], dB[7]);
        dA[8] = ffma2(lA, aA[s0], dA[8]); dB[8] = ffma2(lB, aB[s0], dB[8]);
        const int o = r - 2;
        if ((unsigned)o < (unsigned)rows && active) {
          float2 vA = gA[s2], vB = gB[s2];
          const float2 mA = aA[s2], mB = aB[s2];
          vA.x = (mA.x > 0.f && mA.x < 1.f) ? vA.x : 0.f;
          vA.y = (mA.y > 0.f && mA.y < 1.f) ? vA.y : 0.f;
          vB.x = (mB.x > 0.f && mB.x < 1.f) ? vB.x : 0.f;
          vB.y = (mB.y > 0.f && mB.y < 1.f) ? vB.y : 0.f;
          st8_if(grow, pack4(vA, vB), col_in && (!G.interior || (unsigned)(ih_start + o) < (unsigned)G.H));
          float2 xa, xb;
          unpack4(xr[s2], xa, xb);
          sA = fadd2(sA, vA);
          sB = fadd2(sB, vB);
          qA = ffma2(vA, fadd2(xa, nmuA), qA);
          qB = ffma2(vB, fadd2(xb, nmuB), qB);
        }
        grow += growp;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&bar_empty[s]));
    }
    }   // units
  }
  const int t = threadIdx.x;
  const bool consumer = t < BWD_CONS;
  if (bsums) {
    const float v[8] = {live ? sA.x : 0.f, live ? sA.y : 0.f, live ? sB.x : 0.f, live ? sB.y : 0.f,
                        live ? qA.x : 0.f, live ? qA.y : 0.f, live ? qB.x : 0.f, live ? qB.y : 0.f};
    const float tot = column_reduce<8, BWD_CONS>(red, v, CG, G.TW, consumer);
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
        v[k * 4 + 0] = live ? dA[part * 3 + k].x : 0.f;
        v[k * 4 + 1] = live ? dA[part * 3 + k].y : 0.f;
        v[k * 4 + 2] = live ? dB[part * 3 + k].x : 0.f;
        v[k * 4 + 3] = live ? dB[part * 3 + k].y : 0.f;
      }
      const float tot = column_reduce<12, BWD_CONS>(red, v, CG, G.TW, consumer);
      if (t < CG * 12) {
        const int gg = t / 12, k = t % 12;
        const int ch = (blockIdx.x * CG + gg) * 4 + (k & 3);
        atomicAdd(&dw[ch * 9 + part * 3 + (k >> 2)], 6.f * tot);
      }
      __syncthreads();
    }
  }
}

// chunking: CG 4-channel groups per CTA, TW columns, TW*CG <= ncons_max consumer threads
inline bool s1_plan(int N, int H, int W, int C, int ext, int ncons_max, int stages, bool bwd, int ctas_per_sm, S1Geom* G, dim3* grid,
                    int* threads, size_t* smem, int cols = 1) {
  int CG;
  if (C % 32 == 0) CG = 8;
  else if (C % 48 == 0) CG = 12;
  else if (C % 16 == 0) CG = 4;
  else return false;
  const int He = H + 2 * ext, We = W + 2 * ext;
  int TWmax = cols * (ncons_max / CG);   // cols columns per consumer thread
  if (TWmax + 2 > 256) TWmax = 254 / cols * cols;   // TMA box limit
  const int chunks = C / (CG * 4);
  // persistent CTAs: `per_chunk` of them per channel chunk (ctas_per_sm resident CTAs of this kernel on every SM in
  // total), each taking work units (image, row segment, column tile) round-robin.  The unit shape decides the balance:
  // a CTA's time is (units it takes) x (rows per unit + 2 halo rows), the kernel's time that of the busiest CTA, so
  // the split is searched -- column tilings near the widest one x segment heights with rs + 2 a multiple of the stage
  // depth RB (no loaded row is wasted) -- for the smallest rounds x (rs + 3) (one row of slack per unit boundary).
  // [The first version grew the segment count by doubling until every CTA had ~6 units; 96 units on 74 CTAs, two
  // rounds for 1.3 rounds of work, is what that gave on the 192-channel 64x128 layer.]
  int per_chunk = (ctas_per_sm * s2r_sm_count()) / chunks;
  if (per_chunk < 1) per_chunk = 1;
  const int tmin = s2r_div_up(We, TWmax);
  long best_cost = -1;
  int tiles = tmin, rs = 0;
  for (int t = tmin; t <= tmin + 2; ++t) {
    if (t > tmin && s2r_div_up(We, t) < cols) break;
    for (int r = RB + RB - 2; ; r += RB) {          // 4, 7, 10, ...: r + 2 is a multiple of RB
      const int nsg = s2r_div_up(He, r);
      const long units = (long)t * N * nsg;
      const long rounds = (units + per_chunk - 1) / per_chunk;
      const long cost = rounds * (r + 3);
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && t == tiles && r > rs)) {
        best_cost = cost; tiles = t; rs = r;
      }
      if (r >= He) break;
    }
  }
  const int TW = s2r_div_up(s2r_div_up(We, tiles), cols) * cols;   // balance the column tiles
  if (TW + 2 > 256) return false;
  const int nseg = s2r_div_up(He, rs);
  const long nunits = (long)tiles * N * nseg;
  if (nunits > 0x7fffffffL) return false;
  if (per_chunk > nunits) per_chunk = (int)nunits;
  G->tiles = tiles; G->nunits = (int)nunits;
  G->N = N; G->H = H; G->W = W; G->C = C; G->CG = CG; G->TW = TW; G->rs = rs; G->nseg = nseg; G->ext = ext;
  G->publish = 0;
  const int dy_bytes = RB * (TW + 2) * CG * 8, x_bytes = bwd ? RB * TW * CG * 8 : 0;
  G->xoff = (dy_bytes + 127) / 128 * 128;
  G->stage_bytes = (G->xoff + x_bytes + 127) / 128 * 128;
  *grid = dim3(chunks, per_chunk, 1);
  *threads = (TW / cols * CG + 31) / 32 * 32 + 32;
  if (*threads < (CG * 12 + 31) / 32 * 32) *threads = (CG * 12 + 31) / 32 * 32;   // the final reductions use CG*12 threads
  *smem = (size_t)stages * G->stage_bytes + 128;
  return true;
}

// static + dynamic shared memory can exceed the 48 KB default even when the dynamic part alone does not:
// opt in once per kernel to the largest ring any plan produces
constexpr int S1_SMEM_CAP = 160 * 1024;
template <typename K>
inline int s1_smem_attr(K kernel, size_t smem, int which) {
  static bool flags[32] = {};   // per kernel instance
  bool& done = flags[which];
  S2R_REQUIRE(smem <= (size_t)S1_SMEM_CAP, S2R_ERR_UNSUPPORTED, "dwconv3x3: ring of %zu bytes exceeds the cap", smem);
  if (!done) {
    S2R_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S1_SMEM_CAP));
    done = true;
  }
  return S2R_OK;
}

}  // namespace

// Internal entry points (called from dwconv.cu); return S2R_ERR_UNSUPPORTED when the shape is not covered.
// Dilation d (padding d) is d*d independent dilation-1 problems on the parity planes of the tensor: plane (p, q)
// holds the pixels (d*i + p, d*j + q) and is addressed through a TMA map with d-fold strides.
int s2r_dw_s1_fwd(const void* x, const float* ss, const s2r_bn_tail* in_bn, int halo_const, const float* w, void* y,
                  double* stats, const float* oss, int N, int H, int W, int C, int dil, cudaStream_t stream) {
  if (oss && (stats || (uintptr_t)oss % 16)) return S2R_ERR_UNSUPPORTED;   // output affine: inference only
  // in_bn: the input's BatchNorm is pending -- every launch derives scale / shift from its sums, the first one of a
  // dilated set (dil*dil parity-plane launches) publishes them and updates the running statistics
  const BnTail bt = bn_tail_from(in_bn);
  bool published = false;
  for (int p = 0; p < dil; ++p)
    for (int q = 0; q < dil; ++q) {
      const int Hp = (H - p + dil - 1) / dil, Wp = (W - q + dil - 1) / dil;
      if (Hp <= 0 || Wp <= 0) continue;
      S1Geom G;
      dim3 grid;
      int threads;
      size_t smem;
      // variant: two columns per thread on the large images, one on the small ones (measured per layer,
      // tests/tools/dw_bench.py; S2R_DW_COLS=1|2 forces one)
      static const int force_cols = getenv("S2R_DW_COLS") ? atoi(getenv("S2R_DW_COLS")) : 0;
      const int cols = force_cols == 1 || force_cols == 2 ? force_cols : ((long)Hp * Wp >= 128L * 256 ? 2 : 1);
      if (!s1_plan(N, Hp, Wp, C, 0, cols == 2 ? FWD2_CONS : FWD_CONS, FWD_STAGES, false, 2, &G, &grid, &threads, &smem, cols))
        return S2R_ERR_UNSUPPORTED;
      const long long poff = ((long long)p * W + q) * C;
      G.os_pix = (long long)dil * C; G.os_row = (long long)dil * W * C; G.os_img = (long long)H * W * C;
      G.interior = 0;
      G.publish = published ? 0 : 1;
      published = true;
      CUtensorMap xmap;
      if (!encode_nhwc_view(&xmap, (const __nv_bfloat16*)x + poff, N, Hp, Wp, C, G.os_pix, G.os_row, G.os_img,
                            G.CG * 4, G.TW + 2, RB))
        return S2R_ERR_UNSUPPORTED;
      __nv_bfloat16* yv = (__nv_bfloat16*)y + poff;
#define S2R_DW_FWD(K_, HALO_, CG_, OAFF_, SLOT_)                                              \
  do {                                                                                      \
    int rc = s1_smem_attr(K_<HALO_, CG_, OAFF_>, smem, SLOT_);                              \
    if (rc) return rc;                                                                      \
    S2R_CUDA_OK(s2r_launch(K_<HALO_, CG_, OAFF_>, grid, dim3(threads), smem, stream, xmap, ss, w, yv, stats, G, bt, oss)); \
  } while (0)
#define S2R_DW_FWD_CG(K_, HALO_, OAFF_, SLOT_)                                              \
  do {                                                                                      \
    if (G.CG == 8) S2R_DW_FWD(K_, HALO_, 8, OAFF_, SLOT_);                                  \
    else if (G.CG == 12) S2R_DW_FWD(K_, HALO_, 12, OAFF_, SLOT_ + 1);                       \
    else S2R_DW_FWD(K_, HALO_, 4, OAFF_, SLOT_ + 2);                                        \
  } while (0)
      if (cols == 2) {
        if (halo_const) { if (oss) S2R_DW_FWD_CG(dw_s1_fwd2_kernel, true, true, 0); else S2R_DW_FWD_CG(dw_s1_fwd2_kernel, true, false, 3); }
        else { if (oss) S2R_DW_FWD_CG(dw_s1_fwd2_kernel, false, true, 6); else S2R_DW_FWD_CG(dw_s1_fwd2_kernel, false, false, 9); }
      } else {
        if (halo_const) { if (oss) S2R_DW_FWD_CG(dw_s1_fwd_kernel, true, true, 12); else S2R_DW_FWD_CG(dw_s1_fwd_kernel, true, false, 15); }
        else { if (oss) S2R_DW_FWD_CG(dw_s1_fwd_kernel, false, true, 18); else S2R_DW_FWD_CG(dw_s1_fwd_kernel, false, false, 21); }
      }
#undef S2R_DW_FWD_CG
#undef S2R_DW_FWD
      S2R_LAUNCH_OK();
    }
  return S2R_OK;
}

// ext = 0 or dil (the reference's padded border); g is [N][H+2ext][W+2ext][C]
int s2r_dw_s1_bwd(const void* dy, const void* x, const float* ss, const float* mi, const float* w, int ext,
                  int interior, void* g, double* bsums, float* dw, int N, int H, int W, int C, int dil, cudaStream_t stream) {
  if (!ext) interior = 0;
  const int We = interior ? W : W + 2 * ext, He = interior ? H : H + 2 * ext;   // g layout
  for (int p = 0; p < dil; ++p)
    for (int q = 0; q < dil; ++q) {
      const int Hp = (H - p + dil - 1) / dil, Wp = (W - q + dil - 1) / dil;
      if (Hp <= 0 || Wp <= 0) continue;
      S1Geom G;
      dim3 grid;
      int threads;
      size_t smem;
      if (!s1_plan(N, Hp, Wp, C, ext ? 1 : 0, BWD_CONS, BWD_STAGES, true, 1, &G, &grid, &threads, &smem)) return S2R_ERR_UNSUPPORTED;
      const long long poff = ((long long)p * W + q) * C, goff = ((long long)p * We + q) * C;
      const long long sw = (long long)dil * C, sh = (long long)dil * W * C, sn = (long long)H * W * C;
      G.os_pix = (long long)dil * C; G.os_row = (long long)dil * We * C; G.os_img = (long long)He * We * C;
      G.interior = interior;
      CUtensorMap dymap, xmap;
      if (!encode_nhwc_view(&dymap, (const __nv_bfloat16*)dy + poff, N, Hp, Wp, C, sw, sh, sn, G.CG * 4, G.TW + 2, RB))
        return S2R_ERR_UNSUPPORTED;
      if (!encode_nhwc_view(&xmap, (const __nv_bfloat16*)x + poff, N, Hp, Wp, C, sw, sh, sn, G.CG * 4, G.TW, RB))
        return S2R_ERR_UNSUPPORTED;
#define S2R_DW_BWD(CG_, SLOT_)                                                                                       \
  do {                                                                                                               \
    int rc = s1_smem_attr(dw_s1_bwd_kernel<CG_>, smem, SLOT_);                                                       \
    if (rc) return rc;                                                                                               \
    S2R_CUDA_OK(s2r_launch(dw_s1_bwd_kernel<CG_>, grid, dim3(threads), smem, stream, dymap, xmap, ss, mi, w, (__nv_bfloat16*)g + goff, bsums, dw, G)); \
  } while (0)
      if (G.CG == 8) S2R_DW_BWD(8, 24);
      else if (G.CG == 12) S2R_DW_BWD(12, 25);
      else S2R_DW_BWD(4, 26);
#undef S2R_DW_BWD
      S2R_LAUNCH_OK();
    }
  return S2R_OK;
}
