// tcgen05 / TMA weight-gradient kernel for the dense convolutions (sm_100a).
//
//   dW[t][co][ci] += sum_pix dy[pix][co] * view_t[pix + d_t][ci]
//
// The contraction runs over PIXELS, so both operands are consumed exactly as they sit in HBM (pixel rows
// of channels): TMA boxes [8x8 or 1x64 pixel patch][64 channels] land as 64 rows x 128 B (128B swizzle),
// which is the canonical MN-major UMMA operand (64-element blocks LBO = 8 KB apart, 8-row groups SBO = 1 KB
// apart).  A = dy^T (M = 128 output channels per accumulator, MT accumulators share the B tile), B = the
// tap's shifted input view (N = up to 256 input channels), K = 16 pixels per tcgen05.mma, fp32 accumulators
// in TMEM.  The tap offset is added to the box origin of the input view and TMA zero-fills what falls outside
// (padding, stride-2 parity views, ragged edges), exactly as in the forward kernel.
// Grid = taps x output-channel tiles x input-channel tiles x pixel splits; every CTA walks its pixel range
// through a multi-stage mbarrier pipeline and adds its partial tile into the fp32 gradient with atomics
// (12 epilogue warps, tcgen05.ld 32x32b).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int WG_THREADS = 512;
constexpr int WG_BKP = 64;               // pixels per pipeline stage
constexpr int WG_BOX_BYTES = 64 * 128;   // one [64 pixels][64 channels] box

struct WgTap {
  int map, dh, dw;
  long long wofs;
};

struct WgTcParams {
  WgTap taps[S2R_MAX_TAPS];
  int ntaps, co_tiles, ci_tiles, splits;
  int Cin, Cout;
  int BW, BH, tiles_w, tiles_h, n_chunks, chunks_per_split;
  float* dwt;
  long long s_co, s_ci;
};

struct WgMaps {
  CUtensorMap x[4];
  CUtensorMap dy;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: 64-element (128 B) blocks along M/N are LBO apart, groups of 8 K-rows
// (1024 B) are SBO apart
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}
constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// BN: input channels per tile (N of the MMA), MT: 128-row output-channel accumulators per CTA
template <int BN, int MT, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgTcParams p) {
  constexpr int NBOX_B = (BN + 63) / 64;
  constexpr int A_BYTES = MT * 2 * WG_BOX_BYTES;
  constexpr int B_BYTES = NBOX_B * WG_BOX_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TCOLS = tmem_cols(MT * BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = tap (fastest: CTAs of all taps stream the same pixel range together), then ci tile, then co tile
  int bx = blockIdx.x;
  const int tap = bx % p.ntaps;
  bx /= p.ntaps;
  const int ci0 = (bx % p.ci_tiles) * BN;
  const int co0 = (bx / p.ci_tiles) * (MT * 128);
  const int chunk_beg = blockIdx.y * p.chunks_per_split;
  const int chunk_end = min(p.n_chunks, chunk_beg + p.chunks_per_split);
  const int kiters = chunk_end - chunk_beg;
  const WgTap T = p.taps[tap];

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), 1);
    }
    mbar_init(smem_addr(&bar_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.x[T.map]);
    prefetch_tmap(&maps.dy);
  }
  if (warp == 2) tmem_alloc<TCOLS>(smem_addr(&tmem_base_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_trigger();

  if (kiters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_addr(&bar_empty[s]), ph ^ 1);
          const int c = chunk_beg + it;
          const int w0 = (c % p.tiles_w) * p.BW;
          const int q = c / p.tiles_w;
          const int h0 = (q % p.tiles_h) * p.BH, n = q / p.tiles_h;
          const uint32_t full = smem_addr(&bar_full[s]);
          const uint32_t sa = smem_addr(smem + (size_t)s * STAGE_BYTES);
          mbar_expect_tx(full, STAGE_BYTES);
#pragma unroll
          for (int b = 0; b < MT * 2; ++b) tma_load_4d(sa + b * WG_BOX_BYTES, &maps.dy, full, co0 + b * 64, w0, h0, n);
#pragma unroll
          for (int b = 0; b < NBOX_B; ++b)
            tma_load_4d(sa + A_BYTES + b * WG_BOX_BYTES, &maps.x[T.map], full, ci0 + b * 64, w0 + T.dw, h0 + T.dh, n);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = BN, M = 128
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_addr(&bar_full[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_addr(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < WG_BKP / 16; ++kk) {
            const uint64_t db = umma_desc_mn128(sb + kk * 2048, WG_BOX_BYTES);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              const uint64_t da = umma_desc_mn128(sa + j * 2 * WG_BOX_BYTES + kk * 2048, WG_BOX_BYTES);
              umma_bf16(tmem_base + j * BN, da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_addr(&bar_empty[s]));
        }
        umma_commit(smem_addr(&bar_acc));
      }
    } else if (warp >= 4) {
      const int ew = warp & 3, grp = (warp - 4) >> 2;
      constexpr int CHUNKS = BN / 32;
      mbar_wait(smem_addr(&bar_acc), 0);
      tc_fence_after();
#pragma unroll 1
      for (int u = grp; u < MT * CHUNKS; u += 3) {
        const int j = u / CHUNKS, c0 = (u - j * CHUNKS) * 32;
        if (ci0 + c0 >= p.Cin) continue;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(j * BN + c0), r);
        tmem_ld_wait();
        // transpose through shared memory (the pipeline ring is idle once bar_acc has fired) so that one atomic
        // instruction covers 32 CONSECUTIVE input channels of one output channel -- a single 128-byte line for
        // 1x1 filters -- instead of 32 different output-channel rows
        float* stg = reinterpret_cast<float*>(smem) + (warp - 4) * 1024;   // this warp's 32 rows x 32 floats
        {
          uint4* rp = reinterpret_cast<uint4*>(stg + lane * 32);
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) rp[qq ^ (lane & 7)] = make_uint4(r[4 * qq], r[4 * qq + 1], r[4 * qq + 2], r[4 * qq + 3]);
        }
        __syncwarp();
        const int ci = ci0 + c0 + lane;
        if (ci < p.Cin) {
          float* dst = p.dwt + (long long)ci * p.s_ci + T.wofs;
          const int co_base = co0 + j * 128 + ew * 32;
          const int nrow = min(32, p.Cout - co_base);
#pragma unroll 4
          for (int rr = 0; rr < nrow; ++rr) {
            const float v = stg[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))];
            atomicAdd(dst + (long long)(co_base + rr) * p.s_co, v);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &sym, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
    else
      cudaGetLastError();
  }
  return fn;
}

struct View {
  const void* base;
  long long sn, sh, sw;
  int H, W;
  bool operator==(const View& o) const {
    return base == o.base && sn == o.sn && sh == o.sh && sw == o.sw && H == o.H && W == o.W;
  }
};

bool encode_view(EncodeTiledFn enc, CUtensorMap* m, const View& v, int C, int N, int BW, int BH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = 16;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int MT, int STAGES>
int launch_wg(const WgMaps& maps, const WgTcParams& p, cudaStream_t st) {
  constexpr int smem = STAGES * ((MT * 2 + (BN + 63) / 64) * WG_BOX_BYTES) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    S2R_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<BN, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(p.ntaps * p.ci_tiles * p.co_tiles, p.splits);
  S2R_CUDA_OK(s2r_launch(wgrad_tc_kernel<BN, MT, STAGES>, grid, dim3(WG_THREADS), (size_t)smem, st, maps, p));
  S2R_LAUNCH_OK();
  return 1;
}

int wg_tc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("S2R_WGRAD");
    mode = (e && e[0] == 'm') ? 0 : 1;
  }
  return mode;
}

}  // namespace

// returns 1 when launched here, 0 to fall back to the mma.sync path, <0 on error
int s2r_conv_wgrad_tc(const s2r_wgrad_args* a, cudaStream_t st) {
  if (!wg_tc_mode()) return 0;
  EncodeTiledFn enc = get_encode();
  if (!enc) return 0;
  if (a->ntaps < 1 || a->ntaps > S2R_MAX_TAPS || a->Cin < 1 || a->Cout < 1) return 0;
  if (!a->dy || (uintptr_t)a->dy % 16 || a->dn % 8 || a->dh % 8 || a->dw % 8) return 0;
  View views[4];
  int nviews = 0;
  WgTcParams p;
  for (int i = 0; i < a->ntaps; ++i) {
    const s2r_tap& t = a->taps[i];
    if (!t.base || (uintptr_t)t.base % 16 || t.sn % 8 || t.sh % 8 || t.sw % 8 || t.H < 1 || t.W < 1) return 0;
    View k = {t.base, t.sn, t.sh, t.sw, t.H, t.W};
    int m = -1;
    for (int j = 0; j < nviews; ++j)
      if (views[j] == k) m = j;
    if (m < 0) {
      if (nviews == 4) return 0;
      views[nviews] = k;
      m = nviews++;
    }
    p.taps[i].map = m;
    p.taps[i].dh = t.dh;
    p.taps[i].dw = t.dw;
    p.taps[i].wofs = t.wofs;
  }
  int N = a->N, OH = a->OH, OW = a->OW;
  View dyv = {a->dy, a->dn, a->dh, a->dw, OH, OW};
  // pointwise on pixel-contiguous tensors: flatten the pixel grid so that every 64-pixel chunk is full
  if (a->ntaps == 1 && a->taps[0].dh == 0 && a->taps[0].dw == 0 && views[0].H == OH && views[0].W == OW) {
    View& v = views[0];
    const bool in_flat = v.sh == v.sw * OW && v.sn == v.sh * OH;
    const bool dy_flat = dyv.sh == dyv.sw * OW && dyv.sn == dyv.sh * OH;
    if (in_flat && dy_flat) {
      OW = N * OH * OW;
      OH = 1;
      N = 1;
      v.W = OW; v.H = 1; v.sh = v.sw * OW; v.sn = v.sh;
      dyv.W = OW; dyv.H = 1; dyv.sh = dyv.sw * OW; dyv.sn = dyv.sh;
    }
  }
  int bestBW = 64;
  long long best = -1;
  for (int bw = 64; bw >= 8; bw >>= 1) {
    const long long tiles = (long long)s2r_div_up(OW, bw) * s2r_div_up(OH, 64 / bw);
    if (best < 0 || tiles < best) {
      best = tiles;
      bestBW = bw;
    }
  }
  p.BW = bestBW;
  p.BH = 64 / bestBW;
  p.tiles_w = s2r_div_up(OW, p.BW);
  p.tiles_h = s2r_div_up(OH, p.BH);
  const long long nchunks = (long long)N * p.tiles_w * p.tiles_h;
  if (nchunks >= (1ll << 30)) return 0;
  p.n_chunks = (int)nchunks;
  p.ntaps = a->ntaps;
  p.Cin = a->Cin;
  p.Cout = a->Cout;
  p.dwt = a->dweight;
  p.s_co = a->s_co;
  p.s_ci = a->s_ci;

  const int cin8 = (a->Cin + 7) & ~7, cout8 = (a->Cout + 7) & ~7;
  int BN;
  if (a->Cin > 256 && a->Cin <= 320) BN = 160;   // 257..320 input channels: two 160-wide tiles, not 256 + a sliver
  else if (a->Cin > 128) BN = 256;
  else if (a->Cin > 64) BN = 128;
  else if (a->Cin > 32) BN = 64;
  else BN = 32;
  const int MT = a->Cout > 128 ? 2 : 1;
  p.ci_tiles = s2r_div_up(a->Cin, BN);
  p.co_tiles = s2r_div_up(a->Cout, MT * 128);
  const long long tiles = (long long)p.ntaps * p.ci_tiles * p.co_tiles;
  // split the pixel range so that about two CTAs per SM exist, each with at least 8 chunks
  // Pixel splits trade parallelism against the atomic traffic of the partial tiles (measured, tests/tools/wg_sweep.sh):
  // 1x1 filters (coalesced atomics, one tile per CTA column): one CTA per SM; multi-tap filters re-read the same
  // pixels once per tap out of L2 and like two CTAs per SM, unless the pixel range is short and the strided
  // (uncoalesced) atomics of their large tiles dominate.
  static double knob_ctas = -1;
  static int min_chunks = 8;
  if (knob_ctas < 0) {   // overrides for tuning
    const char* e = getenv("S2R_WG_CTAS_PER_SM");
    knob_ctas = e ? atof(e) : 0.0;
    const char* m = getenv("S2R_WG_MIN_CHUNKS");
    if (m) min_chunks = atoi(m);
  }
  double ctas_per_sm = p.ntaps == 1 ? 1.0 : ((long long)nchunks * WG_BKP >= 50000 ? 4.0 : 0.5);   // re-swept with tap-major (coalesced) atomics
  if (knob_ctas > 0) ctas_per_sm = knob_ctas;
  int splits = s2r_div_up((long)(ctas_per_sm * s2r_sm_count()), tiles);
  const int max_splits = s2r_div_up(nchunks, min_chunks);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.chunks_per_split = s2r_div_up(nchunks, splits);
  p.splits = s2r_div_up(nchunks, p.chunks_per_split);
  if (tiles * p.splits > 65535ll * 4) return 0;

  WgMaps maps;
  for (int i = 0; i < 4; ++i)
    if (!encode_view(enc, &maps.x[i], views[i < nviews ? i : 0], cin8, N, p.BW, p.BH)) return 0;
  if (!encode_view(enc, &maps.dy, dyv, cout8, N, p.BW, p.BH)) return 0;
  if (MT == 2) {
    switch (BN) {
      case 256: return launch_wg<256, 2, 3>(maps, p, st);
      case 160: return launch_wg<160, 2, 3>(maps, p, st);
      case 128: return launch_wg<128, 2, 4>(maps, p, st);
      case 64: return launch_wg<64, 2, 4>(maps, p, st);
      default: return launch_wg<32, 2, 4>(maps, p, st);
    }
  }
  switch (BN) {
    case 256: return launch_wg<256, 1, 4>(maps, p, st);
    case 160: return launch_wg<160, 1, 4>(maps, p, st);
    case 128: return launch_wg<128, 1, 6>(maps, p, st);
    case 64: return launch_wg<64, 1, 8>(maps, p, st);
    default: return launch_wg<32, 1, 8>(maps, p, st);
  }
}
