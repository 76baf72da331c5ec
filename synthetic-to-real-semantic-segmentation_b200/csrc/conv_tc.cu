// tcgen05 / TMA implicit-GEMM convolution for sm_100a (forward and data gradient of every dense
// convolution on the path: pointwise expand/project, ASPP, decoder, discriminator, domain classifier).
//
//   out[pix][co] = epi( sum_t sum_ci  view_t[pix + d_t][ci] * W[t][co][ci] )
//
// Mapping
//   * M tile = a BH x BW patch of output pixels (BH*BW = 128 rows = the 128 TMEM lanes); MT such
//     patches per CTA share every weight tile (MT = 2 halves the weight traffic per FLOP).
//   * A operand: one TMA box [1][BH][BW][64ch] per tap and 64-channel block, fetched from the tap's
//     strided NHWC view with the tap offset added to the box origin; out-of-view rows/columns/
//     channels are zero-filled by TMA, which implements padding, dilation halos, stride-2 parity
//     views and ragged edges without any address arithmetic on the SMs.  The box lands in shared
//     memory as 128 rows x 128 B with the 128B swizzle, i.e. the canonical K-major UMMA layout.
//   * B operand: TMA box [1][BN][64] of the packed bf16 weights [slice][Cout_pad][Kpad].
//   * D: fp32 accumulators in TMEM (MT*BN columns), tcgen05.mma.cta_group::1.kind::f16, M=128,
//     N=BN, K=16 per instruction, issued by one elected thread.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//     warps 4..15 = epilogue (TMEM -> registers via tcgen05.ld 32x32b, one pixel row per thread, three
//     warps per SM sub-partition so their dependent-instruction latencies interleave).
//   * Epilogue: bias, activation, residual add / LeakyReLU mask in registers; the 128 x 32 bf16 unit is staged
//     in shared memory (64B-swizzled rows) and written with ONE TMA store (full 64-byte segments per pixel,
//     ragged edges and channel tails clipped by the tensor map); per-channel BN statistics are column sums of
//     the staged (stored) values, then one fp64 atomic per channel per CTA.
//   * Persistent CTAs (one per SM) loop over tiles.  mbarrier pipelines: full[s] (TMA -> MMA, expect_tx),
//     empty[s] (tcgen05.commit -> TMA), acc_full[a] (last commit of a tile -> epilogue), acc_empty[a]
//     (epilogue -> MMA): with two TMEM accumulators the epilogue of tile i overlaps the mainloop of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TC_THREADS = 512;     // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-15: epilogue
constexpr int TC_EPI_WARPS = 12;
constexpr int TC_BK = 64;            // channels per pipeline stage (128 B of bf16: one swizzle row)
constexpr int TC_A_BYTES = 128 * 128;  // one 128-row A sub-tile

struct TcTap {
  int map;     // index of the tensor map (view) this tap reads
  int dh, dw;  // box origin offset
  int wslice;
};

struct TcParams {
  TcTap taps[S2R_MAX_TAPS];
  int ntaps;
  int kchunks;
  int N, OH, OW, Cout;
  int BW, BH, tiles_w, tiles_h;  // patch geometry
  int n_subtiles;
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

struct TcMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap out;   // output view [N][OH][OW][Cout], box [1][BH][BW][32], 64B swizzle (epilogue TMA stores)
};

constexpr int TC_STG_BYTES = 128 * 64;                 // one epilogue unit: 128 pixel rows x 32 bf16 channels
// NSTG staging buffers per epilogue warp group.  Two: a unit is staged while the TMA store of the previous one is still
// reading its buffer, and the per-unit chain has ONE named barrier instead of two and no wait for the store's read
// (the shallow pointwise convolutions are bound by exactly that chain).  One where the shared memory is needed for the
// operand pipeline of the tensor-bound shapes (MT = 2 with wide tiles).
constexpr int tc_stg_total(int nstg) { return 3 * nstg * TC_STG_BYTES; }

// ------------------------------------------------------------------------------------ PTX wrappers
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand, 128-byte swizzle: 8-row groups are 1024 B apart (SBO), version 1, layout type 2
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
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

// ------------------------------------------------------------------------------------ kernel
constexpr int TC_MAX_COUT = 1024;  // per-CTA statistics staging (channels)

// Persistent: one CTA per SM walks output tiles (tile = m_tile * n_tiles + n_tile, n fastest so that CTAs
// running side by side share the A boxes in L2).  The TMA producer runs ahead across tile boundaries, the
// MMA issuer alternates between NACC TMEM accumulators, the epilogue warps drain one accumulator while the
// next tile is being multiplied.
template <int BN, int MT, int STAGES, int NSTG>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = MT * TC_A_BYTES + B_BYTES;
  constexpr int ACC_COLS = MT * BN;
  constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
  constexpr int TCOLS = tmem_cols(NACC * ACC_COLS);
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_stg = smem + (size_t)STAGES * STAGE_BYTES;   // 1024-aligned: STAGE_BYTES is a multiple of 1024
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc_full[NACC], bar_acc_empty[NACC];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_sum[TC_MAX_COUT], s_sq[TC_MAX_COUT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters = p.ntaps * p.kchunks;
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int n_tiles_m = (p.n_subtiles + MT - 1) / MT;
  const int total_tiles = n_tiles_m * n_tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_addr(&bar_full[s]), 1);
      mbar_init(smem_addr(&bar_empty[s]), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(smem_addr(&bar_acc_full[a]), 1);
      mbar_init(smem_addr(&bar_acc_empty[a]), TC_EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&maps.a[i]);
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.out);
  }
  if (warp == 2) tmem_alloc<TCOLS>(smem_addr(&tmem_base_slot));
  if (p.stats)
    for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) {
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory only from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles_n, n0 = (tile - mt * n_tiles_n) * BN;
        int t_n[MT], t_h[MT], t_w[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          int t = mt * MT + j;
          if (t >= p.n_subtiles) t = p.n_subtiles - 1;  // duplicate load, rows masked in the epilogue
          t_w[j] = (t % p.tiles_w) * p.BW;
          const int q = t / p.tiles_w;
          t_h[j] = (q % p.tiles_h) * p.BH;
          t_n[j] = q / p.tiles_h;
        }
        for (int k = 0; k < kiters; ++k, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_addr(&bar_empty[s]), ph ^ 1);
          const int tap = k / p.kchunks, kc = k - tap * p.kchunks;
          const TcTap T = p.taps[tap];
          const uint32_t full = smem_addr(&bar_full[s]);
          const uint32_t sa = smem_addr(smem + (size_t)s * STAGE_BYTES);
          mbar_expect_tx(full, STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            tma_load_4d(sa + j * TC_A_BYTES, &maps.a[T.map], full, kc * TC_BK, t_w[j] + T.dw, t_h[j] + T.dh, t_n[j]);
          tma_load_3d(sa + MT * TC_A_BYTES, &maps.b, full, kc * TC_BK, n0, T.wslice);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int ab = lt % NACC;
        const uint32_t aph = (lt / NACC) & 1;
        mbar_wait(smem_addr(&bar_acc_empty[ab]), aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + ab * ACC_COLS;
        for (int k = 0; k < kiters; ++k, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_addr(&bar_full[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_addr(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + MT * TC_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            const uint64_t db = umma_desc_k128(sb + kk * 32);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              const uint64_t da = umma_desc_k128(sa + j * TC_A_BYTES + kk * 32);
              umma_bf16(acc + j * BN, da, db, idesc, (k > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_addr(&bar_empty[s]));  // frees the stage when these MMAs retire
        }
        umma_commit(smem_addr(&bar_acc_full[ab]));
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: 12 warps = 3 groups of 4 (128 TMEM lanes
    // each); the (sub-tile, 32-column chunk) units of a tile are dealt round-robin to the groups so that
    // every SM sub-partition interleaves three epilogue warps
    const int ew = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;         // 0..2
    const int row = ew * 32 + lane;
    constexpr int CHUNKS = BN / 32;
    const int act = p.act;
    const int aux_mode = p.aux_mode;
    const int row_h = row / p.BW, row_w = row - row_h * p.BW;   // this thread's pixel inside a patch
    const bool one_n = n_tiles_n == 1;                            // Cout <= BN: tile index = pixel tile
    const bool one_row = p.tiles_h == 1 && p.N == 1;              // flattened pointwise problem: one long pixel row
    int lt = 0, rot = grp;                                        // rot = (grp + 3 - lt % 3) % 3, kept incrementally
    int cnt = 0;                                                  // units this warp group has staged (buffer parity)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt, rot = rot == 0 ? 2 : rot - 1) {
      const int mt = one_n ? tile : tile / n_tiles_n, n0 = one_n ? 0 : (tile - mt * n_tiles_n) * BN;
      const int ab = lt % NACC;
      const uint32_t aph = (lt / NACC) & 1;
      // patch coordinates of the tile's sub-tiles: once per tile, not per unit, and without integer divisions in the
      // common cases (they were a quarter of the epilogue's dependent-instruction chain on the pointwise convolutions)
      int s_tw[MT], s_th[MT], s_tn[MT], s_oh[MT], s_ow[MT];
      bool s_tok[MT], s_rok[MT];
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int t = mt * MT + j;
        s_tok[j] = t < p.n_subtiles;
        const int tt = s_tok[j] ? t : 0;
        const int q = one_row ? 0 : tt / p.tiles_w;
        s_tw[j] = (tt - q * p.tiles_w) * p.BW;
        s_tn[j] = one_row ? 0 : q / p.tiles_h;
        s_th[j] = (q - s_tn[j] * p.tiles_h) * p.BH;
        s_oh[j] = s_th[j] + row_h;
        s_ow[j] = s_tw[j] + row_w;
        s_rok[j] = s_tok[j] && s_oh[j] < p.OH && s_ow[j] < p.OW;
      }
      mbar_wait(smem_addr(&bar_acc_full[ab]), aph);
      tc_fence_after();
      bool released = false;
      // units are dealt round-robin to the three groups, rotating with the tile so that narrow tiles
      // (fewer than three units) still use all epilogue warps
#pragma unroll 1
      for (int u = rot; u < MT * CHUNKS; u += 3) {
        const int j = (MT == 1) ? 0 : u / CHUNKS, c0 = (u - j * CHUNKS) * 32;
        if (n0 + c0 >= p.Cout) continue;
        static_assert(MT <= 2, "sub-tile select below");
#define TC_SEL(a) ((MT == 1 || j == 0) ? a[0] : a[MT - 1])   // register select, no local-memory indexing
        const bool tok = TC_SEL(s_tok);
        const int tw_ = TC_SEL(s_tw), th_ = TC_SEL(s_th), tn_ = TC_SEL(s_tn);
        const int oh_ = TC_SEL(s_oh), ow_ = TC_SEL(s_ow);
        const bool rok = TC_SEL(s_rok);
#undef TC_SEL
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * ACC_COLS + j * BN + c0), r);
        tmem_ld_wait();
        if (u + 3 >= MT * CHUNKS) {
          // last TMEM read of this warp in this tile: hand the accumulator back before the staging / store /
          // statistics part, so that the MMAs of the tile after next do not wait for it
          released = true;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar_acc_empty[ab])) : "memory");
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const bool full = n0 + c0 + 32 <= p.Cout;   // warp-uniform
        if (p.oscale) {   // eval-mode BatchNorm folded into the convolution: acc * scale[co] (+ bias below = shift)
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.oscale + n0 + c0 + i));
              v[i] *= s4.x; v[i + 1] *= s4.y; v[i + 2] *= s4.z; v[i + 3] *= s4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c0 + i < p.Cout) v[i] *= __ldg(p.oscale + n0 + c0 + i);
          }
        }
        if (p.bias) {
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + i));
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c0 + i < p.Cout) v[i] += __ldg(p.bias + n0 + c0 + i);
          }
        }
        // activation / auxiliary operand on the fp32 values of this thread's pixel row
        if (aux_mode == S2R_AUX_NONE) {
          if (act == S2R_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (act == S2R_ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.slope;
          } else if (act == S2R_ACT_RELU6) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 6.f);
          }
        } else if (rok) {
          const __nv_bfloat16* asrc = p.aux + tn_ * p.an + oh_ * p.ah + ow_ * p.aw + n0 + c0;
          if (full && (reinterpret_cast<uintptr_t>(asrc) & 15) == 0) {
            float av[32];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) bf16x8_to_float(ldg16(asrc + qq * 8), av + qq * 8);
            if (aux_mode == S2R_AUX_ADD) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += av[i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= (av[i] > 0.f ? 1.f : p.slope);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c0 + i < p.Cout) {
                const float av = __bfloat162float(asrc[i]);
                v[i] = aux_mode == S2R_AUX_ADD ? v[i] + av : v[i] * (av > 0.f ? 1.f : p.slope);
              }
          }
        }
        if (!rok) {   // rows outside the output (ragged patches, duplicated sub-tile): no statistics, store clipped
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        uint4 pk[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) pk[qq] = float_to_bf16x8(v + qq * 8);
        // stage the 128 x 32 unit in shared memory (64-byte rows, 16-byte chunks XOR-swizzled with (row >> 1) & 3:
        // the layout of CU_TENSOR_MAP_SWIZZLE_64B, conflict-free for these stores) and write it out with one TMA
        // store: full 64-byte segments per pixel instead of 32 scattered 16-byte stores per instruction
        uint8_t* stg = smem_stg + (grp * NSTG + (NSTG == 2 ? (cnt & 1) : 0)) * TC_STG_BYTES;
        ++cnt;
        if (NSTG == 1) {
          if (row == 0) bulk_wait_read0();        // the previous store of this group has finished reading the buffer
          named_bar_sync(1 + grp, 128);
        }
        // NSTG == 2: this buffer was last used two units ago; row 0 waited for that unit's store before the barrier of
        // the unit in between (below), and every thread finished reading its statistics before arriving there
        {
          uint8_t* rp = stg + row * 64;
          const int sw = (row >> 1) & 3;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) *reinterpret_cast<uint4*>(rp + ((qq ^ sw) << 4)) = pk[qq];
        }
        fence_proxy_async();
        if (NSTG == 2 && row == 0) bulk_wait_read0();   // the previous unit's store (the other buffer) has been read
        named_bar_sync(1 + grp, 128);
        if (row == 0 && tok) {
          tma_store_4d(&maps.out, smem_addr(stg), n0 + c0, tw_, th_, tn_);
          bulk_commit();
        }
        if (p.stats) {
          // per-channel sum / sum of squares of the STORED (bf16) values.  thread = (channel pair, row class): warp ew
          // owns the channel pairs 4*ew..4*ew+3, lane = (row class r mod 8, pair), so the 8 partial sums of a channel
          // meet inside ONE warp (3 shuffle steps) and a single lane per channel touches the CTA accumulators.
          // Shared-memory float atomics are CAS loops (ATOMS.CAST.SPIN): with one partial per (channel, 16-row
          // segment) they ran 8-way contended and dominated the epilogue's latency chain.  Rows 8*rr + cls of one
          // load instruction sit in 32 distinct banks (64-byte rows, chunk XOR (row >> 1) & 3).
          const int pr = ew * 4 + (lane & 3), cls = lane >> 2;
          const uint32_t stg_s = smem_addr(stg);
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r2 = rr * 8 + cls;
            uint32_t wv;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv) : "r"(stg_s + (uint32_t)(r2 * 64 + ((((pr >> 2) ^ ((r2 >> 1) & 3))) << 4) + (pr & 3) * 4)));
            const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xffff0000u);
            s0 += f0; s1 += f1;
            q0 = fmaf(f0, f0, q0); q1 = fmaf(f1, f1, q1);
          }
#pragma unroll
          for (int sh = 4; sh < 32; sh <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, sh);
            s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
            q0 += __shfl_xor_sync(0xffffffffu, q0, sh);
            q1 += __shfl_xor_sync(0xffffffffu, q1, sh);
          }
          // all 8 lanes of a channel pair now hold the 4 totals: lanes cls = 0..3 add one of them each (one CAS loop
          // per thread instead of four in sequence)
          const int ch = n0 + c0 + 2 * pr + (cls >> 1);
          const float tv = cls == 0 ? s0 : cls == 1 ? q0 : cls == 2 ? s1 : q1;
          if (cls < 4 && ch < p.Cout) atomicAdd(((cls & 1) ? s_sq : s_sum) + ch, tv);
        }
      }
      if (!released) {   // no unit, or the last one was skipped: all TMEM reads of this accumulator are complete
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar_acc_empty[ab])) : "memory");
      }
    }
    if (row == 0) bulk_wait0();   // this group's TMA stores are complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) {
      atomicAdd(&p.stats[i], (double)s_sum[i]);
      atomicAdd(&p.stats[p.Cout + i], (double)s_sq[i]);
    }
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------ host side
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

struct ViewKey {
  const void* base;
  long long sn, sh, sw;
  int H, W;
  bool operator==(const ViewKey& o) const {
    return base == o.base && sn == o.sn && sh == o.sh && sw == o.sw && H == o.H && W == o.W;
  }
};

bool encode_view(EncodeTiledFn enc, CUtensorMap* m, const ViewKey& v, int C, int N, int BW, int BH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)BW, (cuuint32_t)BH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  // dimensions of extent 1 still need a non-zero 16-byte multiple stride
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = 16;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool encode_weights(EncodeTiledFn enc, CUtensorMap* m, const void* w, int Kpad, int Cout_pad, int nslices, int BN) {
  cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)Cout_pad, (cuuint64_t)nslices};
  cuuint64_t strides[2] = {(cuuint64_t)Kpad * 2, (cuuint64_t)Kpad * 2 * (cuuint64_t)Cout_pad};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)BN, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int MT, int STAGES, int NSTG>
int launch_tc(const TcMaps& maps, const TcParams& p, int n_tiles_n, cudaStream_t st) {
  constexpr int smem = STAGES * (MT * TC_A_BYTES + BN * 128) + tc_stg_total(NSTG) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    S2R_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BN, MT, STAGES, NSTG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long long tiles = (long long)s2r_div_up(p.n_subtiles, MT) * n_tiles_n;
  const int grid = (int)(tiles < s2r_sm_count() ? tiles : s2r_sm_count());  // persistent: one CTA per SM
  S2R_CUDA_OK(s2r_launch(conv_tc_kernel<BN, MT, STAGES, NSTG>, dim3(grid), dim3(TC_THREADS), (size_t)smem, st, maps, p));
  S2R_LAUNCH_OK();
  return 1;
}

int tc_mode() {
  // S2R_CONV=mma forces the mma.sync path (A/B testing); default: tcgen05 where supported
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("S2R_CONV");
    mode = (e && e[0] == 'm') ? 0 : 1;
  }
  return mode;
}

}  // namespace

// returns 1 when the problem was launched here, 0 when the caller should use the generic path, <0 on error
int s2r_conv_fwd_tc(const s2r_conv_args* a, cudaStream_t st) {
  if (!tc_mode()) return 0;
  EncodeTiledFn enc = get_encode();
  if (!enc) return 0;
  if (a->ntaps < 1 || a->ntaps > S2R_MAX_TAPS || a->Cin % 8 || a->Cin < 8) return 0;
  if (a->Kpad % TC_BK || a->Cout_pad % 16 || (uintptr_t)a->w % 16 || a->Kpad < a->Cin) return 0;
  if (a->Cout < 1 || a->N < 1 || a->OH < 1 || a->OW < 1) return 0;
  if (a->stats && a->Cout > TC_MAX_COUT) return 0;

  // distinct views among the taps (<= 4: one per input parity)
  ViewKey views[4];
  int nviews = 0;
  TcParams p;
  int max_slice = 0;
  for (int i = 0; i < a->ntaps; ++i) {
    const s2r_tap& t = a->taps[i];
    if (!t.base || (uintptr_t)t.base % 16 || t.sn % 8 || t.sh % 8 || t.sw % 8 || t.H < 1 || t.W < 1) return 0;
    ViewKey k = {t.base, t.sn, t.sh, t.sw, t.H, t.W};
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
    p.taps[i].wslice = t.wslice;
    if (t.wslice > max_slice) max_slice = t.wslice;
  }
  // pointwise problem on pixel-contiguous tensors: flatten the pixel grid to one long row so that every
  // 128-row tile is full
  int N = a->N, OH = a->OH, OW = a->OW;
  long long on = a->on, oh = a->oh, ow = a->ow, an = a->an, ah = a->ah, aw = a->aw;
  if (a->ntaps == 1 && nviews == 1 && a->taps[0].dh == 0 && a->taps[0].dw == 0 && views[0].H == OH && views[0].W == OW) {
    const ViewKey& v = views[0];
    const bool in_flat = v.sh == v.sw * OW && v.sn == v.sh * OH;
    const bool out_flat = oh == ow * OW && on == oh * OH;
    const bool aux_flat = a->aux_mode == S2R_AUX_NONE || (ah == aw * OW && an == ah * OH);
    if (in_flat && out_flat && aux_flat && (long long)N * OH * OW < (1ll << 31)) {
      OW = N * OH * OW;
      OH = 1;
      N = 1;
      views[0].W = OW;
      views[0].H = 1;
      views[0].sh = views[0].sw * OW;
      views[0].sn = views[0].sh;
      oh = ow * OW;
      on = oh;
      ah = aw * OW;
      an = ah;
    }
  }
  // patch shape: BW*BH = 128, minimise the number of (partially empty) patches
  int bestBW = 128;
  long long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const int bh = 128 / bw;
    const long long tiles = (long long)s2r_div_up(OW, bw) * s2r_div_up(OH, bh);
    if (best < 0 || tiles < best) {
      best = tiles;
      bestBW = bw;
    }
  }
  p.BW = bestBW;
  p.BH = 128 / bestBW;
  p.tiles_w = s2r_div_up(OW, p.BW);
  p.tiles_h = s2r_div_up(OH, p.BH);
  const long long nsub = (long long)N * p.tiles_w * p.tiles_h;
  if (nsub >= (1ll << 30)) return 0;
  p.n_subtiles = (int)nsub;
  p.ntaps = a->ntaps;
  p.kchunks = s2r_div_up(a->Cin, TC_BK);
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = a->Cout;
  p.out = (__nv_bfloat16*)a->out;
  p.on = on; p.oh = oh; p.ow = ow;
  p.bias = a->bias; p.act = a->act; p.slope = a->slope;
  p.oscale = a->oscale;
  if ((a->bias && (uintptr_t)a->bias % 16) || (a->oscale && (uintptr_t)a->oscale % 16)) return 0;   // float4 loads
  p.aux_mode = a->aux_mode;
  p.aux = a->aux_mode == S2R_AUX_NONE ? nullptr : (const __nv_bfloat16*)a->aux;
  p.an = an; p.ah = ah; p.aw = aw;
  p.stats = a->stats;

  // tile width in output channels
  int BN;
  if (a->Cout > 256 && a->Cout <= 320) BN = 160;   // 257..320 channels: two 160-wide tiles instead of 256 + a sliver
  else if (a->Cout > 128) BN = 256;
  else if (a->Cout > 64) BN = 128;
  else if (a->Cout > 32) BN = 64;
  else BN = 32;
  if (a->Cout_pad < BN && a->Cout_pad % BN) {
    // the weight box may run past Cout_pad: TMA zero-fills, nothing to do
  }
  TcMaps maps;
  for (int i = 0; i < 4; ++i) {
    const ViewKey& v = views[i < nviews ? i : 0];
    if (!encode_view(enc, &maps.a[i], v, a->Cin, N, p.BW, p.BH)) {
      s2r_set_error("conv_tc: cuTensorMapEncodeTiled failed for view %d", i);
      return 0;
    }
  }
  if (!encode_weights(enc, &maps.b, a->w, a->Kpad, a->Cout_pad, max_slice + 1, BN)) return 0;
  {
    // output view for the epilogue's TMA stores; needs 16-byte aligned base and strides
    if ((uintptr_t)a->out % 16 || on % 8 || oh % 8 || ow % 8) return 0;
    cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ow * 2, (cuuint64_t)oh * 2, (cuuint64_t)on * 2};
    for (int i = 0; i < 3; ++i)
      if (strides[i] == 0) strides[i] = 16;
    cuuint32_t box[4] = {32u, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&maps.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      s2r_set_error("conv_tc: cuTensorMapEncodeTiled failed for the output view");
      return 0;
    }
  }
  const int ntn = s2r_div_up(a->Cout, BN);
  // two M sub-tiles per CTA (weights fetched once per 256 pixels) when the contraction is deep enough to be
  // tensor-bound and the problem still fills the machine; otherwise one sub-tile and two TMEM accumulators so
  // that the epilogue of a tile overlaps the loads and MMAs of the next
  const bool deep = (long long)p.ntaps * p.kchunks >= 8;
  const bool big = deep && nsub * ntn >= 2ll * s2r_sm_count() * 2;
  switch (BN) {
    // stages x staging buffers within the 227 KB of shared memory: the shallow (MT = 1) shapes, whose epilogue is the
    // bottleneck, give up one or two operand stages for the second staging buffer
    case 256: return big ? launch_tc<256, 2, 3, 1>(maps, p, ntn, st) : launch_tc<256, 1, 3, 2>(maps, p, ntn, st);
    case 160: return big ? launch_tc<160, 2, 3, 2>(maps, p, ntn, st) : launch_tc<160, 1, 4, 2>(maps, p, ntn, st);
    case 128: return big ? launch_tc<128, 2, 4, 1>(maps, p, ntn, st) : launch_tc<128, 1, 5, 2>(maps, p, ntn, st);
    case 64: return big ? launch_tc<64, 2, 4, 2>(maps, p, ntn, st) : launch_tc<64, 1, 6, 2>(maps, p, ntn, st);
    default: return launch_tc<32, 1, 8, 2>(maps, p, ntn, st);
  }
}
