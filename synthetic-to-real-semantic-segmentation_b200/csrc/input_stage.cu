// Device input stage: the reference's per-sample CPU pipeline
//   dataloders/datasets/gtav2cityscapes.py:76-83   encode_segmap (labelId -> trainId table)
//   dataloders/custom_transforms.py:59-71          RandomHorizontalFlip
//   dataloders/custom_transforms.py:108-147        RandomScaleCrop (PIL resize, pad right/bottom, crop window)
//   dataloders/custom_transforms.py:92-105         RandomGaussianBlur (PIL GaussianBlur on the crop)
//   dataloders/custom_transforms.py:17-56          Normalize + ToTensor
// as byte kernels on uint8 HWC images already resident in HBM.  Everything is bit-exact against the reference:
//   * PIL's BILINEAR resize is a two-pass fixed-point convolution (libImaging/Resample.c: int32 accumulation of
//     uint8 * 22-bit coefficients starting at 1 << 21, arithmetic shift, clip, uint8 intermediate between the passes);
//     the coefficient / bounds tables are computed by the host mirror exactly as precompute_coeffs does and passed in;
//   * PIL's NEAREST resize copies in[ytab[y]][xtab[x]] with tables from incremental double additions (Geometry.c);
//   * PIL's GaussianBlur(radius) is three box-blur passes per axis (libImaging/BoxBlur.c) with a fractional box radius;
//     for the reference's radii (< 1) the integer part of that radius is 0, i.e. every pass is the 3-tap fixed-point
//     filter (c*ww + (l + r)*fw + 2^23) >> 24 in uint32 with the line ends replicated and a uint8 result per pass;
//     ww / fw (24-bit weights, derived in single precision) come from the host mirror;
//   * Normalize runs /255 in float32 and (-mean), (/std) in float64 rounded to float32 (numpy's casting of the tuple
//     operands): a 3 x 256 table built per CTA with exactly those operations.
// All kernels are HBM-bound byte movers: coalesced reads along the pixel row, coalesced fp32 plane writes.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[n][y][xx][c] = clip8(2^21 + sum_x in[n][y][src(xmin + x)][c] * k[xx][x]); src mirrors the column when flip != 0
// (the reference flips the PIL image before it is resized)
__global__ void __launch_bounds__(kThreads)
resize_h_kernel(const uint8_t* __restrict__ in, int N, int H, int W, int C, int OW, const int* __restrict__ bounds,
                const int* __restrict__ kk, int ksize, int flip, uint8_t* __restrict__ out) {
  const long long total = (long long)N * H * OW * C;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int xx = (int)(q % OW);
    q /= OW;   // q = n * H + y
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const uint8_t* row = in + q * (long long)W * C + c;
    const int* k = kk + (long long)xx * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < cnt; ++x) {
      const int sx = flip ? (W - 1 - (xmin + x)) : (xmin + x);
      acc += (int)row[(long long)sx * C] * __ldg(k + x);
    }
    out[i] = clip8(acc);
  }
}

// out[n][yy][x][c] = clip8(2^21 + sum_y in[n][ymin + y][x][c] * k[yy][y]); rows are contiguous: thread = byte of a row
__global__ void __launch_bounds__(kThreads)
resize_v_kernel(const uint8_t* __restrict__ in, int N, int H, int WC, int OH, const int* __restrict__ bounds,
                const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
  const long long total = (long long)N * OH * WC;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % WC);
    long long q = i / WC;
    const int yy = (int)(q % OH);
    const int n = (int)(q / OH);
    const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
    const uint8_t* col = in + ((long long)n * H + ymin) * WC + x;
    const int* k = kk + (long long)yy * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < cnt; ++y) acc += (int)col[(long long)y * WC] * __ldg(k + y);
    out[i] = clip8(acc);
  }
}

__global__ void __launch_bounds__(kThreads)
resize_nearest_kernel(const uint8_t* __restrict__ in, int N, int H, int W, const int* __restrict__ xtab,
                      const int* __restrict__ ytab, int OH, int OW, int flip, uint8_t* __restrict__ out) {
  const long long total = (long long)N * OH * OW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % OW);
    long long q = i / OW;
    const int y = (int)(q % OH);
    const int n = (int)(q / OH);
    const int sx = __ldg(xtab + x), sy = __ldg(ytab + y);
    uint8_t v = 0;   // ImagingScaleAffine clears the row first (fill = 1)
    if (sx >= 0 && sy >= 0) v = in[((long long)n * H + sy) * W + (flip ? W - 1 - sx : sx)];
    out[i] = v;
  }
}

struct NormParams {
  double mean[3], std[3];
};

// Crop window (x1, y1) of the (flipped, zero / fill padded) image -> normalised fp32 CHW planes and fp32 labels.
// grid = (ceil(W / kThreads), H, N)
__global__ void __launch_bounds__(kThreads)
input_stage_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ label, int Hs, int Ws, int flip, int x1,
                   int y1, NormParams np, const uint8_t* __restrict__ lut, int fill_label, float* __restrict__ out_img,
                   float* __restrict__ out_label, int H, int W) {
  __shared__ float tab[3][256];
  __shared__ float ltab[256];
  for (int i = threadIdx.x; i < 768; i += kThreads) {
    const int c = i >> 8, u = i & 255;
    const float v = __fdiv_rn((float)u, 255.0f);           // float32 array /= 255.0
    const float a = (float)((double)v - np.mean[c]);        // float64 loop, result stored as float32
    tab[c][u] = (float)((double)a / np.std[c]);
  }
  for (int i = threadIdx.x; i < 256; i += kThreads) ltab[i] = (float)(lut ? lut[i] : (uint8_t)i);
  __syncthreads();
  const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= W) return;
  const int sy = y1 + y, sx0 = x1 + x;
  const bool inside = sy < Hs && sx0 < Ws;                  // pad is on the right / bottom only (ImageOps.expand border)
  const int sx = flip ? Ws - 1 - sx0 : sx0;
  const long long plane = (long long)H * W;
  const long long o = (long long)y * W + x;
  if (img) {
    uint8_t r = 0, g = 0, b = 0;                            // fill = 0 for the image, before normalisation
    if (inside) {
      const uint8_t* p = img + (((long long)n * Hs + sy) * Ws + sx) * 3;
      r = p[0]; g = p[1]; b = p[2];
    }
    float* oi = out_img + (long long)n * 3 * plane + o;
    oi[0] = tab[0][r];
    oi[plane] = tab[1][g];
    oi[2 * plane] = tab[2][b];
  }
  if (label) {
    float v = (float)fill_label;
    if (inside) v = ltab[label[((long long)n * Hs + sy) * Ws + sx]];
    out_label[(long long)n * plane + o] = v;
  }
}


// ---- batched variants: one launch per pass for a whole batch whose samples have different scaled sizes.  blockIdx.y
// walks a device table of per-sample jobs (the arithmetic per element is that of the single-sample kernels above).
__global__ void __launch_bounds__(kThreads)
resize_multi_kernel(const s2r_resize_job* __restrict__ jobs) {
  // A job produces the WINDOW [o0, o0 + on) of the resampled axis for `lines` lines of the other axis (only the crop
  // window of the scaled image is ever used); in points at line 0 / source index `base` of the job's input.
  const s2r_resize_job j = jobs[blockIdx.y];
  const int C = j.C;
  if (j.axis == 1) {
    // columns: in = first needed source row, rows of W*C bytes; out [lines][on][C]
    const long long total = (long long)j.lines * j.on * C;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
      const int c = (int)(i % C);
      long long q = i / C;
      const int xx = j.o0 + (int)(q % j.on);
      q /= j.on;   // line (source row relative to in)
      const int xmin = __ldg(j.bounds + 2 * xx), cnt = __ldg(j.bounds + 2 * xx + 1);
      const uint8_t* row = j.in + q * (long long)j.in_pitch + c;
      const int* k = j.kk + (long long)xx * j.ksize;
      int acc = 1 << (PRECISION_BITS - 1);
      for (int x = 0; x < cnt; ++x) {
        const int sx = j.flip ? (j.W - 1 - (xmin + x)) : (xmin + x);
        acc += (int)row[(long long)sx * C] * __ldg(k + x);
      }
      j.out[i] = clip8(acc);
    }
  } else {
    // rows: in = byte 0 of source row `base`, lines = bytes per output row (window columns * C); out [on][lines]
    const long long total = (long long)j.on * j.lines;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
      const int x = (int)(i % j.lines);
      const int yy = j.o0 + (int)(i / j.lines);
      const int ymin = __ldg(j.bounds + 2 * yy), cnt = __ldg(j.bounds + 2 * yy + 1);
      const uint8_t* col = j.in + (long long)(ymin - j.base) * j.in_pitch + x;
      const int* k = j.kk + (long long)yy * j.ksize;
      int acc = 1 << (PRECISION_BITS - 1);
      for (int y = 0; y < cnt; ++y) acc += (int)col[(long long)y * j.in_pitch] * __ldg(k + y);
      j.out[i] = clip8(acc);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
nearest_multi_kernel(const s2r_nearest_job* __restrict__ jobs) {
  // window [y0, y0 + OH) x [x0, x0 + OW) of the resized label map
  const s2r_nearest_job j = jobs[blockIdx.y];
  const long long total = (long long)j.OH * j.OW;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int x = (int)(i % j.OW), y = (int)(i / j.OW);
    const int sx = __ldg(j.xtab + j.x0 + x), sy = __ldg(j.ytab + j.y0 + y);
    uint8_t v = 0;
    if (sx >= 0 && sy >= 0) v = j.in[(long long)sy * j.W + (j.flip ? j.W - 1 - sx : sx)];
    j.out[i] = v;
  }
}

// grid = (ceil(W / kThreads), H, njobs)
__global__ void __launch_bounds__(kThreads)
input_stage_multi_kernel(const s2r_stage_job* __restrict__ jobs, NormParams np, const uint8_t* __restrict__ lut,
                         int fill_label, int H, int W) {
  __shared__ float tab[3][256];
  __shared__ float ltab[256];
  const s2r_stage_job j = jobs[blockIdx.z];
  for (int i = threadIdx.x; i < 768; i += kThreads) {
    const int c = i >> 8, u = i & 255;
    const float v = __fdiv_rn((float)u, 255.0f);
    const float a = (float)((double)v - np.mean[c]);
    tab[c][u] = (float)((double)a / np.std[c]);
  }
  for (int i = threadIdx.x; i < 256; i += kThreads) ltab[i] = (float)(lut ? lut[i] : (uint8_t)i);
  __syncthreads();
  const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int sy = j.y1 + y, sx0 = j.x1 + x;
  const bool inside = sy < j.Hs && sx0 < j.Ws;
  const int sx = j.flip ? j.Ws - 1 - sx0 : sx0;
  const long long plane = (long long)H * W;
  const long long o = (long long)y * W + x;
  if (j.img) {
    uint8_t r = 0, g = 0, b = 0;
    if (inside) {
      const uint8_t* p = j.img + ((long long)sy * j.Ws + sx) * 3;
      r = p[0]; g = p[1]; b = p[2];
    }
    float* oi = j.out_img + o;
    oi[0] = tab[0][r];
    oi[plane] = tab[1][g];
    oi[2 * plane] = tab[2][b];
  }
  if (j.label) {
    float v = (float)fill_label;
    if (inside) v = ltab[j.label[(long long)sy * j.Ws + sx]];
    j.out_label[o] = v;
  }
}

// ---- RandomGaussianBlur: the crop (cut from the scaled image exactly as input_stage_multi_kernel cuts it: mirror,
// zero padding on the right / bottom) goes through three 3-tap passes along x (blur_rows_kernel -> tmp) and three
// along y (blur_cols_kernel -> out).  P_0 = the line, P_{k+1}(j) = tap(P_k(max(j-1, 0)), P_k(j), P_k(min(j+1, last))).
// A thread produces K CONSECUTIVE positions of a line: it loads the K + 6 source bytes they depend on once and
// evaluates the three passes on shrinking register windows (K + 4, K + 2, K values), so a byte costs (3K + 6) / K taps
// and (K + 6) / K loads -- the first version evaluated P_3 at one position by recursion (13 taps, 27 loads per byte,
// 0.46 ms for the sixteen 512x512 crops of a batch).  The replicated line ends enter through selects on the absolute
// position: window entries that lie off the line are never read.  No intermediate buffer between the passes of an axis.
__device__ __forceinline__ uint32_t blur_tap(uint32_t l, uint32_t c, uint32_t r, uint32_t ww, uint32_t fw) {
  return (c * ww + (l + r) * fw + (1u << 23)) >> 24;
}

// o[u] = P_3(x0 + u), u < K (positions past `last` produce unused values)
template <int K, class Line>
__device__ __forceinline__ void blur_three_passes(Line p0, int x0, int last, uint32_t ww, uint32_t fw, uint8_t* o) {
  uint32_t b0[K + 6], b1[K + 4], b2[K + 2];
#pragma unroll
  for (int i = 0; i < K + 6; ++i) {            // b0[i] = P_0(clamp(x0 - 3 + i))
    const int j = x0 - 3 + i;
    b0[i] = p0(j < 0 ? 0 : (j > last ? last : j));
  }
#pragma unroll
  for (int s = 0; s < K + 4; ++s) b1[s] = blur_tap(b0[s], b0[s + 1], b0[s + 2], ww, fw);   // P_1(x0 - 2 + s) where on the line
#pragma unroll
  for (int t = 0; t < K + 2; ++t) {            // P_2(x0 - 1 + t) where on the line
    const int j = x0 - 1 + t;
    const uint32_t l = j <= 0 ? b1[t + 1] : b1[t], r = j >= last ? b1[t + 1] : b1[t + 2];
    b2[t] = blur_tap(l, b1[t + 1], r, ww, fw);
  }
#pragma unroll
  for (int u = 0; u < K; ++u) {
    const int x = x0 + u;
    const uint32_t l = x <= 0 ? b2[u + 1] : b2[u], r = x >= last ? b2[u + 1] : b2[u + 2];
    o[u] = (uint8_t)blur_tap(l, b2[u + 1], r, ww, fw);
  }
}

constexpr int BLUR_KX = 4;   // pixels (x 3 channels) per thread of the row pass
constexpr int BLUR_KY = 8;   // rows per thread of the column pass

// grid = (blocks over H*W*3 bytes of the crop, njobs)
__global__ void __launch_bounds__(kThreads)
blur_rows_kernel(const s2r_blur_job* __restrict__ jobs, int H, int W) {
  const s2r_blur_job j = jobs[blockIdx.y];
  const int G = (W + BLUR_KX - 1) / BLUR_KX, total = H * G;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const int x0 = (i % G) * BLUR_KX, y = i / G;
    const int sy = j.y1 + y;
    const bool row_inside = sy < j.Hs;
    for (int c = 0; c < 3; ++c) {
      const uint8_t* row = j.img + (long long)sy * j.Ws * 3 + c;
      auto px = [&](int xx) -> uint32_t {
        const int sx0 = j.x1 + xx;
        if (!row_inside || sx0 >= j.Ws) return 0u;             // ImageOps.expand(border, fill = 0) before the crop
        return row[(j.flip ? j.Ws - 1 - sx0 : sx0) * 3];
      };
      uint8_t o[BLUR_KX];
      blur_three_passes<BLUR_KX>(px, x0, W - 1, j.ww, j.fw, o);
      for (int u = 0; u < BLUR_KX; ++u)
        if (x0 + u < W) j.tmp[((long long)y * W + x0 + u) * 3 + c] = o[u];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
blur_cols_kernel(const s2r_blur_job* __restrict__ jobs, int H, int W) {
  const s2r_blur_job j = jobs[blockIdx.y];
  const int pitch = W * 3, total = ((H + BLUR_KY - 1) / BLUR_KY) * pitch;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const int xb = i % pitch, y0 = (i / pitch) * BLUR_KY;
    const uint8_t* col = j.tmp + xb;
    auto px = [&](int yy) -> uint32_t { return col[(long long)yy * pitch]; };
    uint8_t o[BLUR_KY];
    blur_three_passes<BLUR_KY>(px, y0, H - 1, j.ww, j.fw, o);
    for (int u = 0; u < BLUR_KY; ++u)
      if (y0 + u < H) j.out[(long long)(y0 + u) * pitch + xb] = o[u];
  }
}

}  // namespace

extern "C" int s2r_gaussian_blur3_u8_multi(const s2r_blur_job* jobs, int njobs, int H, int W, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535 && H >= 1 && W >= 1 && (long long)H * W * 3 < (1ll << 30), S2R_ERR_SHAPE,
              "gaussian_blur3_u8_multi: bad shape");
  if (njobs == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "gaussian_blur3_u8_multi: null table");
  int gr = s2r_div_up((long long)H * s2r_div_up(W, BLUR_KX), kThreads), gc = s2r_div_up((long long)s2r_div_up(H, BLUR_KY) * W * 3, kThreads);
  if (gr > 4096) gr = 4096;
  if (gc > 4096) gc = 4096;
  blur_rows_kernel<<<dim3(gr, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs, H, W);
  S2R_LAUNCH_OK();
  blur_cols_kernel<<<dim3(gc, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs, H, W);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_resize_bilinear_u8(const uint8_t* in, int N, int H, int W, int C, int axis, int out_size,
                                      const int32_t* bounds, const int32_t* kk, int ksize, int flip, uint8_t* out,
                                      s2r_stream_t stream) {
  S2R_REQUIRE(in && out && bounds && kk && N >= 1 && H >= 1 && W >= 1 && C >= 1 && out_size >= 1 && ksize >= 1,
              S2R_ERR_SHAPE, "resize_bilinear_u8: bad arguments");
  S2R_REQUIRE(axis == 0 || axis == 1, S2R_ERR_SHAPE, "resize_bilinear_u8: axis must be 0 (rows) or 1 (columns)");
  S2R_REQUIRE(axis == 1 || !flip, S2R_ERR_UNSUPPORTED, "resize_bilinear_u8: the mirror is applied by the column pass");
  if (axis == 1) {
    const long long total = (long long)N * H * out_size * C;
    resize_h_kernel<<<s2r_grid(total, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(in, N, H, W, C, out_size, bounds, kk,
                                                                                        ksize, flip, out);
  } else {
    const long long total = (long long)N * out_size * W * C;
    resize_v_kernel<<<s2r_grid(total, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(in, N, H, W * C, out_size, bounds, kk,
                                                                                        ksize, out);
  }
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_resize_nearest_u8(const uint8_t* in, int N, int H, int W, const int32_t* xtab, const int32_t* ytab,
                                     int OH, int OW, int flip, uint8_t* out, s2r_stream_t stream) {
  S2R_REQUIRE(in && out && xtab && ytab && N >= 1 && H >= 1 && W >= 1 && OH >= 1 && OW >= 1, S2R_ERR_SHAPE,
              "resize_nearest_u8: bad arguments");
  const long long total = (long long)N * OH * OW;
  resize_nearest_kernel<<<s2r_grid(total, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(in, N, H, W, xtab, ytab, OH, OW,
                                                                                            flip, out);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_input_stage_u8(const uint8_t* img, const uint8_t* label, int N, int Hs, int Ws, int flip, int x1,
                                  int y1, const double* mean, const double* std_, const uint8_t* lut, int fill_label,
                                  float* out_img, float* out_label, int H, int W, s2r_stream_t stream) {
  S2R_REQUIRE((img || label) && N >= 1 && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1 && x1 >= 0 && y1 >= 0, S2R_ERR_SHAPE,
              "input_stage_u8: bad arguments");
  S2R_REQUIRE((!img || (out_img && mean && std_)) && (!label || out_label), S2R_ERR_SHAPE, "input_stage_u8: null output");
  S2R_REQUIRE(H <= 65535 && N <= 65535, S2R_ERR_SHAPE, "input_stage_u8: grid too large");
  NormParams np;
  for (int c = 0; c < 3; ++c) {
    np.mean[c] = mean ? mean[c] : 0.0;
    np.std[c] = std_ ? std_[c] : 1.0;
  }
  dim3 grid(s2r_div_up(W, kThreads), H, N);
  input_stage_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(img, label, Hs, Ws, flip, x1, y1, np, lut, fill_label,
                                                                  out_img, out_label, H, W);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_resize_bilinear_u8_multi(const s2r_resize_job* jobs, int njobs, int64_t max_elems, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535 && max_elems >= 0, S2R_ERR_SHAPE, "resize_bilinear_u8_multi: bad job count");
  if (njobs == 0 || max_elems == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "resize_bilinear_u8_multi: null table");
  int gx = s2r_div_up(max_elems, kThreads * 4);
  if (gx > 4096) gx = 4096;
  resize_multi_kernel<<<dim3(gx, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_resize_nearest_u8_multi(const s2r_nearest_job* jobs, int njobs, int64_t max_elems, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535 && max_elems >= 0, S2R_ERR_SHAPE, "resize_nearest_u8_multi: bad job count");
  if (njobs == 0 || max_elems == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "resize_nearest_u8_multi: null table");
  int gx = s2r_div_up(max_elems, kThreads * 4);
  if (gx > 4096) gx = 4096;
  nearest_multi_kernel<<<dim3(gx, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs);
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_input_stage_u8_multi(const s2r_stage_job* jobs, int njobs, const double* mean, const double* std_,
                                        const uint8_t* lut, int fill_label, int H, int W, s2r_stream_t stream) {
  S2R_REQUIRE(njobs >= 0 && njobs <= 65535 && H >= 1 && W >= 1 && H <= 65535, S2R_ERR_SHAPE, "input_stage_u8_multi: bad shape");
  if (njobs == 0) return S2R_OK;
  S2R_REQUIRE(jobs != nullptr, S2R_ERR_SHAPE, "input_stage_u8_multi: null table");
  NormParams np;
  for (int c = 0; c < 3; ++c) {
    np.mean[c] = mean ? mean[c] : 0.0;
    np.std[c] = std_ ? std_[c] : 1.0;
  }
  input_stage_multi_kernel<<<dim3(s2r_div_up(W, kThreads), H, njobs), kThreads, 0, (cudaStream_t)stream>>>(jobs, np, lut,
                                                                                                       fill_label, H, W);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
