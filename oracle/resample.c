/* CPU oracle (test infrastructure only), byte path in C: Pillow's resize as the reference's transforms call it
 * (dataloders/custom_transforms.py:124-125, custom_transforms_eval.py:139-140,163-164: Image.resize(BILINEAR) for the
 * image, Image.resize(NEAREST) for the label map).  Third party, not vendored by the reference; restated from
 * libImaging/Resample.c (precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
 * ImagingResampleVertical_8bpc: 22-bit fixed-point coefficients, int32 accumulation from 2^21, uint8 intermediate
 * between the passes) and libImaging/Geometry.c (ImagingScaleAffine: source coordinate advanced by repeated double
 * additions, truncated).  Second, independent restatement next to oracle/input_stage.py; both are pinned against
 * Pillow itself by tests/test_oracle.py. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

/* bounds[out][2] = (first source index, count); kk[out][ksize] integer coefficients.  Returns ksize, 0 on failure. */
static int coefficients(int in_size, int out_size, int** bounds_p, int** kk_p) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  int* bounds = (int*)malloc(sizeof(int) * 2 * (size_t)out_size);
  int* kk = (int*)calloc((size_t)out_size * ksize, sizeof(int));
  double* k = (double*)malloc(sizeof(double) * (size_t)ksize);
  if (!bounds || !kk || !k) {
    free(bounds);
    free(kk);
    free(k);
    return 0;
  }
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < ksize; ++x) k[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      k[x] = bilinear_filter((x + xmin - center + 0.5) * ss);
      ww += k[x];
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x)
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  free(k);
  *bounds_p = bounds;
  *kk_p = kk;
  return ksize;
}

static uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* in [h][w][ch] -> out [oh][ow][ch]; horizontal pass first, each pass skipped when its size does not change. */
int resize_bilinear_u8(const uint8_t* in, int h, int w, int ch, int oh, int ow, uint8_t* out) {
  const uint8_t* cur = in;
  uint8_t* tmp = NULL;
  if (w != ow) {
    int *bounds, *kk;
    const int ksize = coefficients(w, ow, &bounds, &kk);
    if (!ksize) return -1;
    tmp = (uint8_t*)malloc((size_t)h * ow * ch);
    if (!tmp) return -1;
    for (int y = 0; y < h; ++y)
      for (int xx = 0; xx < ow; ++xx)
        for (int c = 0; c < ch; ++c) {
          int acc = 1 << (PRECISION_BITS - 1);
          for (int x = 0; x < bounds[2 * xx + 1]; ++x)
            acc += (int)cur[((size_t)y * w + bounds[2 * xx] + x) * ch + c] * kk[(size_t)xx * ksize + x];
          tmp[((size_t)y * ow + xx) * ch + c] = clip8(acc);
        }
    free(bounds);
    free(kk);
    cur = tmp;
  }
  if (h != oh) {
    int *bounds, *kk;
    const int ksize = coefficients(h, oh, &bounds, &kk);
    if (!ksize) return -1;
    for (int yy = 0; yy < oh; ++yy)
      for (int i = 0; i < ow * ch; ++i) {
        int acc = 1 << (PRECISION_BITS - 1);
        for (int y = 0; y < bounds[2 * yy + 1]; ++y) acc += (int)cur[((size_t)(bounds[2 * yy] + y) * ow) * ch + i] * kk[(size_t)yy * ksize + y];
        out[(size_t)yy * ow * ch + i] = clip8(acc);
      }
    free(bounds);
    free(kk);
  } else {
    memcpy(out, cur, (size_t)oh * ow * ch);
  }
  free(tmp);
  return 0;
}

/* in [h][w] -> out [oh][ow]; positions whose source coordinate falls outside stay 0. */
void resize_nearest_u8(const uint8_t* in, int h, int w, int oh, int ow, uint8_t* out) {
  const double ax = (double)w / ow, ay = (double)h / oh;
  double yo = 0.0 + ay * 0.5;
  memset(out, 0, (size_t)oh * ow);
  for (int y = 0; y < oh; ++y, yo += ay) {
    const int yin = yo < 0.0 ? -1 : (int)yo;
    if (yin < 0 || yin >= h) continue;
    double xo = 0.0 + ax * 0.5;
    for (int x = 0; x < ow; ++x, xo += ax) {
      const int xin = xo < 0.0 ? -1 : (int)xo;
      if (xin >= 0 && xin < w) out[(size_t)y * ow + x] = in[(size_t)yin * w + xin];
    }
  }
}
