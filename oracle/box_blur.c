/* CPU oracle (test infrastructure only), integer path in C: Pillow's box blur as used by ImageFilter.GaussianBlur
 * (third party, not vendored by the reference; restated from libImaging/BoxBlur.c of Pillow 12.2.0: ImagingLineBoxBlur32
 * / ImagingHorizontalBoxBlur / ImagingBoxBlur) -- what dataloders/custom_transforms.py:92-105 (RandomGaussianBlur)
 * calls on the crop.  Second, independent restatement next to oracle/input_stage.py (numpy); both are pinned against
 * Pillow itself by tests/test_oracle.py.
 *
 *   box_blur_lines: one horizontal pass over `lines` lines of `w` pixels with `ch` interleaved uint8 channels.
 *     radius / ww / fw: integer box radius and the 24-bit weights of the inner pixels and of the two fractional outer
 *     pixels (oracle/input_stage.py box_blur_weights).  A running sum over the inner window is kept in uint32;
 *     out = (acc * ww + (left + right) * fw + 2^23) >> 24; positions beyond the line read its first / last pixel.
 *   gaussian_blur_u8: `passes` passes along x, transpose, `passes` passes along y, transpose back. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void blur_line(const uint8_t* in, uint8_t* out, int w, int ch, int c, int radius, uint32_t ww, uint32_t fw) {
  const int last = w - 1;
  const int edge_a = radius + 1 < w ? radius + 1 : w;
  const int edge_b = w - radius - 1 > 0 ? w - radius - 1 : 0;
#define PX(i) ((uint32_t)in[(size_t)(i) * ch + c])
#define EMIT(x, sub, add, left, right)                                   \
  do {                                                                   \
    acc += PX(add) - PX(sub);                                            \
    uint32_t bulk = acc * ww + (PX(left) + PX(right)) * fw;              \
    out[(size_t)(x) * ch + c] = (uint8_t)((bulk + (1u << 23)) >> 24);    \
  } while (0)
  uint32_t acc = PX(0) * (uint32_t)(radius + 1);
  for (int x = 0; x < edge_a - 1; ++x) acc += PX(x);
  acc += PX(last) * (uint32_t)(radius - edge_a + 1);
  if (edge_a <= edge_b) {
    for (int x = 0; x < edge_a; ++x) EMIT(x, 0, x + radius, 0, x + radius + 1);
    for (int x = edge_a; x < edge_b; ++x) EMIT(x, x - radius - 1, x + radius, x - radius - 1, x + radius + 1);
    for (int x = edge_b; x <= last; ++x) EMIT(x, x - radius - 1, last, x - radius - 1, last);
  } else {
    for (int x = 0; x < edge_b; ++x) EMIT(x, 0, x + radius, 0, x + radius + 1);
    for (int x = edge_b; x < edge_a; ++x) EMIT(x, 0, last, 0, last);
    for (int x = edge_a; x <= last; ++x) EMIT(x, x - radius - 1, last, x - radius - 1, last);
  }
#undef EMIT
#undef PX
}

void box_blur_lines(const uint8_t* in, uint8_t* out, int lines, int w, int ch, int radius, uint32_t ww, uint32_t fw) {
  for (int y = 0; y < lines; ++y)
    for (int c = 0; c < ch; ++c) blur_line(in + (size_t)y * w * ch, out + (size_t)y * w * ch, w, ch, c, radius, ww, fw);
}

static void transpose(const uint8_t* in, uint8_t* out, int h, int w, int ch) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) memcpy(out + ((size_t)x * h + y) * ch, in + ((size_t)y * w + x) * ch, (size_t)ch);
}

/* img: [h][w][ch] uint8, blurred in place.  Returns 0, or -1 when out of memory. */
int gaussian_blur_u8(uint8_t* img, int h, int w, int ch, int radius, uint32_t ww, uint32_t fw, int passes) {
  const size_t n = (size_t)h * w * ch;
  uint8_t* a = (uint8_t*)malloc(n);
  uint8_t* b = (uint8_t*)malloc(n);
  if (!a || !b) {
    free(a);
    free(b);
    return -1;
  }
  for (int p = 0; p < passes; ++p) {
    box_blur_lines(img, a, h, w, ch, radius, ww, fw);
    memcpy(img, a, n);
  }
  transpose(img, a, h, w, ch);
  for (int p = 0; p < passes; ++p) {
    box_blur_lines(a, b, w, h, ch, radius, ww, fw);
    memcpy(a, b, n);
  }
  transpose(a, img, w, h, ch);
  free(a);
  free(b);
  return 0;
}
