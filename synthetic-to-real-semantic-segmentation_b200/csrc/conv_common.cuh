// Geometry shared by the dense convolution kernels (direct CUDA-core path and the
// tcgen05 implicit-GEMM path).  A convolution is described as a list of taps: for an
// output pixel (oh, ow), tap t reads source pixel (src(oh, dh_t), src(ow, dw_t)) where
//   forward / stride-s:      src(o, d) = o * s + d            (d = k*dil - pad)
//   transposed (data grad):  src(o, d) = (o + d) / s  if divisible, else no contribution
//                                                              (d = pad - k*dil)
// and contracts its channel vector with weight slice w[t] ([M_pad][K_pad], K contiguous).
#pragma once
#include "common.cuh"

constexpr int S2R_MAX_TAPS = 16;

struct ConvGeom {
  int N, H, W, Cin;   // input tensor (NHWC)
  int Cout, R, S;     // filter
  int stride, pad, dil;
  int Ho, Wo;         // derived
};

static inline int conv_geom_init(ConvGeom* g, int N, int H, int W, int Cin, int Cout, int R, int S,
                                 int stride, int pad, int dil) {
  g->N = N; g->H = H; g->W = W; g->Cin = Cin; g->Cout = Cout; g->R = R; g->S = S;
  g->stride = stride; g->pad = pad; g->dil = dil;
  S2R_REQUIRE(N >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cout >= 1, S2R_ERR_SHAPE, "conv2d: bad tensor shape");
  S2R_REQUIRE(R >= 1 && S >= 1 && R * S <= S2R_MAX_TAPS, S2R_ERR_UNSUPPORTED, "conv2d: filter %dx%d has more than %d taps", R, S, S2R_MAX_TAPS);
  S2R_REQUIRE(stride >= 1 && dil >= 1, S2R_ERR_SHAPE, "conv2d: bad stride/dilation");
  const int he = H + 2 * pad - dil * (R - 1) - 1, we = W + 2 * pad - dil * (S - 1) - 1;
  S2R_REQUIRE(he >= 0 && we >= 0, S2R_ERR_SHAPE, "conv2d: input smaller than the (dilated) filter");
  g->Ho = he / stride + 1;
  g->Wo = we / stride + 1;
  return S2R_OK;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
// packed-weight paddings: contraction dim to 64 (one 128-byte swizzle row), output dim to 16 (UMMA N granule)
static inline int kpad_of(int k) { return round_up(k, 64); }
static inline int mpad_of(int m) { return round_up(m, 16); }
