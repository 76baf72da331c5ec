// Patch extraction (im2col) straight from an NCHW fp32 tensor into an NHWC bf16 patch matrix, for the two
// convolutions whose input has very few channels and very many pixels:
//   * the MobileNetV2 stem, 3x3 stride 2 on the 3-channel image       (modeling/backbone/mobilenet.py:9-14,91)
//   * FCDiscriminator.conv1, 4x4 stride 2 on the 19-channel softmax   (modeling/discriminator.py:11,23)
// As tap-GEMMs these read a 64-channel TMA box per tap of which 3 (19) channels are real; as ONE pointwise
// GEMM over patches the contraction is dense (27 -> 32, 304 -> 304 of 320) and the weight gradient becomes a
// 1x1 weight gradient with coalesced atomics.  The patch matrix also replaces the NCHW->NHWC conversion pass.
//   P[n][oh][ow][k] = x[n][c][oh*s + ky - pad][ow*s + kx - pad],  k = (c*R + ky)*S + kx  (the OIHW order of the
//   filter, so the filter and its gradient are used in place as [Cout][C*R*S]); k in [C*R*S, Kp) is zero.
// HBM-bound: (4*C*H*W + 2*OH*OW*Kp) bytes per image.
#include "common.cuh"

namespace {

constexpr int IC_TP = 64;        // output pixels (one row segment) per CTA
constexpr int IC_THREADS = 256;

__global__ void __launch_bounds__(IC_THREADS)
im2col_nchw_kernel(const float* __restrict__ x, int C, int H, int W, int R, int S, int stride, int pad,
                   __nv_bfloat16* __restrict__ P, int OH, int OW, int Kp, int Wt) {
  extern __shared__ float win[];   // [C*R][Wt]
  const int ow0 = blockIdx.x * IC_TP, oh = blockIdx.y, n = blockIdx.z;
  const int ix0 = ow0 * stride - pad;
  const int rows = C * R;
  // stage the input window (zero outside the image)
  for (int rr = threadIdx.x / 32; rr < rows; rr += IC_THREADS / 32) {
    const int c = rr / R, ky = rr - c * R;
    const int iy = oh * stride + ky - pad;
    const bool row_ok = (unsigned)iy < (unsigned)H;
    const float* src = x + (((long long)n * C + c) * H + (row_ok ? iy : 0)) * W;
    for (int j = threadIdx.x % 32; j < Wt; j += 32) {
      const int ix = ix0 + j;
      win[rr * Wt + j] = (row_ok && (unsigned)ix < (unsigned)W) ? __ldg(src + ix) : 0.f;
    }
  }
  __syncthreads();
  const int nv = Kp / 8;
  const int v = threadIdx.x % nv, p0 = threadIdx.x / nv, pstep = IC_THREADS / nv;
  if (p0 >= pstep) return;   // threads beyond the last full pixel group
  // window offsets of this thread's 8 patch entries
  int off[8];
  const int CRS = C * R * S;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = v * 8 + i;
    if (k < CRS) {
      const int c = k / (R * S), t = k - c * R * S, ky = t / S, kx = t - ky * S;
      off[i] = (c * R + ky) * Wt + kx;
    } else {
      off[i] = -1;
    }
  }
  const int npix = min(IC_TP, OW - ow0);
  __nv_bfloat16* dst = P + (((long long)n * OH + oh) * OW + ow0) * Kp + v * 8;
  for (int p = p0; p < npix; p += pstep) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = off[i] >= 0 ? win[off[i] + p * stride] : 0.f;
    *reinterpret_cast<uint4*>(dst + (long long)p * Kp) = float_to_bf16x8(f);
  }
}

}  // namespace

extern "C" int s2r_im2col_nchw_f32(const float* x, int N, int C, int H, int W, int R, int S, int stride, int pad,
                                   void* P, int Kp, s2r_stream_t stream) {
  S2R_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1 && R >= 1 && S >= 1 && stride >= 1 && pad >= 0, S2R_ERR_SHAPE,
              "im2col: bad shape");
  S2R_REQUIRE(Kp % 8 == 0 && Kp >= C * R * S && Kp <= 8 * IC_THREADS && (uintptr_t)P % 16 == 0, S2R_ERR_SHAPE,
              "im2col: patch pitch %d must be a multiple of 8 in [C*R*S, %d]", Kp, 8 * IC_THREADS);
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  S2R_REQUIRE(OH >= 1 && OW >= 1 && OH <= 65535 && N <= 65535, S2R_ERR_SHAPE, "im2col: bad output shape");
  const int Wt = (IC_TP - 1) * stride + S;
  const size_t smem = (size_t)C * R * Wt * sizeof(float);
  S2R_REQUIRE(smem <= 96 * 1024, S2R_ERR_UNSUPPORTED, "im2col: window of %zu bytes too large", smem);
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(im2col_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  dim3 grid(s2r_div_up(OW, IC_TP), OH, N);
  im2col_nchw_kernel<<<grid, IC_THREADS, smem, (cudaStream_t)stream>>>(x, C, H, W, R, S, stride, pad,
                                                                      (__nv_bfloat16*)P, OH, OW, Kp, Wt);
  S2R_LAUNCH_OK();
  return S2R_OK;
}
