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

// VEC (row pitch a multiple of 16 bytes): the window is fetched with 16-byte cp.async copies from the enclosing
// 4-float-aligned column range (zero-filled outside the image), ~10 copies in flight per thread; otherwise it is
// staged with plain scalar loads.
template <bool VEC>
__global__ void __launch_bounds__(IC_THREADS)
im2col_nchw_kernel(const float* __restrict__ x, int C, int H, int W, int R, int S, int stride, int pad,
                   __nv_bfloat16* __restrict__ P, int OH, int OW, int Kp, int Wt) {
  pdl_wait();      // programmatic dependent launch: see common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) float win[];   // [C*R][Wt]
  const int ow0 = blockIdx.x * IC_TP, oh = blockIdx.y, n = blockIdx.z;
  int ix0 = ow0 * stride - pad;
  const int rows = C * R;
  int shift = 0;   // window column of input column ix0
  if (VEC) {
    const int xs = (ix0 >= 0 ? ix0 : ix0 - 3) / 4 * 4;   // floor to a multiple of 4
    shift = ix0 - xs;
    ix0 = xs;
    const int chunks = Wt / 4;
    for (int t = threadIdx.x; t < rows * chunks; t += IC_THREADS) {
      const int rr = t / chunks, ch = t - rr * chunks;
      const int c = rr / R, ky = rr - c * R;
      const int iy = oh * stride + ky - pad;
      const int ix = ix0 + ch * 4;
      const bool ok = (unsigned)iy < (unsigned)H && ix >= 0 && ix < W;   // W % 4 == 0: a chunk is all in or all out
      const float* src = x + (((long long)n * C + c) * H + (ok ? iy : 0)) * W + (ok ? ix : 0);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(win + rr * Wt + ch * 4);
      const int nbytes = ok ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  // stage the input window (zero outside the image)
  for (int rr = threadIdx.x / 32; !VEC && rr < rows; rr += IC_THREADS / 32) {
    const int c = rr / R, ky = rr - c * R;
    const int iy = oh * stride + ky - pad;
    const bool row_ok = (unsigned)iy < (unsigned)H;
    const float* src = x + (((long long)n * C + c) * H + (row_ok ? iy : 0)) * W;
    for (int j = threadIdx.x % 32; j < Wt; j += 32) {
      const int ix = ix0 + j;
      win[rr * Wt + j] = (row_ok && (unsigned)ix < (unsigned)W) ? __ldg(src + ix) : 0.f;
    }
  }
  // window offset of every patch entry k (built once per CTA; -1: padding entry)
  int* lut = reinterpret_cast<int*>(win + rows * Wt);   // [Kp]
  const int CRS = C * R * S;
  for (int k = threadIdx.x; k < Kp; k += IC_THREADS) {
    int o = -1;
    if (k < CRS) {
      const int c = k / (R * S), t = k - c * R * S, ky = t / S, kx = t - ky * S;
      o = (c * R + ky) * Wt + kx + shift;
    }
    lut[k] = o;
  }
  __syncthreads();
  // warp item = 4 consecutive 16-byte vectors x 8 consecutive pixels: the window reads of a warp spread over
  // 16 banks (a vector-major mapping puts every other lane on the same bank: the rows of consecutive vectors are
  // 2*Wt floats apart) and each pixel still receives a full 64-byte segment per store instruction
  const int nv = Kp / 8, vgroups = (nv + 3) / 4;
  const int npix = min(IC_TP, OW - ow0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vl = lane & 3, pl = lane >> 2;
  __nv_bfloat16* dst0 = P + (((long long)n * OH + oh) * OW + ow0) * Kp;
  for (int item = warp; item < vgroups * (IC_TP / 8); item += IC_THREADS / 32) {
    const int vg = item % vgroups, pg = item / vgroups;
    const int v = vg * 4 + vl, p = pg * 8 + pl;
    if (v >= nv || p >= npix) continue;
    const int4 o0 = *reinterpret_cast<const int4*>(lut + v * 8), o1 = *reinterpret_cast<const int4*>(lut + v * 8 + 4);
    const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = off[i] >= 0 ? win[off[i] + p * stride] : 0.f;
    *reinterpret_cast<uint4*>(dst0 + (long long)p * Kp + v * 8) = float_to_bf16x8(f);
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
  const bool vec = W % 4 == 0 && (uintptr_t)x % 16 == 0;
  // window width: the patch columns, plus up to 3 columns of alignment slack on the vector path, as a multiple of 4
  const int Wt = ((IC_TP - 1) * stride + S + (vec ? 3 : 0) + 3) & ~3;
  const size_t smem = (size_t)C * R * Wt * sizeof(float) + (size_t)Kp * sizeof(int);
  S2R_REQUIRE(smem <= 96 * 1024, S2R_ERR_UNSUPPORTED, "im2col: window of %zu bytes too large", smem);
  static bool attr = false;
  if (!attr) {
    S2R_CUDA_OK(cudaFuncSetAttribute(im2col_nchw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    S2R_CUDA_OK(cudaFuncSetAttribute(im2col_nchw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  dim3 grid(s2r_div_up(OW, IC_TP), OH, N);
  if (vec)
    S2R_CUDA_OK(s2r_launch(im2col_nchw_kernel<true>, dim3(grid), dim3(IC_THREADS), (size_t)(smem), (cudaStream_t)stream, x, C, H, W, R, S, stride, pad,
                                                                            (__nv_bfloat16*)P, OH, OW, Kp, Wt));
  else
    S2R_CUDA_OK(s2r_launch(im2col_nchw_kernel<false>, dim3(grid), dim3(IC_THREADS), (size_t)(smem), (cudaStream_t)stream, x, C, H, W, R, S, stride, pad,
                                                                             (__nv_bfloat16*)P, OH, OW, Kp, Wt));
  S2R_LAUNCH_OK();
  return S2R_OK;
}
