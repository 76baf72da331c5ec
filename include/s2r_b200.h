/* s2r_b200.h -- C ABI of the B200-native segmentation hot path.
 *
 * One shared library (libs2r_b200.so, built from
 * synthetic-to-real-semantic-segmentation_b200/csrc/ by __graft_entry__.build()) exports every
 * symbol declared here.  The reference (haofengsiji/synthetic-to-real-semantic-segmentation) is
 * pure Python on top of PyTorch: it has no FFI of its own, its "operator interface" for this
 * path is the set of torch calls listed next to each entry point below (file:line relative to
 * the reference root).  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to route those calls here.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name ends in _host;
 *   - the callee never allocates, frees or synchronises; it enqueues on `stream`;
 *   - activations are NHWC bf16: element (n,h,w,c) of a tensor with channel pitch `pitch`
 *     and channel offset `off` lives at base[((n*H + h)*W + w)*pitch + off + c];
 *   - return value 0 = success, <0 = error (S2R_ERR_*), message via s2r_last_error().
 */
#ifndef S2R_B200_H_
#define S2R_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* s2r_stream_t; /* a cudaStream_t */

#define S2R_OK 0
#define S2R_ERR_SHAPE (-1)
#define S2R_ERR_UNSUPPORTED (-2)
#define S2R_ERR_CUDA (-3)

/* activation codes */
#define S2R_ACT_NONE 0
#define S2R_ACT_RELU 1
#define S2R_ACT_RELU6 2
#define S2R_ACT_LEAKY 3

int s2r_version(void);
const char* s2r_last_error(void);
/* 1 when the current device is sm_100 (the only architecture the kernels are built for) */
int s2r_device_ok(void);

/* ------------------------------------------------------------------ dense convolutions
 * A convolution is a "tap GEMM":
 *   out[n,oh,ow,co] = epi( sum_t sum_ci src_t[n, oh+dh_t, ow+dw_t, ci] * W[slice_t][co][ci] )
 * Every tap reads a strided VIEW of an NHWC bf16 tensor (out-of-view reads are zero), which
 * expresses padding, dilation, stride-2 (one view per input parity) and transposed (data
 * gradient) convolutions with unit-step pixel boxes.  Replaces nn.Conv2d forward/backward at
 * modeling/backbone/mobilenet.py:11,44,50,58; modeling/assp.py:10,56,59; modeling/decoder.py:19,22,26,30;
 * modeling/discriminator.py:11-15; modeling/domian.py:15,19,23.
 */
#define S2R_MAX_TAPS 16

typedef struct {
  const void* base; /* bf16; channel 0 of view pixel (0,0,0) */
  int64_t sn, sh, sw; /* element strides of the view */
  int32_t H, W;       /* view extent; reads outside are zero */
  int32_t dh, dw;     /* source pixel = (oh + dh, ow + dw) */
  int32_t wslice;     /* fwd: weight slice index; wgrad: unused */
  int32_t _pad;
  int64_t wofs;       /* wgrad: element offset of this tap inside the weight-gradient tensor */
} s2r_tap;

/* A PENDING BatchNorm: the per-channel statistics a producer kernel has accumulated (sums = [2][C] fp64: sum and sum
 * of squares) together with everything needed to turn them into mean / inv-std, scale / shift and the running-statistics
 * update of modeling/sync_batchnorm/batchnorm.py:113-125 (or of the F.batch_norm fallback at :50-53).  There is no
 * finalize launch: the kernel that CONSUMES the normalised tensor (s2r_dwconv3x3_fwd_bn, s2r_bn_apply_act_bn) derives
 * scale / shift from the sums in its prologue -- the kernel boundary after the producer is all the ordering it needs --
 * and its first CTA publishes mean_invstd / scale_shift (for the backward pass) and updates the running statistics.
 * channel >= 0: synchronised BatchNorm on several ranks (batchnorm.py:55-78,90-111): the sums must first be summed over
 * the ranks on that exchange channel (see s2r_comm_*), which waits for the peers and therefore runs as one small kernel
 * of its own: s2r_bn_tail_run (exchange + finalize in one launch), after which the BatchNorm is no longer pending. */
typedef struct s2r_bn_tail {
  double count;        /* elements per channel over all ranks */
  const double* sums;  /* [2][C] */
  const float* gamma;  /* [C] or NULL */
  const float* beta;   /* [C] or NULL */
  float* running_mean; /* [C] or NULL: updated with `momentum` */
  float* running_var;
  float* mean_invstd;  /* out [2][C] */
  float* scale_shift;  /* out [2][C]: gamma*invstd, beta - mean*gamma*invstd */
  float eps, momentum;
  int32_t clamp_mode;  /* 0: (var+eps)^-1/2 (F.batch_norm); 1: max(var,eps)^-1/2 (batchnorm.py:125) */
  int32_t channel;     /* -1: no cross-rank exchange */
} s2r_bn_tail;

#define S2R_AUX_NONE 0
#define S2R_AUX_ADD 1        /* out = act(acc + bias) + aux */
#define S2R_AUX_LEAKY_MASK 2 /* out = (acc + bias) * (aux > 0 ? 1 : slope) */

typedef struct {
  uint32_t struct_size; /* sizeof(s2r_conv_args), checked */
  int32_t ntaps;
  s2r_tap taps[S2R_MAX_TAPS];
  int32_t N, OH, OW; /* output pixel grid, M = N*OH*OW */
  int32_t Cin, Cout; /* contraction width per tap (multiple of 8), output channels */
  const void* w;     /* packed bf16 [nslices][Cout_pad][Kpad] (see s2r_pack_weight) */
  int32_t Cout_pad, Kpad;
  void* out; /* bf16 view */
  int64_t on, oh, ow;
  const float* bias; /* [Cout] or NULL */
  int32_t act;
  float slope;
  int32_t aux_mode;
  int32_t _pad;
  const void* aux; /* bf16 view, same pixel grid as out */
  int64_t an, ah, aw;
  double* stats; /* [2][Cout]: += sum and sum of squares of (acc + bias), or NULL */
  const float* oscale; /* [Cout] or NULL: out = act(acc*oscale[co] + bias[co]) -- an eval-mode BatchNorm
                          (running statistics; batchnorm.py:50-53) folded into the producing convolution */
} s2r_conv_args;

int s2r_conv_fwd(const s2r_conv_args* a, s2r_stream_t stream);
/* the shape-agnostic mma.sync implementation, exported as the on-device cross-check */
int s2r_conv_fwd_mma(const s2r_conv_args* a, s2r_stream_t stream);

typedef struct {
  uint32_t struct_size;
  int32_t ntaps;
  s2r_tap taps[S2R_MAX_TAPS]; /* source views (the conv input) */
  int32_t N, OH, OW;          /* pixel grid of dy */
  int32_t Cin, Cout; /* real (unpadded) channel counts of the weight; views hold >= round8() */
  const void* dy; /* bf16 view [N,OH,OW,>=round8(Cout)] */
  int64_t dn, dh, dw;
  float* dweight;       /* fp32, += ; element (co,ci,tap) at co*s_co + ci*s_ci + taps[t].wofs */
  int64_t s_co, s_ci;
} s2r_wgrad_args;

int s2r_conv_wgrad(const s2r_wgrad_args* a, s2r_stream_t stream);

/* w: fp32 [Cout][Cin][R][S] (OIHW, nn.Conv2d.weight).  packed: bf16 [slices][A_pad][B_pad], zero padded.
 *   mode 0 (forward):       slices = R*S, (A,B) = (Cout,Cin)
 *   mode 1 (data gradient): slices = R*S, (A,B) = (Cin,Cout)
 * Row-tap forms of a 4x4 stride-2 pad-1 filter (modeling/discriminator.py:11) over a zero-padded NHWC buffer with
 * Cp = round8(Cin) channels per pixel, where the four kw taps of a filter row are 4*Cp contiguous channels:
 *   mode 2 (forward):       slices = 4 (kh), (A,B) = (Cout, 4*Cp), b = kw*Cp + c
 *   mode 3/4 (data gradient of the padded rows of parity mode-3): slices = 4 (dy offset (-ta,-tb), slice 2*ta+tb),
 *                           (A,B) = (2*Cp, Cout), a = pw*Cp + c, filter element (kh,kw) = (ph+2*ta, pw+2*tb) */
int s2r_pack_weight(const float* w, int Cout, int Cin, int R, int S, int mode, void* packed,
                    int A_pad, int B_pad, s2r_stream_t stream);
/* w.grad[co][c][kh][kw] += G[co][kh][kw*Cp + c]: a weight gradient accumulated in the mode-2 layout (fp32
 * [Cout][4][4*Cp]) added into the OIHW gradient of the 4x4 filter. */
int s2r_rowtap_wgrad_scatter(const float* G, float* dw, int Cout, int Cin, s2r_stream_t stream);
/* w.grad[co][ci][t] += G[t][co][ci] (row pitch Cp >= Cin): a multi-tap weight gradient accumulated tap-major by
 * s2r_conv_wgrad (s_ci = 1: coalesced atomics) added into the OIHW gradient. */
int s2r_wgrad_scatter_taps(const float* G, float* dw, int Cout, int Cin, int RS, int Cp, s2r_stream_t stream);
/* The same for many filters in one launch.  A job packs the elements [begin, end) of one packed filter
 * (flattened [R*S][A_pad][B_pad] index); the table lives in device memory.  transpose = mode + 16 (modes 0 / 1 only):
 * [begin, end) counts (a, b) pairs of the [A_pad][B_pad] plane instead and the job writes all R*S taps of each pair. */
typedef struct s2r_pack_job {
  const float* w;
  void* packed;
  int32_t Cout, Cin, RS, transpose /* = mode of s2r_pack_weight */, A_pad, B_pad;
  int64_t begin, end;
} s2r_pack_job;
int s2r_pack_weights_multi(const s2r_pack_job* jobs, int njobs, s2r_stream_t stream);


/* ------------------------------------------------------------------ depthwise 3x3
 * groups=C nn.Conv2d of InvertedResidual (modeling/backbone/mobilenet.py:40,54), fused with the
 * BatchNorm+ReLU6 that precedes it (mobilenet.py:51-52 / :12-13) as a load prologue and with
 * the zero padding of fixed_padding (mobilenet.py:17-23,62).
 *   in(n,h,w,c) = act(x*scale[c] + shift[c]) inside the tensor,
 *                 halo_const ? act(shift[c]) : 0 outside (|pad| pixels on every side).
 * in_scale_shift == NULL means in = x (no prologue).  stats (optional) += per-channel sum and
 * sum of squares of the raw output. */
int s2r_dwconv3x3_fwd(const void* x, const float* in_scale_shift, int in_act, int halo_const,
                      const float* w, void* y, double* stats, int N, int H, int W, int C,
                      int stride, int dil, int pad, s2r_stream_t stream);
/* The same with the input's BatchNorm still pending (in_bn: host pointer; NULL = s2r_dwconv3x3_fwd): scale / shift of
 * the prologue are derived from in_bn->sums inside the kernel, which also publishes them (see s2r_bn_tail).
 * out_scale_shift ([2][C] or NULL; inference, excludes stats): y = relu6(conv*scale + shift) -- the BatchNorm + ReLU6
 * that FOLLOWS the convolution (mobilenet.py:41-42,55-56) with its running statistics, applied before the store. */
int s2r_dwconv3x3_fwd_bn(const void* x, const s2r_bn_tail* in_bn, const float* in_scale_shift, int in_act,
                         int halo_const, const float* w, void* y, double* stats, const float* out_scale_shift,
                         int N, int H, int W, int C, int stride, int dil, int pad, s2r_stream_t stream);
/* Data gradient w.r.t. the PRE-prologue tensor's BN output, masked by act':
 *   g = dgrad(dy) * act'(x*scale + shift), written on the domain extended by `ext` pixels on
 *   every side (ext = pad when halo_const, where x counts as 0; else 0):  g[N][H+2ext][W+2ext][C].
 *   bwd_sums (optional, [2][C]) += sum g and sum g*xhat, xhat = (x - mean)*invstd. */
int s2r_dwconv3x3_dgrad(const void* dy, const float* w, const void* x, const float* in_scale_shift,
                        const float* in_mean_invstd, int in_act, int ext, void* g, double* bwd_sums,
                        int N, int H, int W, int C, int stride, int dil, int pad,
                        s2r_stream_t stream);
int s2r_dwconv3x3_wgrad(const void* x, const float* in_scale_shift, int in_act, int halo_const,
                        const void* dy, float* dw, int N, int H, int W, int C, int stride, int dil,
                        int pad, s2r_stream_t stream);
/* Both gradients in one pass over dy and x (autograd of mobilenet.py:40,54 + :51-52 + :62):
 *   g, bwd_sums as s2r_dwconv3x3_dgrad with ext = halo_const ? pad : 0;  dw (optional, fp32 [C][3][3]) +=
 *   weight gradient as s2r_dwconv3x3_wgrad.  g_interior != 0: g is [N][H][W][C] (the border positions of the
 *   extended domain enter bwd_sums but are not stored -- all the producer's BN backward needs). */
int s2r_dwconv3x3_bwd(const void* dy, const float* w, const void* x, const float* in_scale_shift,
                      const float* in_mean_invstd, int in_act, int halo_const, int g_interior, void* g,
                      double* bwd_sums, float* dw, int N, int H, int W, int C, int stride, int dil,
                      int pad, s2r_stream_t stream);

/* Patch matrix of an RxS / stride / pad convolution taken straight from an NCHW fp32 tensor:
 *   P[n][oh][ow][k] = x[n][c][oh*s+ky-pad][ow*s+kx-pad], k = (c*R+ky)*S+kx, zero for k in [C*R*S, Kp).
 * Turns the few-channel convolutions (mobilenet.py:9-14 stem, discriminator.py:11 conv1) into pointwise GEMMs
 * whose filter / filter gradient are the OIHW tensors viewed as [Cout][C*R*S]. */
int s2r_im2col_nchw_f32(const float* x, int N, int C, int H, int W, int R, int S, int stride, int pad,
                        void* P, int Kp, s2r_stream_t stream);

/* ------------------------------------------------------------------ batch norm
 * modeling/sync_batchnorm/batchnorm.py:48-78,113-125 and the F.batch_norm fallback (:50-53). */
int s2r_channel_sums_bf16(const void* x, int64_t P, int C, int pitch, int coff, double* sums,
                          s2r_stream_t stream);
/* clamp_mode 0: invstd = (var+eps)^-1/2 (F.batch_norm); 1: max(var,eps)^-1/2 (batchnorm.py:125) */
int s2r_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                    float eps, int clamp_mode, float momentum, float* running_mean,
                    float* running_var, float* mean_invstd, float* scale_shift, int C,
                    s2r_stream_t stream);
/* [exchange of tail->sums on tail->channel] + finalize as ONE launch of one CTA (what the consumers otherwise do in
 * their prologue): for synchronised BatchNorm and for consumers without a fused prologue. */
int s2r_bn_tail_run(const s2r_bn_tail* tail, int C, s2r_stream_t stream);
int s2r_bn_eval_scale_shift(const float* gamma, const float* beta, const float* running_mean,
                            const float* running_var, float eps, float* mean_invstd,
                            float* scale_shift, int C, s2r_stream_t stream);
/* The same for many BatchNorm layers in one launch (the table lives in device memory): the inference path runs this
 * once per forward pass and folds the scale / shift pairs into the producing kernels' epilogues. */
typedef struct s2r_bn_eval_job {
  const float* gamma;  /* [C] or NULL */
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* mean_invstd;  /* out [2][C] */
  float* scale_shift;  /* out [2][C] */
  int32_t C;
  float eps;
} s2r_bn_eval_job;
int s2r_bn_eval_multi(const s2r_bn_eval_job* jobs, int njobs, s2r_stream_t stream);
/* y = dropout(act(x*scale + shift)) + residual.  The dropout mask is a pure function of
 * (seed + *seed_dev, element index); seed_dev (device, may be NULL) lets a captured CUDA graph
 * draw a fresh mask on every replay. */
int s2r_bn_apply_act(const void* x, int64_t P, int C, int xpitch, int xoff,
                     const float* scale_shift, int act, const void* residual, float drop_p,
                     uint64_t seed, const uint64_t* seed_dev, void* y, int ypitch, int yoff,
                     s2r_stream_t stream);
/* The same for a pending BatchNorm (bn: host pointer; NULL = s2r_bn_apply_act with scale_shift). */
int s2r_bn_apply_act_bn(const void* x, int64_t P, int C, int xpitch, int xoff, const s2r_bn_tail* bn,
                        const float* scale_shift, int act, const void* residual, float drop_p, uint64_t seed,
                        const uint64_t* seed_dev, void* y, int ypitch, int yoff, s2r_stream_t stream);
int s2r_bn_bwd_reduce(const void* dy, int dypitch, int dyoff, const void* x, int xpitch, int xoff,
                      const float* mean_invstd, const float* scale_shift, int act, float drop_p,
                      uint64_t seed, const uint64_t* seed_dev, int64_t P, int C, double* dsums,
                      s2r_stream_t stream);
/* dx = scale*(dy' - mean(dy') - xhat*mean(dy' xhat)); count<=0: frozen statistics.
 * win_pad > 0: dy is the interior window of a [N][win_H+2pad][win_W+2pad] pixel grid. */
int s2r_bn_bwd_apply(const void* dy, int dypitch, int dyoff, const void* x, int xpitch, int xoff,
                     const float* mean_invstd, const float* scale_shift, int act, float drop_p,
                     uint64_t seed, const uint64_t* seed_dev, const double* dsums, double count, int64_t P,
                     int C, void* dx,
                     int dxpitch, int dxoff, float* dgamma /* += */, float* dbeta /* += */, int win_H, int win_W,
                     int win_pad, s2r_stream_t stream);

/* ------------------------------------------------------------------ resampling / layout
 * F.interpolate(mode='bilinear', align_corners=True) at modeling/deeplab.py:31,
 * modeling/decoder.py:39, modeling/assp.py:71; nn.AdaptiveAvgPool2d at modeling/assp.py:55. */
int s2r_upsample_bilinear_nhwc(const void* x, int N, int Hi, int Wi, int C, void* y, int Ho, int Wo,
                               int ypitch, int yoff, s2r_stream_t stream);
int s2r_upsample_bilinear_nhwc_bwd(const void* dy, int dypitch, int dyoff, int N, int Hi, int Wi,
                                   int C, int Ho, int Wo, void* dx, s2r_stream_t stream);
int s2r_upsample_bilinear_nhwc_to_nchw(const void* x, int xpitch, int N, int Hi, int Wi, int C,
                                       float* y, int Ho, int Wo, s2r_stream_t stream);
int s2r_upsample_bilinear_nchw_bwd_to_nhwc(const float* dy, int N, int C, int Ho, int Wo, void* dx,
                                           int dxpitch, int Hi, int Wi, s2r_stream_t stream);
/* The same, multiplied by the device scalar *scale (NULL = 1): a mean-reduced loss hands over its unscaled gradient
 * and the 1/sum-of-weights factor separately (utils/loss.py:21-30 followed by deeplab.py:31's backward). */
int s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled(const float* dy, int N, int C, int Ho, int Wo, void* dx, int dxpitch,
                                                  int Hi, int Wi, const float* scale, s2r_stream_t stream);
int s2r_avgpool_nhwc(const void* x, int N, int HW, int C, int pitch, int coff, float scale,
                     void* y_bf16, float* y_f32, s2r_stream_t stream);
/* y[n,p,c] (+)= v[n,c]*scale */
int s2r_broadcast_nhwc(const void* v, int N, int HW, int C, float scale, int accumulate, void* y,
                       int ypitch, int yoff, s2r_stream_t stream);
int s2r_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int64_t HW, void* y, int ypitch,
                              s2r_stream_t stream);
int s2r_nhwc_bf16_to_nchw_f32(const void* x, int xpitch, int N, int C, int64_t HW, float* y,
                              s2r_stream_t stream);
int s2r_leaky_relu_bwd_bf16(const void* dy, const void* y, void* dx, int64_t n, float slope,
                            s2r_stream_t stream);
/* a[i] += b[i] on bf16 vectors (n multiple of 8) */
int s2r_add_bf16(void* a, const void* b, int64_t n, s2r_stream_t stream);
/* out[c] += (float)sums[c] */
int s2r_add_f64_to_f32(const double* sums, float* out, int n, s2r_stream_t stream);

/* Batched forms of the three entry points above: one launch per pass for a batch whose samples have different scaled
 * sizes (every sample draws its own scale).  The job tables live in device memory; a job is one image or label map, and
 * the resampling jobs produce only the window of the scaled image that the crop keeps.
 * max_elems: the largest output element count among the jobs (sizes the grid). */
typedef struct s2r_resize_job {
  const uint8_t* in;     /* axis 1: first needed source row; axis 0: byte 0 (first window column) of source row `base` */
  uint8_t* out;          /* axis 1: [lines][on][C]; axis 0: [on][lines] */
  const int32_t* bounds; /* tables of the WHOLE axis (s2r_resize_bilinear_u8) */
  const int32_t* kk;
  int32_t W, C, ksize, flip, axis; /* W, flip: axis 1 only (source width, mirror) */
  int32_t o0, on;        /* window [o0, o0 + on) of the resampled axis */
  int32_t lines;         /* axis 1: source rows; axis 0: bytes per output row */
  int32_t in_pitch;      /* bytes between input rows */
  int32_t base;          /* axis 0: source row index `in` points at */
} s2r_resize_job;
typedef struct s2r_nearest_job {
  const uint8_t* in;
  uint8_t* out;          /* [OH][OW]: window [y0, y0 + OH) x [x0, x0 + OW) of the resized map */
  const int32_t* xtab;
  const int32_t* ytab;
  int32_t W, OH, OW, x0, y0, flip;
} s2r_nearest_job;
typedef struct s2r_stage_job {
  const uint8_t* img;   /* u8 [Hs][Ws][3] or NULL */
  const uint8_t* label; /* u8 [Hs][Ws] or NULL */
  float* out_img;       /* f32 [3][H][W] */
  float* out_label;     /* f32 [H][W] */
  int32_t Hs, Ws, flip, x1, y1, _pad;
} s2r_stage_job;
/* RandomGaussianBlur (dataloders/custom_transforms.py:92-105): PIL's ImageFilter.GaussianBlur(radius) applied to the
 * H x W crop.  A job cuts its crop from img exactly as s2r_stage_job does (mirror, window origin (x1, y1), zero padding
 * on the right / bottom) and writes the blurred crop as u8 [H][W][3] (then fed to s2r_input_stage_u8_multi as an
 * H x W image).  PIL runs three box-blur passes per axis (libImaging/BoxBlur.c); ww / fw are that file's 24-bit weights
 * of the centre pixel and of its two neighbours, computed by the caller in single precision as ImagingHorizontalBoxBlur
 * does.  Only box radii with integer part 0 (GaussianBlur radius < 1.41; the reference draws [0, 1)) are supported:
 * the caller rejects larger ones.  Two launches (row passes into tmp, column passes into out). */
typedef struct s2r_blur_job {
  const uint8_t* img;   /* u8 [Hs][Ws][3] */
  uint8_t* tmp;         /* u8 [H][W][3] scratch */
  uint8_t* out;         /* u8 [H][W][3] */
  int32_t Hs, Ws, flip, x1, y1;
  uint32_t ww, fw, _pad;
} s2r_blur_job;
int s2r_gaussian_blur3_u8_multi(const s2r_blur_job* jobs, int njobs, int H, int W, s2r_stream_t stream);
int s2r_resize_bilinear_u8_multi(const s2r_resize_job* jobs, int njobs, int64_t max_elems, s2r_stream_t stream);
int s2r_resize_nearest_u8_multi(const s2r_nearest_job* jobs, int njobs, int64_t max_elems, s2r_stream_t stream);
int s2r_input_stage_u8_multi(const s2r_stage_job* jobs, int njobs, const double* mean, const double* std_,
                             const uint8_t* lut, int fill_label, int H, int W, s2r_stream_t stream);

/* Prediction export: test_adapt.py:118-157 (imgsaver: trainId -> labelId image and palette image, NEAREST resize to the
 * output size) fused with the host argmax at test_adapt.py:170-171.  logits fp32 [N][C][H][W]; xtab/ytab int32 source
 * index per output column/row (PIL NEAREST tables, -1 = outside); id_table u8 [ntab], rgb_table u8 [ntab][3] (device);
 * ids u8 [N][OH][OW] and/or rgb u8 [N][OH][OW][3]. */
int s2r_export_prediction_nchw(const float* logits, int N, int C, int H, int W, const int32_t* xtab, const int32_t* ytab,
                               int OH, int OW, const uint8_t* id_table, const uint8_t* rgb_table, int ntab, uint8_t* ids,
                               uint8_t* rgb, s2r_stream_t stream);

/* ------------------------------------------------------------------ device input stage (uint8 -> network input)
 * The reference's per-sample CPU pipeline as byte kernels on images resident in HBM, bit-exact:
 * dataloders/custom_transforms.py:59-71 (RandomHorizontalFlip), :108-147 (RandomScaleCrop: PIL resize, pad, crop),
 * :17-56 (Normalize, ToTensor); dataloders/datasets/gtav2cityscapes.py:76-83 (encode_segmap).
 * s2r_resize_bilinear_u8: one pass of PIL's BILINEAR resize (libImaging/Resample.c) over in u8 [N][H][W][C] along
 *   axis 1 (columns -> out [N][H][out_size][C], flip != 0 mirrors the source columns first) or axis 0 (rows -> out
 *   [N][out_size][W][C]); bounds int32 [out_size][2] = (first source index, count), kk int32 [out_size][ksize] =
 *   22-bit fixed-point coefficients (device memory, computed as precompute_coeffs / normalize_coeffs_8bpc do).
 * s2r_resize_nearest_u8: PIL's NEAREST resize of u8 [N][H][W] with source index tables (-1 = outside -> 0).
 * s2r_input_stage_u8: crop window (x1, y1) of the flipped image padded on the right/bottom (image 0, label
 *   fill_label) -> out_img fp32 [N][3][H][W] = ((u8/255 - mean)/std with numpy's float32/float64 casting) and out_label
 *   fp32 [N][H][W] = lut[label] (lut u8 [256] in device memory, NULL = identity).  img or label may be NULL.
 *   mean/std are HOST pointers to 3 doubles. */
int s2r_resize_bilinear_u8(const uint8_t* in, int N, int H, int W, int C, int axis, int out_size, const int32_t* bounds,
                           const int32_t* kk, int ksize, int flip, uint8_t* out, s2r_stream_t stream);
int s2r_resize_nearest_u8(const uint8_t* in, int N, int H, int W, const int32_t* xtab, const int32_t* ytab, int OH,
                          int OW, int flip, uint8_t* out, s2r_stream_t stream);
int s2r_input_stage_u8(const uint8_t* img, const uint8_t* label, int N, int Hs, int Ws, int flip, int x1, int y1,
                       const double* mean, const double* std_, const uint8_t* lut, int fill_label, float* out_img,
                       float* out_label, int H, int W, s2r_stream_t stream);

/* ------------------------------------------------------------------ losses
 * F.softmax(x, dim=0) at train_adapt.py:151,166,174; nn.CrossEntropyLoss at utils/loss.py:21-30,
 * 57-69; torch.nn.BCEWithLogitsLoss at train_adapt.py:75,153,168,176. */
int s2r_softmax_dim0_fwd(const float* x, float* y, int B, int64_t M, s2r_stream_t stream);
int s2r_softmax_dim0_bwd(const float* y, const float* dy, float* dx, int B, int64_t M,
                         s2r_stream_t stream);
/* The composite `model_D(F.softmax(x, dim=0))` input stage (train_adapt.py:151,166,174 feeding
 * modeling/discriminator.py:23): x fp32 [B][C][H][W] -> yp bf16 [B][H+2][W+2][Cp], the (optional) batch-axis softmax
 * written into the interior of a zero-padded NHWC buffer (border and pad channels are written as zeros).  softmax != 0
 * needs B <= 8 (S2R_ERR_UNSUPPORTED otherwise).  _bwd: dx fp32 [B][C][H][W] from the gradient gp w.r.t. yp (same padded
 * layout, interior read) with the softmax recomputed from x; softmax == 0: layout conversion only (x may be NULL). */
int s2r_softmax0_nchw_to_nhwc_pad(const float* x, int B, int C, int H, int W, int softmax, void* yp, int Cp,
                                  s2r_stream_t stream);
int s2r_softmax0_nhwc_pad_bwd(const float* x, const void* gp, int B, int C, int H, int W, int Cp, int softmax,
                              float* dx, s2r_stream_t stream);
/* The same over the GLOBAL batch of a data-parallel run -- the reference's nn.DataParallel gathers the logits of all
 * replicas on one device before F.softmax(x, dim=0) (train_adapt.py:87-88,151,166,174), so the softmax runs over the
 * images of ALL ranks.  s2r_softmax0_batch_stats: out[i] = max_b x[b][i] (gmax NULL) or sum_b exp(x[b][i] - gmax[i]) over
 * this rank's B images, i < M = C*H*W; the caller all-reduces the first with MAX, the second with SUM.  _global forward:
 * y_b = exp(x_b - gmax) / gsum into the padded NHWC buffer.  _global backward, two launches around one all-reduce(SUM):
 * tpart != NULL: tpart[i] = this rank's sum_b g_b y_b (dx untouched); tpart == NULL: dx_b = y_b (g_b - tglob). */
int s2r_softmax0_batch_stats(const float* x, int B, int64_t M, const float* gmax, float* out, s2r_stream_t stream);
int s2r_softmax0_nchw_to_nhwc_pad_global(const float* x, int B, int C, int H, int W, const float* gmax,
                                         const float* gsum, void* yp, int Cp, s2r_stream_t stream);
int s2r_softmax0_nhwc_pad_bwd_global(const float* x, const void* gp, int B, int C, int H, int W, int Cp,
                                     const float* gmax, const float* gsum, float* tpart, const float* tglob,
                                     float* dx, s2r_stream_t stream);
/* sums (fp64, zeroed by the caller) += {sum_valid w_t (lse - x_t), sum_valid w_t, #(argmax == t)}; grad_unscaled
 * (optional) = w_t (softmax - onehot).  A target that is neither ignore_index nor in [0, C) makes sums[0] NaN (the
 * reference's nn.CrossEntropyLoss stops with a device assert there, utils/loss.py:27-28). */
int s2r_cross_entropy_nchw(const float* logits, const float* target, int const_target,
                           const float* weight, int N, int C, int64_t HW, int ignore_index,
                           double* sums, float* grad_unscaled, s2r_stream_t stream);
int s2r_ratio(const double* sums, double denom_override, float* out, s2r_stream_t stream);
int s2r_scale_by_ratio(float* g, int64_t n, const float* gout, const double* sums,
                       double denom_override, s2r_stream_t stream);
int s2r_bce_logits_fwd(const float* x, const float* target, float const_target, int64_t n,
                       double* sums, s2r_stream_t stream);
int s2r_bce_logits_bwd(const float* x, const float* target, float const_target, int64_t n,
                       const float* gout, float* dx, s2r_stream_t stream);

/* ------------------------------------------------------------------ Evaluator
 * utils/metrics.py:34-43 (_generate_matrix/add_batch) and the host argmax at val_adapt.py:131-135. */
int s2r_confusion_matrix(const void* gt, int gt_is_i64, const int64_t* pred, int64_t n,
                         int num_class, int64_t* counts, int64_t* bad_pred, s2r_stream_t stream);
int s2r_argmax_confusion_nchw(const float* logits, const float* gt, int N, int C, int64_t HW,
                              int num_class, int64_t* counts, int64_t* pred_out,
                              s2r_stream_t stream);
/* The same from the decoder's LOW-RESOLUTION logits x (NHWC bf16 [N][Hi][Wi][pitch], C <= 32 classes): the final
 * F.interpolate(x, size=(Ho,Wo), mode='bilinear', align_corners=True) of modeling/deeplab.py:31, the argmax of
 * val_adapt.py:133 and the histogram of utils/metrics.py:34-43 in ONE pass; the fp32 [N,C,Ho,Wo] logits are never
 * materialised.  Interpolated values are those of s2r_upsample_bilinear_nhwc_to_nchw bit for bit, so the counts equal
 * s2r_upsample_bilinear_nhwc_to_nchw + s2r_argmax_confusion_nchw exactly.  gt: fp32 [N][Ho][Wo]. */
int s2r_upsample_argmax_confusion_nhwc(const void* x, int xpitch, int N, int Hi, int Wi, int C, const float* gt,
                                       int Ho, int Wo, int num_class, int64_t* counts, s2r_stream_t stream);

/* ------------------------------------------------------------------ optimizers
 * torch.optim.SGD / Adam steps at train_adapt.py:58-60,180-181 and train.py:63-82,202-204, as
 * multi-tensor kernels over a device table of (param, grad, state) pointers. */
typedef struct {
  float* p;
  const float* g;
  float* s0; /* SGD: momentum buffer; Adam: exp_avg */
  float* s1; /* Adam: exp_avg_sq */
  int64_t n;
  float lr_mult; /* per-tensor lr multiplier (1x / 10x groups, deeplab.py:42-72) */
  float _pad;
} s2r_param_slot;

/* hyper (device): [0] = lr, [1] = 1 - beta1^t, [2] = 1 - beta2^t (Adam only).
 * SGD: g' = g*gscale + wd*p; buf = momentum*buf + (1-dampening)*g'; p -= lr*lr_mult*(nesterov ? g' + momentum*buf : buf)
 * (momentum buffers start at zero, which equals torch's first-step rule for dampening == 0). */
int s2r_sgd_step(const s2r_param_slot* slots, int nslots, const float* hyper, float momentum,
                 float dampening, float weight_decay, int nesterov, float gscale,
                 s2r_stream_t stream);
int s2r_adam_step(const s2r_param_slot* slots, int nslots, const float* hyper, float beta1,
                  float beta2, float eps, float weight_decay, float gscale, s2r_stream_t stream);
/* dst[0..n) (device) = host_vals[0..n), n <= 8, stream-ordered.  The values travel as kernel arguments, i.e. they are
 * read from host memory before the call returns: this is how the learning rate (utils/lr_scheduler.py:63-70 writes
 * param_groups[i]['lr'] before every step, train_adapt.py:131-134) and Adam's bias corrections reach `hyper` ahead of
 * a CUDA-graph replay without a host buffer the GPU would read later. */
int s2r_store_f32(float* dst, int n, const float* host_vals, s2r_stream_t stream);

/* ------------------------------------------------------------------ NVLink peer-memory exchange
 * Replaces the master/slave reduce + broadcast of modeling/sync_batchnorm/comm.py:18-129 and
 * batchnorm.py:90-111: a one-kernel, one-shot all-reduce of the small fp64 statistics vectors over CUDA IPC
 * mapped peer memory.  s2r_comm_create writes this rank's 64-byte CUDA IPC handle; the host exchanges the
 * handles (any transport) and passes all of them, indexed by rank, to s2r_comm_open.  Every rank must then issue
 * the same sequence of s2r_allreduce_small_f64 calls.  The sequence counter lives on the device, so the calls may
 * be captured in CUDA graphs. */
int s2r_comm_create(int rank, int world, int slot_doubles, void* handle_out);
int s2r_comm_open(const void* handles);
int s2r_comm_ready(void);   /* world size once opened, else 0 */
int s2r_allreduce_small_f64(double* buf, int n, s2r_stream_t stream);
/* The same on exchange channel 0 or 1: every rank issues the same sequence of calls PER CHANNEL, and calls on
 * different channels may be in flight together (the two streams of the training step). */
int s2r_allreduce_small_f64_ch(double* buf, int n, int channel, s2r_stream_t stream);
int s2r_comm_error(void);   /* non-zero: a bounded wait expired; a host read of a mapped flag (no synchronisation) */
int s2r_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* S2R_B200_H_ */
