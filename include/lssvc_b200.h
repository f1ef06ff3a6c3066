/*
 * lssvc_b200 — C-ABI of the B200-native LSSVC two-layer coding forward pass.
 *
 * Every entry point takes plain pointers and sizes (device pointers unless a
 * parameter is marked "host"), a cudaStream_t passed as void*, and returns 0 on
 * success or a negative lssvc_status.  No torch types cross this boundary.
 *
 * Activations are NHWC fp32 ("pixel-major": all channels of one pixel are
 * contiguous).  A view is a window of `C` channels inside a buffer whose
 * pixels are `pitch` floats apart, so channel-concatenation of the reference
 * (torch.cat(dim=1)) is a matter of pointing producers at slices of one buffer.
 *
 * The reference is a PyTorch code drop with no FFI of its own on this path
 * except the pybind11 entropy coder; each entry point below cites the reference
 * operator (file:line under /root/reference) it replaces.
 */
#ifndef LSSVC_B200_H
#define LSSVC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  LSSVC_OK = 0,
  LSSVC_ERR_ARG = -1,        /* bad shape / alignment / unsupported configuration */
  LSSVC_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed               */
  LSSVC_ERR_NO_DEVICE = -3,  /* no sm_100 device                                   */
  LSSVC_ERR_STREAM = -4      /* rANS stream exhausted / malformed                  */
} lssvc_status;

/* NHWC fp32 window: element (y, x, c) lives at ptr[(y * W + x) * pitch + c]. */
typedef struct {
  float *ptr;
  int32_t H, W, C;
  int32_t pitch;
} lssvc_view;

#define LSSVC_MAX_SRC 3

#define LSSVC_ACT_NONE 0
#define LSSVC_ACT_LRELU 1 /* x > 0 ? x : slope * x (ReLU is slope 0) */

#define LSSVC_IN_NONE 0
#define LSSVC_IN_SQUARE 1 /* x*x, used by GDN's norm pool */
#define LSSVC_IN_LRELU 2

#define LSSVC_PREC_TF32 0   /* reserved (study engines under tools/engines/, not in this library)          */
#define LSSVC_PREC_3XTF32 1 /* reserved                                                                    */
#define LSSVC_PREC_H2 2     /* split-fp16: a = a_hi + a_lo, w = w_hi + w_lo in fp16 (22 significant bits each),
                               3 kind::f16 MMAs per product, hi*hi and the cross terms in separate fp32 accumulators */

#define LSSVC_EPI_PLAIN 0
#define LSSVC_EPI_GDN 1  /* out = gdn_x * rsqrt(acc + bias)   (gdn.py:29-44, video_net_component.py:83-105) */
#define LSSVC_EPI_IGDN 2 /* out = gdn_x * sqrt(acc + bias)                                               */
/* Entropy epilogues (lssvc_conv_hs; see lssvc_conv::ent_*): the convolution that PRODUCES the entropy parameters codes the
 * latent in its own epilogue — no separate pass over the parameters, no extra launch. */
#define LSSVC_EPI_LAPLACE 3 /* conv output = (scale | mean), 2C channels, written to `out` as usual; then
                               q = rint(ent_y - mean), ent_y_hat = q + mean, bits += clamp(-log2(Laplace(0, scale) mass of
                               [q - .5, q + .5] + 1e-5), 0, 50), symbol and CDF-row dumps: exactly lssvc_laplace_quant
                               (LSSVC_net.py:154-167, 288-296; dmc_net.py:421-488) */
#define LSSVC_EPI_FOURPART 5 /* LAPLACE restricted to step ent_step of the four-part spatial prior (LSSVC_net.py:338-443): only the
                                channel quarter / pixel parity pairs of that step are coded (step 0 zeroes the rest of ent_y_hat):
                                exactly lssvc_four_part_step.  The producing convolution is the last 1x1 of a ConvFFN, so this
                                mode (like LAPLACE / BITPARM) takes a LeakyReLU and res1 in front of the entropy arithmetic */
#define LSSVC_EPI_BITPARM 4 /* conv output = z; out = rint(z), bits from the factorised BitEstimator prior (ent_coef),
                               symbol dump: exactly lssvc_bitparm_quant (video_entropy_models.py Bitparm / BitEstimator) */

/*
 * One convolution with its fused epilogue.  Replaces nn.Conv2d / nn.ConvTranspose2d(stride 1)
 * + nn.LeakyReLU / nn.ReLU + residual adds + nn.PixelShuffle(2) as they occur in
 * src/IntraModules/layers.py, src/InterModules/{video_net_component,lssvc_modules}.py,
 * src/models/{dmc_net,LSSVC_net,IntraSS,priors}.py.
 *
 *   acc  = sum over sources, taps, channels  in(src) * weight          (+ bias)
 *   v    = act(acc) * out_scale
 *   v   += res1 + res2                      (either may be absent: ptr == NULL)
 *   out  = v                                (optionally through PixelShuffle(2))
 *   out2 = lrelu(v, slope2)                 (optional second copy, same addressing mode)
 *
 * weight is packed [kh*kw][n_pad][cin_total] (K-major rows, full fp32), bias is [n_pad];
 * cin_total = sum of src[i].C, n_pad = cout rounded up to 16.
 * With pixel_shuffle the packed output-channel order is (2*i + j) * (cout/4) + c for
 * reference channel 4*c + 2*i + j, so that each sub-pixel's channels are contiguous.
 */
typedef struct {
  int32_t n_src;
  lssvc_view src[LSSVC_MAX_SRC];
  const float *weight;
  const float *bias;
  int32_t kh, kw, stride, pad;
  int32_t cout, n_pad, cin_total;
  int32_t in_transform;
  float in_slope;
  int32_t epi;
  int32_t act;
  float slope;
  float out_scale;
  int32_t pixel_shuffle;
  lssvc_view out;
  lssvc_view res1, res2;
  lssvc_view out2;
  float slope2;
  lssvc_view gdn_x;
  /* lssvc_conv_hs: LSSVC_PREC_H2; weight_split is reserved (NULL) */
  int32_t precision;
  const float *weight_split;
  /* lssvc_conv_hs only: fp16 weights [kh*kw][2 (hi, lo)][n_pad][cin_pad16] of w * 2^w_shift (a power of two that
   * moves max|w| to [2^13, 2^14) so that w_lo stays a normal fp16), cin_pad16 = sum of src[i].C rounded up to 16
   * (each source's channels start at a multiple of 16), acc_scale = 2^-w_shift applied to the accumulator. */
  const void *weight_h2;
  int32_t cin_pad16;
  float acc_scale;
  /* entropy epilogue (epi = LSSVC_EPI_LAPLACE / _FOURPART / _BITPARM, lssvc_conv_hs only; needs out_scale = 1, no res2 /
   * PixelShuffle / out2, cout a multiple of 16; LAPLACE / FOURPART beyond one channel tile: see ent_tile):
   *   ent_y      LAPLACE: the latent, C = cout / 2 channels, same H x W as the output
   *   ent_y_hat  LAPLACE: the quantised latent (C channels)
   *   ent_coef   BITPARM: [cout][11] coefficient table (lssvc_bitparm_quant)
   *   ent_bits   device double the bits are added to (may be NULL)
   *   ent_sym / ent_index  int32 NCHW dumps for the rANS coder (may be NULL); ent_thr / ent_n_thr: the scale thresholds
   *                        of the CDF rows (needed with ent_index) */
  lssvc_view ent_y, ent_y_hat;
  const float *ent_coef;
  double *ent_bits;
  int32_t *ent_sym, *ent_index;
  const float *ent_thr;
  int32_t ent_n_thr;
  /* LAPLACE over SEVERAL channel tiles (2C > 128, e.g. the 2 x 96 parameters of the base-layer y): the packed output channels
   * are interleaved per tile so that a channel's scale and mean meet in one accumulator — tile t (ent_tile packed channels)
   * holds [scale of channels t*ent_tile/2 .. | their means]; `out` still receives (scale | mean) in natural order.
   * ent_tile = that tile width (a multiple of 32 dividing cout, = the kernel's own channel tile); 0 = natural order, one tile. */
  int32_t ent_tile;
  int32_t ent_step; /* FOURPART: coding step 0..3 */
} lssvc_conv;

/*
 * Fused ConvFFN of DepthConvBlock (lssvc_modules.py:42-60):
 *   out = in + lrelu(W2 . lrelu(W1 . in + b1, slope1) + b2, slope2)  (+ res2),   W1: C -> hidden, W2: hidden -> C (1x1)
 * in.C = C in {16, 32, 48, 64}, hidden a multiple of 64 with 8 * C * hidden <= 128 KiB (weights stay in shared memory).
 * w1: fp16 [hidden/32][C/16][2 (hi, lo)][32][16]   = split of W1 * 2^shift1 (rows = hidden channel, 16 input channels)
 * w2: fp16 [hidden/32][2][2 (hi, lo)][C][16]       = split of W2 * 2^shift2 (rows = output channel, 16 hidden channels)
 * both with the two 16-byte halves of row r swapped when (r >> 2) & 1 (SWIZZLE_32B image); scaleN = 2^-shiftN.
 */
typedef struct {
  lssvc_view in, out, res2;
  int32_t hidden;
  const void *w1, *w2;
  const float *b1, *b2; /* [hidden], [C] */
  float scale1, scale2;
  float slope1, slope2;
} lssvc_ffn;

/*
 * Pointwise convolution with resident weights, optionally fused with the depthwise 3x3 that precedes it in DepthConv
 * (lssvc_modules.py:15-40):  out = act(W . u + bias) * out_scale (+ res1) (+ res2),  u = in  or  dw3x3(in) + dw_bias.
 * in.C = Cin (multiple of 16, <= 128), out.C = Cout (multiple of 16, <= 64), 2*Cin + 4*Cout <= 512.
 * w: fp16 [Cin/16][2 (hi, lo)][Cout][16] = split of W * 2^shift, SWIZZLE_32B image (see lssvc_ffn); acc_scale = 2^-shift.
 * dw_weight: fp32 [9][Cin] (tap-major, tap = 3*ky + kx), dw_bias: [Cin]; both NULL for a plain 1x1 conv.
 */
typedef struct {
  lssvc_view in, out, res1, res2;
  const void *w;
  const float *bias;
  const float *dw_weight, *dw_bias;
  int32_t act;
  float slope, out_scale, acc_scale;
} lssvc_pw;

/* ---- library ---------------------------------------------------------------------------- */
int32_t lssvc_abi_version(void);
/* 0 when device `dev` is sm_100-class and the driver entry points needed for TMA resolve. */
int32_t lssvc_device_check(int32_t dev);
const char *lssvc_last_error(void);
/* Range guard of the split-fp16 kernels (csrc/range.cu): writes 1.0 to *dst (device double) if, since the last fetch, an
 * operand of conv_hs / conv_pw / conv_ffn reached the fp16 limit (|x| >= 65520, or x^2 under LSSVC_IN_SQUARE) — the
 * outputs of that launch are then NaN — else 0.0; clears the flag.  Stream-ordered, no host synchronisation. */
int32_t lssvc_range_flag_fetch(double *dst, void *stream);
/* number of kernels this library launched since load (bench.py's gpu_launches) */
int64_t lssvc_launch_count(void);
/* kernels launched by replaying a captured CUDA graph are added by the host layer (n per replay) */
void lssvc_launch_count_add(int64_t n);

/* ---- convolutions ------------------------------------------------------------------------ */
/* tcgen05 kind::f16 implicit GEMM on split-fp16 operands (LSSVC_PREC_H2), the convolution of the path: activations are
 * split in flight (fp32 halo tile by TMA -> [fp16 hi | fp16 lo] in place in shared memory, read by the tensor core straight
 * from there: a filter tap = a shifted descriptor), weights are pre-split (weight_h2); 16x8 / 16x16 pixel tiles, any kernel
 * size up to 7x7, stride 1 or 2, up to 3 concatenated sources, input transform (LeakyReLU, GDN's x^2), GDN / IGDN epilogue,
 * residuals, PixelShuffle (csrc/conv_hs.cu).  Needs every source's C, pitch and pointer 16-byte aligned. */
int32_t lssvc_conv_hs(const lssvc_conv *c, void *stream);
/* fused 1x1 -> LeakyReLU -> 1x1 -> LeakyReLU -> + identity block (see lssvc_ffn) */
int32_t lssvc_conv_ffn(const lssvc_ffn *f, void *stream);
/* resident-weight 1x1 conv, optional fused depthwise 3x3 front end (see lssvc_pw) */
int32_t lssvc_conv_pw(const lssvc_pw *f, void *stream);
/* fp32 CUDA-core implicit GEMM: any shape, also hosts the GDN epilogue and input transforms. */
int32_t lssvc_conv_simt(const lssvc_conv *c, void *stream);
/* Narrow heads on the fp32 CUDA cores (csrc/conv_head.cu): k = 3 (2 <= cout <= 4) or k = 7 (cout = 2), stride 1, pad k/2,
 * one source with C % 8 == 0, plain epilogue (bias, LeakyReLU, out_scale, res1, res2) — SpyNet's 7x7 16->2 (video_net_component.py:
 * 213-230), the 3x3 64->2 flow heads and the 3x3 48->3 / 64->3 reconstruction heads (lssvc_modules.py:279-292, 339-365),
 * for which a 128 x 16 MMA tile is 12 % used.  Reads lssvc_conv::weight (fp32 [k*k][n_pad][cin]).
 * lssvc_conv_head_supported: 1 when the descriptor is one this kernel takes, else 0 (no error is set). */
int32_t lssvc_conv_head(const lssvc_conv *c, void *stream);
int32_t lssvc_conv_head_supported(const lssvc_conv *c);
/* depthwise 3x3, pad 1 (lssvc_modules.py:23-24): weight [9][C], bias [C] */
int32_t lssvc_dwconv3x3(const lssvc_view *in, const float *weight, const float *bias,
                        const lssvc_view *out, void *stream);
/* nn.ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1) (dmc_net.py:198-246):
 * weight packed [9][cin][cout]; act/slope as above. */
int32_t lssvc_deconv3x3_s2(const lssvc_view *in, const float *weight, const float *bias,
                           int32_t act, float slope, const lssvc_view *out, void *stream);

/* ---- layout ------------------------------------------------------------------------------ */
/* NCHW contiguous [C][H][W] -> NHWC view (channels C..out.C-1 of the view are zero filled) */
int32_t lssvc_nchw_to_nhwc(const float *src, int32_t C, const lssvc_view *out, void *stream);
int32_t lssvc_nhwc_to_nchw(const lssvc_view *in, float *dst, void *stream);
/* out = lrelu(in, slope) (slope 1 = copy) on views */
int32_t lssvc_lrelu_copy(const lssvc_view *in, float slope, const lssvc_view *out, void *stream);
/* out = a*wa + b*wb (wa/wb: 1-channel views or NULL for plain add) — hybrid context blend,
 * LSSVC_net.py:253-255 with the 2-way softmax of lssvc_modules.py:133-153 folded in:
 * logits is the 2-channel generator output, wa = sigmoid(l0 - l1), wb = 1 - wa. */
int32_t lssvc_softmax2_blend(const lssvc_view *logits, const lssvc_view *a, const lssvc_view *b,
                             const lssvc_view *out, void *stream);

/* ---- warping / resampling ---------------------------------------------------------------- */
/* flow_warp (video_net_component.py:329-352): bilinear, border clamp, align_corners=True,
 * flow in pixels, channel 0 = x.  flow_scale multiplies the flow first. */
int32_t lssvc_flow_warp(const lssvc_view *src, const lssvc_view *flow, float flow_scale,
                        const lssvc_view *out, void *stream);
/* F.interpolate(mode='bilinear', align_corners=False) to out.H x out.W, times `scale`
 * (video_net_component.py:355-368, lssvc_modules.py:360,393,425, layers.py:269,284) */
int32_t lssvc_bilinear_resize(const lssvc_view *in, float scale, const lssvc_view *out, void *stream);
int32_t lssvc_avgpool2(const lssvc_view *in, const lssvc_view *out, void *stream);
int32_t lssvc_maxpool2(const lssvc_view *in, const lssvc_view *out, void *stream);
/* One SpyNet level prologue (video_net_component.py:241-246 / :319-324):
 * flow_up = 2 * bilinear_x2(flow_coarse); out8 = cat(im1, flow_warp(im2, flow_up), flow_up).
 * flow_coarse may be NULL (level 0: zero flow). */
int32_t lssvc_spynet_prep(const lssvc_view *im1, const lssvc_view *im2, const lssvc_view *flow_coarse,
                          const lssvc_view *out8, const lssvc_view *flow_up, void *stream);
/* OffsetDiversity tail (lssvc_modules.py:95-110): `off` is conv_offset's output at half
 * resolution (3*G*O channels = o1 | o2 | mask); it is bilinearly x2-upsampled on the fly,
 * offset = mag * tanh(o) + flow, the feature is warped per (group, offset) and multiplied by
 * sigmoid(mask); result is the (C * O)-channel tensor fed to the grouped 1x1 fusion conv,
 * which is applied here too: out[g*cg + k] = sum_{j<2cg} w[g*cg+k][j] * warped[g*2cg + j] + b.
 * scratch: device buffer of G * H * W * 4 floats (16-byte aligned) for the group-planar copy of x the coalesced gather
 * reads ([G][H][W] float4); NULL selects the direct NHWC gather (same results, ~4x slower at 1080p). */
int32_t lssvc_offset_diversity(const lssvc_view *x, const lssvc_view *off, const lssvc_view *flow,
                               const float *fusion_w, const float *fusion_b, int32_t groups,
                               int32_t offset_num, float magnitude, const lssvc_view *out, float *scratch,
                               void *stream);

/* ---- entropy models ---------------------------------------------------------------------- */
/* Laplace branch (LSSVC_net.py:154-161, 466-473; dmc_net.py:370-377, 429-449):
 *   q = round(y - mean); y_hat = q + mean; bits += clamp(-log2(P(q; scale) + 1e-5), 0, 50)
 * mean may be NULL (zero).  Any of y_q / y_hat / index may be NULL.  bits is a device double
 * accumulator.  index (int32, NCHW order [C][H][W]) and sym (int32, NCHW) feed the rANS coder:
 * index = GaussianEncoder.build_indexes(scale) through `thresholds[n_thr]`
 * (video_entropy_models.py:309-313). */
int32_t lssvc_laplace_quant(const lssvc_view *y, const lssvc_view *mean, const lssvc_view *scale,
                            const lssvc_view *y_q, const lssvc_view *y_hat, double *bits,
                            int32_t *sym_nchw, int32_t *index_nchw, const float *thresholds,
                            int32_t n_thr, void *stream);
/* One step of the 4-step checkerboard/channel-quarter prior (LSSVC_net.py:288-296, 338-443).
 * params8: 8 chunks (4 scales | 4 means) of C/4 channels; step in 0..3 selects the mask pattern
 * (quarter k is coded on checkerboard phase {0123, 3210, 2301, 1032}[step][k], phase = 2*(y&1) + (x&1)).
 * Writes the step's positions of y_hat (y_hat_so_far), y_q, scales_hat; step 0 also zeroes the rest.
 * bits accumulates the Laplace bits of the step's symbols.  sym/index (optional) receive the step's
 * reduced C/4-channel tensors y_q_w_k / build_indexes(scales_w_k) in NCHW order (write=True branch). */
int32_t lssvc_four_part_step(const lssvc_view *y, const lssvc_view *params8, int32_t step,
                             const lssvc_view *y_hat, const lssvc_view *y_q, const lssvc_view *scales_hat,
                             double *bits, int32_t *sym_nchw, int32_t *index_nchw,
                             const float *thresholds, int32_t n_thr, void *stream);
/* Decoder side (LSSVC_net_extend.py:200-263): CDF rows of the step's scales_r ... */
int32_t lssvc_four_part_index(const lssvc_view *params8, int32_t step, int32_t *index_nchw,
                              const float *thresholds, int32_t n_thr, void *stream);
/* ... and y_hat_so_far += (decoded y_q_r + means) on the step's mask */
int32_t lssvc_four_part_dec_step(const int32_t *sym_nchw, const lssvc_view *params8, int32_t step,
                                 const lssvc_view *y_hat, void *stream);
/* GaussianEncoder.build_indexes / GaussianConditional.build_indexes of a whole tensor, NCHW order */
int32_t lssvc_scale_index(const lssvc_view *scale, int32_t *index_nchw, const float *thresholds,
                          int32_t n_thr, void *stream);
/* decoded int32 NCHW symbols -> NHWC fp32 view, optionally + add (means / medians): the
 * `.to(device)` + `y_q + means_hat` of LSSVC_net_extend.py:108-123, dmc_net_extend.py:112-135 */
int32_t lssvc_symbols_to_view(const int32_t *sym_nchw, const lssvc_view *add, const lssvc_view *out,
                              void *stream);
/* Gaussian-conditional branch of the I-frame models (img_entropy_models.py:650-685):
 * y_hat = round(y - mean) + mean; lik = max(Phi(.5-|v|)/s - Phi(-.5-|v|)/s, 1e-9), s = max(scale, .11);
 * bits += -log2(lik).  index uses GaussianConditional.build_indexes (:687-691). */
int32_t lssvc_gaussian_quant(const lssvc_view *y, const lssvc_view *mean, const lssvc_view *scale,
                             const lssvc_view *y_hat, double *bits, int32_t *sym_nchw,
                             int32_t *index_nchw, const float *thresholds, int32_t n_thr, void *stream);
/* Factorised prior of the P-frame models, BitEstimator (video_entropy_models.py:110-166):
 * z_hat = round(z); p = F(z_hat + .5) - F(z_hat - .5); bits += clamp(-log2(p + 1e-5), 0, 50).
 * coef is [C][11] = softplus(h1..4), b1..4, tanh(a1..3) per channel. */
int32_t lssvc_bitparm_quant(const lssvc_view *z, const float *coef, const lssvc_view *z_hat,
                            double *bits, int32_t *sym_nchw, void *stream);
/* Factorised prior of the I-frame models, EntropyBottleneck (img_entropy_models.py:483-554):
 * z_hat = round(z - med) + med, likelihood through the 1-3-3-3-3-1 logistic MLP.
 * coef is [C][59]: softplus(matrices) 3,9,9,9,3 | biases 3,3,3,3,1 | tanh(factors) 3,3,3,3 | median. */
int32_t lssvc_eb_quant(const lssvc_view *z, const float *coef, const lssvc_view *z_hat,
                       double *bits, int32_t *sym_nchw, void *stream);
/* sum of squared error between two views (PSNR statistics gathered over NCCL) */
int32_t lssvc_sse(const lssvc_view *a, const lssvc_view *b, double *out, void *stream);

/* ---- front end of the frame loop (test.py:185-199, 253-254): planar fp32 [C][H][W], the layout of the tensors
 * test.py hands to the models ---------------------------------------------------------------------------------------------- */
/* One 8-bit YUV 4:2:0 frame (YUVReader.read_one_frame, video_reader.py:139-155: y [H][W], uv [2][H/2][W/2], all device
 * pointers) -> RGB [3][Hp][Wp] in [0, 1], zero padded on the right / bottom (F.pad(rgb, P_HR), test.py:192-197):
 * planes / 255, chroma x2 like scipy.ndimage.zoom(uv, (1, 2, 2), order=1), BT.709, clip (ycbcr420_to_rgb, functional.py:42-58). */
int32_t lssvc_yuv420_to_rgb(const uint8_t *y, const uint8_t *uv, int32_t H, int32_t W, float *rgb, int32_t Hp, int32_t Wp,
                            void *stream);
/* One separable pass of the MATLAB-compatible resize (resize_1d, utils/core.py:268-337; imresize runs it over rows, then
 * columns: core.py:417-418): out[c][i][x] = sum_k w[i*K + k] * in[c][taps[i*K + k]][x] for dim = 0 (rows; out is
 * [C][n_out][Wi]) or the same along x for dim = 1 (out is [C][Hi][n_out]).  w / taps: the normalised kernel weights and the
 * reflect-resolved source indices of every output sample (they depend on the sizes only; lssvc_b200/frontend.py builds them
 * as core.py:299-317 does).  clamp01 != 0 applies the .clamp_(0, 1) of test.py:199. */
int32_t lssvc_resample_1d(const float *in, int32_t C, int32_t Hi, int32_t Wi, int32_t dim, const float *w, const int32_t *taps,
                          int32_t K, int32_t n_out, float *out, int32_t clamp01, void *stream);
/* sum of (a - b)^2 over n floats into *out (double; zeroed first): PSNR = 10 log10(n / sum) (test.py:115-118) */
int32_t lssvc_sse_flat(const float *a, const float *b, int64_t n, double *out, void *stream);

/* ---- rANS entropy coder (host code; src/cpp/rans/rans_interface.cpp, src/cpp/ops/ops.cpp) --- */
typedef struct lssvc_rans_encoder lssvc_rans_encoder;
typedef struct lssvc_rans_decoder lssvc_rans_decoder;

/* MLCodec_CXX.pmf_to_quantized_cdf (ops.cpp:24-82): cdf_out has n + 1 entries. host pointers. */
int32_t lssvc_pmf_to_quantized_cdf(const float *pmf, int32_t n, int32_t precision, uint32_t *cdf_out);

/* BufferedRansEncoder (rans_interface.cpp:85-172). All pointers are host pointers;
 * cdfs is row-major [n_rows][cdf_stride], cdf_sizes / offsets have n_rows entries.  Every index is checked against n_rows and
 * every row's size against cdf_stride (LSSVC_ERR_ARG; the reference only asserts, rans_interface.cpp:101-108). */
lssvc_rans_encoder *lssvc_rans_encoder_new(void);
void lssvc_rans_encoder_free(lssvc_rans_encoder *e);
void lssvc_rans_encoder_reset(lssvc_rans_encoder *e);
int32_t lssvc_rans_encode_with_indexes(lssvc_rans_encoder *e, const int32_t *symbols,
                                       const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                       int32_t n_rows, int32_t cdf_stride, const int32_t *cdf_sizes,
                                       const int32_t *offsets);
/* Finishes the stream; returns its size in bytes and a pointer valid until the next
 * reset/flush/free.  Clears the symbol buffer like the reference's flush(). */
int64_t lssvc_rans_encoder_flush(lssvc_rans_encoder *e, const uint8_t **data);

/* RansDecoder (rans_interface.cpp:176-244).  A stream that runs out, or whose bypass digit count is impossible for a 32-bit
 * symbol (> 8 nibbles), returns LSSVC_ERR_STREAM instead of silently desynchronising. */
lssvc_rans_decoder *lssvc_rans_decoder_new(void);
void lssvc_rans_decoder_free(lssvc_rans_decoder *d);
int32_t lssvc_rans_decoder_set_stream(lssvc_rans_decoder *d, const uint8_t *data, int64_t nbytes);
int32_t lssvc_rans_decode_stream(lssvc_rans_decoder *d, const int32_t *indexes, int64_t n,
                                 const int32_t *cdfs, int32_t n_rows, int32_t cdf_stride,
                                 const int32_t *cdf_sizes, const int32_t *offsets, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* LSSVC_B200_H */
