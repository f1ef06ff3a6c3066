// Entropy-model kernels: quantise-with-means, likelihood -> bits (Laplace / Gaussian-erfc / the two
// factorised priors), CDF-row index build and symbol packing for the rANS coder.
// Element-wise over NHWC fp32 views, one thread per element, bits reduced by warp shuffles and one
// double atomicAdd per block.
#include "common.cuh"
#include "entropy_math.cuh"

namespace {

using namespace lssvc_ent;

constexpr int TPB = 256;

inline int blocks_for(long long total) { return static_cast<int>((total + TPB - 1) / TPB); }

__device__ __forceinline__ void block_add(double local, double *out) {
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double warp_sums[TPB / 32];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < TPB / 32; ++i) s += warp_sums[i];
    if (s != 0.0) atomicAdd(out, s);
  }
}

__global__ void laplace_quant_kernel(const float *__restrict__ y, int yp, const float *__restrict__ mean, int mp,
                                     const float *__restrict__ scale, int sp, float *__restrict__ yq, int qp,
                                     float *__restrict__ yhat, int hp, double *__restrict__ bits, int *__restrict__ sym,
                                     int *__restrict__ index, const float *__restrict__ thr, int n_thr, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * C;
  double local = 0.0;
  if (idx < total) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const float m = mean ? mean[pix * mp + c] : 0.f;
    const float s = scale[pix * sp + c];
    const float q = rintf(y[pix * yp + c] - m);
    if (yq) yq[pix * qp + c] = q;
    if (yhat) yhat[pix * hp + c] = q + m;
    local = static_cast<double>(laplace_bits(q, s));
    const long long nchw = static_cast<long long>(c) * H * W + pix;
    if (sym) sym[nchw] = static_cast<int>(q);
    if (index) index[nchw] = scale_index(s, thr, n_thr);
  }
  if (bits) block_add(local, bits);
}

__global__ void four_part_step_kernel(const float *__restrict__ y, int yp, const float *__restrict__ prm, int pp,
                                      int step, float *__restrict__ yhat, int hp, float *__restrict__ yq, int qp,
                                      float *__restrict__ sh, int shp, double *__restrict__ bits, int *__restrict__ sym,
                                      int *__restrict__ index, const float *__restrict__ thr, int n_thr, int H, int W,
                                      int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * C;
  double local = 0.0;
  if (idx < total) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const int x = static_cast<int>(pix % W), yy = static_cast<int>(pix / W);
    const int cq = C >> 2;
    const int quarter = c / cq;
    const int parity = ((yy & 1) << 1) | (x & 1);
    const bool active = four_part_mask(step, quarter) == parity;
    if (active) {
      const float s = prm[pix * pp + c];
      const float m = prm[pix * pp + C + c];
      const float q = rintf(y[pix * yp + c] - m);
      yhat[pix * hp + c] = q + m;
      if (yq) yq[pix * qp + c] = q;
      if (sh) sh[pix * shp + c] = s;
      local = static_cast<double>(laplace_bits(q, s));
      const long long nchw = static_cast<long long>(c - quarter * cq) * H * W + pix;
      if (sym) sym[nchw] = static_cast<int>(q);
      if (index) index[nchw] = scale_index(s, thr, n_thr);
    } else if (step == 0) {
      yhat[pix * hp + c] = 0.f;
      if (yq) yq[pix * qp + c] = 0.f;
      if (sh) sh[pix * shp + c] = 0.f;
    }
  }
  if (bits) block_add(local, bits);
}

__global__ void four_part_index_kernel(const float *__restrict__ prm, int pp, int step, int *__restrict__ index,
                                       const float *__restrict__ thr, int n_thr, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  const int x = static_cast<int>(pix % W), yy = static_cast<int>(pix / W);
  const int cq = C >> 2;
  const int quarter = c / cq;
  if (four_part_mask(step, quarter) != (((yy & 1) << 1) | (x & 1))) return;
  index[static_cast<long long>(c - quarter * cq) * H * W + pix] = scale_index(prm[pix * pp + c], thr, n_thr);
}

__global__ void four_part_dec_kernel(const int *__restrict__ sym, const float *__restrict__ prm, int pp, int step,
                                     float *__restrict__ yhat, int hp, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  const int x = static_cast<int>(pix % W), yy = static_cast<int>(pix / W);
  const int cq = C >> 2;
  const int quarter = c / cq;
  const bool active = four_part_mask(step, quarter) == (((yy & 1) << 1) | (x & 1));
  if (active) {
    const float q = static_cast<float>(sym[static_cast<long long>(c - quarter * cq) * H * W + pix]);
    yhat[pix * hp + c] = q + prm[pix * pp + C + c];
  } else if (step == 0) {
    yhat[pix * hp + c] = 0.f;
  }
}

// 0.5 * erfc(-(2^-0.5) * x)
__device__ __forceinline__ float std_cumulative(float x) { return 0.5f * erfcf(-0.70710678118654752440f * x); }

__global__ void gaussian_quant_kernel(const float *__restrict__ y, int yp, const float *__restrict__ mean, int mp,
                                      const float *__restrict__ scale, int sp, float *__restrict__ yhat, int hp,
                                      double *__restrict__ bits, int *__restrict__ sym, int *__restrict__ index,
                                      const float *__restrict__ thr, int n_thr, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * C;
  double local = 0.0;
  if (idx < total) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const float m = mean[pix * mp + c];
    const float s_raw = scale[pix * sp + c];
    const float q = rintf(y[pix * yp + c] - m);
    const float out = q + m;
    if (yhat) yhat[pix * hp + c] = out;
    // the reference evaluates the likelihood on (dequantised - mean), not on q itself
    const float v = fabsf(out - m);
    const float s = fmaxf(s_raw, 0.11f);
    const float upper = std_cumulative((0.5f - v) / s);
    const float lower = std_cumulative((-0.5f - v) / s);
    const float lik = fmaxf(upper - lower, 1e-9f);
    local = static_cast<double>(-logf(lik) / LN2);
    const long long nchw = static_cast<long long>(c) * H * W + pix;
    if (sym) sym[nchw] = static_cast<int>(q);
    if (index) index[nchw] = scale_index(s_raw, thr, n_thr);
  }
  if (bits) block_add(local, bits);
}

__global__ void bitparm_quant_kernel(const float *__restrict__ z, int zp, const float *__restrict__ coef,
                                     float *__restrict__ zhat, int hp, double *__restrict__ bits, int *__restrict__ sym,
                                     int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * C;
  double local = 0.0;
  if (idx < total) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const float q = rintf(z[pix * zp + c]);
    if (zhat) zhat[pix * hp + c] = q;
    const float *k = coef + c * 11;
    local = static_cast<double>(bitparm_bits(q, k));
    if (sym) sym[static_cast<long long>(c) * H * W + pix] = static_cast<int>(q);
  }
  if (bits) block_add(local, bits);
}

// EntropyBottleneck._logits_cumulative with filters (3,3,3,3)
__device__ __forceinline__ float eb_logits(float x, const float *__restrict__ k) {
  const float *m0 = k, *m1 = k + 3, *m2 = k + 12, *m3 = k + 21, *m4 = k + 30;
  const float *b = k + 33, *f = k + 46;
  float v[3], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = m0[i] * x + b[i];
    v[i] += f[i] * tanhf(v[i]);
  }
  const float *mm[3] = {m1, m2, m3};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // torch.matmul on a [3,3] x [3,n] problem accumulates left to right
      float a = mm[l][i * 3 + 0] * v[0];
      a = fmaf(mm[l][i * 3 + 1], v[1], a);
      a = fmaf(mm[l][i * 3 + 2], v[2], a);
      t[i] = a + b[3 * (l + 1) + i];
      t[i] += f[3 * (l + 1) + i] * tanhf(t[i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = t[i];
  }
  float a = m4[0] * v[0];
  a = fmaf(m4[1], v[1], a);
  a = fmaf(m4[2], v[2], a);
  return a + b[12];
}

__global__ void eb_quant_kernel(const float *__restrict__ z, int zp, const float *__restrict__ coef,
                                float *__restrict__ zhat, int hp, double *__restrict__ bits, int *__restrict__ sym, int H,
                                int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * C;
  double local = 0.0;
  if (idx < total) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const float *k = coef + c * 59;
    const float med = k[58];
    const float q = rintf(z[pix * zp + c] - med);
    const float out = q + med;
    if (zhat) zhat[pix * hp + c] = out;
    const float lower = eb_logits(out - 0.5f, k);
    const float upper = eb_logits(out + 0.5f, k);
    const float sum = lower + upper;
    const float sg = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : -0.f);
    const float lik = fmaxf(fabsf(sigmoidf(sg * upper) - sigmoidf(sg * lower)), 1e-9f);
    local = static_cast<double>(-logf(lik) / LN2);
    if (sym) sym[static_cast<long long>(c) * H * W + pix] = static_cast<int>(q);
  }
  if (bits) block_add(local, bits);
}

bool same_shape(const lssvc_view *a, const lssvc_view *b) { return a->H == b->H && a->W == b->W && a->C == b->C; }

}  // namespace

extern "C" int32_t lssvc_laplace_quant(const lssvc_view *y, const lssvc_view *mean, const lssvc_view *scale,
                                       const lssvc_view *y_q, const lssvc_view *y_hat, double *bits, int32_t *sym_nchw,
                                       int32_t *index_nchw, const float *thresholds, int32_t n_thr, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(y) && lssvc::view_ok(scale) && same_shape(y, scale), "laplace_quant: bad y/scale");
  const bool has_mean = lssvc::view_present(mean), has_q = lssvc::view_present(y_q), has_h = lssvc::view_present(y_hat);
  LSSVC_REQUIRE(!has_mean || same_shape(y, mean), "laplace_quant: mean shape");
  LSSVC_REQUIRE(!has_q || same_shape(y, y_q), "laplace_quant: y_q shape");
  LSSVC_REQUIRE(!has_h || same_shape(y, y_hat), "laplace_quant: y_hat shape");
  LSSVC_REQUIRE(!index_nchw || (thresholds && n_thr > 0), "laplace_quant: index needs thresholds");
  const long long total = static_cast<long long>(y->H) * y->W * y->C;
  laplace_quant_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      y->ptr, y->pitch, has_mean ? mean->ptr : nullptr, has_mean ? mean->pitch : 0, scale->ptr, scale->pitch,
      has_q ? y_q->ptr : nullptr, has_q ? y_q->pitch : 0, has_h ? y_hat->ptr : nullptr, has_h ? y_hat->pitch : 0, bits,
      sym_nchw, index_nchw, thresholds, n_thr, y->H, y->W, y->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_four_part_step(const lssvc_view *y, const lssvc_view *params8, int32_t step,
                                        const lssvc_view *y_hat, const lssvc_view *y_q, const lssvc_view *scales_hat,
                                        double *bits, int32_t *sym_nchw, int32_t *index_nchw, const float *thresholds,
                                        int32_t n_thr, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(y) && lssvc::view_ok(params8) && lssvc::view_ok(y_hat), "four_part_step: bad view");
  LSSVC_REQUIRE(step >= 0 && step < 4 && y->C % 4 == 0, "four_part_step: step=%d C=%d", step, y->C);
  LSSVC_REQUIRE(params8->C == 2 * y->C && params8->H == y->H && params8->W == y->W && same_shape(y, y_hat),
                "four_part_step: params8 must carry 2C channels");
  const bool has_q = lssvc::view_present(y_q), has_s = lssvc::view_present(scales_hat);
  LSSVC_REQUIRE((!has_q || same_shape(y, y_q)) && (!has_s || same_shape(y, scales_hat)), "four_part_step: shape");
  LSSVC_REQUIRE(!index_nchw || (thresholds && n_thr > 0), "four_part_step: index needs thresholds");
  const long long total = static_cast<long long>(y->H) * y->W * y->C;
  four_part_step_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      y->ptr, y->pitch, params8->ptr, params8->pitch, step, y_hat->ptr, y_hat->pitch, has_q ? y_q->ptr : nullptr,
      has_q ? y_q->pitch : 0, has_s ? scales_hat->ptr : nullptr, has_s ? scales_hat->pitch : 0, bits, sym_nchw,
      index_nchw, thresholds, n_thr, y->H, y->W, y->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_four_part_index(const lssvc_view *params8, int32_t step, int32_t *index_nchw,
                                         const float *thresholds, int32_t n_thr, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(params8) && params8->C % 8 == 0 && step >= 0 && step < 4 && index_nchw && thresholds,
                "four_part_index: bad arguments");
  const int C = params8->C / 2;
  const long long total = static_cast<long long>(params8->H) * params8->W * C;
  four_part_index_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      params8->ptr, params8->pitch, step, index_nchw, thresholds, n_thr, params8->H, params8->W, C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_four_part_dec_step(const int32_t *sym_nchw, const lssvc_view *params8, int32_t step,
                                            const lssvc_view *y_hat, void *stream) {
  LSSVC_REQUIRE(sym_nchw && lssvc::view_ok(params8) && lssvc::view_ok(y_hat) && step >= 0 && step < 4,
                "four_part_dec_step: bad arguments");
  LSSVC_REQUIRE(params8->C == 2 * y_hat->C && params8->H == y_hat->H && params8->W == y_hat->W && y_hat->C % 4 == 0,
                "four_part_dec_step: shape");
  const long long total = static_cast<long long>(y_hat->H) * y_hat->W * y_hat->C;
  four_part_dec_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      sym_nchw, params8->ptr, params8->pitch, step, y_hat->ptr, y_hat->pitch, y_hat->H, y_hat->W, y_hat->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_gaussian_quant(const lssvc_view *y, const lssvc_view *mean, const lssvc_view *scale,
                                        const lssvc_view *y_hat, double *bits, int32_t *sym_nchw, int32_t *index_nchw,
                                        const float *thresholds, int32_t n_thr, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(y) && lssvc::view_ok(mean) && lssvc::view_ok(scale) && same_shape(y, mean) &&
                    same_shape(y, scale),
                "gaussian_quant: bad view");
  const bool has_h = lssvc::view_present(y_hat);
  LSSVC_REQUIRE(!has_h || same_shape(y, y_hat), "gaussian_quant: y_hat shape");
  LSSVC_REQUIRE(!index_nchw || (thresholds && n_thr > 0), "gaussian_quant: index needs thresholds");
  const long long total = static_cast<long long>(y->H) * y->W * y->C;
  gaussian_quant_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      y->ptr, y->pitch, mean->ptr, mean->pitch, scale->ptr, scale->pitch, has_h ? y_hat->ptr : nullptr,
      has_h ? y_hat->pitch : 0, bits, sym_nchw, index_nchw, thresholds, n_thr, y->H, y->W, y->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_bitparm_quant(const lssvc_view *z, const float *coef, const lssvc_view *z_hat, double *bits,
                                       int32_t *sym_nchw, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(z) && coef, "bitparm_quant: bad arguments");
  const bool has_h = lssvc::view_present(z_hat);
  LSSVC_REQUIRE(!has_h || same_shape(z, z_hat), "bitparm_quant: z_hat shape");
  const long long total = static_cast<long long>(z->H) * z->W * z->C;
  bitparm_quant_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      z->ptr, z->pitch, coef, has_h ? z_hat->ptr : nullptr, has_h ? z_hat->pitch : 0, bits, sym_nchw, z->H, z->W, z->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_eb_quant(const lssvc_view *z, const float *coef, const lssvc_view *z_hat, double *bits,
                                  int32_t *sym_nchw, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(z) && coef, "eb_quant: bad arguments");
  const bool has_h = lssvc::view_present(z_hat);
  LSSVC_REQUIRE(!has_h || same_shape(z, z_hat), "eb_quant: z_hat shape");
  const long long total = static_cast<long long>(z->H) * z->W * z->C;
  eb_quant_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      z->ptr, z->pitch, coef, has_h ? z_hat->ptr : nullptr, has_h ? z_hat->pitch : 0, bits, sym_nchw, z->H, z->W, z->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

namespace {
__global__ void scale_index_kernel(const float *__restrict__ scale, int sp, int *__restrict__ index,
                                   const float *__restrict__ thr, int n_thr, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  index[static_cast<long long>(c) * H * W + pix] = scale_index(scale[pix * sp + c], thr, n_thr);
}
}  // namespace

extern "C" int32_t lssvc_scale_index(const lssvc_view *scale, int32_t *index_nchw, const float *thresholds,
                                     int32_t n_thr, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(scale) && index_nchw && thresholds && n_thr > 0, "scale_index: bad arguments");
  const long long total = static_cast<long long>(scale->H) * scale->W * scale->C;
  scale_index_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(scale->ptr, scale->pitch, index_nchw,
                                                                             thresholds, n_thr, scale->H, scale->W, scale->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

// NCHW int32 symbols -> NHWC fp32 view (decoder: z_hat / mv_y_q come back from the rANS decoder)
namespace {
__global__ void sym_to_view_kernel(const int *__restrict__ sym, const float *__restrict__ add, int ap,
                                   float *__restrict__ out, int op, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  float v = static_cast<float>(sym[static_cast<long long>(c) * H * W + pix]);
  if (add) v += add[pix * ap + c];
  out[pix * op + c] = v;
}
}  // namespace

extern "C" int32_t lssvc_symbols_to_view(const int32_t *sym_nchw, const lssvc_view *add, const lssvc_view *out,
                                         void *stream) {
  LSSVC_REQUIRE(sym_nchw && lssvc::view_ok(out), "symbols_to_view: bad arguments");
  const bool has_add = lssvc::view_present(add);
  LSSVC_REQUIRE(!has_add || same_shape(add, out), "symbols_to_view: add shape");
  const long long total = static_cast<long long>(out->H) * out->W * out->C;
  sym_to_view_kernel<<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(
      sym_nchw, has_add ? add->ptr : nullptr, has_add ? add->pitch : 0, out->ptr, out->pitch, out->H, out->W, out->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
