// Device arithmetic of the entropy models, shared by the stand-alone kernels (entropy.cu) and by the entropy epilogue of
// the convolution that produces the parameters (conv_hs.cu, LSSVC_EPI_LAPLACE / LSSVC_EPI_BITPARM): one definition, so the
// fused and the stand-alone path produce the same symbols, CDF rows and per-element bits bit for bit.
#pragma once

namespace lssvc_ent {

constexpr float LN2 = 0.693147180559945309f;

// torch.distributions.Laplace(0, s).cdf(v) = 0.5 - 0.5 * sign(v) * expm1(-|v| / s)
__device__ __forceinline__ float laplace_cdf(float v, float s) {
  const float sg = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
  return 0.5f - 0.5f * sg * expm1f(-fabsf(v) / s);
}
// clamp(-log(p + 1e-5) / ln 2, 0, 50)
__device__ __forceinline__ float prob_bits(float p) {
  const float b = -1.0f * logf(p + 1e-5f) / LN2;
  return fminf(fmaxf(b, 0.f), 50.f);
}
__device__ __forceinline__ float laplace_bits(float q, float scale) {
  const float s = fminf(fmaxf(scale, 1e-5f), 1e10f);
  return prob_bits(laplace_cdf(q + 0.5f, s) - laplace_cdf(q - 0.5f, s));
}
// number of thresholds <= s: the CDF-table row (build_indexes is a monotone step function of s)
__device__ __forceinline__ int scale_index(float s, const float *__restrict__ thr, int n) {
  s = fmaxf(s, 1e-5f);
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (thr[mid] <= s) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// BitEstimator: f1..f3: x = x * softplus(h) + b; x += tanh(x) * tanh(a); f4: sigmoid(x * softplus(h) + b)
__device__ __forceinline__ float bitparm_cdf(float x, const float *__restrict__ k) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x = x * k[i] + k[4 + i];
    x = x + tanhf(x) * k[8 + i];
  }
  return sigmoidf(x * k[3] + k[7]);
}
__device__ __forceinline__ float bitparm_bits(float q, const float *__restrict__ k) {
  return prob_bits(bitparm_cdf(q + 0.5f, k) - bitparm_cdf(q - 0.5f, k));
}

// Four-part spatial prior (LSSVC_net.py:338-443): at coding step `step` the channel quarter `quarter` is coded at the pixels of
// parity ((y & 1) << 1 | (x & 1)) == four_part_mask(step, quarter); rows {0,1,2,3}, {3,2,1,0}, {2,3,0,1}, {1,0,3,2} packed 2 bits each
__device__ __forceinline__ int four_part_mask(int step, int quarter) { return (0xb14e1be4u >> (2 * (4 * step + quarter))) & 3; }

}  // namespace lssvc_ent
