// Front end of the reference's frame loop (test.py:185-199, 253-254): 8-bit YUV 4:2:0 planes -> padded RGB
// float planes, the MATLAB-compatible bicubic EL -> BL resize (src/utils/core.py) as two separable passes, and the squared
// error behind PSNR.  Planar (NCHW) fp32 on both sides: this is the layout of the tensors test.py hands to the models.
// All three are streaming kernels; arithmetic follows the reference operation by operation (no FMA contraction) so that
// results can be compared bit for bit with the CPU path.
#include "common.cuh"

namespace {

constexpr int TPB = 256;

// scipy.ndimage.zoom(uv, (1, 2, 2), order=1) as functional.py:49 calls it (grid_mode=False, mode='constant'): output sample o
// sits at input coordinate o * (n_in - 1) / (n_out - 1); linear interpolation evaluated in double, result cast to float.
__device__ __forceinline__ void zoom_coord(int o, double zoom, int n_in, int &i0, int &i1, double &t) {
  const double cc = static_cast<double>(o) * zoom;
  int s = static_cast<int>(floor(cc));
  s = s < 0 ? 0 : (s > n_in - 1 ? n_in - 1 : s);
  i0 = s;
  i1 = s + 1 < n_in ? s + 1 : n_in - 1;
  t = cc - static_cast<double>(s);
}

__global__ void __launch_bounds__(TPB) yuv420_to_rgb_kernel(const uint8_t *__restrict__ yp, const uint8_t *__restrict__ uvp, int H,
                                                            int W, float *__restrict__ rgb, int Hp, int Wp, double zoom_h,
                                                            double zoom_w) {
  const long long idx = static_cast<long long>(blockIdx.x) * TPB + threadIdx.x;
  const long long plane = static_cast<long long>(Hp) * Wp;
  if (idx >= plane) return;
  const int py = static_cast<int>(idx / Wp), px = static_cast<int>(idx - static_cast<long long>(py) * Wp);
  float r = 0.f, g = 0.f, b = 0.f;   // F.pad(..., mode="constant", value=0)  test.py:192-197
  if (py < H && px < W) {
    const int Hc = H >> 1, Wc = W >> 1;
    const float y = __fdiv_rn(static_cast<float>(yp[static_cast<long long>(py) * W + px]), 255.f);  // video_reader.py:152-153
    int y0, y1, x0, x1;
    double ty, tx;
    zoom_coord(py, zoom_h, Hc, y0, y1, ty);
    zoom_coord(px, zoom_w, Wc, x0, x1, tx);
    float c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint8_t *p = uvp + static_cast<long long>(k) * Hc * Wc;
      const double v00 = static_cast<double>(__fdiv_rn(static_cast<float>(p[y0 * Wc + x0]), 255.f));
      const double v01 = static_cast<double>(__fdiv_rn(static_cast<float>(p[y0 * Wc + x1]), 255.f));
      const double v10 = static_cast<double>(__fdiv_rn(static_cast<float>(p[y1 * Wc + x0]), 255.f));
      const double v11 = static_cast<double>(__fdiv_rn(static_cast<float>(p[y1 * Wc + x1]), 255.f));
      // separable order-1 spline: weights (1 - t, t) per axis, products summed in double
      const double wy0 = 1.0 - ty, wx0 = 1.0 - tx;
      c[k] = static_cast<float>(wy0 * wx0 * v00 + wy0 * tx * v01 + ty * wx0 * v10 + ty * tx * v11);
    }
    // functional.py:50-57 in numpy float32 (the Python-float constants are rounded to float32 where they meet an array)
    const float Kr = 0.2126f, Kg = 0.7152f, Kb = 0.0722f;
    const float cr_gain = static_cast<float>(2.0 - 2.0 * 0.2126), cb_gain = static_cast<float>(2.0 - 2.0 * 0.0722);
    r = __fadd_rn(y, __fmul_rn(cr_gain, __fadd_rn(c[1], -0.5f)));
    b = __fadd_rn(y, __fmul_rn(cb_gain, __fadd_rn(c[0], -0.5f)));
    g = __fdiv_rn(__fadd_rn(__fadd_rn(y, -__fmul_rn(Kr, r)), -__fmul_rn(Kb, b)), Kg);
    r = fminf(fmaxf(r, 0.f), 1.f);
    g = fminf(fmaxf(g, 0.f), 1.f);
    b = fminf(fmaxf(b, 0.f), 1.f);
  }
  rgb[idx] = r;
  rgb[plane + idx] = g;
  rgb[2 * plane + idx] = b;
}

// One pass of core.py:268-337 (resize_1d): out[c][i][x] = sum_k w[i][k] * in[c][idx[i][k]][x] (ROWS) or the same along x.
// w and idx come from the host (they depend on the sizes only); the taps are accumulated in order k = 0 .. K-1 with separate
// multiply and add, the order of the reference's (sample * weight).sum(dim=1).
template <bool ROWS>
__global__ void __launch_bounds__(TPB) resample_1d_kernel(const float *__restrict__ in, int Hi, int Wi, const float *__restrict__ w,
                                                          const int *__restrict__ tap, int K, int n_out, float *__restrict__ out,
                                                          int clamp01, long long total) {
  const long long gid = static_cast<long long>(blockIdx.x) * TPB + threadIdx.x;
  if (gid >= total) return;
  const int Ho = ROWS ? n_out : Hi, Wo = ROWS ? Wi : n_out;
  const int x = static_cast<int>(gid % Wo);
  const long long t = gid / Wo;
  const int y = static_cast<int>(t % Ho);
  const long long c = t / Ho;
  const float *plane = in + c * static_cast<long long>(Hi) * Wi;
  const int o = ROWS ? y : x;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const int j = __ldg(tap + o * K + k);
    const float v = ROWS ? __ldg(plane + static_cast<long long>(j) * Wi + x) : __ldg(plane + static_cast<long long>(y) * Wi + j);
    acc = __fadd_rn(acc, __fmul_rn(v, __ldg(w + o * K + k)));
  }
  if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
  out[gid] = acc;
}

// sum of (a - b)^2 over n floats, accumulated in double (PSNR of test.py:115-118 is 10 log10(1 / mean))
__global__ void __launch_bounds__(TPB) sse_flat_kernel(const float *__restrict__ a, const float *__restrict__ b, long long n,
                                                       double *__restrict__ out) {
  double s = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * TPB + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * TPB) {
    const float d = __fadd_rn(a[i], -b[i]);
    s += static_cast<double>(__fmul_rn(d, d));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[TPB / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int i = 0; i < TPB / 32; ++i) tsum += ws[i];
    atomicAdd(out, tsum);
  }
}

}  // namespace

extern "C" int32_t lssvc_yuv420_to_rgb(const uint8_t *y, const uint8_t *uv, int32_t H, int32_t W, float *rgb, int32_t Hp, int32_t Wp,
                                       void *stream) {
  LSSVC_REQUIRE(y && uv && rgb, "yuv420_to_rgb: null pointer");
  LSSVC_REQUIRE(H >= 4 && W >= 4 && H % 2 == 0 && W % 2 == 0, "yuv420_to_rgb: %dx%d must be even and at least 4x4", H, W);
  LSSVC_REQUIRE(Hp >= H && Wp >= W, "yuv420_to_rgb: padded size %dx%d smaller than the picture %dx%d", Hp, Wp, H, W);
  const long long plane = static_cast<long long>(Hp) * Wp;
  const double zoom_h = static_cast<double>(H / 2 - 1) / static_cast<double>(H - 1);
  const double zoom_w = static_cast<double>(W / 2 - 1) / static_cast<double>(W - 1);
  yuv420_to_rgb_kernel<<<static_cast<unsigned>((plane + TPB - 1) / TPB), TPB, 0, lssvc::as_stream(stream)>>>(y, uv, H, W, rgb, Hp, Wp,
                                                                                                       zoom_h, zoom_w);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_resample_1d(const float *in, int32_t C, int32_t Hi, int32_t Wi, int32_t dim, const float *w,
                                     const int32_t *taps, int32_t K, int32_t n_out, float *out, int32_t clamp01, void *stream) {
  LSSVC_REQUIRE(in && w && taps && out, "resample_1d: null pointer");
  LSSVC_REQUIRE(C > 0 && Hi > 0 && Wi > 0 && K > 0 && n_out > 0, "resample_1d: bad shape C=%d %dx%d K=%d n_out=%d", C, Hi, Wi, K, n_out);
  LSSVC_REQUIRE(dim == 0 || dim == 1, "resample_1d: dim %d (0 = rows, 1 = columns)", dim);
  const long long total = static_cast<long long>(C) * (dim == 0 ? static_cast<long long>(n_out) * Wi : static_cast<long long>(Hi) * n_out);
  const unsigned grid = static_cast<unsigned>((total + TPB - 1) / TPB);
  if (dim == 0) resample_1d_kernel<true><<<grid, TPB, 0, lssvc::as_stream(stream)>>>(in, Hi, Wi, w, taps, K, n_out, out, clamp01, total);
  else resample_1d_kernel<false><<<grid, TPB, 0, lssvc::as_stream(stream)>>>(in, Hi, Wi, w, taps, K, n_out, out, clamp01, total);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_sse_flat(const float *a, const float *b, int64_t n, double *out, void *stream) {
  LSSVC_REQUIRE(a && b && out && n > 0, "sse_flat: bad arguments");
  long long blocks = (n + TPB - 1) / TPB;
  if (blocks > 148 * 8) blocks = 148 * 8;
  LSSVC_CUDA(cudaMemsetAsync(out, 0, sizeof(double), lssvc::as_stream(stream)));
  sse_flat_kernel<<<static_cast<unsigned>(blocks), TPB, 0, lssvc::as_stream(stream)>>>(a, b, n, out);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
