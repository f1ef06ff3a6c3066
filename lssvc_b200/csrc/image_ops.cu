// Memory-bound image-space kernels: layout conversion, flow warping, bilinear resampling, pooling,
// the SpyNet level prologue and the OffsetDiversity tail.  All operate on NHWC fp32 views and are
// written for coalesced, 128-bit accesses along the channel axis.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

inline int blocks_for(long long total) { return static_cast<int>((total + TPB - 1) / TPB); }

// ---- layout ------------------------------------------------------------------------------
// src [C][H][W] -> dst NHWC (pitch), zero fill channels C..Cv-1. 32 pixels x 32 channels tile transposed through
// smem so both sides are coalesced.  blockDim = (32, 8).
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ src, int C, float *__restrict__ dst, int Cv, int pitch,
                                    long long HW) {
  __shared__ float tile[32][33];
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long long pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pp < HW) ? src[static_cast<long long>(c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long pp = p0 + i;
    const int c = c0 + threadIdx.x;
    if (pp < HW && c < Cv) dst[pp * pitch + c] = tile[threadIdx.x][i];
  }
}

__global__ void nhwc_to_nchw_kernel(const float *__restrict__ src, int C, int pitch, float *__restrict__ dst,
                                    long long HW) {
  __shared__ float tile[32][33];
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long pp = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (pp < HW && c < C) ? src[pp * pitch + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long long pp = p0 + threadIdx.x;
    if (c < C && pp < HW) dst[static_cast<long long>(c) * HW + pp] = tile[threadIdx.x][i];
  }
}

// C <= 8 (reconstructed frames, flows): the 32 x 32 transpose tile would run 3 of its 32 channel lanes; one thread per pixel
// instead: C strided reads of one sector, C coalesced plane writes.
__global__ void __launch_bounds__(TPB) nhwc_to_nchw_small_kernel(const float *__restrict__ src, int C, int pitch, float *__restrict__ dst,
                                                                 long long HW) {
  const long long pp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pp >= HW) return;
  const float *s = src + pp * pitch;
  for (int c = 0; c < C; ++c) dst[static_cast<long long>(c) * HW + pp] = s[c];
}

__global__ void __launch_bounds__(TPB) nchw_to_nhwc_small_kernel(const float *__restrict__ src, int C, float *__restrict__ dst, int Cv,
                                                                 int pitch, long long HW) {
  const long long pp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pp >= HW) return;
  float *d = dst + pp * pitch;
  for (int c = 0; c < Cv; ++c) d[c] = c < C ? src[static_cast<long long>(c) * HW + pp] : 0.f;
}

__global__ void lrelu_copy_kernel(const float *__restrict__ in, int in_pitch, float slope, float *__restrict__ out,
                                  int out_pitch, int C, long long pixels) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= pixels * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  out[pix * out_pitch + c] = lrelu(in[pix * in_pitch + c], slope);
}

__global__ void softmax2_blend_kernel(const float *__restrict__ logits, int l_pitch, const float *__restrict__ a,
                                      int a_pitch, const float *__restrict__ b, int b_pitch, float *__restrict__ out,
                                      int o_pitch, int C, long long pixels) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= pixels * C) return;
  const long long pix = idx / C;
  const int c = static_cast<int>(idx - pix * C);
  // softmax over two logits, computed the way torch does (subtract the max, exp, normalise)
  const float l0 = logits[pix * l_pitch], l1 = logits[pix * l_pitch + 1];
  const float mx = fmaxf(l0, l1);
  const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
  const float inv = 1.f / (e0 + e1);
  out[pix * o_pitch + c] = a[pix * a_pitch + c] * (e0 * inv) + b[pix * b_pitch + c] * (e1 * inv);
}

// ---- 128-bit variants ---------------------------------------------------------------------------------------
// The streaming kernels below are HBM-bound: each thread moves UNROLL float4 elements that are a whole block apart
// (coalesced 4 KB per block and step), all loads issued before the first use, 32-bit index arithmetic.
constexpr int UNROLL = 4;

__global__ void __launch_bounds__(TPB) lrelu_copy4_kernel(const float4 *__restrict__ in, uint32_t ip4, float slope,
                                                          float4 *__restrict__ out, uint32_t op4, uint32_t cv, uint32_t total) {
  const uint32_t base = blockIdx.x * (TPB * UNROLL) + threadIdx.x;
  float4 v[UNROLL];
  uint32_t o[UNROLL];
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    const uint32_t i = base + k * TPB;
    if (i < total) {
      const uint32_t pix = i / cv, c = i - pix * cv;
      v[k] = __ldg(in + static_cast<size_t>(pix) * ip4 + c);
      o[k] = pix * op4 + c;   // < 2^32 / 4 elements: checked on the host
    }
  }
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    if (base + k * TPB < total)
      out[o[k]] = make_float4(lrelu(v[k].x, slope), lrelu(v[k].y, slope), lrelu(v[k].z, slope), lrelu(v[k].w, slope));
  }
}

__global__ void __launch_bounds__(TPB) softmax2_blend4_kernel(const float *__restrict__ logits, uint32_t l_pitch,
                                                              const float4 *__restrict__ a, uint32_t ap4,
                                                              const float4 *__restrict__ b, uint32_t bp4,
                                                              float4 *__restrict__ out, uint32_t op4, uint32_t cv, uint32_t total) {
  const uint32_t base = blockIdx.x * (TPB * UNROLL) + threadIdx.x;
  float4 va[UNROLL], vb[UNROLL];
  float l0[UNROLL], l1[UNROLL];
  uint32_t o[UNROLL];
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    const uint32_t i = base + k * TPB;
    if (i < total) {
      const uint32_t pix = i / cv, c = i - pix * cv;
      va[k] = __ldg(a + static_cast<size_t>(pix) * ap4 + c);
      vb[k] = __ldg(b + static_cast<size_t>(pix) * bp4 + c);
      l0[k] = __ldg(logits + static_cast<size_t>(pix) * l_pitch);
      l1[k] = __ldg(logits + static_cast<size_t>(pix) * l_pitch + 1);
      o[k] = pix * op4 + c;
    }
  }
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    if (base + k * TPB < total) {
      // softmax over two logits, computed the way torch does (subtract the max, exp, normalise)
      const float mx = fmaxf(l0[k], l1[k]);
      const float e0 = expf(l0[k] - mx), e1 = expf(l1[k] - mx);
      const float inv = 1.f / (e0 + e1);
      const float wa = e0 * inv, wb = e1 * inv;
      out[o[k]] = make_float4(va[k].x * wa + vb[k].x * wb, va[k].y * wa + vb[k].y * wb, va[k].z * wa + vb[k].z * wb,
                              va[k].w * wa + vb[k].w * wb);
    }
  }
}

// ---- warping -----------------------------------------------------------------------------
// grid_sample(bilinear, border, align_corners=True) on grid = base + flow / ((W-1)/2): the sample
// position in pixels is x + fx (up to the fp32 normalise/denormalise round trip the reference does,
// reproduced here so that results track the reference to the last bits that matter).
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  // torch.linspace(-1, 1, n)[i]: symmetric evaluation from both ends
  const float step = 2.f / static_cast<float>(n - 1);
  return (i < n / 2) ? (-1.f + step * static_cast<float>(i)) : (1.f - step * static_cast<float>(n - 1 - i));
}

__device__ __forceinline__ void warp_coords(int x, int y, float fx, float fy, int W, int H, int &x0, int &x1, int &y0,
                                            int &y1, float &wx, float &wy) {
  // reference: g = linspace(-1,1,W)[x] + fx / ((W-1)/2); ATen: ix = ((g + 1) / 2) * (W - 1), clipped to [0, W-1]
  const float sx = (W > 1) ? linspace_pm1(x, W) : -1.f;
  const float sy = (H > 1) ? linspace_pm1(y, H) : -1.f;
  const float gx = sx + fx / ((static_cast<float>(W) - 1.f) / 2.f);
  const float gy = sy + fy / ((static_cast<float>(H) - 1.f) / 2.f);
  float ix = ((gx + 1.f) / 2.f) * static_cast<float>(W - 1);
  float iy = ((gy + 1.f) / 2.f) * static_cast<float>(H - 1);
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(W - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(H - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  x0 = static_cast<int>(fx0);
  y0 = static_cast<int>(fy0);
  x1 = min(x0 + 1, W - 1);
  y1 = min(y0 + 1, H - 1);
  wx = ix - fx0;
  wy = iy - fy0;
}

__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, float wx, float wy) {
  // ATen grid_sampler_2d corner weights: nw = (x_se - ix)(y_se - iy), ne = (ix - x_sw)(y_sw - iy), ...
  const float ex = 1.f - wx, ey = 1.f - wy;
  return v00 * (ex * ey) + v01 * (wx * ey) + v10 * (ex * wy) + v11 * (wx * wy);
}

// float4 variant: two (pixel, 4-channel) elements per thread, the 8 corner loads in flight together
__global__ void __launch_bounds__(TPB) flow_warp4_kernel(const float *__restrict__ src, uint32_t s_pitch,
                                                         const float *__restrict__ flow, uint32_t f_pitch, float fscale,
                                                         float *__restrict__ out, uint32_t o_pitch, int H, int W, uint32_t cv,
                                                         uint32_t total) {
  constexpr int U = 2;
  const uint32_t base = blockIdx.x * (TPB * U) + threadIdx.x;
  float4 a[U], b[U], d[U], e[U];
  float wx[U], wy[U];
  size_t o[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t i = base + k * TPB;
    if (i < total) {
      const uint32_t pix = i / cv, c = (i - pix * cv) * 4;
      const int y = static_cast<int>(pix / static_cast<uint32_t>(W)), x = static_cast<int>(pix - static_cast<uint32_t>(y) * W);
      const float fx = __ldg(flow + static_cast<size_t>(pix) * f_pitch) * fscale;
      const float fy = __ldg(flow + static_cast<size_t>(pix) * f_pitch + 1) * fscale;
      int x0, x1, y0, y1;
      warp_coords(x, y, fx, fy, W, H, x0, x1, y0, y1, wx[k], wy[k]);
      const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
      a[k] = __ldg(reinterpret_cast<const float4 *>(src + (r0 + x0) * s_pitch + c));
      b[k] = __ldg(reinterpret_cast<const float4 *>(src + (r0 + x1) * s_pitch + c));
      d[k] = __ldg(reinterpret_cast<const float4 *>(src + (r1 + x0) * s_pitch + c));
      e[k] = __ldg(reinterpret_cast<const float4 *>(src + (r1 + x1) * s_pitch + c));
      o[k] = static_cast<size_t>(pix) * o_pitch + c;
    }
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    if (base + k * TPB < total)
      *reinterpret_cast<float4 *>(out + o[k]) =
          make_float4(bilerp(a[k].x, b[k].x, d[k].x, e[k].x, wx[k], wy[k]), bilerp(a[k].y, b[k].y, d[k].y, e[k].y, wx[k], wy[k]),
                      bilerp(a[k].z, b[k].z, d[k].z, e[k].z, wx[k], wy[k]), bilerp(a[k].w, b[k].w, d[k].w, e[k].w, wx[k], wy[k]));
  }
}

// Per-pixel-lane variant (the default): `lpp` consecutive threads own one pixel and each moves Q float4 channel quads
// (quad l + j * lpp for lane l), so the flow load, the grid_sample coordinate arithmetic (4 fp32 divisions) and the pixel
// index divisions are paid once per Q quads instead of once per quad (the float4 kernel above issues ~120 instructions per
// 16 output bytes and sits at 65 % issue utilisation, 55 % of the copy bandwidth), and 4 * Q corner loads are in flight
// per thread.  A lane group reads lpp * 16 contiguous bytes per corner and step.
template <int Q>
__global__ void __launch_bounds__(TPB) flow_warp_q_kernel(const float *__restrict__ src, uint32_t s_pitch,
                                                          const float *__restrict__ flow, uint32_t f_pitch, float fscale,
                                                          float *__restrict__ out, uint32_t o_pitch, int H, int W, uint32_t lpp,
                                                          uint32_t n_threads) {
  const uint32_t gid = blockIdx.x * TPB + threadIdx.x;
  if (gid >= n_threads) return;
  const uint32_t pix = gid / lpp, l = gid - pix * lpp;
  const int y = static_cast<int>(pix / static_cast<uint32_t>(W)), x = static_cast<int>(pix - static_cast<uint32_t>(y) * W);
  const float fx = __ldg(flow + static_cast<size_t>(pix) * f_pitch) * fscale;
  const float fy = __ldg(flow + static_cast<size_t>(pix) * f_pitch + 1) * fscale;
  int x0, x1, y0, y1;
  float wx, wy;
  warp_coords(x, y, fx, fy, W, H, x0, x1, y0, y1, wx, wy);
  const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
  const float4 *p00 = reinterpret_cast<const float4 *>(src + (r0 + x0) * s_pitch) + l;
  const float4 *p01 = reinterpret_cast<const float4 *>(src + (r0 + x1) * s_pitch) + l;
  const float4 *p10 = reinterpret_cast<const float4 *>(src + (r1 + x0) * s_pitch) + l;
  const float4 *p11 = reinterpret_cast<const float4 *>(src + (r1 + x1) * s_pitch) + l;
  float4 a[Q], b[Q], d[Q], e[Q];
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    a[j] = __ldg(p00 + j * lpp);
    b[j] = __ldg(p01 + j * lpp);
    d[j] = __ldg(p10 + j * lpp);
    e[j] = __ldg(p11 + j * lpp);
  }
  float4 *o = reinterpret_cast<float4 *>(out + static_cast<size_t>(pix) * o_pitch) + l;
#pragma unroll
  for (int j = 0; j < Q; ++j)
    o[j * lpp] = make_float4(bilerp(a[j].x, b[j].x, d[j].x, e[j].x, wx, wy), bilerp(a[j].y, b[j].y, d[j].y, e[j].y, wx, wy),
                             bilerp(a[j].z, b[j].z, d[j].z, e[j].z, wx, wy), bilerp(a[j].w, b[j].w, d[j].w, e[j].w, wx, wy));
}

template <int VEC>
__global__ void flow_warp_kernel(const float *__restrict__ src, int s_pitch, const float *__restrict__ flow,
                                 int f_pitch, float fscale, float *__restrict__ out, int o_pitch, int H, int W, int C) {
  const int cv = C / VEC;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(H) * W * cv) return;
  const int c = static_cast<int>(idx % cv) * VEC;
  const long long pix = idx / cv;
  const int x = static_cast<int>(pix % W), y = static_cast<int>(pix / W);
  const float fx = flow[pix * f_pitch] * fscale, fy = flow[pix * f_pitch + 1] * fscale;
  int x0, x1, y0, y1;
  float wx, wy;
  warp_coords(x, y, fx, fy, W, H, x0, x1, y0, y1, wx, wy);
  const float *p00 = src + (static_cast<long long>(y0) * W + x0) * s_pitch + c;
  const float *p01 = src + (static_cast<long long>(y0) * W + x1) * s_pitch + c;
  const float *p10 = src + (static_cast<long long>(y1) * W + x0) * s_pitch + c;
  const float *p11 = src + (static_cast<long long>(y1) * W + x1) * s_pitch + c;
  float *o = out + pix * o_pitch + c;
  if (VEC == 4) {
    const float4 a = *reinterpret_cast<const float4 *>(p00), b = *reinterpret_cast<const float4 *>(p01);
    const float4 d = *reinterpret_cast<const float4 *>(p10), e = *reinterpret_cast<const float4 *>(p11);
    *reinterpret_cast<float4 *>(o) = make_float4(bilerp(a.x, b.x, d.x, e.x, wx, wy), bilerp(a.y, b.y, d.y, e.y, wx, wy),
                                                 bilerp(a.z, b.z, d.z, e.z, wx, wy), bilerp(a.w, b.w, d.w, e.w, wx, wy));
  } else {
    o[0] = bilerp(p00[0], p01[0], p10[0], p11[0], wx, wy);
  }
}

// ---- bilinear resize, align_corners=False -----------------------------------------------------
__device__ __forceinline__ void resize_coord(int d, float scale, int in_size, int &i0, int &i1, float &w1) {
  // ATen area_pixel_compute_source_index(align_corners=False): max((d + .5) * scale - .5, 0)
  float s = (static_cast<float>(d) + 0.5f) * scale - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  w1 = s - static_cast<float>(i0);
}

template <int VEC>
__global__ void bilinear_resize_kernel(const float *__restrict__ in, int i_pitch, int Hi, int Wi,
                                       float *__restrict__ out, int o_pitch, int Ho, int Wo, int C, float rh, float rw,
                                       float mul) {
  const int cv = C / VEC;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Ho) * Wo * cv) return;
  const int c = static_cast<int>(idx % cv) * VEC;
  const long long pix = idx / cv;
  const int x = static_cast<int>(pix % Wo), y = static_cast<int>(pix / Wo);
  int x0, x1, y0, y1;
  float wx, wy;
  resize_coord(x, rw, Wi, x0, x1, wx);
  resize_coord(y, rh, Hi, y0, y1, wy);
  const float hx = 1.f - wx, hy = 1.f - wy;
  const float *p00 = in + (static_cast<long long>(y0) * Wi + x0) * i_pitch + c;
  const float *p01 = in + (static_cast<long long>(y0) * Wi + x1) * i_pitch + c;
  const float *p10 = in + (static_cast<long long>(y1) * Wi + x0) * i_pitch + c;
  const float *p11 = in + (static_cast<long long>(y1) * Wi + x1) * i_pitch + c;
  float *o = out + pix * o_pitch + c;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    // ATen upsample_bilinear2d: h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11)
    const float v = hy * (hx * p00[e] + wx * p01[e]) + wy * (hx * p10[e] + wx * p11[e]);
    o[e] = v * mul;
  }
}

// float4 variants (two elements per thread)
__global__ void __launch_bounds__(TPB) bilinear_resize4_kernel(const float *__restrict__ in, uint32_t i_pitch, int Hi, int Wi,
                                                               float *__restrict__ out, uint32_t o_pitch, int Wo, uint32_t cv,
                                                               uint32_t total, float rh, float rw, float mul) {
  constexpr int U = 2;
  const uint32_t base = blockIdx.x * (TPB * U) + threadIdx.x;
  float4 a[U], b[U], d[U], e[U];
  float wx[U], wy[U];
  size_t o[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t i = base + k * TPB;
    if (i < total) {
      const uint32_t pix = i / cv, c = (i - pix * cv) * 4;
      const int y = static_cast<int>(pix / static_cast<uint32_t>(Wo)), x = static_cast<int>(pix - static_cast<uint32_t>(y) * Wo);
      int x0, x1, y0, y1;
      resize_coord(x, rw, Wi, x0, x1, wx[k]);
      resize_coord(y, rh, Hi, y0, y1, wy[k]);
      const size_t r0 = static_cast<size_t>(y0) * Wi, r1 = static_cast<size_t>(y1) * Wi;
      a[k] = __ldg(reinterpret_cast<const float4 *>(in + (r0 + x0) * i_pitch + c));
      b[k] = __ldg(reinterpret_cast<const float4 *>(in + (r0 + x1) * i_pitch + c));
      d[k] = __ldg(reinterpret_cast<const float4 *>(in + (r1 + x0) * i_pitch + c));
      e[k] = __ldg(reinterpret_cast<const float4 *>(in + (r1 + x1) * i_pitch + c));
      o[k] = static_cast<size_t>(pix) * o_pitch + c;
    }
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    if (base + k * TPB < total) {
      const float hx = 1.f - wx[k], hy = 1.f - wy[k], ax = wx[k], ay = wy[k];
      // ATen upsample_bilinear2d: h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11)
      *reinterpret_cast<float4 *>(out + o[k]) =
          make_float4((hy * (hx * a[k].x + ax * b[k].x) + ay * (hx * d[k].x + ax * e[k].x)) * mul,
                      (hy * (hx * a[k].y + ax * b[k].y) + ay * (hx * d[k].y + ax * e[k].y)) * mul,
                      (hy * (hx * a[k].z + ax * b[k].z) + ay * (hx * d[k].z + ax * e[k].z)) * mul,
                      (hy * (hx * a[k].w + ax * b[k].w) + ay * (hx * d[k].w + ax * e[k].w)) * mul);
    }
  }
}

// per-pixel-lane variant of the resize (see flow_warp_q_kernel)
template <int Q>
__global__ void __launch_bounds__(TPB) bilinear_resize_q_kernel(const float *__restrict__ in, uint32_t i_pitch, int Hi, int Wi,
                                                                float *__restrict__ out, uint32_t o_pitch, int Wo, uint32_t lpp,
                                                                uint32_t n_threads, float rh, float rw, float mul) {
  const uint32_t gid = blockIdx.x * TPB + threadIdx.x;
  if (gid >= n_threads) return;
  const uint32_t pix = gid / lpp, l = gid - pix * lpp;
  const int y = static_cast<int>(pix / static_cast<uint32_t>(Wo)), x = static_cast<int>(pix - static_cast<uint32_t>(y) * Wo);
  int x0, x1, y0, y1;
  float ax, ay;
  resize_coord(x, rw, Wi, x0, x1, ax);
  resize_coord(y, rh, Hi, y0, y1, ay);
  const size_t r0 = static_cast<size_t>(y0) * Wi, r1 = static_cast<size_t>(y1) * Wi;
  const float4 *p00 = reinterpret_cast<const float4 *>(in + (r0 + x0) * i_pitch) + l;
  const float4 *p01 = reinterpret_cast<const float4 *>(in + (r0 + x1) * i_pitch) + l;
  const float4 *p10 = reinterpret_cast<const float4 *>(in + (r1 + x0) * i_pitch) + l;
  const float4 *p11 = reinterpret_cast<const float4 *>(in + (r1 + x1) * i_pitch) + l;
  float4 a[Q], b[Q], d[Q], e[Q];
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    a[j] = __ldg(p00 + j * lpp);
    b[j] = __ldg(p01 + j * lpp);
    d[j] = __ldg(p10 + j * lpp);
    e[j] = __ldg(p11 + j * lpp);
  }
  const float hx = 1.f - ax, hy = 1.f - ay;
  float4 *o = reinterpret_cast<float4 *>(out + static_cast<size_t>(pix) * o_pitch) + l;
#pragma unroll
  for (int j = 0; j < Q; ++j)
    // ATen upsample_bilinear2d: h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11)
    o[j * lpp] = make_float4((hy * (hx * a[j].x + ax * b[j].x) + ay * (hx * d[j].x + ax * e[j].x)) * mul,
                             (hy * (hx * a[j].y + ax * b[j].y) + ay * (hx * d[j].y + ax * e[j].y)) * mul,
                             (hy * (hx * a[j].z + ax * b[j].z) + ay * (hx * d[j].z + ax * e[j].z)) * mul,
                             (hy * (hx * a[j].w + ax * b[j].w) + ay * (hx * d[j].w + ax * e[j].w)) * mul);
}

// Exact x2 upsampling (every resampler of the x2 configurations): one lane group per CELL of the input grid, i.e. the 2 x 2
// input pixels (i, j) .. (i + 1, j + 1), i in [-1, Hi - 1], j in [-1, Wi - 1], clamped at the borders.  The four output pixels
// (2i + 1 .. 2i + 2, 2j + 1 .. 2j + 2) interpolate exactly these four inputs (weights 0.75 / 0.25), so each input quad is loaded
// once per four outputs instead of four times.  Every output still evaluates the same expression with the coordinates and
// weights resize_coord gives it (bit-identical to bilinear_resize_kernel); at the borders the clamped neighbour carries weight 0.
template <int Q>
__global__ void __launch_bounds__(TPB) bilinear_up2_q_kernel(const float *__restrict__ in, uint32_t i_pitch, int Hi, int Wi,
                                                             float *__restrict__ out, uint32_t o_pitch, uint32_t lpp,
                                                             uint32_t n_threads, float mul) {
  const uint32_t gid = blockIdx.x * TPB + threadIdx.x;
  if (gid >= n_threads) return;
  const uint32_t cell = gid / lpp, l = gid - cell * lpp;
  const int cw = Wi + 1;
  const int i = static_cast<int>(cell / static_cast<uint32_t>(cw)) - 1, j = static_cast<int>(cell % static_cast<uint32_t>(cw)) - 1;
  const int ra = i < 0 ? 0 : i, rb = i + 1 < Hi ? i + 1 : Hi - 1, ca = j < 0 ? 0 : j, cb = j + 1 < Wi ? j + 1 : Wi - 1;
  const float4 *p00 = reinterpret_cast<const float4 *>(in + (static_cast<size_t>(ra) * Wi + ca) * i_pitch) + l;
  const float4 *p01 = reinterpret_cast<const float4 *>(in + (static_cast<size_t>(ra) * Wi + cb) * i_pitch) + l;
  const float4 *p10 = reinterpret_cast<const float4 *>(in + (static_cast<size_t>(rb) * Wi + ca) * i_pitch) + l;
  const float4 *p11 = reinterpret_cast<const float4 *>(in + (static_cast<size_t>(rb) * Wi + cb) * i_pitch) + l;
  float4 a[Q], b[Q], d[Q], e[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    a[q] = __ldg(p00 + q * lpp);
    b[q] = __ldg(p01 + q * lpp);
    d[q] = __ldg(p10 + q * lpp);
    e[q] = __ldg(p11 + q * lpp);
  }
  const int Ho = 2 * Hi, Wo = 2 * Wi;
#pragma unroll
  for (int dy = 1; dy <= 2; ++dy) {
    const int oy = 2 * i + dy;
    if (oy < 0 || oy >= Ho) continue;
    int y0, y1;
    float ay;
    resize_coord(oy, 0.5f, Hi, y0, y1, ay);
    const float hy = 1.f - ay;
#pragma unroll
    for (int dx = 1; dx <= 2; ++dx) {
      const int ox = 2 * j + dx;
      if (ox < 0 || ox >= Wo) continue;
      int x0, x1;
      float ax;
      resize_coord(ox, 0.5f, Wi, x0, x1, ax);
      const float hx = 1.f - ax;
      float4 *o = reinterpret_cast<float4 *>(out + (static_cast<size_t>(oy) * Wo + ox) * o_pitch) + l;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        o[q * lpp] = make_float4((hy * (hx * a[q].x + ax * b[q].x) + ay * (hx * d[q].x + ax * e[q].x)) * mul,
                                 (hy * (hx * a[q].y + ax * b[q].y) + ay * (hx * d[q].y + ax * e[q].y)) * mul,
                                 (hy * (hx * a[q].z + ax * b[q].z) + ay * (hx * d[q].z + ax * e[q].z)) * mul,
                                 (hy * (hx * a[q].w + ax * b[q].w) + ay * (hx * d[q].w + ax * e[q].w)) * mul);
    }
  }
}

template <bool MAX>
__global__ void __launch_bounds__(TPB) pool2_4_kernel(const float *__restrict__ in, uint32_t i_pitch, int Wi,
                                                      float *__restrict__ out, uint32_t o_pitch, int Wo, uint32_t cv, uint32_t total) {
  constexpr int U = 2;
  const uint32_t base = blockIdx.x * (TPB * U) + threadIdx.x;
  float4 a[U], b[U], d[U], e[U];
  size_t o[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t i = base + k * TPB;
    if (i < total) {
      const uint32_t pix = i / cv, c = (i - pix * cv) * 4;
      const uint32_t y = pix / static_cast<uint32_t>(Wo), x = pix - y * Wo;
      const float *p = in + (static_cast<size_t>(2 * y) * Wi + 2 * x) * i_pitch + c;
      a[k] = __ldg(reinterpret_cast<const float4 *>(p));
      b[k] = __ldg(reinterpret_cast<const float4 *>(p + i_pitch));
      d[k] = __ldg(reinterpret_cast<const float4 *>(p + static_cast<size_t>(Wi) * i_pitch));
      e[k] = __ldg(reinterpret_cast<const float4 *>(p + static_cast<size_t>(Wi) * i_pitch + i_pitch));
      o[k] = static_cast<size_t>(pix) * o_pitch + c;
    }
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    if (base + k * TPB < total) {
      auto f = [](float p, float q, float r, float t) { return MAX ? fmaxf(fmaxf(p, q), fmaxf(r, t)) : (p + q + r + t) * 0.25f; };
      *reinterpret_cast<float4 *>(out + o[k]) = make_float4(f(a[k].x, b[k].x, d[k].x, e[k].x), f(a[k].y, b[k].y, d[k].y, e[k].y),
                                                            f(a[k].z, b[k].z, d[k].z, e[k].z), f(a[k].w, b[k].w, d[k].w, e[k].w));
    }
  }
}

template <bool MAX>
__global__ void pool2_kernel(const float *__restrict__ in, int i_pitch, int Wi, float *__restrict__ out, int o_pitch,
                             int Ho, int Wo, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Ho) * Wo * C) return;
  const int c = static_cast<int>(idx % C);
  const long long pix = idx / C;
  const int x = static_cast<int>(pix % Wo), y = static_cast<int>(pix / Wo);
  const float *p = in + (static_cast<long long>(2 * y) * Wi + 2 * x) * i_pitch + c;
  const float a = p[0], b = p[i_pitch], d = p[static_cast<long long>(Wi) * i_pitch],
              e = p[static_cast<long long>(Wi) * i_pitch + i_pitch];
  out[pix * o_pitch + c] = MAX ? fmaxf(fmaxf(a, b), fmaxf(d, e)) : (a + b + d + e) * 0.25f;
}

// ---- SpyNet level prologue --------------------------------------------------------------------
__global__ void spynet_prep_kernel(const float *__restrict__ im1, int p1, const float *__restrict__ im2, int p2,
                                   const float *__restrict__ fc, int pc, float *__restrict__ out8, int p8,
                                   float *__restrict__ fup, int pf, int H, int W) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= static_cast<long long>(H) * W) return;
  const int x = static_cast<int>(pix % W), y = static_cast<int>(pix / W);
  float fx = 0.f, fy = 0.f;
  if (fc) {
    const int Hc = H / 2, Wc = W / 2;
    int x0, x1, y0, y1;
    float wx, wy;
    resize_coord(x, 0.5f, Wc, x0, x1, wx);
    resize_coord(y, 0.5f, Hc, y0, y1, wy);
    const float hx = 1.f - wx, hy = 1.f - wy;
    const float *q00 = fc + (static_cast<long long>(y0) * Wc + x0) * pc, *q01 = fc + (static_cast<long long>(y0) * Wc + x1) * pc;
    const float *q10 = fc + (static_cast<long long>(y1) * Wc + x0) * pc, *q11 = fc + (static_cast<long long>(y1) * Wc + x1) * pc;
    fx = (hy * (hx * q00[0] + wx * q01[0]) + wy * (hx * q10[0] + wx * q11[0])) * 2.0f;
    fy = (hy * (hx * q00[1] + wx * q01[1]) + wy * (hx * q10[1] + wx * q11[1])) * 2.0f;
  }
  int x0, x1, y0, y1;
  float wx, wy;
  warp_coords(x, y, fx, fy, W, H, x0, x1, y0, y1, wx, wy);
  const float *a = im2 + (static_cast<long long>(y0) * W + x0) * p2, *b = im2 + (static_cast<long long>(y0) * W + x1) * p2;
  const float *d = im2 + (static_cast<long long>(y1) * W + x0) * p2, *e = im2 + (static_cast<long long>(y1) * W + x1) * p2;
  const float *i1 = im1 + pix * p1;
  float *o = out8 + pix * p8;
  const float4 lo = make_float4(i1[0], i1[1], i1[2], bilerp(a[0], b[0], d[0], e[0], wx, wy));
  const float4 hi = make_float4(bilerp(a[1], b[1], d[1], e[1], wx, wy), bilerp(a[2], b[2], d[2], e[2], wx, wy), fx, fy);
  *reinterpret_cast<float4 *>(o) = lo;
  *reinterpret_cast<float4 *>(o + 4) = hi;
  fup[pix * pf] = fx;
  fup[pix * pf + 1] = fy;
}

// ---- OffsetDiversity tail ----------------------------------------------------------------------
// One thread per (pixel, group).  cg = C / groups channels per group; `off` holds, at half resolution,
// [G*O offsets-x/y interleaved as the reference's (o1 | o2) chunks][G*O masks].
// Reference layout (lssvc_modules.py:96-109): offset = cat(o1, o2) viewed as (G*O, 2): warp n uses
// channels (2n, 2n+1) of the 2*G*O-channel tensor; mask n is channel n of the mask chunk; warp n acts on
// x-group (n mod G) ("x.repeat(offset_num)") and its output lands in channels n*cg .. n*cg+cg-1 of the
// (C*O)-channel tensor, which the grouped 1x1 fusion conv (groups=G, 2*cg inputs per group) consumes.
// VEC_X (CG == 3, feature base 8-byte aligned, even pitch): the 12-byte group of a corner is fetched with ONE 8-byte and ONE
// 4-byte load (which of the two comes first follows the address parity) instead of three scalar loads, and the 18 fusion
// weights + 3 biases of a group are read from shared memory instead of global memory.  ncu on the scalar version
// (profiles/r2_kernels_ncu.txt): 1.28 G L1 sectors for 61 M load requests, l1tex 81 % busy = L1-throughput bound; a warp issued
// 24 feature + 21 weight requests of ~16-21 sectors each per 2 pixels.  Same arithmetic in the same order: bit-identical.
constexpr int OD_MAX_FW = 16 * 3 * 6 + 16 * 3;   // fusion weights + biases staged in shared memory (G <= 16, CG = 3, O = 2)

// I32: every element index of the tensors involved fits 31 bits (all sizes up to 4K do): 32-bit index arithmetic — the 64-bit
// divisions and multiplies of the general version were ~200 of its ~900 instructions per thread.
template <int CG, int O, bool VEC_OFF, bool VEC_X, bool I32>
__global__ void __launch_bounds__(TPB) offset_diversity_kernel(const float *__restrict__ xin, int xp, const float *__restrict__ off, int op,
                                        const float *__restrict__ flow, int fp, const float *__restrict__ fw,
                                        const float *__restrict__ fb, int G, float mag, float *__restrict__ out,
                                        int outp, int H, int W) {
  __shared__ float s_fw[VEC_X ? OD_MAX_FW : 1];
  if (VEC_X) {
    const int n_w = G * CG * O * CG, n_b = G * CG;
    for (int i = threadIdx.x; i < n_w + n_b; i += TPB) s_fw[i] = i < n_w ? fw[i] : fb[i - n_w];
    __syncthreads();
  }
  const float *const wsrc = VEC_X ? s_fw : fw;
  const float *const bsrc = VEC_X ? s_fw + G * CG * O * CG : fb;
  using idx_t = typename std::conditional<I32, int, long long>::type;
  using uidx_t = typename std::conditional<I32, unsigned int, long long>::type;
  const uidx_t idx = static_cast<uidx_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<uidx_t>(H) * W * G) return;
  const int g = static_cast<int>(idx % static_cast<uidx_t>(G));
  const idx_t pix = static_cast<idx_t>(idx / static_cast<uidx_t>(G));
  const int y = static_cast<int>(static_cast<uidx_t>(pix) / static_cast<uidx_t>(W)), x = static_cast<int>(pix - static_cast<idx_t>(y) * W);
  const int Hc = H / 2, Wc = W / 2;
  const int n_off = G * O;
  int bx0, bx1, by0, by1;
  float bwx, bwy;
  resize_coord(x, 0.5f, Wc, bx0, bx1, bwx);
  resize_coord(y, 0.5f, Hc, by0, by1, bwy);
  const float hx = 1.f - bwx, hy = 1.f - bwy;
  const float *q00 = off + (static_cast<idx_t>(by0) * Wc + bx0) * op, *q01 = off + (static_cast<idx_t>(by0) * Wc + bx1) * op;
  const float *q10 = off + (static_cast<idx_t>(by1) * Wc + bx0) * op, *q11 = off + (static_cast<idx_t>(by1) * Wc + bx1) * op;
  // this thread's 2*O offset channels (2n, 2n+1 for n = g*O .. g*O+O-1) and O mask channels are contiguous: with O == 2
  // one float4 + one float2 per corner of the x2 upsampling instead of 6 scalar loads
  float upv[3 * O];
  if (O == 2 && VEC_OFF) {
    const float4 a4 = __ldg(reinterpret_cast<const float4 *>(q00 + 4 * g)), b4 = __ldg(reinterpret_cast<const float4 *>(q01 + 4 * g));
    const float4 d4 = __ldg(reinterpret_cast<const float4 *>(q10 + 4 * g)), e4 = __ldg(reinterpret_cast<const float4 *>(q11 + 4 * g));
    const float2 a2 = __ldg(reinterpret_cast<const float2 *>(q00 + 2 * n_off + 2 * g)), b2 = __ldg(reinterpret_cast<const float2 *>(q01 + 2 * n_off + 2 * g));
    const float2 d2 = __ldg(reinterpret_cast<const float2 *>(q10 + 2 * n_off + 2 * g)), e2 = __ldg(reinterpret_cast<const float2 *>(q11 + 2 * n_off + 2 * g));
    auto up4 = [&](float v00, float v01, float v10, float v11) { return hy * (hx * v00 + bwx * v01) + bwy * (hx * v10 + bwx * v11); };
    upv[0] = up4(a4.x, b4.x, d4.x, e4.x); upv[1] = up4(a4.y, b4.y, d4.y, e4.y);
    upv[2] = up4(a4.z, b4.z, d4.z, e4.z); upv[3] = up4(a4.w, b4.w, d4.w, e4.w);
    upv[4] = up4(a2.x, b2.x, d2.x, e2.x); upv[5] = up4(a2.y, b2.y, d2.y, e2.y);
  } else {
    auto up1 = [&](int ch) { return hy * (hx * q00[ch] + bwx * q01[ch]) + bwy * (hx * q10[ch] + bwx * q11[ch]); };
#pragma unroll
    for (int t = 0; t < O; ++t) {
      upv[2 * t] = up1(2 * (g * O + t));
      upv[2 * t + 1] = up1(2 * (g * O + t) + 1);
      upv[2 * O + t] = up1(2 * n_off + g * O + t);
    }
  }
  auto up = [&](int ch) { return ch >= 2 * n_off ? upv[2 * O + (ch - 2 * n_off - g * O)] : upv[ch - 2 * g * O]; };
  const float flx = flow[pix * fp], fly = flow[pix * fp + 1];

  // fusion group g consumes channels [g*2cg, (g+1)*2cg) of the warped tensor, i.e. warps
  // n = (g*2cg + j) / cg for j in [0, 2cg): n = 2g and 2g + 1 when O == 2.
  float acc[CG];
#pragma unroll
  for (int k = 0; k < CG; ++k) acc[k] = bsrc[g * CG + k];
#pragma unroll
  for (int t = 0; t < O; ++t) {
    const int n = g * O + t;          // warp index in the (C*O)-channel tensor
    const int xg = n % G;             // which x group it warps (x.repeat(offset_num, ...))
    const float ox = mag * tanhf(up(2 * n)) + flx;
    const float oy = mag * tanhf(up(2 * n + 1)) + fly;
    const float mk = 1.f / (1.f + expf(-up(2 * n_off + n)));
    int x0, x1, y0, y1;
    float wx, wy;
    warp_coords(x, y, ox, oy, W, H, x0, x1, y0, y1, wx, wy);
    const float *a = xin + (static_cast<idx_t>(y0) * W + x0) * xp + xg * CG;
    const float *b = xin + (static_cast<idx_t>(y0) * W + x1) * xp + xg * CG;
    const float *d = xin + (static_cast<idx_t>(y1) * W + x0) * xp + xg * CG;
    const float *e = xin + (static_cast<idx_t>(y1) * W + x1) * xp + xg * CG;
    float va[CG], vb[CG], vd[CG], ve[CG];
    if (VEC_X && CG == 3) {
      // every corner pointer has the parity of xg * 3 (even pitch, 8-byte aligned base): odd -> [4 B][8 B], even -> [8 B][4 B]
      const bool odd = (xg & 1) != 0;
      const int o8 = odd ? 1 : 0, o4 = odd ? 0 : 2;
      auto gather3 = [&](const float *q, float (&v)[CG]) {
        const float2 w2 = __ldg(reinterpret_cast<const float2 *>(q + o8));
        const float w1 = __ldg(q + o4);
        v[0] = odd ? w1 : w2.x;
        v[1] = odd ? w2.x : w2.y;
        v[CG - 1] = odd ? w2.y : w1;
      };
      gather3(a, va); gather3(b, vb); gather3(d, vd); gather3(e, ve);
    } else {
#pragma unroll
      for (int j = 0; j < CG; ++j) { va[j] = a[j]; vb[j] = b[j]; vd[j] = d[j]; ve[j] = e[j]; }
    }
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      const float v = bilerp(va[j], vb[j], vd[j], ve[j], wx, wy) * mk;
#pragma unroll
      for (int k = 0; k < CG; ++k) acc[k] = fmaf(wsrc[(g * CG + k) * (O * CG) + t * CG + j], v, acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < CG; ++k) out[pix * outp + g * CG + k] = acc[k];
}

// ---- OffsetDiversity, group-planar variant --------------------------------------------------------------------------
// The kernel above gives the 16 groups of a pixel to 16 neighbouring lanes; every (group, offset) has its OWN sampling
// position, so a warp-level load touches 32 scattered 12-byte pieces of the NHWC feature (one 32-byte sector each: 8.6 % of
// the copy bandwidth).  Offsets vary smoothly in SPACE, not across groups: with the feature regrouped as [G][H][W] float4
// (3 channels + pad) and lanes running along x for a fixed group, the four corner loads of a warp hit two runs of
// ~33 contiguous float4 — fully coalesced when the offsets are smooth.  Same formulae as offset_diversity_kernel.
constexpr int OD_PX = 64;  // pixels per block of the regroup kernel

// x [H*W][xp] (C = 3 * G channels) -> planar [G][H*W] float4 = (c0, c1, c2, 0): coalesced on both sides through shared memory
__global__ void __launch_bounds__(TPB) od_regroup_kernel(const float *__restrict__ xin, int xp, float4 *__restrict__ planar, int G,
                                                         long long pixels) {
  __shared__ float tile[OD_PX][49];
  const int C = 3 * G;
  const long long pix0 = static_cast<long long>(blockIdx.x) * OD_PX;
  for (int i = threadIdx.x; i < OD_PX * C; i += TPB) {
    const int p = i / C, c = i - p * C;
    tile[p][c] = pix0 + p < pixels ? xin[(pix0 + p) * xp + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < OD_PX * G; j += TPB) {
    const int g = j / OD_PX, p = j - g * OD_PX;
    if (pix0 + p < pixels)
      planar[static_cast<long long>(g) * pixels + pix0 + p] = make_float4(tile[p][3 * g], tile[p][3 * g + 1], tile[p][3 * g + 2], 0.f);
  }
}

// block = G warps x 32 lanes: warp g codes group g of 32 consecutive pixels of one row; outputs staged in shared memory and
// written as 32 x C contiguous floats.  O == 2, 3 channels per group (the configuration of the reference).
template <bool VEC_OFF>
__global__ void __launch_bounds__(512) od_planar_kernel(const float4 *__restrict__ planar, const float *__restrict__ off, int op,
                                                        const float *__restrict__ flow, int fp, const float *__restrict__ fw,
                                                        const float *__restrict__ fb, int G, float mag, float *__restrict__ out,
                                                        int outp, int H, int W) {
  constexpr int O = 2, CG = 3;
  __shared__ float stage[32][49];
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (W + 31) / 32;
  const int y = blockIdx.x / tiles_x;
  const int x = (blockIdx.x - y * tiles_x) * 32 + lane;
  const long long pixels = static_cast<long long>(H) * W;
  if (x < W && g < G) {
    const long long pix = static_cast<long long>(y) * W + x;
    const int Hc = H / 2, Wc = W / 2;
    const int n_off = G * O;
    int bx0, bx1, by0, by1;
    float bwx, bwy;
    resize_coord(x, 0.5f, Wc, bx0, bx1, bwx);
    resize_coord(y, 0.5f, Hc, by0, by1, bwy);
    const float hx = 1.f - bwx, hy = 1.f - bwy;
    const float *q00 = off + (static_cast<long long>(by0) * Wc + bx0) * op, *q01 = off + (static_cast<long long>(by0) * Wc + bx1) * op;
    const float *q10 = off + (static_cast<long long>(by1) * Wc + bx0) * op, *q11 = off + (static_cast<long long>(by1) * Wc + bx1) * op;
    float upv[3 * O];
    if (VEC_OFF) {
      const float4 a4 = __ldg(reinterpret_cast<const float4 *>(q00 + 4 * g)), b4 = __ldg(reinterpret_cast<const float4 *>(q01 + 4 * g));
      const float4 d4 = __ldg(reinterpret_cast<const float4 *>(q10 + 4 * g)), e4 = __ldg(reinterpret_cast<const float4 *>(q11 + 4 * g));
      const float2 a2 = __ldg(reinterpret_cast<const float2 *>(q00 + 2 * n_off + 2 * g)), b2 = __ldg(reinterpret_cast<const float2 *>(q01 + 2 * n_off + 2 * g));
      const float2 d2 = __ldg(reinterpret_cast<const float2 *>(q10 + 2 * n_off + 2 * g)), e2 = __ldg(reinterpret_cast<const float2 *>(q11 + 2 * n_off + 2 * g));
      auto up4 = [&](float v00, float v01, float v10, float v11) { return hy * (hx * v00 + bwx * v01) + bwy * (hx * v10 + bwx * v11); };
      upv[0] = up4(a4.x, b4.x, d4.x, e4.x); upv[1] = up4(a4.y, b4.y, d4.y, e4.y);
      upv[2] = up4(a4.z, b4.z, d4.z, e4.z); upv[3] = up4(a4.w, b4.w, d4.w, e4.w);
      upv[4] = up4(a2.x, b2.x, d2.x, e2.x); upv[5] = up4(a2.y, b2.y, d2.y, e2.y);
    } else {
      auto up1 = [&](int ch) { return hy * (hx * q00[ch] + bwx * q01[ch]) + bwy * (hx * q10[ch] + bwx * q11[ch]); };
#pragma unroll
      for (int t = 0; t < O; ++t) {
        upv[2 * t] = up1(2 * (g * O + t));
        upv[2 * t + 1] = up1(2 * (g * O + t) + 1);
        upv[2 * O + t] = up1(2 * n_off + g * O + t);
      }
    }
    const float flx = flow[pix * fp], fly = flow[pix * fp + 1];
    float acc[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k) acc[k] = fb[g * CG + k];
#pragma unroll
    for (int t = 0; t < O; ++t) {
      const int n = g * O + t;
      const int xg = n % G;
      const float ox = mag * tanhf(upv[2 * t]) + flx;
      const float oy = mag * tanhf(upv[2 * t + 1]) + fly;
      const float mk = 1.f / (1.f + expf(-upv[2 * O + t]));
      int x0, x1, y0, y1;
      float wx, wy;
      warp_coords(x, y, ox, oy, W, H, x0, x1, y0, y1, wx, wy);
      const float4 *base = planar + static_cast<long long>(xg) * pixels;
      const float4 a = __ldg(base + static_cast<long long>(y0) * W + x0), b = __ldg(base + static_cast<long long>(y0) * W + x1);
      const float4 d = __ldg(base + static_cast<long long>(y1) * W + x0), e = __ldg(base + static_cast<long long>(y1) * W + x1);
      const float v[CG] = {bilerp(a.x, b.x, d.x, e.x, wx, wy) * mk, bilerp(a.y, b.y, d.y, e.y, wx, wy) * mk,
                           bilerp(a.z, b.z, d.z, e.z, wx, wy) * mk};
#pragma unroll
      for (int j = 0; j < CG; ++j) {
#pragma unroll
        for (int k = 0; k < CG; ++k) acc[k] = fmaf(fw[(g * CG + k) * (O * CG) + t * CG + j], v[j], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < CG; ++k) stage[lane][g * CG + k] = acc[k];
  }
  __syncthreads();
  const int C = G * CG;
  const int x_base = (blockIdx.x - y * tiles_x) * 32;
  for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) {
    const int p = i / C, c = i - p * C;
    if (x_base + p < W) out[(static_cast<long long>(y) * W + x_base + p) * outp + c] = stage[p][c];
  }
}

__global__ void sse_kernel(const float *__restrict__ a, int ap, const float *__restrict__ b, int bp, int C,
                           long long pixels, double *__restrict__ out) {
  double local = 0.0;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < pixels * C;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = idx / C;
    const int c = static_cast<int>(idx - pix * C);
    const float d = a[pix * ap + c] - b[pix * bp + c];
    local += static_cast<double>(d) * d;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double warp_sums[TPB / 32];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < TPB / 32; ++i) s += warp_sums[i];
    atomicAdd(out, s);
  }
}

// 32-bit element indexing is safe: pixels * pitch below 2^31 for both pitches
bool fits32(long long pixels, int pitch_a, int pitch_b) {
  const long long m = pitch_a > pitch_b ? pitch_a : pitch_b;
  return pixels * m < (1LL << 31);
}

bool aligned4(const lssvc_view *v) { return v->C % 4 == 0 && v->pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0; }

// Channel quads per thread (Q) and lanes per pixel (cv / Q) of the per-pixel-lane gather kernels: the widest Q that still
// leaves 4 lanes (64 contiguous bytes per corner) per pixel.
int pick_quads(uint32_t cv, uint32_t *lpp) {
  for (int q = 4; q >= 2; --q) {
    if (cv % q == 0 && cv / q >= 4) {
      *lpp = cv / q;
      return q;
    }
  }
  *lpp = cv;
  return 1;
}

// A/B switch for tools/mem_bench.py: LSSVC_GATHER_LEGACY=1 selects the earlier two-elements-per-thread float4 kernels
bool legacy_gather() {
  static const bool on = getenv("LSSVC_GATHER_LEGACY") != nullptr;
  return on;
}

}  // namespace

extern "C" int32_t lssvc_nchw_to_nhwc(const float *src, int32_t C, const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(src && lssvc::view_ok(out) && C >= 1 && C <= out->C, "nchw_to_nhwc: bad arguments");
  const long long HW = static_cast<long long>(out->H) * out->W;
  if (out->C <= 8) {
    nchw_to_nhwc_small_kernel<<<blocks_for(HW), TPB, 0, lssvc::as_stream(stream)>>>(src, C, out->ptr, out->C, out->pitch, HW);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), lssvc::ceil_div(out->C, 32));
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, lssvc::as_stream(stream)>>>(src, C, out->ptr, out->C, out->pitch, HW);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_nhwc_to_nchw(const lssvc_view *in, float *dst, void *stream) {
  LSSVC_REQUIRE(dst && lssvc::view_ok(in), "nhwc_to_nchw: bad arguments");
  const long long HW = static_cast<long long>(in->H) * in->W;
  if (in->C <= 8) {
    nhwc_to_nchw_small_kernel<<<blocks_for(HW), TPB, 0, lssvc::as_stream(stream)>>>(in->ptr, in->C, in->pitch, dst, HW);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), lssvc::ceil_div(in->C, 32));
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, lssvc::as_stream(stream)>>>(in->ptr, in->C, in->pitch, dst, HW);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_lrelu_copy(const lssvc_view *in, float slope, const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(in) && lssvc::view_ok(out), "lrelu_copy: bad view");
  LSSVC_REQUIRE(in->H == out->H && in->W == out->W && in->C == out->C, "lrelu_copy: shape mismatch");
  const long long pixels = static_cast<long long>(in->H) * in->W;
  if (aligned4(in) && aligned4(out) && fits32(pixels, in->pitch, out->pitch)) {
    const uint32_t cv = in->C / 4, total = static_cast<uint32_t>(pixels) * cv;
    lrelu_copy4_kernel<<<(total + TPB * UNROLL - 1) / (TPB * UNROLL), TPB, 0, lssvc::as_stream(stream)>>>(
        reinterpret_cast<const float4 *>(in->ptr), in->pitch / 4, slope, reinterpret_cast<float4 *>(out->ptr), out->pitch / 4, cv, total);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  lrelu_copy_kernel<<<blocks_for(pixels * in->C), TPB, 0, lssvc::as_stream(stream)>>>(in->ptr, in->pitch, slope, out->ptr,
                                                                                       out->pitch, in->C, pixels);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_softmax2_blend(const lssvc_view *logits, const lssvc_view *a, const lssvc_view *b,
                                        const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(logits) && lssvc::view_ok(a) && lssvc::view_ok(b) && lssvc::view_ok(out),
                "softmax2_blend: bad view");
  LSSVC_REQUIRE(logits->C == 2 && a->C == b->C && a->C == out->C, "softmax2_blend: channel mismatch");
  LSSVC_REQUIRE(a->H == out->H && b->H == out->H && logits->H == out->H && a->W == out->W && b->W == out->W &&
                    logits->W == out->W,
                "softmax2_blend: size mismatch");
  const long long pixels = static_cast<long long>(out->H) * out->W;
  if (aligned4(a) && aligned4(b) && aligned4(out) && fits32(pixels, a->pitch, out->pitch) && fits32(pixels, b->pitch, logits->pitch)) {
    const uint32_t cv = out->C / 4, total = static_cast<uint32_t>(pixels) * cv;
    softmax2_blend4_kernel<<<(total + TPB * UNROLL - 1) / (TPB * UNROLL), TPB, 0, lssvc::as_stream(stream)>>>(
        logits->ptr, logits->pitch, reinterpret_cast<const float4 *>(a->ptr), a->pitch / 4, reinterpret_cast<const float4 *>(b->ptr),
        b->pitch / 4, reinterpret_cast<float4 *>(out->ptr), out->pitch / 4, cv, total);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  softmax2_blend_kernel<<<blocks_for(pixels * out->C), TPB, 0, lssvc::as_stream(stream)>>>(
      logits->ptr, logits->pitch, a->ptr, a->pitch, b->ptr, b->pitch, out->ptr, out->pitch, out->C, pixels);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_flow_warp(const lssvc_view *src, const lssvc_view *flow, float flow_scale,
                                   const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(src) && lssvc::view_ok(flow) && lssvc::view_ok(out), "flow_warp: bad view");
  LSSVC_REQUIRE(flow->C >= 2 && src->H == out->H && src->W == out->W && flow->H == out->H && flow->W == out->W &&
                    src->C == out->C,
                "flow_warp: shape mismatch");
  const long long pixels = static_cast<long long>(out->H) * out->W;
  cudaStream_t s = lssvc::as_stream(stream);
  if (aligned4(src) && aligned4(out) && fits32(pixels, src->pitch, out->pitch)) {
    const uint32_t cv = src->C / 4, total = static_cast<uint32_t>(pixels) * cv;
    uint32_t lpp = cv;
    const int q = pick_quads(cv, &lpp);
    const uint32_t n_threads = static_cast<uint32_t>(pixels) * lpp, grid = (n_threads + TPB - 1) / TPB;
#define LSSVC_WARP_Q(QQ)                                                                                                   \
  flow_warp_q_kernel<QQ><<<grid, TPB, 0, s>>>(src->ptr, src->pitch, flow->ptr, flow->pitch, flow_scale, out->ptr, out->pitch, \
                                              out->H, out->W, lpp, n_threads)
    if (legacy_gather()) {
      flow_warp4_kernel<<<(total + TPB * 2 - 1) / (TPB * 2), TPB, 0, s>>>(src->ptr, src->pitch, flow->ptr, flow->pitch, flow_scale,
                                                                         out->ptr, out->pitch, out->H, out->W, cv, total);
    } else if (q == 4) {
      LSSVC_WARP_Q(4);
    } else if (q == 3) {
      LSSVC_WARP_Q(3);
    } else if (q == 2) {
      LSSVC_WARP_Q(2);
    } else {
      LSSVC_WARP_Q(1);
    }
#undef LSSVC_WARP_Q
  } else if (aligned4(src) && aligned4(out)) {
    flow_warp_kernel<4><<<blocks_for(pixels * (src->C / 4)), TPB, 0, s>>>(src->ptr, src->pitch, flow->ptr, flow->pitch,
                                                                         flow_scale, out->ptr, out->pitch, out->H, out->W,
                                                                         src->C);
  } else {
    flow_warp_kernel<1><<<blocks_for(pixels * src->C), TPB, 0, s>>>(src->ptr, src->pitch, flow->ptr, flow->pitch,
                                                                   flow_scale, out->ptr, out->pitch, out->H, out->W, src->C);
  }
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_bilinear_resize(const lssvc_view *in, float scale, const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(in) && lssvc::view_ok(out) && in->C == out->C, "bilinear_resize: bad view");
  const float rh = static_cast<float>(in->H) / static_cast<float>(out->H);
  const float rw = static_cast<float>(in->W) / static_cast<float>(out->W);
  const long long pixels = static_cast<long long>(out->H) * out->W;
  cudaStream_t s = lssvc::as_stream(stream);
  if (aligned4(in) && aligned4(out) && fits32(pixels, in->pitch, out->pitch)) {
    const uint32_t cv = in->C / 4, total = static_cast<uint32_t>(pixels) * cv;
    uint32_t lpp = cv;
    const int q = pick_quads(cv, &lpp);
    const uint32_t n_threads = static_cast<uint32_t>(pixels) * lpp, grid = (n_threads + TPB - 1) / TPB;
#define LSSVC_RESIZE_Q(QQ)                                                                                               \
  bilinear_resize_q_kernel<QQ><<<grid, TPB, 0, s>>>(in->ptr, in->pitch, in->H, in->W, out->ptr, out->pitch, out->W, lpp, n_threads, \
                                                    rh, rw, scale)
    if (!legacy_gather() && out->H == 2 * in->H && out->W == 2 * in->W && in->H > 1 && in->W > 1 &&
        static_cast<long long>(in->H + 1) * (in->W + 1) * lpp < (1LL << 32)) {
      const uint32_t cells = static_cast<uint32_t>(in->H + 1) * static_cast<uint32_t>(in->W + 1) * lpp, cgrid = (cells + TPB - 1) / TPB;
#define LSSVC_UP2_Q(QQ) \
  bilinear_up2_q_kernel<QQ><<<cgrid, TPB, 0, s>>>(in->ptr, in->pitch, in->H, in->W, out->ptr, out->pitch, lpp, cells, scale)
      if (q == 4) LSSVC_UP2_Q(4);
      else if (q == 3) LSSVC_UP2_Q(3);
      else if (q == 2) LSSVC_UP2_Q(2);
      else LSSVC_UP2_Q(1);
#undef LSSVC_UP2_Q
    } else if (legacy_gather()) {
      bilinear_resize4_kernel<<<(total + TPB * 2 - 1) / (TPB * 2), TPB, 0, s>>>(in->ptr, in->pitch, in->H, in->W, out->ptr, out->pitch,
                                                                               out->W, cv, total, rh, rw, scale);
    } else if (q == 4) {
      LSSVC_RESIZE_Q(4);
    } else if (q == 3) {
      LSSVC_RESIZE_Q(3);
    } else if (q == 2) {
      LSSVC_RESIZE_Q(2);
    } else {
      LSSVC_RESIZE_Q(1);
    }
#undef LSSVC_RESIZE_Q
  } else if (aligned4(in) && aligned4(out)) {
    bilinear_resize_kernel<4><<<blocks_for(pixels * (in->C / 4)), TPB, 0, s>>>(in->ptr, in->pitch, in->H, in->W, out->ptr,
                                                                              out->pitch, out->H, out->W, in->C, rh, rw, scale);
  } else {
    bilinear_resize_kernel<1><<<blocks_for(pixels * in->C), TPB, 0, s>>>(in->ptr, in->pitch, in->H, in->W, out->ptr,
                                                                        out->pitch, out->H, out->W, in->C, rh, rw, scale);
  }
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

static int32_t pool2(const lssvc_view *in, const lssvc_view *out, void *stream, bool is_max) {
  LSSVC_REQUIRE(lssvc::view_ok(in) && lssvc::view_ok(out) && in->C == out->C, "pool2: bad view");
  LSSVC_REQUIRE(out->H == in->H / 2 && out->W == in->W / 2, "pool2: output must be half the input");
  const long long total = static_cast<long long>(out->H) * out->W * out->C;
  cudaStream_t s = lssvc::as_stream(stream);
  if (aligned4(in) && aligned4(out) && fits32(static_cast<long long>(in->H) * in->W, in->pitch, out->pitch)) {
    const uint32_t cv = out->C / 4, tot = static_cast<uint32_t>(out->H) * out->W * cv;
    const int blocks = static_cast<int>((tot + TPB * 2 - 1) / (TPB * 2));
    if (is_max) pool2_4_kernel<true><<<blocks, TPB, 0, s>>>(in->ptr, in->pitch, in->W, out->ptr, out->pitch, out->W, cv, tot);
    else pool2_4_kernel<false><<<blocks, TPB, 0, s>>>(in->ptr, in->pitch, in->W, out->ptr, out->pitch, out->W, cv, tot);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  if (is_max) pool2_kernel<true><<<blocks_for(total), TPB, 0, s>>>(in->ptr, in->pitch, in->W, out->ptr, out->pitch, out->H, out->W, out->C);
  else pool2_kernel<false><<<blocks_for(total), TPB, 0, s>>>(in->ptr, in->pitch, in->W, out->ptr, out->pitch, out->H, out->W, out->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
extern "C" int32_t lssvc_avgpool2(const lssvc_view *in, const lssvc_view *out, void *stream) { return pool2(in, out, stream, false); }
extern "C" int32_t lssvc_maxpool2(const lssvc_view *in, const lssvc_view *out, void *stream) { return pool2(in, out, stream, true); }

extern "C" int32_t lssvc_spynet_prep(const lssvc_view *im1, const lssvc_view *im2, const lssvc_view *flow_coarse,
                                     const lssvc_view *out8, const lssvc_view *flow_up, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(im1) && lssvc::view_ok(im2) && lssvc::view_ok(out8) && lssvc::view_ok(flow_up),
                "spynet_prep: bad view");
  LSSVC_REQUIRE(im1->C >= 3 && im2->C >= 3 && out8->C == 8 && flow_up->C >= 2, "spynet_prep: channel counts");
  LSSVC_REQUIRE(out8->pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(out8->ptr) & 15) == 0, "spynet_prep: out8 alignment");
  const int H = im1->H, W = im1->W;
  LSSVC_REQUIRE(im2->H == H && im2->W == W && out8->H == H && out8->W == W && flow_up->H == H && flow_up->W == W,
                "spynet_prep: size mismatch");
  const float *fc = nullptr;
  int pc = 0;
  if (lssvc::view_present(flow_coarse)) {
    LSSVC_REQUIRE(flow_coarse->H * 2 == H && flow_coarse->W * 2 == W && flow_coarse->C >= 2, "spynet_prep: coarse flow size");
    fc = flow_coarse->ptr;
    pc = flow_coarse->pitch;
  }
  spynet_prep_kernel<<<blocks_for(static_cast<long long>(H) * W), TPB, 0, lssvc::as_stream(stream)>>>(
      im1->ptr, im1->pitch, im2->ptr, im2->pitch, fc, pc, out8->ptr, out8->pitch, flow_up->ptr, flow_up->pitch, H, W);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_offset_diversity(const lssvc_view *x, const lssvc_view *off, const lssvc_view *flow,
                                          const float *fusion_w, const float *fusion_b, int32_t groups,
                                          int32_t offset_num, float magnitude, const lssvc_view *out, float *scratch,
                                          void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(x) && lssvc::view_ok(off) && lssvc::view_ok(flow) && lssvc::view_ok(out),
                "offset_diversity: bad view");
  LSSVC_REQUIRE(groups > 0 && x->C % groups == 0 && x->C / groups == 3 && offset_num == 2,
                "offset_diversity: only 3 channels per group and 2 offsets are built (C=%d G=%d O=%d)", x->C, groups,
                offset_num);
  LSSVC_REQUIRE(off->C == 3 * groups * offset_num && off->H * 2 == x->H && off->W * 2 == x->W,
                "offset_diversity: offset tensor shape");
  LSSVC_REQUIRE(flow->C >= 2 && flow->H == x->H && flow->W == x->W && out->H == x->H && out->W == x->W && out->C == x->C,
                "offset_diversity: size mismatch");
  const long long total = static_cast<long long>(x->H) * x->W * groups;
  const bool vec_off = off->pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(off->ptr) & 15) == 0;
  static const bool legacy = getenv("LSSVC_GATHER_LEGACY") != nullptr;
  if (scratch != nullptr && !legacy && groups <= 16) {
    // group-planar path: regroup the feature once ([G][H][W] float4 in `scratch`), then gather with lanes along x
    LSSVC_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "offset_diversity: scratch must be 16-byte aligned");
    const long long pixels = static_cast<long long>(x->H) * x->W;
    float4 *planar = reinterpret_cast<float4 *>(scratch);
    od_regroup_kernel<<<static_cast<int>((pixels + OD_PX - 1) / OD_PX), TPB, 0, lssvc::as_stream(stream)>>>(x->ptr, x->pitch, planar,
                                                                                                         groups, pixels);
    LSSVC_LAUNCHED();
    const int blocks = x->H * ((x->W + 31) / 32);
    if (vec_off)
      od_planar_kernel<true><<<blocks, 32 * groups, 0, lssvc::as_stream(stream)>>>(planar, off->ptr, off->pitch, flow->ptr, flow->pitch,
                                                                                  fusion_w, fusion_b, groups, magnitude, out->ptr,
                                                                                  out->pitch, x->H, x->W);
    else
      od_planar_kernel<false><<<blocks, 32 * groups, 0, lssvc::as_stream(stream)>>>(planar, off->ptr, off->pitch, flow->ptr, flow->pitch,
                                                                                   fusion_w, fusion_b, groups, magnitude, out->ptr,
                                                                                   out->pitch, x->H, x->W);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  // 8-byte gathers + shared-memory fusion weights (see the kernel); LSSVC_GATHER_LEGACY keeps the scalar version for A/B
  const bool vec_x = !legacy && groups <= 16 && x->pitch % 2 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 7) == 0;
  const long long px = static_cast<long long>(x->H) * x->W;
  const int max_pitch = std::max(std::max(x->pitch, out->pitch), std::max(off->pitch, flow->pitch));
  const bool i32 = !legacy && px * max_pitch < (1ll << 31) && total < (1ll << 31);
#define OD_LAUNCH(VO, VX, I3)                                                                                                \
  offset_diversity_kernel<3, 2, VO, VX, I3><<<blocks_for(total), TPB, 0, lssvc::as_stream(stream)>>>(                        \
      x->ptr, x->pitch, off->ptr, off->pitch, flow->ptr, flow->pitch, fusion_w, fusion_b, groups, magnitude, out->ptr,       \
      out->pitch, x->H, x->W)
  if (vec_off && vec_x && i32) OD_LAUNCH(true, true, true);
  else if (vec_off && vec_x) OD_LAUNCH(true, true, false);
  else if (vec_off) OD_LAUNCH(true, false, false);
  else if (vec_x) OD_LAUNCH(false, true, false);
  else OD_LAUNCH(false, false, false);
#undef OD_LAUNCH
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_sse(const lssvc_view *a, const lssvc_view *b, double *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(a) && lssvc::view_ok(b) && out, "sse: bad arguments");
  LSSVC_REQUIRE(a->H == b->H && a->W == b->W && a->C == b->C, "sse: shape mismatch");
  const long long pixels = static_cast<long long>(a->H) * a->W;
  int blocks = blocks_for(pixels * a->C);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sse_kernel<<<blocks, TPB, 0, lssvc::as_stream(stream)>>>(a->ptr, a->pitch, b->ptr, b->pitch, a->C, pixels, out);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
