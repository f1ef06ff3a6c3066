// fp32 CUDA-core convolutions: the generic implicit GEMM (any kernel size / stride / channel count,
// input transforms, GDN epilogue), depthwise 3x3 and the stride-2 transposed 3x3.
// These cover the layers the tensor-core kernel does not take (odd channel counts, GDN's
// squared-input norm pool which wants full fp32, tiny heads) and serve as the on-device
// cross-check of the tcgen05 path.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 64;  // output pixels per block (8 x 8 patch)
constexpr int BN = 64;  // output channels per block
constexpr int BK = 16;  // input channels per step
constexpr int PATCH = 8;

struct SimtParams {
  int n_src;
  const float *src[LSSVC_MAX_SRC];
  int src_c[LSSVC_MAX_SRC];
  int src_pitch[LSSVC_MAX_SRC];
  int Hin, Win, Ho, Wo;
  const float *weight;
  const float *bias;
  int kh, kw, stride, pad;
  int cout, n_pad, cin_total;
  int in_transform;
  float in_slope;
  int epi, act;
  float slope, out_scale;
  int pixel_shuffle;
  float *out;
  int out_pitch;
  const float *res1;
  int res1_pitch;
  const float *res2;
  int res2_pitch;
  float *out2;
  int out2_pitch;
  float slope2;
  const float *gdn_x;
  int gdn_pitch;
  int tiles_x;
};

__device__ __forceinline__ long long simt_out_offset(const SimtParams &p, int oy, int ox, int ch, int P) {
  if (!p.pixel_shuffle) return (static_cast<long long>(oy) * p.Wo + ox) * P + ch;
  const int cq = p.cout >> 2;
  const int sub = ch / cq;
  const int c = ch - sub * cq;
  return (static_cast<long long>(2 * oy + (sub >> 1)) * (2 * p.Wo) + (2 * ox + (sub & 1))) * P + c;
}

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int oy0 = ty * PATCH, ox0 = tx * PATCH;
  const int n0 = blockIdx.y * BN;

  // loader roles
  const int lm = tid >> 2;         // 0..63: pixel (A) or output channel (B)
  const int lk = (tid & 3) * 4;    // 0,4,8,12: channel sub-offset
  const int l_oy = oy0 + (lm >> 3), l_ox = ox0 + (lm & 7);
  // compute roles
  const int tm = (tid >> 4) * 4;   // pixel group
  const int tn = (tid & 15) * 4;   // channel group

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int coff = 0;
  for (int j = 0; j < p.n_src; ++j) {
    const float *src = p.src[j];
    const int C = p.src_c[j];
    const int pitch = p.src_pitch[j];
    for (int r = 0; r < p.kh; ++r) {
      const int iy = l_oy * p.stride + r - p.pad;
      for (int s = 0; s < p.kw; ++s) {
        const int ix = l_ox * p.stride + s - p.pad;
        const bool in_img = (iy >= 0) && (iy < p.Hin) && (ix >= 0) && (ix < p.Win) && (l_oy < p.Ho) && (l_ox < p.Wo);
        const float *a_ptr = src + (static_cast<long long>(iy) * p.Win + ix) * pitch;
        const int tap = r * p.kw + s;
        const float *b_ptr = p.weight + (static_cast<long long>(tap) * p.n_pad + (n0 + lm)) * p.cin_total + coff;
        const bool b_row_ok = (n0 + lm) < p.n_pad;
        for (int c0 = 0; c0 < C; c0 += BK) {
          float a[4], b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = c0 + lk + e;
            float v = (in_img && ch < C) ? a_ptr[ch] : 0.f;
            if (p.in_transform == LSSVC_IN_SQUARE) v = v * v;
            else if (p.in_transform == LSSVC_IN_LRELU) v = v > 0.f ? v : v * p.in_slope;
            a[e] = v;
            b[e] = (b_row_ok && ch < C) ? b_ptr[ch] : 0.f;
          }
          __syncthreads();
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            As[lk + e][lm] = a[e];
            Bs[lk + e][lm] = b[e];
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k][tm]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tn]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(aa[i], bb[jj], acc[i][jj]);
          }
        }
      }
    }
    coff += C;
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tm + i;
    const int oy = oy0 + (m >> 3), ox = ox0 + (m & 7);
    if (oy >= p.Ho || ox >= p.Wo) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int ch = n0 + tn + jj;
      if (ch >= p.cout) continue;
      float v = acc[i][jj] + p.bias[ch];
      if (p.epi == LSSVC_EPI_GDN) {
        v = p.gdn_x[(static_cast<long long>(oy) * p.Wo + ox) * p.gdn_pitch + ch] * rsqrtf(v);
      } else if (p.epi == LSSVC_EPI_IGDN) {
        v = p.gdn_x[(static_cast<long long>(oy) * p.Wo + ox) * p.gdn_pitch + ch] * sqrtf(v);
      }
      if (p.act) v = v > 0.f ? v : v * p.slope;
      v *= p.out_scale;
      if (p.res1) v += p.res1[simt_out_offset(p, oy, ox, ch, p.res1_pitch)];
      if (p.res2) v += p.res2[simt_out_offset(p, oy, ox, ch, p.res2_pitch)];
      p.out[simt_out_offset(p, oy, ox, ch, p.out_pitch)] = v;
      if (p.out2) p.out2[simt_out_offset(p, oy, ox, ch, p.out2_pitch)] = v > 0.f ? v : v * p.slope2;
    }
  }
}

// depthwise 3x3, pad 1: one thread per (pixel, 4 channels)
__global__ void dwconv3x3_kernel(const float *__restrict__ in, int in_pitch, const float *__restrict__ w,
                                 const float *__restrict__ bias, float *__restrict__ out, int out_pitch, int H, int W,
                                 int C) {
  const int c4 = C >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(H) * W * c4;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % c4) * 4;
  const long long pix = idx / c4;
  const int x = static_cast<int>(pix % W), y = static_cast<int>(pix / W);
  float4 acc = *reinterpret_cast<const float4 *>(bias + c);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = y + r - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ix = x + s - 1;
      if (ix < 0 || ix >= W) continue;
      const float4 v = *reinterpret_cast<const float4 *>(in + (static_cast<long long>(iy) * W + ix) * in_pitch + c);
      const float4 k = *reinterpret_cast<const float4 *>(w + (r * 3 + s) * C + c);
      acc.x = fmaf(v.x, k.x, acc.x);
      acc.y = fmaf(v.y, k.y, acc.y);
      acc.z = fmaf(v.z, k.z, acc.z);
      acc.w = fmaf(v.w, k.w, acc.w);
    }
  }
  *reinterpret_cast<float4 *>(out + pix * out_pitch + c) = acc;
}

// same, two horizontally adjacent output pixels per thread: the 3 x 4 input window is loaded once (12 float4 loads, all
// in flight, for 2 outputs instead of 18) — the kernel is HBM/L2-bound, so bytes in flight per thread are what count
__global__ void __launch_bounds__(256) dwconv3x3_x2_kernel(const float *__restrict__ in, uint32_t in_pitch,
                                                           const float *__restrict__ w, const float *__restrict__ bias,
                                                           float *__restrict__ out, uint32_t out_pitch, int H, int W, int C,
                                                           uint32_t total) {
  const uint32_t c4 = static_cast<uint32_t>(C) >> 2, W2 = static_cast<uint32_t>(W + 1) >> 1;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const uint32_t pp = idx / c4, c = (idx - pp * c4) * 4;
  const int y = static_cast<int>(pp / W2), x0 = static_cast<int>(pp - static_cast<uint32_t>(y) * W2) * 2;
  float4 v[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = y + r - 1;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int ix = x0 + s - 1;
      v[r][s] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                    ? __ldg(reinterpret_cast<const float4 *>(in + (static_cast<size_t>(iy) * W + ix) * in_pitch + c))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
  float4 a0 = b, a1 = b;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const float4 k = __ldg(reinterpret_cast<const float4 *>(w + (r * 3 + s) * C + c));
      // the taps of out-of-image pixels are skipped in the one-pixel kernel; here they multiply an exact zero (same sum:
      // fmaf(0, k, acc) == acc), so both kernels produce identical results
      a0.x = fmaf(v[r][s].x, k.x, a0.x); a0.y = fmaf(v[r][s].y, k.y, a0.y);
      a0.z = fmaf(v[r][s].z, k.z, a0.z); a0.w = fmaf(v[r][s].w, k.w, a0.w);
      a1.x = fmaf(v[r][s + 1].x, k.x, a1.x); a1.y = fmaf(v[r][s + 1].y, k.y, a1.y);
      a1.z = fmaf(v[r][s + 1].z, k.z, a1.z); a1.w = fmaf(v[r][s + 1].w, k.w, a1.w);
    }
  }
  float *o = out + (static_cast<size_t>(y) * W + x0) * out_pitch + c;
  *reinterpret_cast<float4 *>(o) = a0;
  if (x0 + 1 < W) *reinterpret_cast<float4 *>(o + out_pitch) = a1;
}

// same arithmetic, a strip of R output rows x 2 columns per thread: the 3 x 4 input window slides down the strip (4 new float4
// loads per row instead of 12), the nine weight quads stay in registers.  2 (R + 2) / R loads per output instead of 6; taps
// are accumulated in the order of the kernels above (bias, then r = 0..2, s = 0..2), so results are identical.
template <int R>
__global__ void __launch_bounds__(256) dwconv3x3_strip_kernel(const float *__restrict__ in, uint32_t in_pitch,
                                                              const float *__restrict__ w, const float *__restrict__ bias,
                                                              float *__restrict__ out, uint32_t out_pitch, int H, int W, int C,
                                                              uint32_t total) {
  const uint32_t c4 = static_cast<uint32_t>(C) >> 2, W2 = static_cast<uint32_t>(W + 1) >> 1;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const uint32_t pp = idx / c4, c = (idx - pp * c4) * 4;
  const int ys = static_cast<int>(pp / W2) * R, x0 = static_cast<int>(pp - (pp / W2) * W2) * 2;
  float4 k[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = __ldg(reinterpret_cast<const float4 *>(w + t * C + c));
  const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
  auto load_row = [&](int iy, float4 (&row)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int ix = x0 + s - 1;
      row[s] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                   ? __ldg(reinterpret_cast<const float4 *>(in + (static_cast<size_t>(iy) * W + ix) * in_pitch + c))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 v[3][4];
  load_row(ys - 1, v[0]);
  load_row(ys, v[1]);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int y = ys + j;
    if (y >= H) break;
    load_row(y + 1, v[(j + 2) % 3]);
    float4 a0 = b, a1 = b;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float4(&row)[4] = v[(j + r) % 3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4 kk = k[r * 3 + s];
        a0.x = fmaf(row[s].x, kk.x, a0.x); a0.y = fmaf(row[s].y, kk.y, a0.y);
        a0.z = fmaf(row[s].z, kk.z, a0.z); a0.w = fmaf(row[s].w, kk.w, a0.w);
        a1.x = fmaf(row[s + 1].x, kk.x, a1.x); a1.y = fmaf(row[s + 1].y, kk.y, a1.y);
        a1.z = fmaf(row[s + 1].z, kk.z, a1.z); a1.w = fmaf(row[s + 1].w, kk.w, a1.w);
      }
    }
    float *o = out + (static_cast<size_t>(y) * W + x0) * out_pitch + c;
    *reinterpret_cast<float4 *>(o) = a0;
    if (x0 + 1 < W) *reinterpret_cast<float4 *>(o + out_pitch) = a1;
  }
}

// ConvTranspose2d(k=3, s=2, p=1, output_padding=1): out[oy, ox] gathers in[(oy + 1 - ky) / 2, ...] where
// (oy + 1 - ky) is even.  weight [9][cin][cout].  One block = 32 output pixels x 32 output channels.
__global__ void __launch_bounds__(256) deconv3x3_s2_kernel(const float *__restrict__ in, int in_pitch, int Hi, int Wi,
                                                         int Cin, const float *__restrict__ w,
                                                         const float *__restrict__ bias, int Cout, int act, float slope,
                                                         float *__restrict__ out, int out_pitch) {
  const int Ho = Hi * 2, Wo = Wi * 2;
  const int co = blockIdx.y * 32 + (threadIdx.x & 31);
  const long long pix = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (pix >= static_cast<long long>(Ho) * Wo) return;
  const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>(pix / Wo);
  if (co >= Cout) return;
  float acc = bias[co];
  for (int ky = 0; ky < 3; ++ky) {
    const int ny = oy + 1 - ky;
    if (ny < 0 || (ny & 1)) continue;
    const int iy = ny >> 1;
    if (iy >= Hi) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int nx = ox + 1 - kx;
      if (nx < 0 || (nx & 1)) continue;
      const int ix = nx >> 1;
      if (ix >= Wi) continue;
      const float *ip = in + (static_cast<long long>(iy) * Wi + ix) * in_pitch;
      const float *wp = w + static_cast<long long>(ky * 3 + kx) * Cin * Cout + co;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ip[ci], wp[static_cast<long long>(ci) * Cout], acc);
    }
  }
  if (act) acc = acc > 0.f ? acc : acc * slope;
  out[pix * out_pitch + co] = acc;
}

}  // namespace

extern "C" int32_t lssvc_conv_simt(const lssvc_conv *c, void *stream) {
  LSSVC_REQUIRE(c != nullptr, "conv_simt: null descriptor");
  LSSVC_REQUIRE(c->n_src >= 1 && c->n_src <= LSSVC_MAX_SRC, "conv_simt: n_src=%d", c->n_src);
  SimtParams p;
  memset(&p, 0, sizeof(p));
  p.n_src = c->n_src;
  p.Hin = c->src[0].H;
  p.Win = c->src[0].W;
  int cin_total = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    LSSVC_REQUIRE(lssvc::view_ok(&v), "conv_simt: bad source view %d", j);
    LSSVC_REQUIRE(v.H == p.Hin && v.W == p.Win, "conv_simt: source %d size mismatch", j);
    p.src[j] = v.ptr;
    p.src_c[j] = v.C;
    p.src_pitch[j] = v.pitch;
    cin_total += v.C;
  }
  LSSVC_REQUIRE(cin_total == c->cin_total, "conv_simt: cin_total %d != sum of sources %d", c->cin_total, cin_total);
  p.Ho = (p.Hin + 2 * c->pad - c->kh) / c->stride + 1;
  p.Wo = (p.Win + 2 * c->pad - c->kw) / c->stride + 1;
  const int ps = c->pixel_shuffle ? 2 : 1;
  LSSVC_REQUIRE(lssvc::view_ok(&c->out), "conv_simt: bad output view");
  LSSVC_REQUIRE(c->out.H == p.Ho * ps && c->out.W == p.Wo * ps, "conv_simt: output view %dx%d, expected %dx%d",
                c->out.H, c->out.W, p.Ho * ps, p.Wo * ps);
  LSSVC_REQUIRE(!c->pixel_shuffle || c->cout % 4 == 0, "conv_simt: pixel shuffle needs cout %% 4 == 0");
  LSSVC_REQUIRE(c->out.C == (c->pixel_shuffle ? c->cout / 4 : c->cout), "conv_simt: output channel mismatch (%d vs %d)",
                c->out.C, c->cout);
  p.weight = c->weight; p.bias = c->bias;
  p.kh = c->kh; p.kw = c->kw; p.stride = c->stride; p.pad = c->pad;
  p.cout = c->cout; p.n_pad = c->n_pad; p.cin_total = c->cin_total;
  p.in_transform = c->in_transform; p.in_slope = c->in_slope;
  p.epi = c->epi; p.act = c->act; p.slope = c->slope; p.out_scale = c->out_scale;
  p.pixel_shuffle = c->pixel_shuffle;
  p.out = c->out.ptr; p.out_pitch = c->out.pitch;
  auto opt = [&](const lssvc_view &v, const float **ptr, int *pitch) -> bool {
    if (!v.ptr) { *ptr = nullptr; *pitch = 0; return true; }
    if (v.H != c->out.H || v.W != c->out.W || v.C != c->out.C) return false;
    *ptr = v.ptr; *pitch = v.pitch;
    return true;
  };
  LSSVC_REQUIRE(opt(c->res1, &p.res1, &p.res1_pitch), "conv_simt: res1 shape mismatch");
  LSSVC_REQUIRE(opt(c->res2, &p.res2, &p.res2_pitch), "conv_simt: res2 shape mismatch");
  const float *o2 = nullptr;
  LSSVC_REQUIRE(opt(c->out2, &o2, &p.out2_pitch), "conv_simt: out2 shape mismatch");
  p.out2 = const_cast<float *>(o2);
  p.slope2 = c->slope2;
  if (c->epi != LSSVC_EPI_PLAIN) {
    LSSVC_REQUIRE(!c->pixel_shuffle && lssvc::view_ok(&c->gdn_x) && c->gdn_x.H == p.Ho && c->gdn_x.W == p.Wo &&
                      c->gdn_x.C == c->cout,
                  "conv_simt: GDN epilogue needs a matching gdn_x view");
    p.gdn_x = c->gdn_x.ptr;
    p.gdn_pitch = c->gdn_x.pitch;
  }
  p.tiles_x = lssvc::ceil_div(p.Wo, PATCH);
  const int tiles = p.tiles_x * lssvc::ceil_div(p.Ho, PATCH);
  dim3 grid(tiles, lssvc::ceil_div(c->cout, BN));
  conv_simt_kernel<<<grid, 256, 0, lssvc::as_stream(stream)>>>(p);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_dwconv3x3(const lssvc_view *in, const float *weight, const float *bias, const lssvc_view *out,
                                   void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(in) && lssvc::view_ok(out), "dwconv3x3: bad view");
  LSSVC_REQUIRE(in->H == out->H && in->W == out->W && in->C == out->C, "dwconv3x3: shape mismatch");
  LSSVC_REQUIRE(in->C % 4 == 0 && in->pitch % 4 == 0 && out->pitch % 4 == 0, "dwconv3x3: channels must be 4-aligned");
  const long long total = static_cast<long long>(in->H) * in->W * (in->C / 4);
  const int pitch_max = in->pitch > out->pitch ? in->pitch : out->pitch;
  if (static_cast<long long>(in->H) * in->W * pitch_max < (1LL << 31) && (reinterpret_cast<uintptr_t>(in->ptr) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out->ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
    const uint32_t tot2 = static_cast<uint32_t>(in->H) * static_cast<uint32_t>((in->W + 1) / 2) * static_cast<uint32_t>(in->C / 4);
    // 4-row strips per thread (sliding window): 0.043 -> 0.036 ms at C = 128, 288x480 (50 -> 60 % of the copy bandwidth),
    // bit-identical results; LSSVC_DW_STRIPS=0 selects the two-pixel kernel (A/B)
    static const bool strips = [] {
      const char *e = getenv("LSSVC_DW_STRIPS");
      return e == nullptr || atoi(e) != 0;
    }();
    if (strips && in->H >= 16) {
      constexpr int R = 4;
      const uint32_t tot = static_cast<uint32_t>((in->H + R - 1) / R) * static_cast<uint32_t>((in->W + 1) / 2) * static_cast<uint32_t>(in->C / 4);
      dwconv3x3_strip_kernel<R><<<(tot + 255) / 256, 256, 0, lssvc::as_stream(stream)>>>(in->ptr, in->pitch, weight, bias, out->ptr,
                                                                                        out->pitch, in->H, in->W, in->C, tot);
      LSSVC_LAUNCHED();
      return LSSVC_OK;
    }
    dwconv3x3_x2_kernel<<<(tot2 + 255) / 256, 256, 0, lssvc::as_stream(stream)>>>(in->ptr, in->pitch, weight, bias, out->ptr,
                                                                                 out->pitch, in->H, in->W, in->C, tot2);
    LSSVC_LAUNCHED();
    return LSSVC_OK;
  }
  const int blocks = static_cast<int>((total + 255) / 256);
  dwconv3x3_kernel<<<blocks, 256, 0, lssvc::as_stream(stream)>>>(in->ptr, in->pitch, weight, bias, out->ptr, out->pitch,
                                                                in->H, in->W, in->C);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

extern "C" int32_t lssvc_deconv3x3_s2(const lssvc_view *in, const float *weight, const float *bias, int32_t act,
                                      float slope, const lssvc_view *out, void *stream) {
  LSSVC_REQUIRE(lssvc::view_ok(in) && lssvc::view_ok(out), "deconv3x3_s2: bad view");
  LSSVC_REQUIRE(out->H == 2 * in->H && out->W == 2 * in->W, "deconv3x3_s2: output must be 2x the input");
  const long long pixels = static_cast<long long>(out->H) * out->W;
  dim3 grid(static_cast<unsigned>((pixels + 7) / 8), lssvc::ceil_div(out->C, 32));
  deconv3x3_s2_kernel<<<grid, 256, 0, lssvc::as_stream(stream)>>>(in->ptr, in->pitch, in->H, in->W, in->C, weight, bias,
                                                                   out->C, act, slope, out->ptr, out->pitch);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
