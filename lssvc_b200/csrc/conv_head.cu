// Narrow "head" convolutions on the fp32 CUDA cores: k x k, stride 1, same padding; 3x3 with Cout 2..4, 7x7 with Cout 2.
//
// The layers that end a branch in 2 or 3 channels — flow heads (SpyNet's 7x7 16->2 at five scales, the 3x3 64->2 of the
// motion decoders / MvResampler: video_net_component.py:213-230, lssvc_modules.py:339-365) and the reconstruction heads
// (3x3 48->3, 64->3: lssvc_modules.py:279-292, dmc_net.py) — gave the tensor-core kernel an N = 32 + 16 MMA pair per
// 128 pixels and K = 16 slice for 2 useful columns: 14-18 TFLOP/s, 0.33-0.36 ms per full-resolution launch, ~2.3 ms per
// P-frame.  Their arithmetic is tiny (1152-1568 FMA per pixel): as a register-blocked direct convolution they are bound by
// the fp32 FMA rate / the read of their input instead.
//
// One CTA = 4 warps = a 64 x 16 pixel tile.  Input channels are walked one quad (4 channels) at a time; a quad's halo tile
// ((16 + k - 1) x (64 + k - 1) pixels) is copied with cp.async (zero fill outside the image = the conv's padding) into one
// of TWO shared-memory buffers laid out [row][column parity][column / 2] float4, so that the lanes of a warp, which own the
// pixel pairs (2 lane, 2 lane + 1), read consecutive float4 for every window column: conflict-free LDS.128; the copy of
// quad q + 1 is in flight while quad q is computed (thread x < 64 + k - 1 owns halo column x and walks its rows with a running
// pointer: ~3 instructions per copy).  A thread owns 2 (x) x 4 (y) pixels x Cout accumulators and streams the input rows of
// its strip once: row ir is loaded (k + 1 float4) and added into every output row ir - r it belongs to.  Weights sit in
// shared memory as [tap][quad][cout] float4 (broadcast reads); for 3x3 with Cout = 2 the 18 float4 of a quad are held in
// registers.  3-4 CTAs per SM.  Measured variants and what bounds the kernel: DESIGN.md 3.2, profiles/r2_ab_results.txt.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int HT_W = 64;   // tile width  (2 pixels per lane)
constexpr int HT_H = 16;   // tile height (4 rows per warp)
constexpr int HT_PY = 4;
constexpr int HT_CK = 4;   // input channels per chunk (one float4 per pixel)
constexpr int HT_THREADS = 128;

struct HeadParams {
  const float *in;
  int in_pitch, H, W, Cin;
  const float *weight;  // [k*k][n_pad][Cin] fp32
  const float *bias;    // [n_pad]
  int n_pad;
  int act;
  float slope, out_scale;
  float *out;
  int out_pitch;
  const float *res1;
  int res1_pitch;
  const float *res2;
  int res2_pitch;
  int tiles_x;
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void *src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__device__ __forceinline__ void fma4(float &acc, const float4 &a, const float4 &w) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  acc = fmaf(a.w, w.w, acc);
}

template <int K, int COUT>
__global__ void __launch_bounds__(HT_THREADS, (K == 3 ? 4 : 3)) conv_head_kernel(const HeadParams p) {
  constexpr int R = HT_H + K - 1;        // halo rows
  constexpr int WC = HT_W + K - 1;       // halo columns (even)
  constexpr int HP = WC / 2;             // float4 per column parity
  constexpr int ROW4 = 2 * HP;           // float4 per halo row
  constexpr int QUAD4 = R * ROW4;        // float4 per chunk buffer (one channel quad)
  constexpr int PAD = K / 2;
  constexpr int NWIN = K + 1;            // window columns a thread reads per row (2 pixels)
  constexpr bool WREG = (K == 3 && COUT == 2);   // the 9 x COUT weight float4 of a quad live in registers
  static_assert(WC <= HT_THREADS, "one fill thread per halo column");
  extern __shared__ float4 smem4[];
  float4 *const tile = smem4;                          // [2 buffers][R][2][HP]
  float4 *const wsm = smem4 + 2 * QUAD4;               // [K*K][Cin/4][COUT]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
  const int oy0 = ty * HT_H, ox0 = tx * HT_W;
  const int nq = p.Cin >> 2;

  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  // ---- fill plan: thread x < WC owns halo column x and walks its R rows with a running pointer; only the channel offset
  // changes from chunk to chunk ---------------------------------------------------------------------------------------
  const uint32_t tile_s = ptx::smem_u32(tile);
  const int fx = threadIdx.x;
  const int ixA = ox0 + fx - PAD;
  const bool vxA = fx < WC && ixA >= 0 && ixA < p.W;
  const long long row_stride = static_cast<long long>(p.W) * p.in_pitch;
  const float *srcA = p.in + (static_cast<long long>(oy0 - PAD) * p.W + (vxA ? ixA : 0)) * p.in_pitch;
  const uint32_t dstA = tile_s + 16u * static_cast<uint32_t>((fx & 1) * HP + (fx >> 1));
  auto fill = [&](int chunk, int buf) {
    if (fx < WC) {
      const uint32_t dst = dstA + static_cast<uint32_t>(buf) * (QUAD4 * 16u);
      const float *src = srcA + chunk * 4;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool valid = vxA && static_cast<unsigned>(oy0 - PAD + r) < static_cast<unsigned>(p.H);
        cp_async_16_zfill(dst + static_cast<uint32_t>(r * ROW4 * 16), valid ? static_cast<const void *>(src) : static_cast<const void *>(p.in), valid);
        src += row_stride;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill(0, 0);

  // weights: [tap][co][ci] in global memory -> [tap][quad][co] float4
  for (int i = threadIdx.x; i < K * K * nq * COUT; i += HT_THREADS) {
    const int co = i % COUT, t2 = i / COUT, q = t2 % nq, tap = t2 / nq;
    wsm[i] = __ldg(reinterpret_cast<const float4 *>(p.weight + (static_cast<long long>(tap) * p.n_pad + co) * p.Cin) + q);
  }

  float acc[HT_PY][2][COUT];
#pragma unroll
  for (int a = 0; a < HT_PY; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[a][b][c] = 0.f;

#pragma unroll 1
  for (int q = 0; q < nq; ++q) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's copies of chunk q have landed
    __syncthreads();                                        // ... everyone's have, and chunk q - 1 has been consumed
    if (q + 1 < nq) fill(q + 1, (q + 1) & 1);               // lands while chunk q is being computed
    const float4 *const tq = tile + (q & 1) * QUAD4 + (warp * HT_PY) * ROW4 + lane;
    const float4 *const wq = wsm + q * COUT;                // + tap * nq * COUT + co
    float4 wr[WREG ? K * K : 1][COUT];
    if (WREG) {
#pragma unroll
      for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int co = 0; co < COUT; ++co) wr[WREG ? t : 0][co] = wq[t * nq * COUT + co];
    }
#pragma unroll
    for (int ir = 0; ir < HT_PY + K - 1; ++ir) {
      float4 a[NWIN];
#pragma unroll
      for (int j = 0; j < NWIN; ++j) a[j] = tq[ir * ROW4 + (j & 1) * HP + (j >> 1)];
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const int orow = ir - r;
        if (orow < 0 || orow >= HT_PY) continue;
#pragma unroll
        for (int s = 0; s < K; ++s) {
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float4 w = WREG ? wr[WREG ? r * K + s : 0][co] : wq[(r * K + s) * nq * COUT + co];
            fma4(acc[orow][0][co], a[s], w);
            fma4(acc[orow][1][co], a[s + 1], w);
          }
        }
      }
    }
  }

  // ---- epilogue: bias, activation, scale, residuals -----------------------------------------------------------------
  float b[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) b[co] = __ldg(p.bias + co);
#pragma unroll
  for (int orow = 0; orow < HT_PY; ++orow) {
    const int oy = oy0 + warp * HT_PY + orow;
    if (oy >= p.H) continue;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int ox = ox0 + 2 * lane + px;
      if (ox >= p.W) continue;
      const long long pix = static_cast<long long>(oy) * p.W + ox;
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        float v = acc[orow][px][co] + b[co];
        if (p.act) v = v > 0.f ? v : v * p.slope;
        v *= p.out_scale;
        if (p.res1) v += p.res1[pix * p.res1_pitch + co];
        if (p.res2) v += p.res2[pix * p.res2_pitch + co];
        p.out[pix * p.out_pitch + co] = v;
      }
    }
  }
}

template <int K, int COUT>
int launch_head(const HeadParams &p, int grid, cudaStream_t s) {
  constexpr int R = HT_H + K - 1, WC = HT_W + K - 1;
  const size_t smem = static_cast<size_t>(2 * R * WC + K * K * (p.Cin / 4) * COUT) * 16;
  static bool attr_set = false;
  if (!attr_set) {
    LSSVC_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(conv_head_kernel<K, COUT>),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    attr_set = true;
  }
  LSSVC_REQUIRE(smem <= 72 * 1024, "conv_head: %zu bytes of shared memory", smem);
  LSSVC_CUDA(lssvc::launch_pdl(conv_head_kernel<K, COUT>, grid, HT_THREADS, smem, s, p));
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}

}  // namespace

extern "C" int32_t lssvc_conv_head_supported(const lssvc_conv *c) {
  if (c == nullptr || c->n_src != 1) return 0;
  const lssvc_view &v = c->src[0];
  const bool k_ok = c->kh == c->kw && (c->kh == 3 || c->kh == 7) && c->stride == 1 && c->pad == c->kh / 2;
  const bool c_ok = c->cout >= 2 && c->cout <= (c->kh == 7 ? 2 : 4) && v.C % HT_CK == 0 && v.C >= HT_CK && v.C <= 128 && v.C == c->cin_total;
  const bool a_ok = v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(c->weight) & 15) == 0 && c->cin_total % 4 == 0;
  const bool e_ok = c->epi == LSSVC_EPI_PLAIN && !c->pixel_shuffle && c->out2.ptr == nullptr && c->in_transform == LSSVC_IN_NONE;
  const size_t smem = static_cast<size_t>(2 * (HT_H + c->kh - 1) * (HT_W + c->kw - 1) + c->kh * c->kw * (v.C / 4) * c->cout) * 16;
  return (k_ok && c_ok && a_ok && e_ok && smem <= 72 * 1024) ? 1 : 0;
}

extern "C" int32_t lssvc_conv_head(const lssvc_conv *c, void *stream) {
  LSSVC_REQUIRE(c != nullptr, "conv_head: null descriptor");
  LSSVC_REQUIRE(lssvc_conv_head_supported(c), "conv_head: unsupported layer (k=%dx%d stride=%d pad=%d cout=%d cin=%d n_src=%d)", c->kh,
                c->kw, c->stride, c->pad, c->cout, c->cin_total, c->n_src);
  const lssvc_view &v = c->src[0];
  LSSVC_REQUIRE(lssvc::view_ok(&v) && lssvc::view_ok(&c->out), "conv_head: bad view");
  LSSVC_REQUIRE(c->out.H == v.H && c->out.W == v.W && c->out.C == c->cout, "conv_head: output view %dx%dx%d, expected %dx%dx%d",
                c->out.H, c->out.W, c->out.C, v.H, v.W, c->cout);
  HeadParams p;
  memset(&p, 0, sizeof(p));
  p.in = v.ptr; p.in_pitch = v.pitch; p.H = v.H; p.W = v.W; p.Cin = v.C;
  p.weight = c->weight; p.bias = c->bias; p.n_pad = c->n_pad;
  p.act = c->act; p.slope = c->slope; p.out_scale = c->out_scale;
  p.out = c->out.ptr; p.out_pitch = c->out.pitch;
  auto opt = [&](const lssvc_view &r, const float **ptr, int *pitch) -> bool {
    if (!r.ptr) { *ptr = nullptr; *pitch = 0; return true; }
    if (r.H != c->out.H || r.W != c->out.W || r.C != c->out.C) return false;
    *ptr = r.ptr; *pitch = r.pitch;
    return true;
  };
  LSSVC_REQUIRE(opt(c->res1, &p.res1, &p.res1_pitch), "conv_head: res1 shape mismatch");
  LSSVC_REQUIRE(opt(c->res2, &p.res2, &p.res2_pitch), "conv_head: res2 shape mismatch");
  p.tiles_x = lssvc::ceil_div(v.W, HT_W);
  const int grid = p.tiles_x * lssvc::ceil_div(v.H, HT_H);
  cudaStream_t s = lssvc::as_stream(stream);
  const int key = c->kh * 10 + c->cout;
  switch (key) {
    case 32: return launch_head<3, 2>(p, grid, s);
    case 33: return launch_head<3, 3>(p, grid, s);
    case 34: return launch_head<3, 4>(p, grid, s);
    case 72: return launch_head<7, 2>(p, grid, s);
  }
  LSSVC_REQUIRE(false, "conv_head: no instance for k=%d cout=%d", c->kh, c->cout);
}
