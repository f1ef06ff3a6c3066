// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM, operands staged by TMA) for NHWC fp32 activations.
//
//   D[M = 128 output pixels (8 x 16 patch), N = output channels] +=
//        A[M, K = KC input channels of one filter tap] * B[K, N]
//
// A is fetched per (source, channel chunk, tap) with one 5-D TMA box load of the input patch
// shifted by the tap offset; out-of-image pixels are zero-filled by TMA, which is the conv's
// zero padding.  The 5-D view [H/s][s][W/s][s][C] of the NHWC tensor makes a stride-s conv a
// unit-stride box load (coordinates (c, px, x, py, y)).  B is the packed weight
// [tap][n_pad][cin_total] fetched with a 3-D box.  Both land in the canonical K-major
// SWIZZLE_{128,64,32}B layout (row = KC * 4 bytes), so the UMMA descriptors are the plain ones.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue
// (TMEM -> registers -> bias / activation / residual / pixel-shuffle -> global), warps 6..9 operand
// splitters (3xTF32 mode only).  Persistent over tiles; two TMEM accumulator buffers overlap the
// epilogue of tile i with the MMAs of tile i + 1.
//
// Precision modes
//   TF32   : operands are used as they are (the tensor core keeps 10 mantissa bits of each fp32).
//   3xTF32 : error-compensated.  Weights are pre-split on the host into w_hi = rn_tf32(w) and
//            w_lo = rn_tf32(w - w_hi); the splitter warps rewrite each landed A tile in shared memory as
//            a_hi = rn_tf32(a) and a_lo = rn_tf32(a - a_hi), and the MMA warp issues
//            a_hi*w_hi + a_lo*w_hi + a_hi*w_lo into the same fp32 TMEM accumulator (the dropped a_lo*w_lo
//            term is below 2^-22 relative).  This is what gives the fp32-reference parity of the symbols.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int MAX_STAGES = 8;
constexpr int NUM_THREADS = 320;
constexpr int SPLIT_THREADS = 128;

struct alignas(64) TcParams {
  CUtensorMap a_map[LSSVC_MAX_SRC];
  CUtensorMap b_map;
  int n_src;
  int chunks[LSSVC_MAX_SRC];  // C / KC per source
  int coff[LSSVC_MAX_SRC];    // channel offset of the source inside the packed weight
  int kh, kw, stride, pad;
  int Ho, Wo;
  int tiles_x, tiles_y, n_tiles, n_tile;
  int cout;
  int stages, stage_bytes, tmem_cols;
  int split;  // 1: 3xTF32
  const float *bias;
  int act;
  float slope;
  float out_scale;
  int pixel_shuffle;
  int vec_ok;
  float *out;
  int out_pitch;
  const float *res1;
  int res1_pitch;
  const float *res2;
  int res2_pitch;
  float *out2;
  int out2_pitch;
  float slope2;
};

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// element offset of output channel `ch` of conv-output pixel (oy, ox) in a buffer of pitch P
__device__ __forceinline__ long long out_offset(const TcParams &p, int oy, int ox, int ch, int P) {
  if (!p.pixel_shuffle) return (static_cast<long long>(oy) * p.Wo + ox) * P + ch;
  const int cq = p.cout >> 2;
  const int sub = ch / cq;
  const int c = ch - sub * cq;
  const int i = sub >> 1, j = sub & 1;
  return (static_cast<long long>(2 * oy + i) * (2 * p.Wo) + (2 * ox + j)) * P + c;
}

template <int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int ROW_BYTES = KC * 4;
  constexpr int A_BYTES = TILE_H * TILE_W * ROW_BYTES;
  constexpr uint32_t LAYOUT = KC == 32 ? 2u : (KC == 16 ? 4u : 6u);
  constexpr uint32_t SBO = 8 * ROW_BYTES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t ready_bar[MAX_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_bytes = p.n_tile * ROW_BYTES;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.a_map[j]);
    ptx::prefetch_tensormap(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&ready_bar[s]), SPLIT_THREADS);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[b]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[b]), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), static_cast<uint32_t>(p.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_per_n = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_n * p.n_tiles;
  int iters = 0;
  for (int j = 0; j < p.n_src; ++j) iters += p.chunks[j];
  iters *= p.kh * p.kw;

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / tiles_per_n;
        const int rem = tile - nt * tiles_per_n;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int oy0 = ty * TILE_H, ox0 = tx * TILE_W, n0 = nt * p.n_tile;
        for (int j = 0; j < p.n_src; ++j) {
          for (int c = 0; c < p.chunks[j]; ++c) {
            for (int r = 0; r < p.kh; ++r) {
              const int dy = r - p.pad;
              const int qy = floor_div(dy, p.stride);
              const int py = dy - qy * p.stride;
              for (int s = 0; s < p.kw; ++s) {
                const int dx = s - p.pad;
                const int qx = floor_div(dx, p.stride);
                const int px = dx - qx * p.stride;
                const uint32_t full = ptx::smem_u32(&full_bar[stage]);
                ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
                ptx::mbar_expect_tx(full, static_cast<uint32_t>(A_BYTES + (p.split ? 2 : 1) * b_bytes));
                const uint32_t a_dst = smem_base + static_cast<uint32_t>(stage * p.stage_bytes);
                const uint32_t b_dst = a_dst + (p.split ? 2 : 1) * A_BYTES;
                ptx::tma_load_5d(a_dst, &p.a_map[j], full, c * KC, px, ox0 + qx, py, oy0 + qy);
                ptx::tma_load_3d(b_dst, &p.b_map, full, p.coff[j] + c * KC, n0, r * p.kw + s);
                if (p.split)
                  ptx::tma_load_3d(b_dst + b_bytes, &p.b_map, full, p.coff[j] + c * KC, n0,
                                   p.kh * p.kw + r * p.kw + s);
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1u;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ------------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = ptx::make_idesc_tf32_m128(static_cast<uint32_t>(p.n_tile));
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(ptx::smem_u32(&tempty_bar[buf]), acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.n_tile);
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(ptx::smem_u32(p.split ? &ready_bar[stage] : &full_bar[stage]), phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_base + static_cast<uint32_t>(stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + (p.split ? 2 : 1) * A_BYTES;
#pragma unroll
          for (int kk = 0; kk < KC / 8; ++kk) {
            const uint64_t a_desc = ptx::make_kmajor_desc(a_addr + kk * 32, SBO, LAYOUT);
            const uint64_t b_desc = ptx::make_kmajor_desc(b_addr + kk * 32, SBO, LAYOUT);
            ptx::mma_tf32(d_tmem, a_desc, b_desc, idesc, (it | kk) != 0 ? 1u : 0u);
            if (p.split) {
              const uint64_t al_desc = ptx::make_kmajor_desc(a_addr + A_BYTES + kk * 32, SBO, LAYOUT);
              const uint64_t bl_desc = ptx::make_kmajor_desc(b_addr + b_bytes + kk * 32, SBO, LAYOUT);
              ptx::mma_tf32(d_tmem, al_desc, b_desc, idesc, 1u);
              ptx::mma_tf32(d_tmem, a_desc, bl_desc, idesc, 1u);
            }
          }
          ptx::mma_commit(ptx::smem_u32(&empty_bar[stage]));
          if (it == iters - 1) ptx::mma_commit(ptx::smem_u32(&tfull_bar[buf]));
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      buf ^= 1;
      if (buf == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 6) {
    // ------------------------------- operand splitters (3xTF32) ---------------------------
    if (p.split) {
      const int t = threadIdx.x - 6 * 32;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int it = 0; it < iters; ++it) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          uint8_t *a_hi = smem_raw + (smem_base - ptx::smem_u32(smem_raw)) + stage * p.stage_bytes;
          uint8_t *a_lo = a_hi + A_BYTES;
#pragma unroll
          for (int off = 0; off < A_BYTES; off += SPLIT_THREADS * 16) {
            const float4 v = *reinterpret_cast<const float4 *>(a_hi + off + t * 16);
            // hi = rn_tf32(a); lo = rn_tf32(a - hi): both exactly representable, so the tensor core's own
            // fp32 -> TF32 conversion is the identity and the split error is unbiased (< 2^-22 |a|)
            const float hx = ptx::rn_tf32(v.x), hy = ptx::rn_tf32(v.y), hz = ptx::rn_tf32(v.z), hw = ptx::rn_tf32(v.w);
            *reinterpret_cast<float4 *>(a_hi + off + t * 16) = make_float4(hx, hy, hz, hw);
            *reinterpret_cast<float4 *>(a_lo + off + t * 16) =
                make_float4(ptx::rn_tf32(v.x - hx), ptx::rn_tf32(v.y - hy), ptx::rn_tf32(v.z - hz), ptx::rn_tf32(v.w - hw));
          }
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(ptx::smem_u32(&ready_bar[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ------------------------------- epilogue ---------------------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int h = m / TILE_W, w = m % TILE_W;
    int buf = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile / tiles_per_n;
      const int rem = tile - nt * tiles_per_n;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int oy = ty * TILE_H + h, ox = tx * TILE_W + w, n0 = nt * p.n_tile;
      const bool valid = (oy < p.Ho) && (ox < p.Wo);
      ptx::mbar_wait(ptx::smem_u32(&tfull_bar[buf]), acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(buf * p.n_tile);
      for (int n = 0; n < p.n_tile; n += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(t_row + static_cast<uint32_t>(n), r);
        ptx::tmem_ld_wait();
        const int cg = n0 + n;
        if (valid && cg < p.cout) {
          if (p.vec_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int ch = cg + 4 * g;
              if (ch < p.cout) {
                const float4 b4 = *reinterpret_cast<const float4 *>(p.bias + ch);
                float v[4] = {__uint_as_float(r[4 * g + 0]) + b4.x, __uint_as_float(r[4 * g + 1]) + b4.y,
                              __uint_as_float(r[4 * g + 2]) + b4.z, __uint_as_float(r[4 * g + 3]) + b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (p.act) v[e] = v[e] > 0.f ? v[e] : v[e] * p.slope;
                  v[e] *= p.out_scale;
                }
                if (p.res1) {
                  const float4 t = *reinterpret_cast<const float4 *>(p.res1 + out_offset(p, oy, ox, ch, p.res1_pitch));
                  v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                if (p.res2) {
                  const float4 t = *reinterpret_cast<const float4 *>(p.res2 + out_offset(p, oy, ox, ch, p.res2_pitch));
                  v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                *reinterpret_cast<float4 *>(p.out + out_offset(p, oy, ox, ch, p.out_pitch)) =
                    make_float4(v[0], v[1], v[2], v[3]);
                if (p.out2) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * p.slope2;
                  *reinterpret_cast<float4 *>(p.out2 + out_offset(p, oy, ox, ch, p.out2_pitch)) =
                      make_float4(v[0], v[1], v[2], v[3]);
                }
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int ch = cg + e;
              if (ch < p.cout) {
                float v = __uint_as_float(r[e]) + p.bias[ch];
                if (p.act) v = v > 0.f ? v : v * p.slope;
                v *= p.out_scale;
                if (p.res1) v += p.res1[out_offset(p, oy, ox, ch, p.res1_pitch)];
                if (p.res2) v += p.res2[out_offset(p, oy, ox, ch, p.res2_pitch)];
                p.out[out_offset(p, oy, ox, ch, p.out_pitch)] = v;
                if (p.out2) p.out2[out_offset(p, oy, ox, ch, p.out2_pitch)] = v > 0.f ? v : v * p.slope2;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[buf]));
      buf ^= 1;
      if (buf == 0) acc_phase ^= 1u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
bool g_attr_set[3] = {false, false, false};

int resolve_driver() {
  if (g_encode) return 0;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    lssvc::set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return LSSVC_ERR_NO_DEVICE;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  return 0;
}

CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace

extern "C" int32_t lssvc_device_check(int32_t dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    lssvc::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return LSSVC_ERR_NO_DEVICE;
  }
  if (prop.major != 10) {
    lssvc::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    return LSSVC_ERR_NO_DEVICE;
  }
  return resolve_driver();
}

extern "C" int32_t lssvc_conv_tc(const lssvc_conv *c, void *stream) {
  LSSVC_REQUIRE(c != nullptr, "conv_tc: null descriptor");
  LSSVC_REQUIRE(c->n_src >= 1 && c->n_src <= LSSVC_MAX_SRC, "conv_tc: n_src=%d", c->n_src);
  LSSVC_REQUIRE(c->in_transform == LSSVC_IN_NONE && c->epi == LSSVC_EPI_PLAIN,
                "conv_tc: input transforms / GDN epilogue are SIMT-only");
  LSSVC_REQUIRE(c->stride == 1 || c->stride == 2, "conv_tc: stride %d", c->stride);

  const int Hin = c->src[0].H, Win = c->src[0].W;
  int kc = 32;
  int cin_total = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    LSSVC_REQUIRE(lssvc::view_ok(&v), "conv_tc: bad source view %d", j);
    LSSVC_REQUIRE(v.H == Hin && v.W == Win, "conv_tc: source %d is %dx%d, expected %dx%d", j, v.H, v.W, Hin, Win);
    LSSVC_REQUIRE(v.C % 8 == 0, "conv_tc: source %d has %d channels (need a multiple of 8)", j, v.C);
    LSSVC_REQUIRE(v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0,
                  "conv_tc: source %d is not 16-byte aligned", j);
    while (v.C % kc) kc >>= 1;
    cin_total += v.C;
  }
  LSSVC_REQUIRE(cin_total == c->cin_total, "conv_tc: cin_total %d != sum of sources %d", c->cin_total, cin_total);
  LSSVC_REQUIRE(Hin % c->stride == 0 && Win % c->stride == 0, "conv_tc: %dx%d not divisible by stride", Hin, Win);
  const int Ho = (Hin + 2 * c->pad - c->kh) / c->stride + 1;
  const int Wo = (Win + 2 * c->pad - c->kw) / c->stride + 1;
  const int ps = c->pixel_shuffle ? 2 : 1;
  LSSVC_REQUIRE(lssvc::view_ok(&c->out), "conv_tc: bad output view");
  LSSVC_REQUIRE(c->out.H == Ho * ps && c->out.W == Wo * ps, "conv_tc: output view %dx%d, expected %dx%d",
                c->out.H, c->out.W, Ho * ps, Wo * ps);
  LSSVC_REQUIRE(!c->pixel_shuffle || c->cout % 4 == 0, "conv_tc: pixel shuffle needs cout %% 4 == 0");
  const int c_store = c->pixel_shuffle ? c->cout / 4 : c->cout;
  LSSVC_REQUIRE(c->out.C == c_store, "conv_tc: output view has %d channels, expected %d", c->out.C, c_store);
  LSSVC_REQUIRE(c->n_pad % 16 == 0 && c->n_pad >= c->cout, "conv_tc: n_pad=%d cout=%d", c->n_pad, c->cout);
  LSSVC_REQUIRE((reinterpret_cast<uintptr_t>(c->weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(c->bias) & 15) == 0,
                "conv_tc: weight/bias not 16-byte aligned");
  LSSVC_REQUIRE(c->precision == LSSVC_PREC_TF32 || c->precision == LSSVC_PREC_3XTF32, "conv_tc: precision %d", c->precision);
  LSSVC_REQUIRE(c->precision != LSSVC_PREC_3XTF32 ||
                    (c->weight_split && (reinterpret_cast<uintptr_t>(c->weight_split) & 15) == 0),
                "conv_tc: 3xTF32 needs the hi|lo split weights");

  int n_tile = c->n_pad;
  if (n_tile > 256) {
    n_tile = 256;
    while (n_tile >= 16 && (c->n_pad % n_tile)) n_tile -= 16;
    LSSVC_REQUIRE(n_tile >= 16, "conv_tc: cannot tile n_pad=%d", c->n_pad);
  }

  {
    const int ps_ = c->pixel_shuffle ? 2 : 1;
    auto shape_ok = [&](const lssvc_view &v) {
      return !v.ptr || (v.H == Ho * ps_ && v.W == Wo * ps_ && v.C == c_store);
    };
    LSSVC_REQUIRE(shape_ok(c->res1) && shape_ok(c->res2) && shape_ok(c->out2), "conv_tc: res1/res2/out2 shape mismatch");
  }
  if (int rc = resolve_driver()) return rc;

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_src = c->n_src;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const int st = c->stride;
  int coff = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    p.chunks[j] = v.C / kc;
    p.coff[j] = coff;
    coff += v.C;
    const cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(st),
                                static_cast<cuuint64_t>(Win / st), static_cast<cuuint64_t>(st),
                                static_cast<cuuint64_t>(Hin / st)};
    const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
    const cuuint64_t strides[4] = {px, px * st, px * Win, px * Win * st};
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(kc), 1, TILE_W, 1, TILE_H};
    CUresult r = g_encode(&p.a_map[j], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, v.ptr, dims, strides, box, ones,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_tc: cuTensorMapEncodeTiled(A%d) failed with %d (C=%d pitch=%d %dx%d stride=%d)", j,
                       static_cast<int>(r), v.C, v.pitch, Hin, Win, st);
      return LSSVC_ERR_CUDA;
    }
  }
  {
    const int taps = c->kh * c->kw;
    const bool split = c->precision == LSSVC_PREC_3XTF32;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cin_total), static_cast<cuuint64_t>(c->n_pad),
                                static_cast<cuuint64_t>(split ? 2 * taps : taps)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(cin_total) * 4,
                                   static_cast<cuuint64_t>(cin_total) * 4 * c->n_pad};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(n_tile), 1};
    const float *wsrc = split ? c->weight_split : c->weight;
    CUresult r = g_encode(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(wsrc), dims, strides,
                          box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_tc: cuTensorMapEncodeTiled(B) failed with %d", static_cast<int>(r));
      return LSSVC_ERR_CUDA;
    }
  }
  p.kh = c->kh; p.kw = c->kw; p.stride = st; p.pad = c->pad;
  p.Ho = Ho; p.Wo = Wo;
  p.tiles_x = lssvc::ceil_div(Wo, TILE_W);
  p.tiles_y = lssvc::ceil_div(Ho, TILE_H);
  p.n_tiles = c->n_pad / n_tile;
  p.n_tile = n_tile;
  p.cout = c->cout;
  const int row_bytes = kc * 4;
  p.split = c->precision == LSSVC_PREC_3XTF32 ? 1 : 0;
  p.stage_bytes = (((TILE_H * TILE_W + n_tile) * row_bytes * (p.split ? 2 : 1)) + 1023) & ~1023;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  LSSVC_REQUIRE(p.stages >= 2, "conv_tc: stage of %d bytes does not fit twice", p.stage_bytes);
  int cols = 32;
  while (cols < 2 * n_tile) cols <<= 1;
  p.tmem_cols = cols;
  p.bias = c->bias;
  p.act = c->act; p.slope = c->slope; p.out_scale = c->out_scale;
  p.pixel_shuffle = c->pixel_shuffle;
  p.out = c->out.ptr; p.out_pitch = c->out.pitch;
  bool vec = (c->cout % 4 == 0) && (c->out.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(c->out.ptr) & 15) == 0);
  if (c->pixel_shuffle) vec = vec && ((c->cout / 4) % 4 == 0);
  auto opt = [&](const lssvc_view &v, const float **ptr, int *pitch) -> bool {
    if (!v.ptr) { *ptr = nullptr; *pitch = 0; return true; }
    if (v.H != c->out.H || v.W != c->out.W || v.C != c->out.C) return false;
    *ptr = v.ptr; *pitch = v.pitch;
    vec = vec && (v.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0);
    return true;
  };
  LSSVC_REQUIRE(opt(c->res1, &p.res1, &p.res1_pitch), "conv_tc: res1 shape mismatch");
  LSSVC_REQUIRE(opt(c->res2, &p.res2, &p.res2_pitch), "conv_tc: res2 shape mismatch");
  const float *o2 = nullptr;
  LSSVC_REQUIRE(opt(c->out2, &o2, &p.out2_pitch), "conv_tc: out2 shape mismatch");
  p.out2 = const_cast<float *>(o2);
  p.slope2 = c->slope2;
  p.vec_ok = vec ? 1 : 0;

  const int total_tiles = p.tiles_x * p.tiles_y * p.n_tiles;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  const int ki = kc == 32 ? 0 : (kc == 16 ? 1 : 2);
  cudaStream_t s = lssvc::as_stream(stream);
  if (!g_attr_set[ki]) {
    const void *fn = kc == 32 ? reinterpret_cast<const void *>(conv_tc_kernel<32>)
                              : (kc == 16 ? reinterpret_cast<const void *>(conv_tc_kernel<16>)
                                          : reinterpret_cast<const void *>(conv_tc_kernel<8>));
    LSSVC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 205 * 1024));
    g_attr_set[ki] = true;
  }
  if (kc == 32) conv_tc_kernel<32><<<grid, NUM_THREADS, smem, s>>>(p);
  else if (kc == 16) conv_tc_kernel<16><<<grid, NUM_THREADS, smem, s>>>(p);
  else conv_tc_kernel<8><<<grid, NUM_THREADS, smem, s>>>(p);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
