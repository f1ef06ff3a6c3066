// Pointwise (1x1) convolution with shared-memory-resident weights, optionally preceded by the depthwise 3x3 of
// DepthConv (reference: src/InterModules/lssvc_modules.py:15-40):
//
//     y = act(W . u + b) * out_scale (+ res1) (+ res2),        u = x   or   u = dw3x3(x) + bd   (zero padding 1)
//
// These layers move 2*C*4 bytes per pixel for 2*Cin*Cout FLOPs per pixel: they are HBM-bound, and in the general
// implicit-GEMM kernel (conv_hs.cu) they pay a per-tile weight reload and a deep pipeline built for 3x3 taps.
// Here the split-fp16 weights (4*Cin*Cout bytes) are loaded once per CTA, the tile pipeline is
//   TMA tile (fp32, +1-pixel apron when the depthwise conv is fused) -> 16 operand warps (dw3x3 in fp32 registers,
//   split to fp16 hi/lo, tcgen05.st into a double-buffered A operand in TMEM) -> Cin/16 x 2 MMAs -> double-buffered
//   accumulators -> 8 epilogue warps -> swizzled smem staging -> TMA store,
// and fusing the depthwise conv removes its own kernel plus one write + read of the C-channel intermediate.
// Same arithmetic as conv_hs.cu: x = x_hi + x_lo in fp16, D1 += A_hi*W_hi, D2 += A_hi*W_lo + A_lo*W_hi.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int MAX_IN_BUFS = 3;
constexpr int NUM_THREADS = 896;
constexpr int TMEM_COLS = 512;

struct alignas(64) PwParams {
  CUtensorMap in_map;   // x [H][W][Cin] fp32, box (slab_w, 16 (+2), 8 (+2))
  CUtensorMap out_map;  // y [H][W][Cout]
  CUtensorMap res_map;  // res1 [H][W][Cout], TMA-loaded into the staging tile ahead of the accumulator (res_tma)
  int res_tma;
  const void *w;        // fp16 [Cin/16][2 (hi, lo)][Cout][16], 32-byte rows pre-swizzled (SWIZZLE_32B)
  const float *bias;    // [Cout]
  const float *dw_w;    // [9][Cin] (tap-major) or null
  const float *dw_b;    // [Cin]
  const float *res1;
  const float *res2;
  int res1_pitch, res2_pitch;
  int Cin, Cout;
  int H, W, tiles_x, tiles_y;
  int in_slab_w, in_slabs, in_rows, in_w, in_slab_stride, in_bytes, in_tx;  // input tile geometry in shared memory
  int out_slab_w, out_slabs;
  int w_bytes, dw_off, in_off, stage_off, in_bufs;
  int act;
  float slope, out_scale, acc_scale;
  unsigned int *range_flag;  // raised when a split operand leaves the fp16 range (range.cu)
};

__device__ __forceinline__ void split_pair_p(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// address of channel c (multiple of 4) of row r in a swizzled [slab][rows][slab_w] fp32 tile
__device__ __forceinline__ uint32_t slab_addr(uint32_t base, int slab_w, uint32_t slab_stride, int r, int c) {
  const int slab = c / slab_w;
  const uint32_t row = base + static_cast<uint32_t>(slab) * slab_stride + static_cast<uint32_t>(r) * (slab_w * 4u);
  const uint32_t swz = (slab_w == 32 ? static_cast<uint32_t>(r & 7) : static_cast<uint32_t>((r >> 1) & 3)) << 4;
  return row + ((static_cast<uint32_t>(c % slab_w) << 2) ^ swz);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_pw_kernel(const __grid_constant__ PwParams p) {
  ptx::pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t in_full[MAX_IN_BUFS], in_empty[MAX_IN_BUFS];
  __shared__ uint64_t a_full[2], a_empty[2], d_full[2], d_empty[2], w_full, res_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::pin((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  uint8_t *const smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int Cin = p.Cin, Cout = p.Cout, KS = Cin >> 4;
  const bool dw = p.dw_w != nullptr;
  const uint32_t w_s = smem_base, dw_s = smem_base + static_cast<uint32_t>(p.dw_off);
  const uint32_t in_s = smem_base + static_cast<uint32_t>(p.in_off), stage_s = smem_base + static_cast<uint32_t>(p.stage_off);
  const uint32_t in_bytes = static_cast<uint32_t>(p.in_bytes), in_slab_stride = static_cast<uint32_t>(p.in_slab_stride);
  const uint32_t b_in_full = ptx::pin(ptx::smem_u32(in_full)), b_in_empty = ptx::pin(ptx::smem_u32(in_empty));
  const uint32_t b_a_full = ptx::pin(ptx::smem_u32(a_full)), b_a_empty = ptx::pin(ptx::smem_u32(a_empty));
  const uint32_t b_d_full = ptx::pin(ptx::smem_u32(d_full)), b_d_empty = ptx::pin(ptx::smem_u32(d_empty));
  const uint32_t b_w_full = ptx::pin(ptx::smem_u32(&w_full));
  const uint32_t b_res_full = ptx::pin(ptx::smem_u32(&res_full));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.in_map);
    ptx::prefetch_tensormap(&p.out_map);
    if (p.res_tma) ptx::prefetch_tensormap(&p.res_map);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < p.in_bufs; ++b) {
      ptx::mbar_init(b_in_full + 8 * b, 1);
      ptx::mbar_init(b_in_empty + 8 * b, 16);  // the 16 operand warps
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(b_a_full + 8 * b, 16);
      ptx::mbar_init(b_a_empty + 8 * b, 1);
      ptx::mbar_init(b_d_full + 8 * b, 1);
      ptx::mbar_init(b_d_empty + 8 * b, 8);
    }
    ptx::mbar_init(b_w_full, 1);
    ptx::mbar_init(b_res_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::pdl_wait();  // first global-memory access below (ptx.cuh: programmatic dependent launch)
  if (dw) {
    // depthwise weights [9][Cin] + bias [Cin] -> shared memory (read as broadcast float4 by the operand warps)
    float *dst = reinterpret_cast<float *>(smem_gen + p.dw_off);
    for (int i = threadIdx.x; i < 10 * Cin; i += NUM_THREADS) dst[i] = i < 9 * Cin ? p.dw_w[i] : p.dw_b[i - 9 * Cin];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // TMEM columns: A[2] Cin each | D[2] 2*Cout each
  const uint32_t t_a = tmem_base, t_d = tmem_base + static_cast<uint32_t>(2 * Cin);
  const int total_tiles = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------- input TMA producer -----------------------------------
    if (ptx::elect_one()) {
      int ib = 0;
      uint32_t iph = 0;
      const int org = dw ? -1 : 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        ptx::mbar_wait(b_in_empty + 8 * ib, iph ^ 1u);
        ptx::mbar_expect_tx(b_in_full + 8 * ib, static_cast<uint32_t>(p.in_tx));
        for (int s = 0; s < p.in_slabs; ++s)
          ptx::tma_load_3d(in_s + static_cast<uint32_t>(ib) * in_bytes + static_cast<uint32_t>(s) * in_slab_stride, &p.in_map,
                           b_in_full + 8 * ib, s * p.in_slab_w, tx * TILE_W + org, ty * TILE_H + org);
        if (++ib == p.in_bufs) {
          ib = 0;
          iph ^= 1u;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------- weight loader (once) ---------------------------------
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(b_w_full, static_cast<uint32_t>(p.w_bytes));
      const uint8_t *g = reinterpret_cast<const uint8_t *>(p.w);
      for (int off = 0; off < p.w_bytes; off += 16384) {
        const int n = p.w_bytes - off < 16384 ? p.w_bytes - off : 16384;
        ptx::bulk_load_1d(w_s + off, g + off, n, b_w_full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------------------
    const uint32_t idesc_2n = ptx::make_idesc_f16_m128(static_cast<uint32_t>(2 * Cout));
    const uint32_t idesc_n = ptx::make_idesc_f16_m128(static_cast<uint32_t>(Cout));
    const uint32_t w_sub = 2u * static_cast<uint32_t>(Cout) * 32u;  // bytes of one [2][Cout][16] sub-tile
    int b = 0;
    uint32_t ph = 0;
    ptx::mbar_wait(b_w_full, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(b_a_full + 8 * b, ph);
      ptx::mbar_wait(b_d_empty + 8 * b, ph ^ 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t a = t_a + static_cast<uint32_t>(b * Cin), d = t_d + static_cast<uint32_t>(b * 2 * Cout);
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t bd = ptx::make_kmajor_desc(w_s + static_cast<uint32_t>(ks) * w_sub, 256, 6u);
          ptx::mma_f16_ts(d, a + ks * 8, bd, idesc_2n, ks != 0 ? 1u : 0u);
          ptx::mma_f16_ts(d + Cout, a + (Cin >> 1) + ks * 8, bd, idesc_n, 1u);
        }
        ptx::mma_commit(b_a_empty + 8 * b);
        ptx::mma_commit(b_d_full + 8 * b);
      }
      __syncwarp();
      if (++b == 2) {
        b = 0;
        ph ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 20) {
    // ------------------------------- operand warps: tile -> (dw3x3) -> split -> A (TMEM) ---
    const int ow = warp - 4;
    const int q = ow & 3;    // TMEM lane quarter (== warp % 4)
    const int sg = ow >> 2;  // K-slice group: slices sg, sg + 4, ...
    const int m = q * 32 + lane;
    const int h = m / TILE_W, w = m % TILE_W;
    const int in_w = p.in_w;
    const int rc = dw ? (h + 1) * in_w + (w + 1) : m;  // row of this pixel in the input tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool s32 = p.in_slab_w == 32;
    const uint32_t row_bytes = s32 ? 128u : 64u;
    int ib = 0, b = 0;
    uint32_t iph = 0, ph = 0;
    float amax = 0.f;  // running max |operand| of this thread (range guard)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(b_in_full + 8 * ib, iph);
      const uint32_t tile_s = in_s + static_cast<uint32_t>(ib) * in_bytes;
      ptx::mbar_wait(b_a_empty + 8 * b, ph ^ 1u);
      ptx::tc_fence_after();
      const uint32_t dst = t_a + static_cast<uint32_t>(b * Cin) + lane_off;
      for (int ks = sg; ks < KS; ks += 4) {
        float4 v[4];
        // a 16-channel slice lies inside one slab: slab base and chunk offset once per slice, row base and swizzle once per
        // tap (slab_addr's divisions in the tap loop were 20 % of the kernel's instructions)
        const uint32_t c0 = static_cast<uint32_t>(ks) * 16u;
        const uint32_t slab_base = tile_s + (s32 ? (c0 >> 5) : (c0 >> 4)) * in_slab_stride;
        const uint32_t coff0 = s32 ? ((c0 & 31u) << 2) : 0u;
        if (dw) {
          const uint32_t wb = dw_s + c0 * 4u;
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = ptx::lds_f4(wb + static_cast<uint32_t>(9 * Cin + 4 * i) * 4u);  // bias
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t r = static_cast<uint32_t>(rc + (t / 3 - 1) * in_w + (t % 3 - 1));
            const uint32_t rb = slab_base + r * row_bytes;
            const uint32_t sw = (s32 ? (r & 7u) : ((r >> 1) & 3u)) << 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 x4 = ptx::lds_f4(rb + ((coff0 + 16u * i) ^ sw));
              const float4 w4 = ptx::lds_f4(wb + static_cast<uint32_t>(t * Cin + 4 * i) * 4u);
              v[i].x = fmaf(x4.x, w4.x, v[i].x);
              v[i].y = fmaf(x4.y, w4.y, v[i].y);
              v[i].z = fmaf(x4.z, w4.z, v[i].z);
              v[i].w = fmaf(x4.w, w4.w, v[i].w);
            }
          }
        } else {
          const uint32_t r = static_cast<uint32_t>(rc);
          const uint32_t rb = slab_base + r * row_bytes;
          const uint32_t sw = (s32 ? (r & 7u) : ((r >> 1) & 3u)) << 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = ptx::lds_f4(rb + ((coff0 + 16u * i) ^ sw));
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
          split_pair_p(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
          split_pair_p(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
        }
        ptx::tmem_st8(dst + ks * 8, hi);
        ptx::tmem_st8(dst + (Cin >> 1) + ks * 8, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(b_a_full + 8 * b);
        ptx::mbar_arrive(b_in_empty + 8 * ib);
      }
      if (++ib == p.in_bufs) {
        ib = 0;
        iph ^= 1u;
      }
      if (++b == 2) {
        b = 0;
        ph ^= 1u;
      }
    }
    if (amax >= lssvc::kSplitRangeLimit && p.range_flag) atomicOr(p.range_flag, 1u);
  } else if (warp >= 20) {
    // ------------------------------- epilogue ----------------------------------------------
    const int ew = warp - 20;
    const int q = ew & 3;
    const int eset = ew >> 2;  // 16-channel chunks eset, eset + 2, ...
    const int m = q * 32 + lane;
    const int h = m / TILE_W, w = m % TILE_W;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float acc_scale = p.acc_scale, out_scale = p.out_scale, slope = p.slope;
    const bool has_act = p.act != 0;
    const bool store_thread = warp == 20 && lane == 0;
    const uint32_t out_slab_stride = 128u * static_cast<uint32_t>(p.out_slab_w) * 4u;
    int b = 0;
    uint32_t ph = 0, res_ph = 0;
    // LeakyReLU as max(v, slope v) (exact for 0 <= slope <= 1; slope = 1 without activation): no per-element branches.  The
    // epilogue warps are instruction-latency bound, so the lean form is what lets the kernel follow HBM.
    const float sl = has_act ? slope : 1.f;
    const bool lean_act = !has_act || (slope >= 0.f && slope <= 1.f);
    const bool r_tma = p.res_tma != 0;
    const float *const res1 = r_tma ? nullptr : p.res1;
    const float *const res2 = p.res2;
    const int out_slab_w = p.out_slab_w;
    const uint32_t sswz = (out_slab_w == 32 ? static_cast<uint32_t>(m & 7) : static_cast<uint32_t>((m >> 1) & 3)) << 4;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int oy = ty * TILE_H + h, ox = tx * TILE_W + w;
      const bool valid = oy < p.H && ox < p.W;
      const long long pix = static_cast<long long>(oy) * p.W + ox;
      if (store_thread) {
        ptx::bulk_wait_read_all();  // staging free: the previous tile's TMA store has read it
        if (r_tma) {                // the residual tile lands in the staging buffer while the MMAs of this tile run
          ptx::mbar_expect_tx(b_res_full, static_cast<uint32_t>(p.out_slabs) * out_slab_stride);
          for (int s = 0; s < p.out_slabs; ++s)
            ptx::tma_load_3d(stage_s + static_cast<uint32_t>(s) * out_slab_stride, &p.res_map, b_res_full, s * out_slab_w, tx * TILE_W,
                             ty * TILE_H);
        }
      }
      ptx::named_bar_sync(2, 256);
      ptx::mbar_wait(b_d_full + 8 * b, ph);
      if (r_tma) {
        ptx::mbar_wait(b_res_full, res_ph);
        res_ph ^= 1u;
      }
      ptx::tc_fence_after();
      const uint32_t src = t_d + static_cast<uint32_t>(b * 2 * Cout) + lane_off;
      for (int n = 16 * eset; n < Cout; n += 32) {
        uint32_t r1[16], r2[16];
        ptx::tmem_ld16(src + n, r1);
        ptx::tmem_ld16(src + Cout + n, r2);
        float4 bv[4];
        const uint32_t srow = static_cast<uint32_t>(n / out_slab_w) * out_slab_stride + static_cast<uint32_t>(m) * (out_slab_w * 4u);
        const uint32_t spiece = static_cast<uint32_t>(n % out_slab_w) << 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bv[i] = __ldg(reinterpret_cast<const float4 *>(p.bias + n) + i);
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[4];
          v[0] = (__uint_as_float(r1[4 * i + 0]) + __uint_as_float(r2[4 * i + 0])) * acc_scale + bv[i].x;
          v[1] = (__uint_as_float(r1[4 * i + 1]) + __uint_as_float(r2[4 * i + 1])) * acc_scale + bv[i].y;
          v[2] = (__uint_as_float(r1[4 * i + 2]) + __uint_as_float(r2[4 * i + 2])) * acc_scale + bv[i].z;
          v[3] = (__uint_as_float(r1[4 * i + 3]) + __uint_as_float(r2[4 * i + 3])) * acc_scale + bv[i].w;
          if (lean_act) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], v[e] * sl) * out_scale;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (v[e] > 0.f ? v[e] : v[e] * slope) * out_scale;
          }
          if (r_tma) {
            const float4 t = ptx::lds_f4(stage_s + srow + ((spiece + 16u * i) ^ sswz));
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          } else if (res1 && valid) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(res1 + pix * p.res1_pitch + n) + i);
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          }
          if (res2 && valid) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(res2 + pix * p.res2_pitch + n) + i);
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          }
          ptx::sts_u4(stage_s + srow + ((spiece + 16u * i) ^ sswz), __float_as_uint(v[0]), __float_as_uint(v[1]),
                      __float_as_uint(v[2]), __float_as_uint(v[3]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(b_d_empty + 8 * b);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(3, 256);
      if (store_thread) {
        for (int s = 0; s < p.out_slabs; ++s)
          ptx::tma_store_3d(&p.out_map, stage_s + static_cast<uint32_t>(s) * out_slab_stride, s * p.out_slab_w, tx * TILE_W,
                            ty * TILE_H);
        ptx::bulk_commit();
      }
      if (++b == 2) {
        b = 0;
        ph ^= 1u;
      }
    }
    if (store_thread) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
bool g_attr_set = false;

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

}  // namespace

extern "C" int32_t lssvc_conv_pw(const lssvc_pw *f, void *stream) {
  LSSVC_REQUIRE(f != nullptr, "conv_pw: null descriptor");
  const int Cin = f->in.C, Cout = f->out.C;
  LSSVC_REQUIRE(lssvc::view_ok(&f->in) && lssvc::view_ok(&f->out), "conv_pw: bad views");
  LSSVC_REQUIRE(f->out.H == f->in.H && f->out.W == f->in.W, "conv_pw: output view mismatch");
  LSSVC_REQUIRE(Cin % 16 == 0 && Cin >= 16 && Cin <= 128, "conv_pw: Cin=%d (multiple of 16 up to 128)", Cin);
  LSSVC_REQUIRE(Cout % 16 == 0 && Cout >= 16 && Cout <= 64, "conv_pw: Cout=%d (multiple of 16 up to 64)", Cout);
  LSSVC_REQUIRE(2 * Cin + 4 * Cout <= TMEM_COLS, "conv_pw: Cin=%d Cout=%d do not fit in tensor memory", Cin, Cout);
  auto aligned = [](const lssvc_view &v) { return v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0; };
  LSSVC_REQUIRE(aligned(f->in) && aligned(f->out), "conv_pw: views must be 16-byte aligned");
  LSSVC_REQUIRE(f->w && f->bias && (reinterpret_cast<uintptr_t>(f->w) & 15) == 0 && (reinterpret_cast<uintptr_t>(f->bias) & 15) == 0,
                "conv_pw: weights / bias must be non-null and 16-byte aligned");
  for (const lssvc_view *r : {&f->res1, &f->res2}) {
    if (r->ptr) LSSVC_REQUIRE(r->H == f->in.H && r->W == f->in.W && r->C == Cout && aligned(*r), "conv_pw: residual view mismatch");
  }
  LSSVC_REQUIRE((f->dw_weight == nullptr) == (f->dw_bias == nullptr), "conv_pw: dw_weight and dw_bias go together");
  if (int rc = resolve_driver()) return rc;

  PwParams p;
  memset(&p, 0, sizeof(p));
  const bool dw = f->dw_weight != nullptr;
  p.Cin = Cin; p.Cout = Cout;
  p.H = f->in.H; p.W = f->in.W;
  p.tiles_x = lssvc::ceil_div(p.W, TILE_W);
  p.tiles_y = lssvc::ceil_div(p.H, TILE_H);
  p.in_slab_w = Cin % 32 == 0 ? 32 : 16;
  p.in_slabs = Cin / p.in_slab_w;
  p.in_w = dw ? TILE_W + 2 : TILE_W;
  p.in_rows = p.in_w * (dw ? TILE_H + 2 : TILE_H);
  p.in_tx = p.in_rows * Cin * 4;
  p.in_slab_stride = (p.in_rows * p.in_slab_w * 4 + 1023) & ~1023;
  p.in_bytes = p.in_slabs * p.in_slab_stride;
  p.out_slab_w = Cout % 32 == 0 ? 32 : 16;
  p.out_slabs = Cout / p.out_slab_w;
  p.w_bytes = Cin * Cout * 4;
  p.w = f->w; p.bias = f->bias; p.dw_w = f->dw_weight; p.dw_b = f->dw_bias;
  p.res1 = f->res1.ptr; p.res1_pitch = f->res1.pitch;
  p.res2 = f->res2.ptr; p.res2_pitch = f->res2.pitch;
  p.act = f->act; p.slope = f->slope; p.out_scale = f->out_scale; p.acc_scale = f->acc_scale;
  p.range_flag = lssvc::range_flag();
  p.dw_off = (p.w_bytes + 1023) & ~1023;
  p.in_off = (p.dw_off + (dw ? 10 * Cin * 4 : 0) + 1023) & ~1023;
  size_t smem = 0;
  for (p.in_bufs = MAX_IN_BUFS; p.in_bufs >= 2; --p.in_bufs) {
    p.stage_off = p.in_off + p.in_bufs * p.in_bytes;
    smem = static_cast<size_t>(p.stage_off) + 128 * Cout * 4 + 1024;
    if (smem <= 226 * 1024) break;
  }
  LSSVC_REQUIRE(p.in_bufs >= 2, "conv_pw: Cin=%d Cout=%d dw=%d needs %zu bytes of shared memory", Cin, Cout, dw ? 1 : 0, smem);

  const cuuint32_t ones[3] = {1, 1, 1};
  auto make_map = [&](CUtensorMap *m, const lssvc_view &v, int slab_w, int bw, int bh) -> CUresult {
    const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(v.W), static_cast<cuuint64_t>(v.H)};
    const cuuint64_t strides[2] = {px, px * v.W};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(slab_w), static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh)};
    return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    slab_w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = make_map(&p.in_map, f->in, p.in_slab_w, p.in_w, dw ? TILE_H + 2 : TILE_H);
  if (r == CUDA_SUCCESS) r = make_map(&p.out_map, f->out, p.out_slab_w, TILE_W, TILE_H);
  if (r == CUDA_SUCCESS && f->res1.ptr) {
    r = make_map(&p.res_map, f->res1, p.out_slab_w, TILE_W, TILE_H);
    p.res_tma = 1;
  }
  if (r != CUDA_SUCCESS) {
    lssvc::set_error("conv_pw: cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
    return LSSVC_ERR_CUDA;
  }
  if (!g_attr_set) {
    LSSVC_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(conv_pw_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    226 * 1024));
    g_attr_set = true;
  }
  const int total_tiles = p.tiles_x * p.tiles_y;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  LSSVC_CUDA(lssvc::launch_pdl(conv_pw_kernel, grid, NUM_THREADS, smem, lssvc::as_stream(stream), p));
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
