// Implicit-GEMM convolution on the 5th-gen tensor cores, split-fp16 arithmetic, BOTH operands read from shared memory.
//
//   D[M = 128 output pixels (16 rows x 8 columns), N = output channels] += A[M, K] * W[K, N],   K = (tap, input channel)
//
// Split-fp16 arithmetic (x = x_hi + x_lo in fp16; D1 += A_hi*W_hi, D2 += A_hi*W_lo + A_lo*W_hi in separate fp32
// TMEM accumulators, [W_hi | W_lo] stacked so that A_hi feeds one MMA of width 2N).  How the A operand reaches the tensor
// core:
//
//   * per (source, KC-channel chunk, stride-parity plane) ONE fp32 halo tile ((16 + kh - 1) x (8*MT + kw - 1) pixels, one
//     pixel = one 4*KC-byte row) is fetched by a 5-D TMA box load with the 128-byte (KC = 32) / 64-byte (KC = 16) swizzle;
//   * 8 converter warps rewrite it IN PLACE, once: pixel row fp32 x KC  ->  [fp16 hi x KC | fp16 lo x KC], same swizzle;
//   * the converted halo IS a K-major swizzled UMMA operand whose "rows" are pixels: 8 consecutive pixels of an image row
//     form one 8-row core group, the next image row of the patch is the next group, SBO = halo row pitch.  The A tile of
//     filter tap (r, s) is the same buffer with the descriptor start address advanced by (r * halo_w + s) pixel rows (the
//     swizzle is a function of the absolute shared-memory address, so a shifted start still reads what TMA wrote); the
//     K = 16 slices of A_hi / A_lo are byte offsets 0, 32 / 2*KC, 2*KC + 32 inside the pixel row.
//
// So a tap costs no data movement at all besides the MMA's own operand reads: no per-tap copies, no TMEM operand
// staging, no splitter warps (the earlier TMEM-operand kernel, tools/engines/conv_h2.cu, spent most of its issue slots and
// shared-memory bandwidth there).
// MT = 2 sub-tiles (16 x 16 pixels) share every weight stage, which halves the weight traffic L2 -> shared memory.
// With several channel tiles and a halo ring that holds a whole pixel tile, the converted halos stay resident while every
// channel tile is computed from them (A-resident mode, HsParams::a_res).
//
// Warp roles (640 threads): 0 halo TMA producer, 1 MMA issuer, 2 TMEM allocator + weight TMA producer, 4..11 halo
// converters, 12..19 epilogue (two independent sets of 4 warps; a set owns whole 128-pixel accumulators: sub-tile j of
// every tile when MT = 2, every other tile when MT = 1): TMEM -> registers -> bias / GDN / activation / residuals ->
// swizzled staging -> TMA store (or direct float4 stores).  Persistent over tiles; 2 or 4 accumulator slots in TMEM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "entropy_math.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_H = 16;
constexpr int SUB_W = 8;
constexpr int MAX_SLOTS = 8;   // ring of weight stages
constexpr int MAX_HALO = 6;    // ring of halo tiles
constexpr int MAX_TAPS = 49;
constexpr int MAX_GROUPS = 4;  // stride-2: one halo per input parity plane
constexpr int STAGE_K = 64;    // input channels x taps of one weight stage: KC = 32 -> 2 taps, KC = 16 -> 4 taps
constexpr int NUM_THREADS = 640;
constexpr int TMEM_COLS = 512;
constexpr int MAX_ACC = 4;
constexpr int MAX_ENTRIES = 32;  // weight stages per chunk (7x7 stride 2: 25)

struct alignas(64) HsParams {
  CUtensorMap a_map[LSSVC_MAX_SRC];
  CUtensorMap b_map;
  CUtensorMap out_map, out2_map;  // TMA-store epilogue (use_tma)
  CUtensorMap res1_map, res2_map;  // residual tiles TMA-loaded into the staging buffers ahead of the accumulator (res_tma bits 0, 1)
  int res_tma, res_tx;  // res_tx: number of residual tensors fetched by TMA (0..2)
  int use_tma, slab_w, n_slabs, stage_off, stage2_delta, stage_stride;  // staging per epilogue set: [out, out2][n_slabs][128 px][slab_w]
  int n_src;
  int chunks[LSSVC_MAX_SRC];  // ceil(C / KC) per source
  int coff[LSSVC_MAX_SRC];    // first packed input channel of the source
  int n_groups;
  int g_px[MAX_GROUPS], g_py[MAX_GROUPS];
  int g_tap0[MAX_GROUPS + 1];  // taps of group g: [g_tap0[g], g_tap0[g + 1])
  int g_step[MAX_GROUPS];      // taps per stage in group g
  int q0x, q0y;                // halo origin relative to the patch origin (plane coordinates)
  int halo_w;                  // halo width in pixels
  int halo_tx;                 // bytes of one halo box
  int halo_rows;               // pixels of one halo box
  unsigned char tap_w[MAX_TAPS];     // weight tap index r * kw + s
  unsigned short tap16[MAX_TAPS + 7];  // (byte offset of the tap's window origin inside the halo) >> 4, zero padded
  // one entry per weight stage of a chunk (identical for every chunk): x = tap16[0] | tap16[1] << 16, y = tap16[2] | tap16[3] << 16,
  // z = items | FIRST_OF_HALO << 8 | LAST_OF_HALO << 9, w = weight tap indices (4 x u8).  Copied to shared memory at start so
  // that the single-thread producers walk it with one LDS per stage instead of chains of parameter-space loads.
  uint4 stage_tab[MAX_ENTRIES];
  int n_entries, total_chunks;
  // Work decomposition.  Normal: work item = one (channel tile, pixel tile) pair, `outer` = all of them, `inner` = 1.
  // A-resident (a_res): work item = one PIXEL tile, whose converted halo tiles stay in shared memory while `inner` = n_tiles
  // channel tiles are computed from them one after the other — the input is fetched and converted once instead of n_tiles times.
  int a_res, outer, inner;
  int mt;                      // sub-tiles per tile (1 or 2)
  int n_acc;                   // accumulator slots in TMEM (2 or 4), each 2 * n_tile columns
  int Ho, Wo;
  int tiles_x, tiles_y, n_tiles, n_tile, cout;
  int slots, halo_bufs, halo_bytes, b_bytes;
  int in_transform;
  float in_slope;
  float acc_scale;
  const float *bias;
  int epi;
  const float *gdn_x;
  int gdn_pitch;
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
  unsigned int *range_flag;  // raised when a split operand leaves the fp16 range (range.cu)
  // entropy epilogue (LSSVC_EPI_LAPLACE / LSSVC_EPI_BITPARM): see epilogue_entropy
  const float *ent_y;
  int ent_y_pitch;
  float *ent_y_hat;
  int ent_h_pitch;
  const float *ent_coef;
  double *ent_bits;
  int *ent_sym, *ent_index;
  const float *ent_thr;
  int ent_n_thr;
  int ent_step;
  long long *prof;  // DBG variant: [5 roles][8] cycle counters of CTA 0
  int dbg;  // LSSVC_HS_DBG: bottleneck-isolation switches (results are wrong when non-zero); see lssvc_conv_hs
};

__device__ __forceinline__ long long out_offset(const HsParams &p, int oy, int ox, int ch, int P) {
  if (!p.pixel_shuffle) return (static_cast<long long>(oy) * p.Wo + ox) * P + ch;
  const int cq = p.cout >> 2;
  const int sub = ch / cq;
  const int c = ch - sub * cq;
  const int i = sub >> 1, j = sub & 1;
  return (static_cast<long long>(2 * oy + i) * (2 * p.Wo) + (2 * ox + j)) * P + c;
}

// x = hi + lo, both fp16 (packed two values per register, even channel in the low half)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

template <int NV>
__device__ __forceinline__ void transform_row(float4 (&v)[NV], int in_transform, float in_slope) {
  if (in_transform == LSSVC_IN_SQUARE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x *= v[i].x; v[i].y *= v[i].y; v[i].z *= v[i].z; v[i].w *= v[i].w;
    }
  } else if (in_transform == LSSVC_IN_LRELU) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = v[i].x > 0.f ? v[i].x : v[i].x * in_slope;
      v[i].y = v[i].y > 0.f ? v[i].y : v[i].y * in_slope;
      v[i].z = v[i].z > 0.f ? v[i].z : v[i].z * in_slope;
      v[i].w = v[i].w > 0.f ? v[i].w : v[i].w * in_slope;
    }
  }
}

// Epilogue of one 128-pixel accumulator on the common path (plain epilogue, TMA store, residuals — if any — already in the
// staging buffers): no per-element branches.  LeakyReLU is max(v, slope * v) (exact for 0 <= slope <= 1; slope = 1 when
// the layer has no activation), the output scale is always applied (x 1.0 is exact), so the only compile-time variants
// are the two residual sources.  ~35 instructions per float4 instead of the ~120 of the general path below, which is
// what bounds the HBM-bound 1x1 layers (the epilogue warps are instruction-latency bound).
// EPI = LSSVC_EPI_GDN / IGDN: v = x * rsqrt(v) / x * sqrt(v) with x (gdn_x) read from global memory, row pointer gx_row
// (clamped to a valid pixel for the overhang of border tiles: those rows are clipped by the TMA store).
// R2: 0 no second residual, 1 staged in `stage2` by TMA, 2 read from global memory (row pointer r2_row, clamped like gx_row;
// needs cout % 16 == 0) — the case of a 64-channel 16 x 16 tile, where a second staging buffer does not fit next to the halo
// ring and the general path below cost twice the time of the layer (3x3 64->64 + two residuals: 0.24 ms against 0.13).
template <bool R1, int R2, int EPI>
__device__ __forceinline__ void epilogue_lean(uint32_t t_row, int n_tile, int n0, int cout, const float *__restrict__ bias,
                                              float acc_scale, float slope, float out_scale, uint32_t stage, uint32_t stage2,
                                              uint32_t slab_w, int m, const float *__restrict__ gx_row,
                                              const float *__restrict__ r2_row = nullptr) {
  const uint32_t sswz = (slab_w == 32 ? static_cast<uint32_t>(m & 7) : static_cast<uint32_t>((m >> 1) & 3)) << 4;
  const uint32_t row_b = static_cast<uint32_t>(m) * (slab_w * 4u);
  for (int n = 0; n < n_tile; n += 16) {
    const int cg = n0 + n;
    if (cg >= cout) break;
    uint32_t r1[16], r2[16];
    ptx::tmem_ld16(t_row + static_cast<uint32_t>(n), r1);
    ptx::tmem_ld16(t_row + static_cast<uint32_t>(n_tile + n), r2);
    float4 b4v[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b4v[g] = __ldg(reinterpret_cast<const float4 *>(bias + cg) + g);  // bias is padded to n_pad
    float4 gxv[4];
    if (EPI != LSSVC_EPI_PLAIN) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gxv[g] = __ldg(reinterpret_cast<const float4 *>(gx_row + cg) + g);  // cout % 16 == 0 on this path
    }
    const uint32_t srow = static_cast<uint32_t>(n / static_cast<int>(slab_w)) * (128u * slab_w * 4u) + row_b;
    const uint32_t spiece = static_cast<uint32_t>(n % static_cast<int>(slab_w)) << 2;
    float4 a1[4], a2[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t soff = srow + ((spiece + 16u * g) ^ sswz);
      if (R1) a1[g] = ptx::lds_f4(stage + soff);
      if (R2 == 1) a2[g] = ptx::lds_f4(stage2 + soff);
      if (R2 == 2) a2[g] = __ldg(reinterpret_cast<const float4 *>(r2_row + cg) + g);
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[4];
      v[0] = (__uint_as_float(r1[4 * g + 0]) + __uint_as_float(r2[4 * g + 0])) * acc_scale + b4v[g].x;
      v[1] = (__uint_as_float(r1[4 * g + 1]) + __uint_as_float(r2[4 * g + 1])) * acc_scale + b4v[g].y;
      v[2] = (__uint_as_float(r1[4 * g + 2]) + __uint_as_float(r2[4 * g + 2])) * acc_scale + b4v[g].z;
      v[3] = (__uint_as_float(r1[4 * g + 3]) + __uint_as_float(r2[4 * g + 3])) * acc_scale + b4v[g].w;
      if (EPI == LSSVC_EPI_GDN) {
        v[0] = gxv[g].x * rsqrtf(v[0]); v[1] = gxv[g].y * rsqrtf(v[1]);
        v[2] = gxv[g].z * rsqrtf(v[2]); v[3] = gxv[g].w * rsqrtf(v[3]);
      } else if (EPI == LSSVC_EPI_IGDN) {
        v[0] = gxv[g].x * sqrtf(v[0]); v[1] = gxv[g].y * sqrtf(v[1]);
        v[2] = gxv[g].z * sqrtf(v[2]); v[3] = gxv[g].w * sqrtf(v[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], v[e] * slope) * out_scale;
      // (a residual read from global memory is added first: the order of the general path, bit for bit)
      if (R2 == 2) { v[0] += a2[g].x; v[1] += a2[g].y; v[2] += a2[g].z; v[3] += a2[g].w; }
      if (R1) { v[0] += a1[g].x; v[1] += a1[g].y; v[2] += a1[g].z; v[3] += a1[g].w; }
      if (R2 == 1) { v[0] += a2[g].x; v[1] += a2[g].y; v[2] += a2[g].z; v[3] += a2[g].w; }
      const uint32_t soff = srow + ((spiece + 16u * g) ^ sswz);
      ptx::sts_u4(stage + soff, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
  }
}

// Entropy epilogue: the convolution that produces the entropy parameters codes the latent while the parameters are still in
// registers (direct global stores, no staging).  The arithmetic is the stand-alone kernels' own (entropy_math.cuh), applied to
// the very values the plain epilogue would have stored ((D1 + D2) * acc_scale + bias, LeakyReLU, + res1), so symbols, CDF rows,
// quantised latents and per-element bits are bit-identical to conv + lssvc_laplace_quant / lssvc_four_part_step /
// lssvc_bitparm_quant; only the order in which the per-element bits are summed (warp shuffles, one double atomicAdd per warp)
// differs.
//   LAPLACE:  a channel tile holds [scale of latent channels c0 .. c0 + Ct | their means] (one tile: the natural (scale | mean)
//             order; several tiles: interleaved at pack time, lssvc_conv::ent_tile); both go to `out` in natural order;
//             q = rint(y - mean), y_hat = q + mean, bits, symbol / CDF-row dumps (NCHW int32).
//   FOURPART: the same for coding step p.ent_step of the four-part prior: a 16-channel chunk lies in one channel quarter and is
//             coded only at the pixels of that step's parity; step 0 zeroes y_hat elsewhere.
//   BITPARM:  accumulator = z: out = rint(z), bits from the per-channel BitEstimator coefficients, symbol dump.
// (__noinline__: its register needs — two 16-channel parameter chunks live at once — stay out of the kernel's 96-register budget)
template <int MODE>
__device__ __noinline__ void epilogue_entropy(const HsParams &p, uint32_t t_row, int n_tile, int n0, bool valid, long long pix) {
  const int cout = p.cout;
  const float acc_scale = p.acc_scale, slope = p.slope;
  const bool has_act = p.act != 0;
  const long long HW = static_cast<long long>(p.Ho) * p.Wo;
  const float *const res_row = (p.res1 && valid) ? p.res1 + pix * p.res1_pitch : nullptr;
  double local = 0.0;
  // the plain epilogue's value for accumulator columns col .. col + 15 of this channel tile (packed channels n0 + col ..);
  // nat = the natural (reference) index of the first of these channels, where the residual lives
  auto load16 = [&](int col, int nat, float (&v)[16]) {
    uint32_t r1[16], r2[16];
    ptx::tmem_ld16(t_row + static_cast<uint32_t>(col), r1);
    ptx::tmem_ld16(t_row + static_cast<uint32_t>(n_tile + col), r2);
    float4 rv[4];
    if (res_row) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rv[g] = __ldg(reinterpret_cast<const float4 *>(res_row + nat) + g);
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float t = (__uint_as_float(r1[e]) + __uint_as_float(r2[e])) * acc_scale + __ldg(p.bias + n0 + col + e);
      if (has_act) t = t > 0.f ? t : t * slope;
      v[e] = t;
    }
    if (res_row) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        v[4 * g] += rv[g].x; v[4 * g + 1] += rv[g].y; v[4 * g + 2] += rv[g].z; v[4 * g + 3] += rv[g].w;
      }
    }
  };
  if (MODE == LSSVC_EPI_BITPARM) {
    for (int n = 0; n < n_tile && n0 + n < cout; n += 16) {
      float z[16];
      load16(n, n0 + n, z);
      if (valid) {
        float *const o = p.out + pix * p.out_pitch + n0 + n;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float q[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = n0 + n + 4 * g + e;
            q[e] = rintf(z[4 * g + e]);
            local += static_cast<double>(lssvc_ent::bitparm_bits(q[e], p.ent_coef + c * 11));
            if (p.ent_sym) p.ent_sym[static_cast<long long>(c) * HW + pix] = static_cast<int>(q[e]);
          }
          reinterpret_cast<float4 *>(o)[g] = make_float4(q[0], q[1], q[2], q[3]);
        }
      }
    }
  } else {
    const int C = cout >> 1, Ct = n_tile >> 1, c0 = n0 >> 1, cq = C >> 2;
    const int step = MODE == LSSVC_EPI_FOURPART ? p.ent_step : -1;
    int parity = 0;
    if (MODE == LSSVC_EPI_FOURPART && valid) {
      const int oy = static_cast<int>(pix / p.Wo), ox = static_cast<int>(pix - static_cast<long long>(oy) * p.Wo);
      parity = ((oy & 1) << 1) | (ox & 1);
    }
    for (int n = 0; n < Ct; n += 16) {
      float sc[16], mu[16];
      load16(n, c0 + n, sc);
      load16(Ct + n, C + c0 + n, mu);
      if (valid) {
        float *const o = p.out + pix * p.out_pitch + c0;
        float4 *const yh = reinterpret_cast<float4 *>(p.ent_y_hat + pix * p.ent_h_pitch + c0 + n);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          reinterpret_cast<float4 *>(o + n)[g] = make_float4(sc[4 * g], sc[4 * g + 1], sc[4 * g + 2], sc[4 * g + 3]);
          reinterpret_cast<float4 *>(o + C + n)[g] = make_float4(mu[4 * g], mu[4 * g + 1], mu[4 * g + 2], mu[4 * g + 3]);
        }
        const int quarter = MODE == LSSVC_EPI_FOURPART ? (c0 + n) / cq : 0;   // (a 16-channel chunk never straddles two quarters)
        const bool active = MODE != LSSVC_EPI_FOURPART || lssvc_ent::four_part_mask(step, quarter) == parity;
        if (active) {
          const float4 *const yq = reinterpret_cast<const float4 *>(p.ent_y + pix * p.ent_y_pitch + c0 + n);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 y4 = __ldg(yq + g);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
            float h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c0 + n + 4 * g + e;
              const float m = mu[4 * g + e], sv = sc[4 * g + e];
              const float q = rintf(yv[e] - m);
              h[e] = q + m;
              local += static_cast<double>(lssvc_ent::laplace_bits(q, sv));
              const long long nchw = static_cast<long long>(c - quarter * cq) * HW + pix;   // (FOURPART dumps hold one quarter)
              if (p.ent_sym) p.ent_sym[nchw] = static_cast<int>(q);
              if (p.ent_index) p.ent_index[nchw] = lssvc_ent::scale_index(sv, p.ent_thr, p.ent_n_thr);
            }
            yh[g] = make_float4(h[0], h[1], h[2], h[3]);
          }
        } else if (step == 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) yh[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  if (p.ent_bits) {
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(p.ent_bits, local);
  }
}

// DBG = true: the instrumented variant (LSSVC_HS_DBG switches + per-role wait-time counters), never on the product path
template <int KC, int MT, bool DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_hs_kernel(const __grid_constant__ HsParams p) {
  ptx::pdl_launch_dependents();  // the next kernel of the stream may take SMs as they fall idle (it waits in its own pdl_wait)
  const int dbgf = DBG ? p.dbg : 0;
  long long prof[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = DBG ? clock64() : 0;
#define HS_WAIT(slot, bar, parity)                    \
  do {                                                \
    if (DBG && (dbgf & 2048)) {                       \
      ptx::mbar_spin(bar, parity);                    \
    } else if (DBG && (dbgf & 64)) {                  \
      const long long t0__ = clock64();               \
      ptx::mbar_wait(bar, parity);                    \
      prof[slot] += clock64() - t0__;                 \
    } else {                                          \
      ptx::mbar_wait(bar, parity);                    \
    }                                                 \
  } while (0)
  constexpr int ROWB = KC * 4;         // bytes of one halo pixel (fp32, or fp16 hi | fp16 lo after conversion)
  constexpr int NV = KC / 4;           // 16-byte chunks per halo pixel
  constexpr int KS = KC / 16;          // K = 16 MMA slices per tap
  constexpr int TAPS_PER_STAGE = STAGE_K / KC;
  constexpr uint32_t LO_OFF = KC * 2;  // byte offset of the lo half inside a pixel row
  constexpr uint32_t A_LAYOUT = KC == 32 ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t B_ROWB = KC * 2;  // bytes of one weight row (fp16)
  constexpr uint32_t B_LAYOUT = KC == 32 ? 4u : 6u;  // SWIZZLE_64B : SWIZZLE_32B
  constexpr uint32_t B_SBO = 8 * B_ROWB;
  constexpr uint32_t SWZ = KC == 32 ? 0x70u : 0x30u;  // 16-byte-chunk XOR of the TMA swizzle: address bits 7.. -> bits 4..

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_s[MAX_SLOTS];
  __shared__ uint64_t empty_s[MAX_SLOTS];
  __shared__ uint64_t halo_full[MAX_HALO];
  __shared__ uint64_t halo_empty[MAX_HALO];
  __shared__ uint64_t halo_conv[MAX_HALO];
  __shared__ uint64_t tfull_bar[MAX_ACC];
  __shared__ uint64_t tempty_bar[MAX_ACC];
  __shared__ uint64_t res_full[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint4 stage_tab_s[MAX_ENTRIES];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::pin((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  // layout: [halo_bufs x halo_bytes][slots x b_bytes][2 sets x staging]
  const uint32_t b_base = smem_base + static_cast<uint32_t>(p.halo_bufs * p.halo_bytes);
  const uint32_t bar_full = ptx::pin(ptx::smem_u32(full_s));
  const uint32_t bar_empty = ptx::pin(ptx::smem_u32(empty_s));
  const uint32_t bar_halo_full = ptx::pin(ptx::smem_u32(halo_full)), bar_halo_empty = ptx::pin(ptx::smem_u32(halo_empty));
  const uint32_t bar_halo_conv = ptx::pin(ptx::smem_u32(halo_conv));
  const uint32_t bar_tfull = ptx::pin(ptx::smem_u32(tfull_bar)), bar_tempty = ptx::pin(ptx::smem_u32(tempty_bar));
  const uint32_t bar_res = ptx::pin(ptx::smem_u32(res_full));

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.a_map[j]);
    ptx::prefetch_tensormap(&p.b_map);
    if (p.res_tma & 1) ptx::prefetch_tensormap(&p.res1_map);
    if (p.res_tma & 2) ptx::prefetch_tensormap(&p.res2_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.slots; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int h = 0; h < p.halo_bufs; ++h) {
      ptx::mbar_init(bar_halo_full + 8 * h, 1);
      ptx::mbar_init(bar_halo_empty + 8 * h, 1);
      ptx::mbar_init(bar_halo_conv + 8 * h, 8);
    }
    for (int b = 0; b < MAX_ACC; ++b) {
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tempty + 8 * b, 4);
    }
    ptx::mbar_init(bar_res, 1);
    ptx::mbar_init(bar_res + 8, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && lane < p.n_entries) stage_tab_s[lane] = p.stage_tab[lane];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // everything above touched only parameters, shared memory and TMEM; from here on global memory is read and written
  ptx::pdl_wait();

  const int tiles_per_n = p.tiles_x * p.tiles_y;
  const int outer = p.outer, inner = p.inner;
  const bool a_res = p.a_res != 0;
  constexpr int mt = MT;
  constexpr int tile_w = SUB_W * MT;
  // work item `item`, inner step `ni`  ->  channel tile nt, pixel tile rem
#define HS_DECODE(item, ni, nt, rem)                 \
  const int nt = a_res ? (ni) : (item) / tiles_per_n; \
  const int rem = a_res ? (item) : (item) - nt * tiles_per_n

  if (warp == 0) {
    // ------------------------------- halo TMA producer -----------------------------------
    if (ptx::elect_one()) {
      int hb = 0;
      uint32_t hph = 0;
      for (int item = blockIdx.x; item < outer; item += gridDim.x) {
        const int rem = a_res ? item : item % tiles_per_n;      // (A-resident: the halos are loaded once per pixel tile)
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int oy0 = ty * TILE_H + p.q0y, ox0 = tx * tile_w + p.q0x;
        for (int j = 0; j < p.n_src; ++j) {
          for (int c = 0; c < p.chunks[j]; ++c) {
            for (int g = 0; g < p.n_groups; ++g) {
              const uint32_t full = bar_halo_full + 8 * hb;
              HS_WAIT(0, bar_halo_empty + 8 * hb, hph ^ 1u);
              if (dbgf & 16) {
                ptx::mbar_arrive(full);
              } else {
                ptx::mbar_expect_tx(full, static_cast<uint32_t>(p.halo_tx));
                ptx::tma_load_5d(smem_base + static_cast<uint32_t>(hb * p.halo_bytes), &p.a_map[j], full, c * KC, p.g_px[g],
                                 ox0, p.g_py[g], oy0);
              }
              if (++hb == p.halo_bufs) {
                hb = 0;
                hph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------- weight TMA producer ---------------------------------
    // A stage is up to STAGE_K / KC consecutive taps of one halo group: their weight tiles [2][n_tile][KC] land back
    // to back in the stage's slot and complete on one barrier.
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tile_bytes = static_cast<uint32_t>(2 * p.n_tile) * B_ROWB;
      const int n_entries = p.n_entries;
      const uint32_t tab = ptx::smem_u32(stage_tab_s);
      for (int item = blockIdx.x; item < outer; item += gridDim.x)
      for (int ni = 0; ni < inner; ++ni) {
        const int n0 = (a_res ? ni : item / tiles_per_n) * p.n_tile;
        for (int j = 0; j < p.n_src; ++j) {
          for (int c = 0; c < p.chunks[j]; ++c) {
            const int k0 = p.coff[j] + c * KC;
            for (int e = 0; e < n_entries; ++e) {
              uint32_t e_x, e_y, e_z, e_w;
              ptx::lds_u4(tab + 16u * static_cast<uint32_t>(e), e_x, e_y, e_z, e_w);
              const int items = static_cast<int>(e_z & 0xffu);
              const uint32_t full = bar_full + 8 * s;
              HS_WAIT(0, bar_empty + 8 * s, ph ^ 1u);
              if (dbgf & 32) {
                ptx::mbar_arrive(full);
              } else {
                ptx::mbar_expect_tx(full, tile_bytes * static_cast<uint32_t>(items));
#pragma unroll
                for (int i = 0; i < TAPS_PER_STAGE; ++i)
                  if (i < items)
                    ptx::tma_load_4d(b_base + static_cast<uint32_t>(s * p.b_bytes) + static_cast<uint32_t>(i) * tile_bytes, &p.b_map,
                                     full, k0, n0, 0, static_cast<int>((e_w >> (8 * i)) & 0xffu));
              }
              if (++s == p.slots) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ------------------------------------------
    // One elected thread runs the whole loop.  It is a serial instruction stream (every dependent instruction costs
    // ~5 cycles), so the per-MMA work is kept to a couple of 32-bit adds: descriptors are (lo, hi) word pairs whose
    // hi word is constant and whose lo word is (address >> 4) + constant; tap offsets come pre-shifted from the
    // parameter block and are fetched before the stage's barrier wait.
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      int hb = 0;
      uint32_t hph = 0;
      int slot0 = 0;  // first accumulator slot of the current tile
      uint32_t acc_ph = 0;
      const uint32_t n_tile = static_cast<uint32_t>(p.n_tile);
      const uint32_t idesc_2n = ptx::make_idesc_f16_m128(2u * n_tile);
      const uint32_t idesc_n = ptx::make_idesc_f16_m128(n_tile);
      const int slots = p.slots, halo_bufs = p.halo_bufs, n_acc = p.n_acc;
      const uint32_t b16_stage = static_cast<uint32_t>(p.b_bytes) >> 4;
      const uint32_t halo16 = static_cast<uint32_t>(p.halo_bytes) >> 4;
      const uint32_t tile16 = (2u * n_tile * B_ROWB) >> 4;
      constexpr uint32_t DESC_LO = 1u << 16;  // LBO field = 1 (unused for swizzled K-major)
      // DBG 16384: SBO forced to 2048 B (8-row groups on 1024-byte swizzle-atom boundaries; reads the wrong pixels)
      const uint32_t a_hi_word = ((DBG && (dbgf & 16384)) ? (2048u >> 4) : ((static_cast<uint32_t>(p.halo_w) * ROWB) >> 4)) | (1u << 14) | (A_LAYOUT << 29);
      constexpr uint32_t b_hi_word = (B_SBO >> 4) | (1u << 14) | (B_LAYOUT << 29);
      const uint32_t a16_base = (smem_base >> 4) | DESC_LO, b16_base = (b_base >> 4) | DESC_LO;
      constexpr uint32_t SUB16 = (SUB_W * ROWB) >> 4, LO16 = LO_OFF >> 4;
      const int n_entries = p.n_entries, total_chunks = p.total_chunks;
      const uint32_t tab = ptx::smem_u32(stage_tab_s);
      uint32_t e_x, e_y, e_z, e_w;
      ptx::lds_u4(tab, e_x, e_y, e_z, e_w);
      int hb0 = 0;  // A-resident: first halo buffer of the current pixel tile
      for (int item = blockIdx.x; item < outer; item += gridDim.x)
      for (int ni = 0; ni < inner; ++ni) {
        const bool first_pass = ni == 0, last_pass = ni == inner - 1;
        if (first_pass) hb0 = hb;
        int hb_r = hb0;  // halo buffer of the current walk when the halos are already resident (ni > 0)
#pragma unroll
        for (int j = 0; j < MT; ++j) HS_WAIT(0, bar_tempty + 8 * (slot0 + j), acc_ph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tile = tmem_base + static_cast<uint32_t>(slot0) * 2u * n_tile;
        uint32_t acc = 0;  // 0 only for the first K slice of the tile
        uint32_t a16_halo = 0;
        for (int ch = 0; ch < total_chunks; ++ch) {
          for (int e = 0; e < n_entries; ++e) {
            const uint32_t c_x = e_x, c_y = e_y, c_z = e_z;
            {  // next stage's entry: in flight while this stage's barrier wait and MMAs are issued
              const int en = e + 1 == n_entries ? 0 : e + 1;
              ptx::lds_u4(tab + 16u * static_cast<uint32_t>(en), e_x, e_y, e_z, e_w);
            }
            if (c_z & 0x100u) {  // first stage of a halo tile
              if (first_pass) {
                HS_WAIT(1, bar_halo_conv + 8 * hb, hph);
                a16_halo = a16_base + static_cast<uint32_t>(hb) * halo16;
              } else {  // converted by this pixel tile's first pass and still resident
                a16_halo = a16_base + static_cast<uint32_t>(hb_r) * halo16;
              }
            }
            const int items = static_cast<int>(c_z & 0xffu);
            uint32_t tap16[4] = {c_x & 0xffffu, c_x >> 16, c_y & 0xffffu, c_y >> 16};
            if (DBG && (dbgf & 8192)) {  // tap windows rounded down to 1024-byte swizzle atoms (wrong pixels: timing only)
#pragma unroll
              for (int i = 0; i < 4; ++i) tap16[i] &= ~63u;
            }
            HS_WAIT(2, bar_full + 8 * s, ph);
            ptx::tc_fence_after();
            const long long ti0 = (DBG && (dbgf & 64)) ? clock64() : 0;
            const uint32_t b16_s = b16_base + static_cast<uint32_t>(s) * b16_stage;
#pragma unroll
            for (int i = 0; i < TAPS_PER_STAGE; ++i) {
              if (i < items) {
                const uint32_t a16_tap = a16_halo + tap16[i];
                const uint32_t b16_tap = b16_s + static_cast<uint32_t>(i) * tile16;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                  // issue order hi(0), hi(1), lo(0), lo(1): consecutive MMAs share the weight operand and never
                  // target the accumulator the previous one wrote (measured 4-5 % faster than hi/lo per sub-tile)
                  // [D1 | D2] (+)= A_hi * [W_hi | W_lo]   then   D2 += A_lo * W_hi
#pragma unroll
                  for (int j = 0; j < MT; ++j) {
                    if (!DBG || !(dbgf & 1))
                      ptx::mma_f16_ss2(d_tile + static_cast<uint32_t>(j) * 2u * n_tile, a16_tap + j * SUB16 + ks * 2, a_hi_word,
                                       b16_tap + ks * 2, b_hi_word, (DBG && (dbgf & 256)) ? idesc_n : idesc_2n, acc);
                  }
#pragma unroll
                  for (int j = 0; j < MT; ++j) {
                    const uint32_t a16 = a16_tap + j * SUB16 + ks * 2;
                    const uint32_t d1 = d_tile + static_cast<uint32_t>(j) * 2u * n_tile;
                    if (!DBG || !(dbgf & 2))
                      ptx::mma_f16_ss2((DBG && (dbgf & 4096)) ? ((d1 + n_tile + 256u) & 0xffff01ffu) : d1 + n_tile,
                                       (DBG && (dbgf & 512)) ? a16 : a16 + LO16, a_hi_word, b16_tap + ks * 2, b_hi_word,
                                       (DBG && (dbgf & 1024)) ? idesc_2n : idesc_n, 1u);
                  }
                  acc = 1u;
                }
              }
            }
            ptx::mma_commit(bar_empty + 8 * s);
            if (DBG && (dbgf & 64)) prof[3] += clock64() - ti0;
            if (++s == slots) {
              s = 0;
              ph ^= 1u;
            }
            if (c_z & 0x200u) {  // last stage of the halo tile
              // it is free once every MMA that reads it has completed: after the LAST channel tile of the pixel tile
              if (last_pass) ptx::mma_commit(bar_halo_empty + 8 * (first_pass ? hb : hb_r));
              if (first_pass) {
                if (++hb == halo_bufs) {
                  hb = 0;
                  hph ^= 1u;
                }
              } else if (++hb_r == halo_bufs) {
                hb_r = 0;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < MT; ++j) ptx::mma_commit(bar_tfull + 8 * (slot0 + j));
        slot0 += MT;
        if (slot0 == n_acc) {
          slot0 = 0;
          acc_ph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------- converters: fp32 halo -> fp16 hi | lo, in place -------
    // One thread per halo pixel: the 4*KC-byte fp32 row becomes [hi fp16 x KC | lo fp16 x KC] with the same
    // 16-byte-chunk swizzle, i.e. a K-major UMMA operand row.
    const int ct = threadIdx.x - 128;  // 0..255
    const int halo_rows = p.halo_rows, halo_bufs = p.halo_bufs;
    const int in_transform = p.in_transform;
    const float in_slope = p.in_slope;
    const uint32_t halo_bytes = static_cast<uint32_t>(p.halo_bytes);
    int hb = 0;
    uint32_t hph = 0;
    float amax = 0.f;  // running max |operand| of this thread (range guard)
    for (int item = blockIdx.x; item < outer; item += gridDim.x) {      // (A-resident: once per pixel tile)
      for (int j = 0; j < p.n_src; ++j) {
        for (int c = 0; c < p.chunks[j]; ++c) {
          for (int g = 0; g < p.n_groups; ++g) {
            HS_WAIT(0, bar_halo_full + 8 * hb, hph);
            const long long tc0 = (DBG && (dbgf & 64)) ? clock64() : 0;
            const uint32_t halo = smem_base + static_cast<uint32_t>(hb) * halo_bytes;
            for (int r = ct; r < halo_rows && !(dbgf & 4); r += 256) {
              const uint32_t x = halo + static_cast<uint32_t>(r) * ROWB;
              const uint32_t a0 = x | ((x >> 3) & SWZ);
              float4 v[NV];
#pragma unroll
              for (int i = 0; i < NV; ++i) v[i] = ptx::lds_f4(a0 ^ (i << 4));
              transform_row<NV>(v, in_transform, in_slope);
              uint32_t hi[2 * NV], lo[2 * NV];
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
                split_pair(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
                split_pair(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
              }
#pragma unroll
              for (int i = 0; i < NV / 2; ++i) {
                ptx::sts_u4(a0 ^ (i << 4), hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                ptx::sts_u4(a0 ^ ((NV / 2 + i) << 4), lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
            }
            ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_halo_conv + 8 * hb);
            if (DBG && (dbgf & 64)) prof[1] += clock64() - tc0;
            if (++hb == halo_bufs) {
              hb = 0;
              hph ^= 1u;
            }
          }
        }
      }
    }
    if (amax >= lssvc::kSplitRangeLimit && p.range_flag) atomicOr(p.range_flag, 1u);
  } else if (warp >= 12) {
    // ------------------------------- epilogue ---------------------------------------------
    // two independent sets of 4 warps; set e owns the accumulator units u = tile_counter * mt + j with (u & 1) == e
    const int eset = (warp - 12) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int h = m / SUB_W, w = m % SUB_W;
    const int n_tile = p.n_tile, cout = p.cout, n_acc = p.n_acc, Wo = p.Wo;
    const float acc_scale = p.acc_scale, out_scale = p.out_scale, slope = p.slope, slope2 = p.slope2;
    const bool has_act = p.act != 0, fast = p.vec_ok != 0, ps = p.pixel_shuffle != 0;
    const int epi = p.epi;
    const int cq = cout >> 2;
    const bool chunk_uniform = !ps || (cq & 15) == 0;  // a 16-channel chunk never straddles two sub-pixels
    float *const out = p.out;
    float *const out2 = p.out2;
    const float *const res1 = p.res1;
    const float *const res2 = p.res2;
    const float *const gdn_x = p.gdn_x;
    const float *const bias = p.bias;
    const int out_pitch = p.out_pitch, out2_pitch = p.out2_pitch, res1_pitch = p.res1_pitch, res2_pitch = p.res2_pitch,
              gdn_pitch = p.gdn_pitch;
    const bool use_tma = p.use_tma != 0 && fast && chunk_uniform;
    const uint32_t slab_w = static_cast<uint32_t>(p.slab_w);
    const uint32_t stage = smem_base + static_cast<uint32_t>(p.stage_off) + static_cast<uint32_t>(eset * p.stage_stride);
    const uint32_t stage2 = stage + static_cast<uint32_t>(p.stage2_delta);
    const bool store_thread = q == 0 && lane == 0;
    // residual tiles fetched by TMA into the staging buffers (res1 -> out staging, res2 -> out2 staging) while the
    // unit's MMAs are still running: the accumulator is then added in place, no global-load latency in the epilogue
    const bool r1_tma = use_tma && (p.res_tma & 1), r2_tma = use_tma && (p.res_tma & 2);
    uint32_t res_ph = 0;
    // the common case runs the branch-free epilogue_lean: TMA store, plain epilogue, one output, residuals only via staging,
    // LeakyReLU slope within [0, 1] (max(v, slope v) form)
    const bool r2_glob = res2 != nullptr && !r2_tma && epi == LSSVC_EPI_PLAIN && (cout & 15) == 0 && !ps;  // lean with res2 from global
    const bool lean = use_tma && (epi == LSSVC_EPI_PLAIN || ((cout & 15) == 0 && !r2_tma && (gdn_pitch & 3) == 0)) && out2 == nullptr &&
                      (res1 == nullptr || r1_tma) && (res2 == nullptr || r2_tma || r2_glob) &&
                      !(r2_tma && !r1_tma) && (!has_act || (slope >= 0.f && slope <= 1.f)) && (cout & 3) == 0;
    int u = 0;  // running unit counter (all units, both sets)
    for (int item = blockIdx.x; item < outer; item += gridDim.x)
    for (int ni = 0; ni < inner; ++ni) {
      HS_DECODE(item, ni, nt, rem);
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      for (int j = 0; j < mt; ++j, ++u) {
        if ((u & 1) != eset) continue;
        const int slot = u % n_acc;
        const uint32_t acc_ph = static_cast<uint32_t>(u / n_acc) & 1u;
        const int oy = ty * TILE_H + h, ox = tx * tile_w + j * SUB_W + w, n0 = nt * n_tile;
        const bool valid = (oy < p.Ho) && (ox < Wo);
        const long long pix = static_cast<long long>(oy) * Wo + ox;
        if (use_tma) {
          // the staging tile is free once the TMA stores issued from it (this set's previous unit) have read it
          const long long ts0 = (DBG && (dbgf & 64)) ? clock64() : 0;
          if (store_thread) {
            ptx::bulk_wait_read_all();
            if (r1_tma || r2_tma) {
              const uint32_t bar = bar_res + 8 * eset;
              const int oy0 = ty * TILE_H, ox0 = tx * tile_w + j * SUB_W;
              int live = (cout - n0 + static_cast<int>(slab_w) - 1) / static_cast<int>(slab_w);  // slabs of this channel tile that hold real channels
              live = live < p.n_slabs ? live : p.n_slabs;
              ptx::mbar_expect_tx(bar, static_cast<uint32_t>(live) * (128u * slab_w * 4u) * static_cast<uint32_t>(p.res_tx));
              for (int k = 0; k < p.n_slabs; ++k) {
                const int pc = n0 + k * static_cast<int>(slab_w);
                if (pc >= cout) break;
                const uint32_t dst = static_cast<uint32_t>(k) * (128u * slab_w * 4u);
                if (r1_tma) ptx::tma_load_3d(stage + dst, &p.res1_map, bar, pc, ox0, oy0);
                if (r2_tma) ptx::tma_load_3d(stage2 + dst, &p.res2_map, bar, pc, ox0, oy0);
              }
            }
          }
          ptx::named_bar_sync(2 + eset, 128);
          if (DBG && (dbgf & 64)) prof[2] += clock64() - ts0;
        }
        HS_WAIT(0, bar_tfull + 8 * slot, acc_ph);
        if (r1_tma || r2_tma) {
          HS_WAIT(0, bar_res + 8 * eset, res_ph);
          res_ph ^= 1u;
        }
        const long long te0 = (DBG && (dbgf & 64)) ? clock64() : 0;
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slot * 2 * n_tile);
        if (lean && !(dbgf & 8)) {
          const float sl = has_act ? slope : 1.f;
          if (epi != LSSVC_EPI_PLAIN) {
            // gdn_x row of this thread's pixel, clamped into the image for the overhang of border tiles
            const int cy = oy < p.Ho ? oy : p.Ho - 1, cx = ox < Wo ? ox : Wo - 1;
            const float *gx_row = gdn_x + (static_cast<long long>(cy) * Wo + cx) * gdn_pitch;
            if (epi == LSSVC_EPI_GDN) {
              if (r1_tma) epilogue_lean<true, 0, LSSVC_EPI_GDN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, gx_row);
              else epilogue_lean<false, 0, LSSVC_EPI_GDN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, gx_row);
            } else {
              if (r1_tma) epilogue_lean<true, 0, LSSVC_EPI_IGDN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, gx_row);
              else epilogue_lean<false, 0, LSSVC_EPI_IGDN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, gx_row);
            }
          } else if (r2_glob) {
            const int cy = oy < p.Ho ? oy : p.Ho - 1, cx = ox < Wo ? ox : Wo - 1;  // overhang rows are clipped by the TMA store
            const float *r2_row = res2 + (static_cast<long long>(cy) * Wo + cx) * res2_pitch;
            if (r1_tma) epilogue_lean<true, 2, LSSVC_EPI_PLAIN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, nullptr, r2_row);
            else epilogue_lean<false, 2, LSSVC_EPI_PLAIN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, nullptr, r2_row);
          } else if (r1_tma && r2_tma) epilogue_lean<true, 1, LSSVC_EPI_PLAIN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, nullptr);
          else if (r1_tma) epilogue_lean<true, 0, LSSVC_EPI_PLAIN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, nullptr);
          else epilogue_lean<false, 0, LSSVC_EPI_PLAIN>(t_row, n_tile, n0, cout, bias, acc_scale, sl, out_scale, stage, stage2, slab_w, m, nullptr);
        }
        const bool ent = epi == LSSVC_EPI_LAPLACE || epi == LSSVC_EPI_BITPARM || epi == LSSVC_EPI_FOURPART;
        if (ent && !(dbgf & 8)) {
          if (epi == LSSVC_EPI_LAPLACE) epilogue_entropy<LSSVC_EPI_LAPLACE>(p, t_row, n_tile, n0, valid, pix);
          else if (epi == LSSVC_EPI_FOURPART) epilogue_entropy<LSSVC_EPI_FOURPART>(p, t_row, n_tile, n0, valid, pix);
          else epilogue_entropy<LSSVC_EPI_BITPARM>(p, t_row, n_tile, n0, valid, pix);
        }
        for (int n = 0; n < n_tile && !(dbgf & 8) && !lean && !ent; n += 16) {
          const int cg = n0 + n;
          uint32_t r1[16], r2[16];
          ptx::tmem_ld16(t_row + static_cast<uint32_t>(n), r1);
          ptx::tmem_ld16(t_row + static_cast<uint32_t>(n_tile + n), r2);
          if (fast && chunk_uniform) {
            // ---- the 16 channels of this chunk are contiguous in every tensor involved
            long long opix = pix;
            int c0 = cg;
            if (ps) {
              const int sub = (cg >= cq) + (cg >= 2 * cq) + (cg >= 3 * cq);
              c0 = cg - sub * cq;
              opix = static_cast<long long>(2 * oy + (sub >> 1)) * (2 * Wo) + (2 * ox + (sub & 1));
            }
            const bool live = (valid || use_tma) && cg < cout;
            float4 b4v[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
              b4v[g] = (live && cg + 4 * g < cout) ? __ldg(reinterpret_cast<const float4 *>(bias + cg) + g)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            // residuals that do not ride the staging buffers: all loads of the chunk in flight before the accumulator wait
            const float4 *const q1 = (live && res1 && valid && !r1_tma) ? reinterpret_cast<const float4 *>(res1 + opix * res1_pitch + c0) : nullptr;
            const float4 *const q2 = (live && res2 && valid && !r2_tma) ? reinterpret_cast<const float4 *>(res2 + opix * res2_pitch + c0) : nullptr;
            const float4 *const gq = (live && epi != LSSVC_EPI_PLAIN && valid) ? reinterpret_cast<const float4 *>(gdn_x + pix * gdn_pitch + cg) : nullptr;
            float4 q1v[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) q1v[g] = (q1 && cg + 4 * g < cout) ? __ldg(q1 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            ptx::tmem_ld_wait();
            if (live) {
              float4 *const o = reinterpret_cast<float4 *>(out + opix * out_pitch + c0);
              float4 *const o2 = out2 ? reinterpret_cast<float4 *>(out2 + opix * out2_pitch + c0) : nullptr;
              // staging row of this pixel for the slab holding channels n .. n+15 (TMA-store path)
              const uint32_t srow = static_cast<uint32_t>(n / slab_w) * (128u * slab_w * 4u) + static_cast<uint32_t>(m) * (slab_w * 4u);
              const uint32_t sswz = (slab_w == 32 ? static_cast<uint32_t>(m & 7) : static_cast<uint32_t>((m >> 1) & 3)) << 4;
              const uint32_t spiece = static_cast<uint32_t>(n % slab_w) << 2;  // byte offset of the chunk inside the row
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (cg + 4 * g < cout) {
                  const float4 b4 = b4v[g];
                  float v[4];
                  v[0] = (__uint_as_float(r1[4 * g + 0]) + __uint_as_float(r2[4 * g + 0])) * acc_scale + b4.x;
                  v[1] = (__uint_as_float(r1[4 * g + 1]) + __uint_as_float(r2[4 * g + 1])) * acc_scale + b4.y;
                  v[2] = (__uint_as_float(r1[4 * g + 2]) + __uint_as_float(r2[4 * g + 2])) * acc_scale + b4.z;
                  v[3] = (__uint_as_float(r1[4 * g + 3]) + __uint_as_float(r2[4 * g + 3])) * acc_scale + b4.w;
                  if (epi != LSSVC_EPI_PLAIN) {
                    const float4 gx = gq ? gq[g] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (epi == LSSVC_EPI_GDN) {
                      v[0] = gx.x * rsqrtf(v[0]); v[1] = gx.y * rsqrtf(v[1]);
                      v[2] = gx.z * rsqrtf(v[2]); v[3] = gx.w * rsqrtf(v[3]);
                    } else {
                      v[0] = gx.x * sqrtf(v[0]); v[1] = gx.y * sqrtf(v[1]);
                      v[2] = gx.z * sqrtf(v[2]); v[3] = gx.w * sqrtf(v[3]);
                    }
                  }
                  if (has_act) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
                  }
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] *= out_scale;
                  if (q1) {
                    const float4 t = q1v[g];
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                  }
                  if (q2) {
                    const float4 t = __ldg(q2 + g);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                  }
                  const uint32_t soff = srow + ((spiece + 16u * g) ^ sswz);
                  if (r1_tma) {
                    const float4 t = ptx::lds_f4(stage + soff);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                  }
                  if (r2_tma) {
                    const float4 t = ptx::lds_f4(stage2 + soff);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                  }
                  if (use_tma) {
                    ptx::sts_u4(stage + soff, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                  } else {
                    o[g] = make_float4(v[0], v[1], v[2], v[3]);
                  }
                  if (out2) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope2;
                    if (use_tma) {
                      ptx::sts_u4(stage2 + soff, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                    } else {
                      o2[g] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                  }
                }
              }
            }
          } else {
            ptx::tmem_ld_wait();
            if (valid && cg < cout) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int ch = cg + e;
                if (ch < cout) {
                  float v = (__uint_as_float(r1[e]) + __uint_as_float(r2[e])) * acc_scale + bias[ch];
                  if (epi == LSSVC_EPI_GDN) {
                    v = gdn_x[pix * gdn_pitch + ch] * rsqrtf(v);
                  } else if (epi == LSSVC_EPI_IGDN) {
                    v = gdn_x[pix * gdn_pitch + ch] * sqrtf(v);
                  }
                  if (has_act) v = v > 0.f ? v : v * slope;
                  v *= out_scale;
                  if (res1) v += res1[out_offset(p, oy, ox, ch, p.res1_pitch)];
                  if (res2) v += res2[out_offset(p, oy, ox, ch, p.res2_pitch)];
                  out[out_offset(p, oy, ox, ch, p.out_pitch)] = v;
                  if (out2) out2[out_offset(p, oy, ox, ch, p.out2_pitch)] = v > 0.f ? v : v * slope2;
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * slot);
        if (DBG && (dbgf & 64)) prof[1] += clock64() - te0;
        if (use_tma) {
          ptx::fence_proxy_async_smem();  // staging writes (generic proxy) -> visible to the TMA store (async proxy)
          ptx::named_bar_sync(4 + eset, 128);
          if (store_thread) {
            const int oy0 = ty * TILE_H, ox0 = tx * tile_w + j * SUB_W;
            for (int k = 0; k < p.n_slabs; ++k) {
              const int pc = n0 + k * static_cast<int>(slab_w);  // first packed channel of the slab
              if (pc >= cout) break;
              const uint32_t src = static_cast<uint32_t>(k) * (128u * slab_w * 4u);
              if (ps) {
                const int sub = pc / cq, c0 = pc - sub * cq;
                ptx::tma_store_5d(&p.out_map, stage + src, c0, sub & 1, ox0, sub >> 1, oy0);
                if (out2) ptx::tma_store_5d(&p.out2_map, stage2 + src, c0, sub & 1, ox0, sub >> 1, oy0);
              } else {
                ptx::tma_store_3d(&p.out_map, stage + src, pc, ox0, oy0);
                if (out2) ptx::tma_store_3d(&p.out2_map, stage2 + src, pc, ox0, oy0);
              }
            }
            ptx::bulk_commit();
          }
        }
      }
    }
    if (use_tma && store_thread) ptx::bulk_wait_all();
  }

  if (DBG && p.prof && blockIdx.x == 0 && lane == 0) {
    // rows: 0 halo producer, 1 MMA issuer, 2 weight producer, 3 converter warp 4, 4 epilogue warp 12, 5 epilogue warp 16
    const int row = warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 2 : warp == 4 ? 3 : warp == 12 ? 4 : warp == 16 ? 5 : -1;
    if (row >= 0) {
      for (int i = 0; i < 6; ++i) p.prof[row * 8 + i] = prof[i];
      p.prof[row * 8 + 7] = clock64() - t_begin;
    }
  }
#undef HS_WAIT
#undef HS_DECODE
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
bool g_attr_set[2] = {false, false};

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

int floor_div_h(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
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

extern "C" int32_t lssvc_conv_hs(const lssvc_conv *c, void *stream) {
  LSSVC_REQUIRE(c != nullptr, "conv_hs: null descriptor");
  LSSVC_REQUIRE(c->n_src >= 1 && c->n_src <= LSSVC_MAX_SRC, "conv_hs: n_src=%d", c->n_src);
  LSSVC_REQUIRE(c->stride == 1 || c->stride == 2, "conv_hs: stride %d", c->stride);
  LSSVC_REQUIRE(c->kh >= 1 && c->kw >= 1 && c->kh * c->kw <= MAX_TAPS, "conv_hs: kernel %dx%d", c->kh, c->kw);
  LSSVC_REQUIRE(c->weight_h2 != nullptr && (reinterpret_cast<uintptr_t>(c->weight_h2) & 15) == 0,
                "conv_hs: needs the split fp16 weights (weight_h2)");
  LSSVC_REQUIRE(c->acc_scale > 0.f, "conv_hs: acc_scale %g", static_cast<double>(c->acc_scale));

  const int Hin = c->src[0].H, Win = c->src[0].W;
  int kc = 32;
  int cin16 = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    LSSVC_REQUIRE(lssvc::view_ok(&v), "conv_hs: bad source view %d", j);
    LSSVC_REQUIRE(v.H == Hin && v.W == Win, "conv_hs: source %d is %dx%d, expected %dx%d", j, v.H, v.W, Hin, Win);
    LSSVC_REQUIRE(v.C % 4 == 0, "conv_hs: source %d has %d channels (need a multiple of 4)", j, v.C);
    LSSVC_REQUIRE(v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0,
                  "conv_hs: source %d is not 16-byte aligned", j);
    if (v.C % 32) kc = 16;
    cin16 += (v.C + 15) / 16 * 16;
  }
  LSSVC_REQUIRE(cin16 == c->cin_pad16, "conv_hs: cin_pad16 %d != %d from the sources", c->cin_pad16, cin16);
  const int st = c->stride;
  LSSVC_REQUIRE(Hin % st == 0 && Win % st == 0, "conv_hs: %dx%d not divisible by stride", Hin, Win);
  const int Ho = (Hin + 2 * c->pad - c->kh) / st + 1;
  const int Wo = (Win + 2 * c->pad - c->kw) / st + 1;
  const int ps = c->pixel_shuffle ? 2 : 1;
  LSSVC_REQUIRE(lssvc::view_ok(&c->out), "conv_hs: bad output view");
  LSSVC_REQUIRE(c->out.H == Ho * ps && c->out.W == Wo * ps, "conv_hs: output view %dx%d, expected %dx%d", c->out.H,
                c->out.W, Ho * ps, Wo * ps);
  LSSVC_REQUIRE(!c->pixel_shuffle || c->cout % 4 == 0, "conv_hs: pixel shuffle needs cout %% 4 == 0");
  const int c_store = c->pixel_shuffle ? c->cout / 4 : c->cout;
  LSSVC_REQUIRE(c->out.C == c_store, "conv_hs: output view has %d channels, expected %d", c->out.C, c_store);
  LSSVC_REQUIRE(c->n_pad % 16 == 0 && c->n_pad >= c->cout, "conv_hs: n_pad=%d cout=%d", c->n_pad, c->cout);
  LSSVC_REQUIRE((reinterpret_cast<uintptr_t>(c->bias) & 15) == 0, "conv_hs: bias not 16-byte aligned");

  // output-channel tiling: equal tiles of at most 128 channels
  int n_tile = c->n_pad;
  const char *nt_str = getenv("LSSVC_HS_NTILE");  // A/B switch: cap of the channel tile
  const int nt_cap = nt_str && atoi(nt_str) >= 16 ? atoi(nt_str) : 128;
  if (n_tile > nt_cap) {
    n_tile = nt_cap;
    while (n_tile >= 16 && (c->n_pad % n_tile)) n_tile -= 16;
    LSSVC_REQUIRE(n_tile >= 16, "conv_hs: cannot tile n_pad=%d", c->n_pad);
  }
  if (int rc = resolve_driver()) return rc;

  HsParams p;
  memset(&p, 0, sizeof(p));
  p.n_src = c->n_src;
  p.n_tiles = c->n_pad / n_tile;

  // sub-tiles per tile: 2 (weights shared by 256 pixels) when that still leaves every SM a tile
  const char *mt_str = getenv("LSSVC_HS_MT");  // A/B switch, read per call so tests can flip it
  const int mt_env = mt_str ? atoi(mt_str) : 0;
  // (wide channel tiles keep MT = 1: two 2 x 128-column accumulators per sub-tile would leave no double buffering)
  int mt = (n_tile <= 64 && lssvc::ceil_div(Wo, 2 * SUB_W) * lssvc::ceil_div(Ho, TILE_H) * p.n_tiles >= g_num_sms) ? 2 : 1;
  if (mt_env == 1 || mt_env == 2) mt = mt_env;
  p.mt = mt;
  p.n_acc = (mt == 2 && 8 * n_tile <= TMEM_COLS) ? 4 : 2;

  // ---- taps grouped by input parity plane ------------------------------------------------------
  int q0y = 1 << 20, q1y = -(1 << 20), q0x = 1 << 20, q1x = -(1 << 20);
  for (int r = 0; r < c->kh; ++r) {
    const int q = floor_div_h(r - c->pad, st);
    q0y = q < q0y ? q : q0y;
    q1y = q > q1y ? q : q1y;
  }
  for (int s = 0; s < c->kw; ++s) {
    const int q = floor_div_h(s - c->pad, st);
    q0x = q < q0x ? q : q0x;
    q1x = q > q1x ? q : q1x;
  }
  const int halo_h = TILE_H + (q1y - q0y), halo_w = SUB_W * mt + (q1x - q0x);
  LSSVC_REQUIRE(halo_w <= 256 && halo_h <= 256, "conv_hs: halo %dx%d", halo_h, halo_w);
  p.q0x = q0x; p.q0y = q0y; p.halo_w = halo_w;
  int n_taps = 0;
  p.n_groups = 0;
  for (int py = 0; py < st; ++py) {
    for (int px = 0; px < st; ++px) {
      const int first = n_taps;
      for (int r = 0; r < c->kh; ++r) {
        const int dy = r - c->pad, qy = floor_div_h(dy, st);
        if (dy - qy * st != py) continue;
        for (int s = 0; s < c->kw; ++s) {
          const int dx = s - c->pad, qx = floor_div_h(dx, st);
          if (dx - qx * st != px) continue;
          p.tap_w[n_taps] = static_cast<unsigned char>(r * c->kw + s);
          p.tap16[n_taps] = static_cast<unsigned short>((((qy - q0y) * halo_w + (qx - q0x)) * kc * 4) >> 4);
          ++n_taps;
        }
      }
      if (n_taps > first) {
        p.g_px[p.n_groups] = px;
        p.g_py[p.n_groups] = py;
        p.g_tap0[p.n_groups] = first;
        ++p.n_groups;
      }
    }
  }
  p.g_tap0[p.n_groups] = n_taps;
  LSSVC_REQUIRE(n_taps == c->kh * c->kw, "conv_hs: tap enumeration");

  // ---- tensor maps --------------------------------------------------------------------------------
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const int row_bytes = kc * 4;
  int coff = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    p.chunks[j] = (v.C + kc - 1) / kc;
    p.coff[j] = coff;
    coff += (v.C + 15) / 16 * 16;
    const cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(st),
                                static_cast<cuuint64_t>(Win / st), static_cast<cuuint64_t>(st),
                                static_cast<cuuint64_t>(Hin / st)};
    const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
    const cuuint64_t strides[4] = {px, px * st, px * Win, px * Win * st};
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(kc), 1, static_cast<cuuint32_t>(halo_w), 1,
                               static_cast<cuuint32_t>(halo_h)};
    CUresult r = g_encode(&p.a_map[j], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, v.ptr, dims, strides, box, ones,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_hs: cuTensorMapEncodeTiled(A%d) failed with %d (C=%d pitch=%d %dx%d stride=%d halo %dx%d)", j,
                       static_cast<int>(r), v.C, v.pitch, Hin, Win, st, halo_h, halo_w);
      return LSSVC_ERR_CUDA;
    }
  }
  {
    const int taps = c->kh * c->kw;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin16), static_cast<cuuint64_t>(c->n_pad), 2,
                                static_cast<cuuint64_t>(taps)};
    const cuuint64_t rb = static_cast<cuuint64_t>(cin16) * 2;
    const cuuint64_t strides[3] = {rb, rb * c->n_pad, rb * c->n_pad * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(n_tile), 2, 1};
    CUresult r = g_encode(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(c->weight_h2), dims, strides,
                          box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_hs: cuTensorMapEncodeTiled(B) failed with %d (cin16=%d n_pad=%d taps=%d)",
                       static_cast<int>(r), cin16, c->n_pad, taps);
      return LSSVC_ERR_CUDA;
    }
  }

  // ---- pipeline geometry --------------------------------------------------------------------------
  // build_stages(tps): the per-chunk table of weight stages with at most `tps` taps per stage; returns false on overflow
  auto build_stages = [&](int tps) -> bool {
  for (int g = 0; g < p.n_groups; ++g) {
    const int nt = p.g_tap0[g + 1] - p.g_tap0[g];
    const int n_st = (nt + tps - 1) / tps;
    p.g_step[g] = (nt + n_st - 1) / n_st;
  }
  {
    int ne = 0, total_chunks = 0;
    for (int j = 0; j < c->n_src; ++j) total_chunks += p.chunks[j];
    for (int g = 0; g < p.n_groups; ++g) {
      const int t1 = p.g_tap0[g + 1], step = p.g_step[g];
      for (int t = p.g_tap0[g]; t < t1; t += step) {
        const int items = t1 - t < step ? t1 - t : step;
        if (ne >= MAX_ENTRIES || items > 4) return false;
        uint32_t t16[4] = {0, 0, 0, 0}, tw = 0;
        for (int i = 0; i < items; ++i) {
          t16[i] = p.tap16[t + i];
          tw |= static_cast<uint32_t>(p.tap_w[t + i]) << (8 * i);
        }
        uint32_t flags = static_cast<uint32_t>(items);
        if (t == p.g_tap0[g]) flags |= 0x100u;
        if (t + step >= t1) flags |= 0x200u;
        p.stage_tab[ne++] = make_uint4(t16[0] | (t16[1] << 16), t16[2] | (t16[3] << 16), flags, tw);
      }
    }
    p.n_entries = ne;
    p.total_chunks = total_chunks;
  }
  // a weight stage holds the taps of ONE pipeline step: a 1x1 conv (one tap per step) needs half / a quarter of the room of a
  // 3x3 one, and the shared memory it gives back goes to the halo ring below — 1x1 layers are latency-bound on that ring
  // (one 32-channel chunk feeds only two MMA pairs per TMA round trip)
  int max_step = 1;
  for (int g = 0; g < p.n_groups; ++g) max_step = p.g_step[g] > max_step ? p.g_step[g] : max_step;
  p.b_bytes = max_step * (2 * n_tile * kc * 2);  // n_tile % 16 == 0 keeps every tile 1024-byte aligned
  return true;
  };
  LSSVC_REQUIRE(build_stages(STAGE_K / kc), "conv_hs: stage table overflow");
  p.Ho = Ho; p.Wo = Wo;
  p.tiles_x = lssvc::ceil_div(Wo, SUB_W * mt);
  p.tiles_y = lssvc::ceil_div(Ho, TILE_H);
  p.n_tile = n_tile;
  p.cout = c->cout;
  p.halo_tx = halo_w * halo_h * row_bytes;
  p.halo_rows = halo_w * halo_h;
  p.halo_bytes = (p.halo_tx + 1023) & ~1023;
  p.in_transform = c->in_transform;
  p.in_slope = c->in_slope;
  p.range_flag = lssvc::range_flag();
  p.acc_scale = c->acc_scale;
  p.bias = c->bias;
  p.epi = c->epi;
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
  LSSVC_REQUIRE(opt(c->res1, &p.res1, &p.res1_pitch), "conv_hs: res1 shape mismatch");
  LSSVC_REQUIRE(opt(c->res2, &p.res2, &p.res2_pitch), "conv_hs: res2 shape mismatch");
  const float *o2 = nullptr;
  LSSVC_REQUIRE(opt(c->out2, &o2, &p.out2_pitch), "conv_hs: out2 shape mismatch");
  p.out2 = const_cast<float *>(o2);
  p.slope2 = c->slope2;
  const bool ent_epi = c->epi == LSSVC_EPI_LAPLACE || c->epi == LSSVC_EPI_BITPARM || c->epi == LSSVC_EPI_FOURPART;
  if (ent_epi) {
    LSSVC_REQUIRE(c->out_scale == 1.f && !c->pixel_shuffle && !p.res2 && !p.out2,
                  "conv_hs: an entropy epilogue takes no output scale, second residual, PixelShuffle or second output");
    LSSVC_REQUIRE(c->cout % 16 == 0 && c->cout == c->n_pad && vec,
                  "conv_hs: entropy epilogue needs cout %% 16 == 0 and 16-byte aligned views (cout=%d)", c->cout);
    if (c->epi != LSSVC_EPI_BITPARM) {
      const int C = c->cout / 2;
      if (c->epi == LSSVC_EPI_FOURPART) {
        LSSVC_REQUIRE(c->ent_step >= 0 && c->ent_step < 4 && C % 64 == 0, "conv_hs: four-part epilogue: step %d, C = %d (needs C %% 64 == 0)",
                      c->ent_step, C);
        p.ent_step = c->ent_step;
      }
      // several channel tiles: the packed channels must have been interleaved for exactly this tile width (lssvc_conv::ent_tile)
      LSSVC_REQUIRE(n_tile % 32 == 0 && (p.n_tiles == 1 ? (c->ent_tile == 0 || c->ent_tile == n_tile) : c->ent_tile == n_tile),
                    "conv_hs: Laplace epilogue over %d channel tiles of %d needs weights interleaved for that tile (ent_tile=%d)",
                    p.n_tiles, n_tile, c->ent_tile);
      LSSVC_REQUIRE(C % 16 == 0 && lssvc::view_ok(&c->ent_y) && lssvc::view_ok(&c->ent_y_hat) && c->ent_y.H == Ho && c->ent_y.W == Wo &&
                        c->ent_y.C == C && c->ent_y_hat.H == Ho && c->ent_y_hat.W == Wo && c->ent_y_hat.C == C,
                    "conv_hs: Laplace epilogue needs ent_y / ent_y_hat views of %dx%dx%d", Ho, Wo, C);
      LSSVC_REQUIRE(c->ent_y.pitch % 4 == 0 && c->ent_y_hat.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(c->ent_y.ptr) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(c->ent_y_hat.ptr) & 15) == 0,
                    "conv_hs: ent_y / ent_y_hat must be 16-byte aligned");
      LSSVC_REQUIRE(!c->ent_index || (c->ent_thr && c->ent_n_thr > 0), "conv_hs: ent_index needs the scale thresholds");
      p.ent_y = c->ent_y.ptr; p.ent_y_pitch = c->ent_y.pitch;
      p.ent_y_hat = c->ent_y_hat.ptr; p.ent_h_pitch = c->ent_y_hat.pitch;
      p.ent_index = c->ent_index; p.ent_thr = c->ent_thr; p.ent_n_thr = c->ent_n_thr;
    } else {
      LSSVC_REQUIRE(c->ent_coef != nullptr, "conv_hs: BitEstimator epilogue needs ent_coef");
      p.ent_coef = c->ent_coef;
    }
    p.ent_bits = c->ent_bits;
    p.ent_sym = c->ent_sym;
  } else if (c->epi != LSSVC_EPI_PLAIN) {
    LSSVC_REQUIRE(!c->pixel_shuffle && lssvc::view_ok(&c->gdn_x) && c->gdn_x.H == Ho && c->gdn_x.W == Wo &&
                      c->gdn_x.C == c->cout,
                  "conv_hs: GDN epilogue needs a matching gdn_x view");
    p.gdn_x = c->gdn_x.ptr;
    p.gdn_pitch = c->gdn_x.pitch;
    vec = vec && (c->gdn_x.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(c->gdn_x.ptr) & 15) == 0);
  }
  p.vec_ok = vec ? 1 : 0;

  // ---- TMA-store epilogue: each epilogue set stages its 128-pixel unit in shared memory as 128-byte (or 64-byte)
  // swizzled rows per slab of 32 (16) channels and writes it with cp.async.bulk.tensor stores (coalesced, clipped at the
  // image border and at the view's last channel).  With PixelShuffle a slab must not straddle two sub-pixels.
  const int cq = c->cout / 4;
  int slab_w = 32;
  static const bool no_tma_env = getenv("LSSVC_H2_NOTMA") != nullptr;
  bool use_tma = vec && !no_tma_env && (!c->pixel_shuffle || cq % 16 == 0) && !ent_epi;  // (entropy epilogues store directly)
  if (n_tile % 32 != 0 || (c->pixel_shuffle && cq % 32 != 0)) slab_w = 16;
  p.slab_w = slab_w;
  p.n_slabs = (n_tile + slab_w - 1) / slab_w;
  const int stage_bytes = p.n_slabs * 128 * slab_w * 4;
  // residuals ride the staging buffers: res1 lands in the out staging, res2 in the out2 staging when that one is free
  const bool r1_tma = use_tma && p.res1 && !c->pixel_shuffle;
  bool r2_tma = use_tma && p.res2 && !c->pixel_shuffle && !p.out2;
  int per_set = stage_bytes * ((p.out2 || r2_tma) ? 2 : 1);
  const int smem_budget = 225 * 1024;
  auto fits = [&](int halos, int slots, bool staging) {
    return halos * p.halo_bytes + slots * p.b_bytes + (staging ? 2 * per_set : 0) + 1024 <= smem_budget;
  };
  if (r2_tma && !fits(2, 3, true)) {  // no room for a second staging buffer: res2 is read from global in the epilogue
    r2_tma = false;
    per_set = stage_bytes * (p.out2 ? 2 : 1);
  }
  if (use_tma && !fits(2, 2, true) && STAGE_K / kc > 1) {
    // a 128-channel tile of a 3x3 layer: 128 KB of staging + two halos + two 2-tap weight stages exceed the budget by a few
    // KB.  One tap per stage (twice the stages, half the slot) keeps the TMA-store / lean epilogue, which is worth far more
    // than the longer stage list (3x3 64->128 + residual at 576x960 ran the general epilogue: 228 TFLOP/s)
    const bool ok = build_stages(1) && fits(2, 2, true);
    if (!ok) build_stages(STAGE_K / kc);  // (7x7: 49 one-tap stages overflow the table) back to the default
  }
  if (use_tma && !fits(2, 2, true)) use_tma = false;
  LSSVC_REQUIRE(fits(2, 2, use_tma), "conv_hs: pipeline does not fit in shared memory (halo %d B, weight stage %d B, staging %d B)",
                p.halo_bytes, p.b_bytes, use_tma ? 2 * per_set : 0);
  p.halo_bufs = 2;
  p.slots = 2;
  const char *hb_str = getenv("LSSVC_HS_HALOS");  // A/B switch: cap of the halo ring
  const int halo_cap = hb_str && atoi(hb_str) >= 2 && atoi(hb_str) <= MAX_HALO ? atoi(hb_str) : MAX_HALO;
  for (bool grown = true; grown;) {
    grown = false;
    if (p.slots < MAX_SLOTS && p.slots <= p.halo_bufs + 1 && fits(p.halo_bufs, p.slots + 1, use_tma)) { ++p.slots; grown = true; }
    if (p.halo_bufs < halo_cap && fits(p.halo_bufs + 1, p.slots, use_tma)) { ++p.halo_bufs; grown = true; }
  }
  p.use_tma = use_tma ? 1 : 0;
  p.stage_off = p.halo_bufs * p.halo_bytes + p.slots * p.b_bytes;
  p.stage2_delta = stage_bytes;
  p.stage_stride = per_set;
  if (use_tma) {
    auto make_out_map = [&](CUtensorMap *m, const lssvc_view &v) -> CUresult {
      const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
      const CUtensorMapSwizzle sw = slab_w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      if (c->pixel_shuffle) {
        // [2Ho][2Wo][cq] viewed as [Ho][2][Wo][2][cq]: one box per sub-pixel (i, j)
        const cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.C), 2, static_cast<cuuint64_t>(Wo), 2, static_cast<cuuint64_t>(Ho)};
        const cuuint64_t strides[4] = {px, px * 2, px * 2 * Wo, px * 2 * Wo * 2};
        const cuuint32_t box[5] = {static_cast<cuuint32_t>(slab_w), 1, SUB_W, 1, TILE_H};
        return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho)};
      const cuuint64_t strides[2] = {px, px * Wo};
      const cuuint32_t box[3] = {static_cast<cuuint32_t>(slab_w), SUB_W, TILE_H};
      return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = make_out_map(&p.out_map, c->out);
    if (r == CUDA_SUCCESS && p.out2) r = make_out_map(&p.out2_map, c->out2);
    if (r == CUDA_SUCCESS && r1_tma && use_tma) { r = make_out_map(&p.res1_map, c->res1); p.res_tma |= 1; }
    if (r == CUDA_SUCCESS && r2_tma && use_tma) { r = make_out_map(&p.res2_map, c->res2); p.res_tma |= 2; }
    p.res_tx = (p.res_tma & 1) + ((p.res_tma >> 1) & 1);  // residual tensors loaded per slab
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_hs: cuTensorMapEncodeTiled(out) failed with %d (C=%d pitch=%d %dx%d ps=%d)", static_cast<int>(r),
                       c->out.C, c->out.pitch, Ho, Wo, c->pixel_shuffle);
      return LSSVC_ERR_CUDA;
    }
  }

  // Instrumented variant (tools/conv_bench.py): LSSVC_HS_DBG = bit mask, 1 no A_hi MMAs, 2 no A_lo MMAs, 4 no halo
  // conversion, 8 no epilogue, 16 no halo TMA, 32 no weight TMA (output is garbage when non-zero), 64 = just the
  // per-role wait counters of CTA 0, printed to stderr after a device sync.  Never set outside profiling.
  const char *dbg_str = getenv("LSSVC_HS_DBG");
  const bool dbg_on = dbg_str != nullptr && atoi(dbg_str) != 0;
  long long *prof_dev = nullptr;
  if (dbg_on) {
    p.dbg = atoi(dbg_str);
    if (p.dbg & 64) {
      LSSVC_CUDA(cudaMalloc(&prof_dev, 48 * sizeof(long long)));
      LSSVC_CUDA(cudaMemset(prof_dev, 0, 48 * sizeof(long long)));
      p.prof = prof_dev;
    }
  }
  const int total_tiles = p.tiles_x * p.tiles_y * p.n_tiles;
  // A-resident mode: several channel tiles, every halo of a pixel tile fits the ring at once, and there are enough pixel
  // tiles to give every SM one.  (1x1 128 -> 512 at 288x480: the input was fetched and converted four times.)
  {
    static const bool ares_off = getenv("LSSVC_HS_NO_ARES") != nullptr;  // A/B switch
    const int halos_per_tile = p.total_chunks * p.n_groups;
    p.a_res = (!ares_off && p.n_tiles > 1 && halos_per_tile <= p.halo_bufs && p.tiles_x * p.tiles_y >= g_num_sms) ? 1 : 0;
    p.outer = p.a_res ? p.tiles_x * p.tiles_y : total_tiles;
    p.inner = p.a_res ? p.n_tiles : 1;
  }
  const int grid = p.outer < g_num_sms ? p.outer : g_num_sms;
  const size_t smem = static_cast<size_t>(p.stage_off) + (use_tma ? 2 * static_cast<size_t>(per_set) : 0) + 1024;
  const int ki = kc == 32 ? 0 : 1;
  cudaStream_t s = lssvc::as_stream(stream);
  typedef void (*KernelFn)(const HsParams);
  static const KernelFn fns[8] = {conv_hs_kernel<32, 1, false>, conv_hs_kernel<32, 2, false>, conv_hs_kernel<16, 1, false>,
                                  conv_hs_kernel<16, 2, false>, conv_hs_kernel<32, 1, true>,  conv_hs_kernel<32, 2, true>,
                                  conv_hs_kernel<16, 1, true>,  conv_hs_kernel<16, 2, true>};
  if (!g_attr_set[0]) {
    for (KernelFn fn : fns)
      LSSVC_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    g_attr_set[0] = true;
  }
  LSSVC_CUDA(lssvc::launch_pdl(fns[(dbg_on ? 4 : 0) + ki * 2 + (mt - 1)], grid, NUM_THREADS, smem, s, p));
  LSSVC_LAUNCHED();
  if (prof_dev) {
    long long h[48];
    LSSVC_CUDA(cudaStreamSynchronize(s));
    LSSVC_CUDA(cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(prof_dev);
    static const char *names[6] = {"halo_tma  [wait empty]", "mma       [wait tempty, wait halo_conv, wait w_full, issue]",
                                   "weight_tma[wait empty]", "convert   [wait halo_full, convert]",
                                   "epilogue0 [wait tfull, work, wait staging]", "epilogue1 [wait tfull, work, wait staging]"};
    fprintf(stderr, "conv_hs prof (CTA 0, cycles; mt=%d halos=%d slots=%d n_acc=%d tma_store=%d tiles/cta~%d):\n", mt, p.halo_bufs, p.slots,
            p.n_acc, p.use_tma, (total_tiles + grid - 1) / grid);
    for (int r = 0; r < 6; ++r)
      fprintf(stderr, "  %-62s total %8lld | %8lld %8lld %8lld %8lld\n", names[r], h[r * 8 + 7], h[r * 8], h[r * 8 + 1], h[r * 8 + 2], h[r * 8 + 3]);
  }
  return LSSVC_OK;
}
