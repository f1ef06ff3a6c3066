// Fused ConvFFN of DepthConvBlock (reference: src/InterModules/lssvc_modules.py:42-60):
//
//     y = o + lrelu(W2 . lrelu(W1 . o + b1, s1) + b2, s2)  (+ res2)        W1: C -> Hd, W2: Hd -> C, both 1x1
//
// as ONE kernel: the Hd-channel intermediate (4x the width of o: 2.26 GB per tensor at 1080p) never leaves the SM.
// Two chained GEMMs per 128-pixel tile in the split-fp16 arithmetic of conv_hs.cu (x = x_hi + x_lo, three
// kind::f16 MMAs per product, hi*hi and cross terms in separate fp32 TMEM accumulators):
//
//   o tile (TMA, fp32, stays in smem for the residual) --splitters--> A_o[tile & 1] in TMEM
//   for each chunk j of 32 hidden channels (global chunk counter g over the CTA's tiles, buffer g % 4):
//       D_h[g%4]  = A_o . W1[j]                         (K = C,  N = 32, accumulators 64 TMEM columns)
//       A_h[g%4]  = split(lrelu(D_h * 2^-s + b1))       (16 "hidden" warps: TMEM -> regs -> TMEM, IN PLACE over D_h)
//       D_y      += A_h[g%4] . W2[j]                    (K = 32, N = C)
//   y = o + lrelu(D_y * 2^-s + b2) -> swizzled smem staging -> TMA store
//
// Pipeline (round 2).  The first version ran ONE issuing thread through a fixed interleaving of GEMM-a and GEMM-b with two
// chunks in flight: 10 k cycles per 128-pixel tile for 4 k of tensor time — the thread needed 6.2-6.7 k cycles just to issue
// 96 MMAs, 24 barrier waits, commits and fences (LSSVC_FFN_DBG counters), waited 1.8 k for the single A_o buffer, and the
// hidden warps idled 71-83 %.  Now: TWO issuing threads (warp 1: every GEMM-a, warp 3: every GEMM-b) that only meet at the
// barriers, so the issue work is halved per thread and GEMM-a runs ahead as far as the buffers allow (across tile
// boundaries); THREE hidden chunks in flight with A_h written in place over the D_h columns already read (a buffer costs 64
// TMEM columns instead of 96); A_o double-buffered; D_y single-buffered with EIGHT output warps (two per TMEM lane quarter)
// so that its epilogue — now between GEMM-b of consecutive tiles — is short.
// TMEM: A_o[2] 2C | D_h/A_h[3] 192 | D_y 2C = 4C + 192 <= 512 columns.
//
// Both weight matrices (pre-split, pre-scaled, pre-swizzled fp16, 8*C*Hd bytes) are loaded ONCE per CTA and stay
// resident in shared memory; per tile only o is read and y written: 2*C*4 bytes per pixel, the HBM minimum.
// Warp roles (896 threads): 0 input TMA, 1 GEMM-a issuer, 2 TMEM alloc + weight loader, 3 GEMM-b issuer, 4..7 splitters,
// 8..19 hidden epilogue (3 chunks in flight x 4 lane quarters), 20..27 output epilogue (4 lane quarters x 2 column sets).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int HC = 32;        // hidden channels per chunk
constexpr int IN_BUFS = 2;
constexpr int NB = 3;         // hidden chunks in flight (D_h / A_h buffers)
constexpr int OUT_WARPS = 8;  // output epilogue: 4 lane quarters x 2 column interleaves
constexpr int NUM_THREADS = 896;
constexpr int TMEM_COLS = 512;

struct alignas(64) FfnParams {
  CUtensorMap in_map;   // o   [H][W][C] fp32, box (slab_w, 16, 8)
  CUtensorMap out_map;  // y
  const void *w1;       // fp16 [n_chunks][C/16][2 (hi, lo)][32][16], 32-byte rows pre-swizzled (SWIZZLE_32B)
  const void *w2;       // fp16 [n_chunks][2][2 (hi, lo)][C][16]
  const float *b1;      // [Hd]
  const float *b2;      // [C]
  const float *res2;
  int res2_pitch;
  int C, n_chunks;
  int H, W, tiles_x, tiles_y;
  int slab_w, n_slabs;
  int w1_bytes, w2_bytes;
  int in_off, stage_off;  // smem offsets (after the resident weights)
  float scale1, scale2;   // 2^-shift of W1 / W2
  float slope1, slope2;
  unsigned int *range_flag;  // raised when a split operand leaves the fp16 range (range.cu)
  long long *prof;  // DBG variant (LSSVC_FFN_DBG=1): [6 roles][8] wait-cycle counters of CTA 0
};

__device__ __forceinline__ void split_pair_f(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// shared-memory address of 16-byte chunk `chunk` (of the slab row) of pixel m in a swizzled [slab][128][slab_w] tile
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int slab_w, int m, int c) {
  const int slab = c / slab_w;
  const uint32_t row = base + static_cast<uint32_t>(slab) * (128u * slab_w * 4u) + static_cast<uint32_t>(m) * (slab_w * 4u);
  const uint32_t swz = (slab_w == 32 ? static_cast<uint32_t>(m & 7) : static_cast<uint32_t>((m >> 1) & 3)) << 4;
  return row + ((static_cast<uint32_t>(c % slab_w) << 2) ^ swz);
}

// DBG = true: instrumented variant (per-role wait-time counters), never on the product path
template <bool DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_ffn_kernel(const __grid_constant__ FfnParams p) {
  ptx::pdl_launch_dependents();
  long long prof[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = DBG ? clock64() : 0;
#define FFN_WAIT(slot, bar, parity)       \
  do {                                    \
    if (DBG) {                            \
      const long long t0__ = clock64();   \
      ptx::mbar_wait(bar, parity);        \
      prof[slot] += clock64() - t0__;     \
    } else {                              \
      ptx::mbar_wait(bar, parity);        \
    }                                     \
  } while (0)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t in_full[IN_BUFS], in_empty[IN_BUFS];
  __shared__ uint64_t ao_full[2], ao_empty[2], w_full;
  __shared__ uint64_t dh_full[NB], ah_full[NB], ah_empty[NB], dy_full, dy_empty;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::pin((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  const int C = p.C, n_chunks = p.n_chunks, KS1 = C >> 4;
  const uint32_t w1_s = smem_base, w2_s = smem_base + static_cast<uint32_t>(p.w1_bytes);
  const uint32_t in_s = smem_base + static_cast<uint32_t>(p.in_off), stage_s = smem_base + static_cast<uint32_t>(p.stage_off);
  const uint32_t in_bytes = 128u * static_cast<uint32_t>(p.n_slabs * p.slab_w) * 4u;
  const uint32_t b_in_full = ptx::pin(ptx::smem_u32(in_full)), b_in_empty = ptx::pin(ptx::smem_u32(in_empty));
  const uint32_t b_ao_full = ptx::pin(ptx::smem_u32(ao_full)), b_ao_empty = ptx::pin(ptx::smem_u32(ao_empty));
  const uint32_t b_w_full = ptx::pin(ptx::smem_u32(&w_full));
  const uint32_t b_dh_full = ptx::pin(ptx::smem_u32(dh_full));
  const uint32_t b_ah_full = ptx::pin(ptx::smem_u32(ah_full)), b_ah_empty = ptx::pin(ptx::smem_u32(ah_empty));
  const uint32_t b_dy_full = ptx::pin(ptx::smem_u32(&dy_full)), b_dy_empty = ptx::pin(ptx::smem_u32(&dy_empty));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.in_map);
    ptx::prefetch_tensormap(&p.out_map);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < IN_BUFS; ++b) {
      ptx::mbar_init(b_in_full + 8 * b, 1);
      ptx::mbar_init(b_in_empty + 8 * b, 4 + OUT_WARPS);  // splitter + output-epilogue warps read the tile
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(b_ao_full + 8 * b, 4);
      ptx::mbar_init(b_ao_empty + 8 * b, 1);
    }
    ptx::mbar_init(b_w_full, 1);
    for (int b = 0; b < NB; ++b) {
      ptx::mbar_init(b_dh_full + 8 * b, 1);
      ptx::mbar_init(b_ah_full + 8 * b, 4);   // the four lane-quarter warps of the buffer
      ptx::mbar_init(b_ah_empty + 8 * b, 1);
    }
    ptx::mbar_init(b_dy_full, 1);
    ptx::mbar_init(b_dy_empty, OUT_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  ptx::pdl_wait();  // first global-memory access below (ptx.cuh: programmatic dependent launch)
  // TMEM columns: A_o[2] C each | D_h[NB] 64 each (A_h written in place: chunk half h -> hi at +16h, lo at +16h + 8) | D_y 2C
  const uint32_t t_ao = tmem_base;
  const uint32_t t_dh = tmem_base + static_cast<uint32_t>(2 * C);
  const uint32_t t_dy = t_dh + 64u * NB;

  const int total_tiles = p.tiles_x * p.tiles_y;
  const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------- input TMA producer -----------------------------------
    if (ptx::elect_one()) {
      int ib = 0;
      uint32_t iph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        FFN_WAIT(0, b_in_empty + 8 * ib, iph ^ 1u);
        ptx::mbar_expect_tx(b_in_full + 8 * ib, in_bytes);
        for (int s = 0; s < p.n_slabs; ++s)
          ptx::tma_load_3d(in_s + static_cast<uint32_t>(ib) * in_bytes + static_cast<uint32_t>(s) * (128u * p.slab_w * 4u),
                           &p.in_map, b_in_full + 8 * ib, s * p.slab_w, tx * TILE_W, ty * TILE_H);
        if (++ib == IN_BUFS) {
          ib = 0;
          iph ^= 1u;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------- weight loader (once) ---------------------------------
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(b_w_full, static_cast<uint32_t>(p.w1_bytes + p.w2_bytes));
      const uint8_t *g1 = reinterpret_cast<const uint8_t *>(p.w1);
      for (int off = 0; off < p.w1_bytes; off += 16384) {
        const int n = p.w1_bytes - off < 16384 ? p.w1_bytes - off : 16384;
        ptx::bulk_load_1d(w1_s + off, g1 + off, n, b_w_full);
      }
      const uint8_t *g2 = reinterpret_cast<const uint8_t *>(p.w2);
      for (int off = 0; off < p.w2_bytes; off += 16384) {
        const int n = p.w2_bytes - off < 16384 ? p.w2_bytes - off : 16384;
        ptx::bulk_load_1d(w2_s + off, g2 + off, n, b_w_full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- GEMM-a issuer: D_h[g % NB] = A_o . W1[j] ------------------
    // One elected thread; descriptors are (lo, hi) 32-bit words advanced by adds (see conv_hs.cu).
    if (ptx::elect_one()) {
      const uint32_t idesc_a1 = ptx::make_idesc_f16_m128(2 * HC), idesc_a2 = ptx::make_idesc_f16_m128(HC);
      constexpr uint32_t B_HI = (256u >> 4) | (1u << 14) | (6u << 29);  // SBO = 256 B, version 1, SWIZZLE_32B
      const uint32_t w1_sub16 = (2u * HC * 32u) >> 4;  // one W1 sub-tile [2][32][16] fp16, in 16-byte units
      const uint32_t w1_16 = (w1_s >> 4) | (1u << 16);
      const uint32_t half_c = static_cast<uint32_t>(C) >> 1;
      ptx::mbar_wait(b_w_full, 0);
      uint32_t b = 0, ph = 1;  // buffer of the next chunk; parity to wait on ah_empty[b] (1 on a fresh barrier: free)
      for (int ti = 0; ti < my_tiles; ++ti) {
        FFN_WAIT(0, b_ao_full + 8 * (ti & 1), static_cast<uint32_t>(ti >> 1) & 1u);
        const uint32_t ao = t_ao + static_cast<uint32_t>((ti & 1) * C);
        uint32_t w1p = w1_16;
        for (int j = 0; j < n_chunks; ++j) {
          FFN_WAIT(2, b_ah_empty + 8 * b, ph);  // GEMM-b of the buffer's previous chunk has read A_h
          ptx::tc_fence_after();
          const uint32_t d = t_dh + 64u * b;
          for (int ks = 0; ks < KS1; ++ks) {
            ptx::mma_f16_ts2(d, ao + ks * 8, w1p, B_HI, idesc_a1, ks != 0 ? 1u : 0u);
            ptx::mma_f16_ts2(d + HC, ao + half_c + ks * 8, w1p, B_HI, idesc_a2, 1u);
            w1p += w1_sub16;
          }
          ptx::mma_commit(b_dh_full + 8 * b);
          if (++b == NB) {
            b = 0;
            ph ^= 1u;
          }
        }
        ptx::mma_commit(b_ao_empty + 8 * (ti & 1));  // every GEMM-a of this tile has been issued
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------- GEMM-b issuer: D_y += A_h[g % NB] . W2[j] -----------------
    if (ptx::elect_one()) {
      const uint32_t idesc_b1 = ptx::make_idesc_f16_m128(static_cast<uint32_t>(2 * C));
      const uint32_t idesc_b2 = ptx::make_idesc_f16_m128(static_cast<uint32_t>(C));
      constexpr uint32_t B_HI = (256u >> 4) | (1u << 14) | (6u << 29);
      const uint32_t w2_sub16 = (2u * static_cast<uint32_t>(C) * 32u) >> 4;  // one W2 sub-tile [2][C][16]
      const uint32_t w2_16 = (w2_s >> 4) | (1u << 16);
      ptx::mbar_wait(b_w_full, 0);
      uint32_t b = 0, ph = 0;  // buffer of the next chunk; parity to wait on ah_full[b]
      for (int ti = 0; ti < my_tiles; ++ti) {
        FFN_WAIT(1, b_dy_empty, (static_cast<uint32_t>(ti) & 1u) ^ 1u);  // the previous tile's epilogue has read D_y
        uint32_t w2p = w2_16;
        for (int j = 0; j < n_chunks; ++j) {
          FFN_WAIT(3, b_ah_full + 8 * b, ph);
          ptx::tc_fence_after();
          const uint32_t a = t_dh + 64u * b;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            ptx::mma_f16_ts2(t_dy, a + ks * 16, w2p, B_HI, idesc_b1, (j | ks) != 0 ? 1u : 0u);
            ptx::mma_f16_ts2(t_dy + C, a + ks * 16 + 8, w2p, B_HI, idesc_b2, 1u);
            w2p += w2_sub16;
          }
          ptx::mma_commit(b_ah_empty + 8 * b);
          if (++b == NB) {
            b = 0;
            ph ^= 1u;
          }
        }
        ptx::mma_commit(b_dy_full);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------- splitters: o tile -> A_o (TMEM) ----------------------
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int ib = 0, ti = 0;
    uint32_t iph = 0;
    float amax = 0.f;  // running max |operand| of this thread (range guard)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      FFN_WAIT(0, b_in_full + 8 * ib, iph);
      const uint32_t tile_s = in_s + static_cast<uint32_t>(ib) * in_bytes;
      const uint32_t lane_addr = t_ao + static_cast<uint32_t>((ti & 1) * C) + (static_cast<uint32_t>(q * 32) << 16);
      FFN_WAIT(1, b_ao_empty + 8 * (ti & 1), (static_cast<uint32_t>(ti >> 1) & 1u) ^ 1u);  // GEMM-a of tile ti - 2 has read this A_o
      ptx::tc_fence_after();
      for (int ks = 0; ks < KS1; ++ks) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = ptx::lds_f4(tile_addr(tile_s, p.slab_w, m, ks * 16 + 4 * i));
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v[i].x), fabsf(v[i].y))), fmaxf(fabsf(v[i].z), fabsf(v[i].w)));
          split_pair_f(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
          split_pair_f(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
        }
        ptx::tmem_st8(lane_addr + ks * 8, hi);
        ptx::tmem_st8(lane_addr + (C >> 1) + ks * 8, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(b_ao_full + 8 * (ti & 1));
        ptx::mbar_arrive(b_in_empty + 8 * ib);
      }
      if (++ib == IN_BUFS) {
        ib = 0;
        iph ^= 1u;
      }
    }
    if (amax >= lssvc::kSplitRangeLimit && p.range_flag) atomicOr(p.range_flag, 1u);
  } else if (warp >= 8 && warp < 8 + 4 * NB) {
    // ------------------------------- hidden epilogue: D_h -> lrelu -> split -> A_h (in place) ---------
    // 12 warps = NB buffers x 4 TMEM lane quarters; a warp owns the 32 lanes x 64 columns of its buffer: D1 = columns
    // [0, 32), D2 = [32, 64) of the chunk's 32 hidden channels.  Half h (channels 16h .. 16h+15) is read (D1[16h..], D2[32+16h..])
    // and its split activations written back over the D1 columns just read: hi at 16h, lo at 16h + 8.
    const int hw = warp - 8;
    const int q = hw & 3;          // TMEM lane quarter (== warp % 4)
    const int b = hw >> 2;         // buffer
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t buf = t_dh + 64u * b + lane_off;
    const float scale1 = p.scale1, slope1 = p.slope1;
    const int n_all = my_tiles * n_chunks;
    uint32_t use = 0;
    float amax = 0.f;  // running max |hidden activation| of this thread (range guard)
    for (int g = b; g < n_all; g += NB, ++use) {
      const int j = g % n_chunks;
      FFN_WAIT(0, b_dh_full + 8 * b, use & 1u);
      ptx::tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float4 *bq = reinterpret_cast<const float4 *>(p.b1 + j * HC + 16 * half);
        float4 bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = __ldg(bq + i);
        uint32_t r1[16], r2[16];
        ptx::tmem_ld16(buf + 16u * half, r1);
        ptx::tmem_ld16(buf + HC + 16u * half, r2);
        ptx::tmem_ld_wait();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float h0 = (__uint_as_float(r1[4 * i + 0]) + __uint_as_float(r2[4 * i + 0])) * scale1 + bv[i].x;
          float h1 = (__uint_as_float(r1[4 * i + 1]) + __uint_as_float(r2[4 * i + 1])) * scale1 + bv[i].y;
          float h2 = (__uint_as_float(r1[4 * i + 2]) + __uint_as_float(r2[4 * i + 2])) * scale1 + bv[i].z;
          float h3 = (__uint_as_float(r1[4 * i + 3]) + __uint_as_float(r2[4 * i + 3])) * scale1 + bv[i].w;
          h0 = h0 > 0.f ? h0 : h0 * slope1;
          h1 = h1 > 0.f ? h1 : h1 * slope1;
          h2 = h2 > 0.f ? h2 : h2 * slope1;
          h3 = h3 > 0.f ? h3 : h3 * slope1;
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(h0), fabsf(h1))), fmaxf(fabsf(h2), fabsf(h3)));
          split_pair_f(h0, h1, hi[2 * i], lo[2 * i]);
          split_pair_f(h2, h3, hi[2 * i + 1], lo[2 * i + 1]);
        }
        // columns [16h, 16h + 16) of D1 have just been read by this warp (its own lanes): safe to overwrite
        ptx::tmem_st8(buf + 16u * half, hi);
        ptx::tmem_st8(buf + 16u * half + 8u, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(b_ah_full + 8 * b);
    }
    if (amax >= lssvc::kSplitRangeLimit && p.range_flag) atomicOr(p.range_flag, 1u);
  } else if (warp >= 8 + 4 * NB) {
    // ------------------------------- output epilogue ---------------------------------------
    // 8 warps: lane quarter q = warp % 4 (the TMEM lanes a warp may read), column set cs = 0 / 1: 16-channel groups
    // cs, cs + 2, ... of the tile
    const int q = warp & 3;
    const int cs = (warp - (8 + 4 * NB)) >> 2;
    const int m = q * 32 + lane;
    const int h = m / TILE_W, w = m % TILE_W;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float scale2 = p.scale2, slope2 = p.slope2;
    const bool store_thread = warp == 8 + 4 * NB && lane == 0;
    int ib = 0, ti = 0;
    uint32_t iph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int oy = ty * TILE_H + h, ox = tx * TILE_W + w;
      FFN_WAIT(0, b_in_full + 8 * ib, iph);  // acquire the TMA-written o tile for the residual reads below
      const bool valid = oy < p.H && ox < p.W;
      const uint32_t tile_s = in_s + static_cast<uint32_t>(ib) * in_bytes;
      // staging is free once the previous tile's TMA store has read it
      if (store_thread) ptx::bulk_wait_read_all();
      ptx::named_bar_sync(2, 32 * OUT_WARPS);
      FFN_WAIT(1, b_dy_full, static_cast<uint32_t>(ti) & 1u);
      ptx::tc_fence_after();
      const uint32_t src = t_dy + lane_off;
      for (int n = 16 * cs; n < C; n += 32) {
        uint32_t r1[16], r2[16];
        ptx::tmem_ld16(src + n, r1);
        ptx::tmem_ld16(src + C + n, r2);
        float4 bv[4], ov[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bv[i] = __ldg(reinterpret_cast<const float4 *>(p.b2 + n) + i);
          ov[i] = ptx::lds_f4(tile_addr(tile_s, p.slab_w, m, n + 4 * i));
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[4];
          v[0] = (__uint_as_float(r1[4 * i + 0]) + __uint_as_float(r2[4 * i + 0])) * scale2 + bv[i].x;
          v[1] = (__uint_as_float(r1[4 * i + 1]) + __uint_as_float(r2[4 * i + 1])) * scale2 + bv[i].y;
          v[2] = (__uint_as_float(r1[4 * i + 2]) + __uint_as_float(r2[4 * i + 2])) * scale2 + bv[i].z;
          v[3] = (__uint_as_float(r1[4 * i + 3]) + __uint_as_float(r2[4 * i + 3])) * scale2 + bv[i].w;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope2;
          v[0] += ov[i].x; v[1] += ov[i].y; v[2] += ov[i].z; v[3] += ov[i].w;
          if (p.res2 && valid) {
            const float4 t = *reinterpret_cast<const float4 *>(
                p.res2 + (static_cast<long long>(oy) * p.W + ox) * p.res2_pitch + n + 4 * i);
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          }
          ptx::sts_u4(tile_addr(stage_s, p.slab_w, m, n + 4 * i), __float_as_uint(v[0]), __float_as_uint(v[1]),
                      __float_as_uint(v[2]), __float_as_uint(v[3]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(b_dy_empty);
        ptx::mbar_arrive(b_in_empty + 8 * ib);
      }
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(3, 32 * OUT_WARPS);
      if (store_thread) {
        for (int s = 0; s < p.n_slabs; ++s)
          ptx::tma_store_3d(&p.out_map, stage_s + static_cast<uint32_t>(s) * (128u * p.slab_w * 4u), s * p.slab_w,
                            tx * TILE_W, ty * TILE_H);
        ptx::bulk_commit();
      }
      if (++ib == IN_BUFS) {
        ib = 0;
        iph ^= 1u;
      }
    }
    if (store_thread) ptx::bulk_wait_all();
  }

  if (DBG && p.prof && blockIdx.x == 0 && lane == 0) {
    // rows: 0 input TMA, 1 GEMM-a issuer, 2 splitter warp 4, 3 hidden warp 8 (buffer 0), 4 GEMM-b issuer, 5 first output warp
    const int row = warp == 0 ? 0 : warp == 1 ? 1 : warp == 4 ? 2 : warp == 8 ? 3 : warp == 3 ? 4 : warp == 8 + 4 * NB ? 5 : -1;
    if (row >= 0) {
      for (int i = 0; i < 6; ++i) p.prof[row * 8 + i] = prof[i];
      p.prof[row * 8 + 7] = clock64() - t_begin;
    }
  }
#undef FFN_WAIT
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

extern "C" int32_t lssvc_conv_ffn(const lssvc_ffn *f, void *stream) {
  LSSVC_REQUIRE(f != nullptr, "conv_ffn: null descriptor");
  const int C = f->in.C, Hd = f->hidden;
  LSSVC_REQUIRE(lssvc::view_ok(&f->in) && lssvc::view_ok(&f->out), "conv_ffn: bad views");
  LSSVC_REQUIRE(f->out.H == f->in.H && f->out.W == f->in.W && f->out.C == C, "conv_ffn: output view mismatch");
  LSSVC_REQUIRE(C % 16 == 0 && C >= 16 && C <= 64, "conv_ffn: C=%d (multiple of 16 up to 64)", C);
  LSSVC_REQUIRE(Hd % (2 * HC) == 0 && Hd >= 2 * HC, "conv_ffn: hidden=%d (multiple of %d)", Hd, 2 * HC);
  auto aligned = [](const lssvc_view &v) {
    return v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0;
  };
  LSSVC_REQUIRE(aligned(f->in) && aligned(f->out), "conv_ffn: views must be 16-byte aligned");
  LSSVC_REQUIRE(f->w1 && f->w2 && f->b1 && f->b2, "conv_ffn: null weights");
  LSSVC_REQUIRE((reinterpret_cast<uintptr_t>(f->w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(f->w2) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(f->b1) & 15) == 0 && (reinterpret_cast<uintptr_t>(f->b2) & 15) == 0,
                "conv_ffn: weights / biases must be 16-byte aligned");
  if (f->res2.ptr) {
    LSSVC_REQUIRE(f->res2.H == f->in.H && f->res2.W == f->in.W && f->res2.C == C && aligned(f->res2),
                  "conv_ffn: res2 view mismatch");
  }
  if (int rc = resolve_driver()) return rc;

  FfnParams p;
  memset(&p, 0, sizeof(p));
  p.C = C;
  p.n_chunks = Hd / HC;
  p.H = f->in.H; p.W = f->in.W;
  p.tiles_x = lssvc::ceil_div(p.W, TILE_W);
  p.tiles_y = lssvc::ceil_div(p.H, TILE_H);
  p.slab_w = C % 32 == 0 ? 32 : 16;
  p.n_slabs = C / p.slab_w;
  p.w1_bytes = Hd * C * 4;  // 2 (hi, lo) x fp16
  p.w2_bytes = Hd * C * 4;
  p.w1 = f->w1; p.w2 = f->w2; p.b1 = f->b1; p.b2 = f->b2;
  p.res2 = f->res2.ptr; p.res2_pitch = f->res2.pitch;
  p.scale1 = f->scale1; p.scale2 = f->scale2; p.slope1 = f->slope1; p.slope2 = f->slope2;
  p.range_flag = lssvc::range_flag();
  const int tile_bytes = 128 * C * 4;
  p.in_off = (p.w1_bytes + p.w2_bytes + 1023) & ~1023;
  p.stage_off = p.in_off + IN_BUFS * tile_bytes;
  const size_t smem = static_cast<size_t>(p.stage_off) + tile_bytes + 1024;
  LSSVC_REQUIRE(smem <= 226 * 1024, "conv_ffn: C=%d hidden=%d needs %zu bytes of shared memory", C, Hd, smem);
  LSSVC_REQUIRE(64 * NB + 4 * C <= TMEM_COLS, "conv_ffn: C=%d does not fit in tensor memory", C);
  static_assert(8 + 4 * NB + OUT_WARPS == NUM_THREADS / 32, "conv_ffn: warp roles");

  const cuuint32_t ones[3] = {1, 1, 1};
  auto make_map = [&](CUtensorMap *m, const lssvc_view &v) -> CUresult {
    const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(v.W), static_cast<cuuint64_t>(v.H)};
    const cuuint64_t strides[2] = {px, px * v.W};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(p.slab_w), TILE_W, TILE_H};
    return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    p.slab_w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = make_map(&p.in_map, f->in);
  if (r == CUDA_SUCCESS) r = make_map(&p.out_map, f->out);
  if (r != CUDA_SUCCESS) {
    lssvc::set_error("conv_ffn: cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
    return LSSVC_ERR_CUDA;
  }
  if (!g_attr_set) {
    LSSVC_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(conv_ffn_kernel<false>),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LSSVC_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(conv_ffn_kernel<true>),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    g_attr_set = true;
  }
  const int total_tiles = p.tiles_x * p.tiles_y;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  // instrumented variant (tools/ffn_bench.py): LSSVC_FFN_DBG=1 prints the per-role wait counters of CTA 0 after a device sync
  const char *dbg_str = getenv("LSSVC_FFN_DBG");
  if (dbg_str != nullptr && atoi(dbg_str) != 0) {
    long long *prof_dev = nullptr;
    LSSVC_CUDA(cudaMalloc(&prof_dev, 48 * sizeof(long long)));
    LSSVC_CUDA(cudaMemset(prof_dev, 0, 48 * sizeof(long long)));
    p.prof = prof_dev;
    conv_ffn_kernel<true><<<grid, NUM_THREADS, smem, lssvc::as_stream(stream)>>>(p);
    LSSVC_LAUNCHED();
    long long h[48];
    LSSVC_CUDA(cudaStreamSynchronize(lssvc::as_stream(stream)));
    LSSVC_CUDA(cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(prof_dev);
    static const char *names[6] = {"input_tma [wait in_empty]", "gemm-a    [wait ao_full, -, ah_empty, -]", "splitter  [wait in_full, ao_empty]",
                                   "hidden b0 [wait dh_full]", "gemm-b    [-, wait dy_empty, -, ah_full]", "output    [wait in_full, dy_full]"};
    fprintf(stderr, "conv_ffn prof (CTA 0, cycles; C=%d hidden=%d tiles/cta~%d):\n", C, Hd, (total_tiles + grid - 1) / grid);
    for (int r = 0; r < 6; ++r)
      fprintf(stderr, "  %-58s total %8lld | %8lld %8lld %8lld %8lld\n", names[r], h[r * 8 + 7], h[r * 8], h[r * 8 + 1], h[r * 8 + 2], h[r * 8 + 3]);
    return LSSVC_OK;
  }
  LSSVC_CUDA(lssvc::launch_pdl(conv_ffn_kernel<false>, grid, NUM_THREADS, smem, lssvc::as_stream(stream), p));
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
