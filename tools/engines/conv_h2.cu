// Implicit-GEMM convolution on the 5th-gen tensor cores with fp32-reference accuracy at fp16 MMA rate.
//
//   D[M = 128 output pixels (8 x 16 patch), N = output channels] += A[M, K] * W[K, N],   K = (tap, input channel)
//
// Split-fp16 arithmetic ("h2").  Every fp32 operand is the sum of two fp16 numbers, x = x_hi + x_lo with
// x_hi = rn_f16(x) and x_lo = rn_f16(x - x_hi) (22 significant bits).  The product is evaluated as
//     D1 += A_hi * W_hi                 (tcgen05.mma kind::f16, fp32 accumulate)
//     D2 += A_hi * W_lo + A_lo * W_hi   (the dropped A_lo * W_lo term is below 2^-22 relative)
// with D1 and D2 in SEPARATE tensor-memory accumulators that the epilogue adds in fp32: the tensor core
// truncates when it adds into a large accumulator, so keeping the 2^-11-sized cross terms out of D1 is what
// keeps the result at fp32-reference level.  Weights are pre-split on the host (scaled by a power of two so that
// W_lo is a normal fp16); W_hi and W_lo rows are stacked so that A_hi * [W_hi | W_lo] is ONE MMA of width 2N.
//
// Data movement (the kernel is bound by shared-memory and L2->SM bandwidth, not by the MMA rate, so every byte
// counts):
//   * activations stay fp32 NHWC in HBM.  Per (source, KC-channel chunk) ONE halo tile (patch + kernel apron)
//     is fetched by a 5-D TMA box load; out-of-image pixels are zero-filled by TMA (= the conv's zero padding);
//     the 5-D view [H/s][s][W/s][s][C] turns a stride-s conv into unit-stride boxes per parity plane.
//   * 8 splitter warps read, per filter tap, the shifted window of the halo tile from shared memory (each thread
//     owns one output pixel = one TMEM lane), split it to fp16 hi/lo in registers and write it with tcgen05.st
//     straight into tensor memory.  The A operand of the MMA is read from TMEM (tcgen05.mma [d], [a], b_desc), so
//     neither the 9x re-fetch of the taps from L2 nor any A traffic of the MMA touches shared memory.
//   * the stacked weight tile [2][n_tile][KC] fp16 of each (chunk, tap) is TMA-loaded into a swizzled K-major ring.
//
// Warp roles (896 threads): 0 halo TMA producer, 1 MMA issuer, 2 TMEM allocator + weight TMA producer,
// 4..11 splitters (two sets of 4 warps, alternate stages), 12..19 epilogue (two sets, alternate 16-channel
// chunks: TMEM -> registers -> bias / activation / GDN / residuals / pixel-shuffle -> global), 20..27 halo
// converters (fp32 -> fp16 hi | lo in place).  Persistent over tiles; the accumulators are double
// buffered when 4 * n_tile + ring fits in the 512 TMEM columns.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int MAX_SLOTS = 8;   // ring of (A slot in TMEM, weight tile in smem)
constexpr int MAX_HALO = 8;    // ring of halo tiles
constexpr int MAX_TAPS = 49;
constexpr int MAX_GROUPS = 4;  // stride-2: one halo per input parity plane
constexpr int STAGE_K = 64;    // input channels x taps of one pipeline stage: KC = 32 -> 2 taps, KC = 16 -> 4 taps
constexpr int NUM_THREADS = 896;
constexpr int TMEM_COLS = 512;

struct alignas(64) H2Params {
  CUtensorMap a_map[LSSVC_MAX_SRC];
  CUtensorMap b_map;
  CUtensorMap out_map, out2_map;  // TMA-store epilogue (use_tma)
  int use_tma, slab_w, n_slabs, stage_off, stage2_off, stage_bufs, stage_stride;  // staging: [bufs][out, out2][n_slabs][128 px][slab_w fp32]
  int n_src;
  int chunks[LSSVC_MAX_SRC];  // ceil(C / KC) per source
  int coff[LSSVC_MAX_SRC];    // first packed input channel of the source
  int n_groups;
  int g_px[MAX_GROUPS], g_py[MAX_GROUPS];
  int g_tap0[MAX_GROUPS + 1];  // taps of group g: [g_tap0[g], g_tap0[g + 1])
  int g_step[MAX_GROUPS];      // taps per stage in group g: the group's taps split evenly over ceil(taps / max) stages
  int q0x, q0y;                // halo origin relative to the patch origin (plane coordinates)
  int halo_w;                  // halo width in pixels
  int halo_tx;                 // bytes of one halo box
  int halo_rows;               // pixels of one halo box
  unsigned char tap_w[MAX_TAPS];   // weight tap index r * kw + s
  unsigned char tap_ry[MAX_TAPS];  // row / column of the tap's window origin inside the halo
  unsigned char tap_rx[MAX_TAPS];
  unsigned short tap_off[MAX_TAPS];  // tap_ry * halo_w + tap_rx
  int stages_per_tile;
  int Ho, Wo;
  int tiles_x, tiles_y, n_tiles, n_tile, cout;
  int slots, halo_bufs, halo_bytes, b_bytes;
  int d_bufs;   // accumulator buffers (2 when they fit)
  int a_col0;   // first TMEM column of the A ring
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
  int dbg;  // LSSVC_H2_DBG: bottleneck-isolation switches (results are wrong when non-zero); see lssvc_conv_h2
};

__device__ __forceinline__ long long out_offset(const H2Params &p, int oy, int ox, int ch, int P) {
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

// fp32 halo row (already in registers) -> packed fp16 hi / lo
template <int NV>
__device__ __forceinline__ void split_row(const float4 (&v)[NV], uint32_t (&hi)[2 * NV], uint32_t (&lo)[2 * NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    split_pair(v[i].x, v[i].y, hi[2 * i], lo[2 * i]);
    split_pair(v[i].z, v[i].w, hi[2 * i + 1], lo[2 * i + 1]);
  }
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

template <int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_h2_kernel(const __grid_constant__ H2Params p) {
  constexpr int TAPS_PER_STAGE = STAGE_K / KC;  // taps of one halo group sharing a stage (one barrier round trip)
  constexpr int ROWB = KC * 4;        // bytes of one halo pixel (fp32, or fp16 hi | fp16 lo after conversion)
  constexpr int NV = KC / 4;          // 16-byte chunks per halo pixel
  constexpr int KS = KC / 16;         // K = 16 MMA slices per stage
  constexpr int HALF = KC / 2;        // TMEM columns of A_hi (and of A_lo) per slot
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
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::pin((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  // layout: [halo_bufs x halo_bytes][slots x b_bytes]
  const uint32_t b_base = smem_base + static_cast<uint32_t>(p.halo_bufs * p.halo_bytes);
  // pinned: keeps the compiler from re-deriving the shared-window address (S2UR + ULEA) at every use
  const uint32_t bar_full = ptx::pin(ptx::smem_u32(full_s));
  const uint32_t bar_empty = ptx::pin(ptx::smem_u32(empty_s));
  const uint32_t bar_halo_full = ptx::pin(ptx::smem_u32(halo_full)), bar_halo_empty = ptx::pin(ptx::smem_u32(halo_empty));
  const uint32_t bar_halo_conv = ptx::pin(ptx::smem_u32(halo_conv));
  const uint32_t bar_tfull = ptx::pin(ptx::smem_u32(tfull_bar)), bar_tempty = ptx::pin(ptx::smem_u32(tempty_bar));

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.a_map[j]);
    ptx::prefetch_tensormap(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.slots; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 5);  // weight TMA producer (+ its bytes) and the 4 splitter warps of the stage
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int h = 0; h < p.halo_bufs; ++h) {
      ptx::mbar_init(bar_halo_full + 8 * h, 1);
      ptx::mbar_init(bar_halo_empty + 8 * h, 8);
      ptx::mbar_init(bar_halo_conv + 8 * h, 8);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tempty + 8 * b, 8);
    }
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

  const int tiles_per_n = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_n * p.n_tiles;

  if (warp == 0) {
    // ------------------------------- halo TMA producer -----------------------------------
    if (ptx::elect_one()) {
      int hb = 0;
      uint32_t hph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int rem = tile % tiles_per_n;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int oy0 = ty * TILE_H + p.q0y, ox0 = tx * TILE_W + p.q0x;
        for (int j = 0; j < p.n_src; ++j) {
          for (int c = 0; c < p.chunks[j]; ++c) {
            for (int g = 0; g < p.n_groups; ++g) {
              const uint32_t full = bar_halo_full + 8 * hb;
              ptx::mbar_wait(bar_halo_empty + 8 * hb, hph ^ 1u);
              if (p.dbg & 16) {
                ptx::mbar_arrive(full);
              } else {
                ptx::mbar_expect_tx(full, static_cast<uint32_t>(p.halo_tx));
                ptx::tma_load_5d(smem_base + static_cast<uint32_t>(hb * p.halo_bytes), &p.a_map[j], full, c * KC,
                                 p.g_px[g], ox0, p.g_py[g], oy0);
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
    // A stage is up to TAPS_PER_STAGE consecutive taps of one halo group: their weight tiles land back to back
    // in the stage's slot and complete on one barrier.
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tile_bytes = static_cast<uint32_t>(2 * p.n_tile) * B_ROWB;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile / tiles_per_n) * p.n_tile;
        for (int j = 0; j < p.n_src; ++j) {
          for (int c = 0; c < p.chunks[j]; ++c) {
            const int k0 = p.coff[j] + c * KC;
            for (int g = 0; g < p.n_groups; ++g) {
              const int t1 = p.g_tap0[g + 1];
              const int step = p.g_step[g];
              for (int t = p.g_tap0[g]; t < t1; t += step) {
                const int items = t1 - t < step ? t1 - t : step;
                const uint32_t full = bar_full + 8 * s;
                ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                if (p.dbg & 32) {
                  ptx::mbar_arrive(full);
                } else {
                  ptx::mbar_expect_tx(full, tile_bytes * static_cast<uint32_t>(items));
                  for (int i = 0; i < items; ++i)
                    ptx::tma_load_4d(b_base + static_cast<uint32_t>(s * p.b_bytes) + static_cast<uint32_t>(i) * tile_bytes,
                                     &p.b_map, full, k0, n0, 0, static_cast<int>(p.tap_w[t + i]));
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
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ------------------------------------------
    // The whole warp walks the loop (all lanes poll the barriers); one elected lane issues.  elect.sync lets
    // the compiler emit the UTCHMMA / UTCBAR sequence straight, without a per-instruction election loop.
    int s = 0;
    uint32_t ph = 0;
    int buf = 0;
    uint32_t acc_ph = 0;
    const uint32_t idesc_2n = ptx::make_idesc_f16_m128(static_cast<uint32_t>(2 * p.n_tile));
    const uint32_t idesc_n = ptx::make_idesc_f16_m128(static_cast<uint32_t>(p.n_tile));
    const int slots = p.slots, d_bufs = p.d_bufs;
    const uint32_t n_tile = static_cast<uint32_t>(p.n_tile);
    const uint32_t a_ring = tmem_base + static_cast<uint32_t>(p.a_col0);
    const uint32_t b_bytes = static_cast<uint32_t>(p.b_bytes);
    const bool no_mma = (p.dbg & 1) != 0;
    const uint32_t tile_bytes = 2u * n_tile * B_ROWB;
    const int n_src = p.n_src, n_groups = p.n_groups;
    // the poll of the NEXT stage's barrier is issued before this stage's MMAs so that its latency is hidden
    bool ready = total_tiles > static_cast<int>(blockIdx.x) ? ptx::mbar_try_wait(bar_full, 0) : true;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(bar_tempty + 8 * buf, acc_ph ^ 1u);
      const uint32_t d1 = tmem_base + static_cast<uint32_t>(buf) * 2u * n_tile;
      const uint32_t d2 = d1 + n_tile;
      uint32_t acc = 0;  // 0 only for the very first MMA of the tile
      for (int j = 0; j < n_src; ++j) {
        for (int c = 0; c < p.chunks[j]; ++c) {
          for (int g = 0; g < n_groups; ++g) {
            const int t1 = p.g_tap0[g + 1];
            const int step = p.g_step[g];
            for (int t = p.g_tap0[g]; t < t1; t += step) {
              const int items = t1 - t < step ? t1 - t : step;
              if (!ready) ptx::mbar_wait_slow(bar_full + 8 * s, ph);
              ptx::tc_fence_after();
              int s_next = s + 1;
              uint32_t ph_next = ph;
              if (s_next == slots) {
                s_next = 0;
                ph_next ^= 1u;
              }
              ready = ptx::mbar_try_wait(bar_full + 8 * s_next, ph_next);
              if (ptx::elect_one()) {
                if (!no_mma) {
                  for (int i = 0; i < items; ++i) {
                    const uint32_t a_hi = a_ring + static_cast<uint32_t>((s * TAPS_PER_STAGE + i) * KC);
                    const uint32_t b_addr = b_base + static_cast<uint32_t>(s) * b_bytes + static_cast<uint32_t>(i) * tile_bytes;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                      const uint64_t b_desc = ptx::make_kmajor_desc(b_addr + ks * 32, B_SBO, B_LAYOUT);
                      // [D1 | D2] (+)= A_hi * [W_hi | W_lo]   then   D2 += A_lo * W_hi
                      ptx::mma_f16_ts(d1, a_hi + ks * 8, b_desc, idesc_2n, acc);
                      ptx::mma_f16_ts(d2, a_hi + HALF + ks * 8, b_desc, idesc_n, 1u);
                      acc = 1u;
                    }
                  }
                }
                ptx::mma_commit(bar_empty + 8 * s);
              }
              __syncwarp();
              s = s_next;
              ph = ph_next;
            }
          }
        }
      }
      if (ptx::elect_one()) ptx::mma_commit(bar_tfull + 8 * buf);
      __syncwarp();
      if (++buf == d_bufs) {
        buf = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp >= 20) {
    // ------------------------------- converters: fp32 halo -> fp16 hi | lo, in place -------
    // One thread per halo pixel: the 4*KC-byte fp32 row becomes [hi fp16 x KC | lo fp16 x KC] with the same
    // 16-byte-chunk swizzle, so that a tap of the splitters is NV LDS.128 feeding tcgen05.st with no arithmetic.
    // Groups with a single tap (1x1 convs) are left as they are (the splitters convert on the fly).
    const int ct = threadIdx.x - 640;  // 0..255
    const int halo_rows = p.halo_rows, halo_bufs = p.halo_bufs;
    const int in_transform = p.in_transform;
    const float in_slope = p.in_slope;
    const uint32_t halo_bytes = static_cast<uint32_t>(p.halo_bytes);
    const bool no_conv = (p.dbg & 2) != 0;
    int hb = 0;
    uint32_t hph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_src; ++j) {
        for (int c = 0; c < p.chunks[j]; ++c) {
          for (int g = 0; g < p.n_groups; ++g) {
            ptx::mbar_wait(bar_halo_full + 8 * hb, hph);
            if (p.g_tap0[g + 1] - p.g_tap0[g] > 1 && !no_conv) {
              const uint32_t halo = smem_base + static_cast<uint32_t>(hb) * halo_bytes;
              for (int r = ct; r < halo_rows; r += 256) {
                const uint32_t x = halo + static_cast<uint32_t>(r) * ROWB;
                const uint32_t a0 = x | ((x >> 3) & SWZ);
                float4 v[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) v[i] = ptx::lds_f4(a0 ^ (i << 4));
                transform_row<NV>(v, in_transform, in_slope);
                uint32_t hi[HALF], lo[HALF];
                split_row<NV>(v, hi, lo);
#pragma unroll
                for (int i = 0; i < NV / 2; ++i) {
                  ptx::sts_u4(a0 ^ (i << 4), hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                  ptx::sts_u4(a0 ^ ((NV / 2 + i) << 4), lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                }
              }
              ptx::fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_halo_conv + 8 * hb);
            if (++hb == halo_bufs) {
              hb = 0;
              hph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------- splitters: halo window -> TMEM -----------------------
    // Each thread owns one output pixel (= one TMEM lane): per tap it reads the shifted halo row (already fp16
    // hi | lo) and stores it into the stage's A slot.  Two sets of 4 warps alternate stages; a stage is published
    // (wait::st + arrive) after the loads of the warp's next stage have been issued.
    const int set = (warp - 4) >> 2;
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const int m = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(p.a_col0);
    const int slots = p.slots, halo_bufs = p.halo_bufs;
    const uint32_t pix_boff = static_cast<uint32_t>((m / TILE_W) * p.halo_w + (m % TILE_W)) * ROWB;
    const int in_transform = p.in_transform;
    const float in_slope = p.in_slope;
    const uint32_t halo_bytes = static_cast<uint32_t>(p.halo_bytes);
    const bool no_st = (p.dbg & 4) != 0;
    int s = 0;
    uint32_t ph = 0;
    int hb = 0;
    uint32_t hph = 0;
    uint32_t parity = 0;  // global stage counter & 1
    int pending = -1;     // slot whose TMEM store has been issued but not yet published
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_src; ++j) {
        for (int c = 0; c < p.chunks[j]; ++c) {
          for (int g = 0; g < p.n_groups; ++g) {
            ptx::mbar_wait(bar_halo_conv + 8 * hb, hph);
            const uint32_t halo = smem_base + static_cast<uint32_t>(hb) * halo_bytes + pix_boff;
            const int t0 = p.g_tap0[g], t1 = p.g_tap0[g + 1];
            const bool pre = (t1 - t0) > 1;
            const int step = p.g_step[g];
            for (int t = t0; t < t1; t += step) {
              if (parity == static_cast<uint32_t>(set)) {
                const int items = t1 - t < step ? t1 - t : step;  // warp-uniform
                const uint32_t dst = lane_base + static_cast<uint32_t>(s * TAPS_PER_STAGE * KC);
#pragma unroll
                for (int e = 0; e < TAPS_PER_STAGE; ++e) {
                  if (e < items) {
                    const uint32_t x = halo + static_cast<uint32_t>(p.tap_off[t + e]) * ROWB;
                    const uint32_t a0 = x | ((x >> 3) & SWZ);
                    uint32_t hi[HALF], lo[HALF];
                    if (pre) {
#pragma unroll
                      for (int i = 0; i < NV / 2; ++i)
                        ptx::lds_u4(a0 ^ (i << 4), hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
#pragma unroll
                      for (int i = 0; i < NV / 2; ++i)
                        ptx::lds_u4(a0 ^ ((NV / 2 + i) << 4), lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    } else {
                      float4 v[NV];
#pragma unroll
                      for (int i = 0; i < NV; ++i) v[i] = ptx::lds_f4(a0 ^ (i << 4));
                      transform_row<NV>(v, in_transform, in_slope);
                      split_row<NV>(v, hi, lo);
                    }
                    if (e == 0) {
                      if (pending >= 0) {  // publish the previous stage of this warp
                        ptx::tmem_st_wait();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(bar_full + 8 * pending);
                      }
                      // the slot is free once the MMAs of its previous use have completed
                      ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                      ptx::tc_fence_after();
                    }
                    if (!no_st) {
                      if constexpr (KC == 32) {
                        ptx::tmem_st16(dst + e * KC, hi);
                        ptx::tmem_st16(dst + e * KC + HALF, lo);
                      } else {
                        ptx::tmem_st8(dst + e * KC, hi);
                        ptx::tmem_st8(dst + e * KC + HALF, lo);
                      }
                    }
                  }
                }
                pending = s;
              }
              parity ^= 1u;
              if (++s == slots) {
                s = 0;
                ph ^= 1u;
              }
            }
            if (pending >= 0) {
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(bar_full + 8 * pending);
              pending = -1;
            }
            // generic-proxy accesses of this buffer are done before the next TMA (async proxy) write
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_halo_empty + 8 * hb);
            if (++hb == halo_bufs) {
              hb = 0;
              hph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= 12 && warp < 20) {
    // ------------------------------- epilogue ---------------------------------------------
    // two sets of 4 warps; set e takes the 16-channel chunks e, e + 2, ... of every tile
    const int eset = (warp - 12) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int h = m / TILE_W, w = m % TILE_W;
    const int n_tile = p.n_tile, cout = p.cout, d_bufs = p.d_bufs, Wo = p.Wo;
    const float acc_scale = p.acc_scale, out_scale = p.out_scale, slope = p.slope, slope2 = p.slope2;
    const bool has_act = p.act != 0, fast = p.vec_ok != 0, ps = p.pixel_shuffle != 0;
    const bool no_store = (p.dbg & 8) != 0;
    const int epi = p.epi;
    const int cq = cout >> 2;
    const bool chunk_uniform = !ps || (cq & 15) == 0;  // a 16-channel chunk never straddles two sub-pixels
    float *const out = p.out;
    float *const out2 = p.out2;
    const float *const res1 = p.res1;
    const float *const res2 = p.res2;
    const float *const gdn_x = p.gdn_x;
    const float *const bias = p.bias;
    const long long out_pitch = p.out_pitch, out2_pitch = p.out2_pitch, res1_pitch = p.res1_pitch,
                    res2_pitch = p.res2_pitch, gdn_pitch = p.gdn_pitch;
    const bool use_tma = p.use_tma != 0 && fast && chunk_uniform;
    const uint32_t slab_w = static_cast<uint32_t>(p.slab_w);
    const uint32_t stage_base = smem_base + static_cast<uint32_t>(p.stage_off), stage2_delta = static_cast<uint32_t>(p.stage2_off - p.stage_off);
    const uint32_t stage_stride = static_cast<uint32_t>(p.stage_stride);
    const int stage_bufs = p.stage_bufs;
    int sb = 0;
    const bool store_thread = warp == 12 && lane == 0;
    int buf = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile / tiles_per_n;
      const int rem = tile - nt * tiles_per_n;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int oy = ty * TILE_H + h, ox = tx * TILE_W + w, n0 = nt * n_tile;
      const bool valid = (oy < p.Ho) && (ox < Wo) && !no_store;
      const long long pix = static_cast<long long>(oy) * Wo + ox;
      const uint32_t stage = stage_base + static_cast<uint32_t>(sb) * stage_stride, stage2 = stage + stage2_delta;
      if (use_tma) {
        // the staging tile is free once the TMA stores issued from it (1 or 2 tiles ago) have read it
        if (store_thread) {
          if (stage_bufs == 2) ptx::bulk_wait_read_1();
          else ptx::bulk_wait_read_all();
        }
        ptx::named_bar_sync(2, 256);
        if (++sb == stage_bufs) sb = 0;
      }
      ptx::mbar_wait(bar_tfull + 8 * buf, acc_ph);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 2 * n_tile);
      for (int n = 16 * eset; n < n_tile; n += 32) {
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
          ptx::tmem_ld_wait();
          if (live) {
            float4 *const o = reinterpret_cast<float4 *>(out + opix * out_pitch + c0);
            float4 *const o2 = out2 ? reinterpret_cast<float4 *>(out2 + opix * out2_pitch + c0) : nullptr;
            const float4 *const q1 = (res1 && valid) ? reinterpret_cast<const float4 *>(res1 + opix * res1_pitch + c0) : nullptr;
            const float4 *const q2 = (res2 && valid) ? reinterpret_cast<const float4 *>(res2 + opix * res2_pitch + c0) : nullptr;
            const float4 *const gq = (epi != LSSVC_EPI_PLAIN && valid) ? reinterpret_cast<const float4 *>(gdn_x + pix * gdn_pitch + cg) : nullptr;
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
                  const float4 t = q1[g];
                  v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                if (q2) {
                  const float4 t = q2[g];
                  v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                const uint32_t soff = srow + ((spiece + 16u * g) ^ sswz);
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
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);
      if (use_tma) {
        ptx::fence_proxy_async_smem();  // staging writes (generic proxy) -> visible to the TMA store (async proxy)
        ptx::named_bar_sync(3, 256);
        if (store_thread) {
          const int oy0 = ty * TILE_H, ox0 = tx * TILE_W;
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
      if (++buf == d_bufs) {
        buf = 0;
        acc_ph ^= 1u;
      }
    }
    if (use_tma && store_thread) ptx::bulk_wait_all();
  }

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

extern "C" int32_t lssvc_conv_h2(const lssvc_conv *c, void *stream) {
  LSSVC_REQUIRE(c != nullptr, "conv_h2: null descriptor");
  LSSVC_REQUIRE(c->n_src >= 1 && c->n_src <= LSSVC_MAX_SRC, "conv_h2: n_src=%d", c->n_src);
  LSSVC_REQUIRE(c->stride == 1 || c->stride == 2, "conv_h2: stride %d", c->stride);
  LSSVC_REQUIRE(c->kh >= 1 && c->kw >= 1 && c->kh * c->kw <= MAX_TAPS, "conv_h2: kernel %dx%d", c->kh, c->kw);
  LSSVC_REQUIRE(c->weight_h2 != nullptr && (reinterpret_cast<uintptr_t>(c->weight_h2) & 15) == 0,
                "conv_h2: needs the split fp16 weights (weight_h2)");
  LSSVC_REQUIRE(c->acc_scale > 0.f, "conv_h2: acc_scale %g", static_cast<double>(c->acc_scale));

  const int Hin = c->src[0].H, Win = c->src[0].W;
  int kc = 32;
  int cin16 = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    LSSVC_REQUIRE(lssvc::view_ok(&v), "conv_h2: bad source view %d", j);
    LSSVC_REQUIRE(v.H == Hin && v.W == Win, "conv_h2: source %d is %dx%d, expected %dx%d", j, v.H, v.W, Hin, Win);
    LSSVC_REQUIRE(v.C % 4 == 0, "conv_h2: source %d has %d channels (need a multiple of 4)", j, v.C);
    LSSVC_REQUIRE(v.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0,
                  "conv_h2: source %d is not 16-byte aligned", j);
    if (v.C % 32) kc = 16;
    cin16 += (v.C + 15) / 16 * 16;
  }
  LSSVC_REQUIRE(cin16 == c->cin_pad16, "conv_h2: cin_pad16 %d != %d from the sources", c->cin_pad16, cin16);
  const int st = c->stride;
  LSSVC_REQUIRE(Hin % st == 0 && Win % st == 0, "conv_h2: %dx%d not divisible by stride", Hin, Win);
  const int Ho = (Hin + 2 * c->pad - c->kh) / st + 1;
  const int Wo = (Win + 2 * c->pad - c->kw) / st + 1;
  const int ps = c->pixel_shuffle ? 2 : 1;
  LSSVC_REQUIRE(lssvc::view_ok(&c->out), "conv_h2: bad output view");
  LSSVC_REQUIRE(c->out.H == Ho * ps && c->out.W == Wo * ps, "conv_h2: output view %dx%d, expected %dx%d", c->out.H,
                c->out.W, Ho * ps, Wo * ps);
  LSSVC_REQUIRE(!c->pixel_shuffle || c->cout % 4 == 0, "conv_h2: pixel shuffle needs cout %% 4 == 0");
  const int c_store = c->pixel_shuffle ? c->cout / 4 : c->cout;
  LSSVC_REQUIRE(c->out.C == c_store, "conv_h2: output view has %d channels, expected %d", c->out.C, c_store);
  LSSVC_REQUIRE(c->n_pad % 16 == 0 && c->n_pad >= c->cout, "conv_h2: n_pad=%d cout=%d", c->n_pad, c->cout);
  LSSVC_REQUIRE((reinterpret_cast<uintptr_t>(c->bias) & 15) == 0, "conv_h2: bias not 16-byte aligned");

  // output-channel tiling: equal tiles of at most 128 channels
  int n_tile = c->n_pad;
  if (n_tile > 128) {
    n_tile = 128;
    while (n_tile >= 16 && (c->n_pad % n_tile)) n_tile -= 16;
    LSSVC_REQUIRE(n_tile >= 16, "conv_h2: cannot tile n_pad=%d", c->n_pad);
  }
  if (int rc = resolve_driver()) return rc;

  H2Params p;
  memset(&p, 0, sizeof(p));
  p.n_src = c->n_src;

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
  const int halo_h = TILE_H + (q1y - q0y), halo_w = TILE_W + (q1x - q0x);
  LSSVC_REQUIRE(halo_w <= 256 && halo_h <= 256, "conv_h2: halo %dx%d", halo_h, halo_w);
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
          p.tap_ry[n_taps] = static_cast<unsigned char>(qy - q0y);
          p.tap_rx[n_taps] = static_cast<unsigned char>(qx - q0x);
          p.tap_off[n_taps] = static_cast<unsigned short>((qy - q0y) * halo_w + (qx - q0x));
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
  LSSVC_REQUIRE(n_taps == c->kh * c->kw, "conv_h2: tap enumeration");

  // ---- tensor maps --------------------------------------------------------------------------------
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const int row_bytes = kc * 4;
  int coff = 0, total_chunks = 0;
  for (int j = 0; j < c->n_src; ++j) {
    const lssvc_view &v = c->src[j];
    p.chunks[j] = (v.C + kc - 1) / kc;
    total_chunks += p.chunks[j];
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
      lssvc::set_error("conv_h2: cuTensorMapEncodeTiled(A%d) failed with %d (C=%d pitch=%d %dx%d stride=%d halo %dx%d)", j,
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
      lssvc::set_error("conv_h2: cuTensorMapEncodeTiled(B) failed with %d (cin16=%d n_pad=%d taps=%d)",
                       static_cast<int>(r), cin16, c->n_pad, taps);
      return LSSVC_ERR_CUDA;
    }
  }

  // ---- pipeline geometry --------------------------------------------------------------------------
  {
    int per_chunk = 0;
    for (int g = 0; g < p.n_groups; ++g) {
      const int nt = p.g_tap0[g + 1] - p.g_tap0[g], tps_max = STAGE_K / kc;
      const int n_st = (nt + tps_max - 1) / tps_max;
      p.g_step[g] = (nt + n_st - 1) / n_st;
      per_chunk += (nt + p.g_step[g] - 1) / p.g_step[g];
    }
    p.stages_per_tile = total_chunks * per_chunk;
  }
  p.Ho = Ho; p.Wo = Wo;
  p.tiles_x = lssvc::ceil_div(Wo, TILE_W);
  p.tiles_y = lssvc::ceil_div(Ho, TILE_H);
  p.n_tiles = c->n_pad / n_tile;
  p.n_tile = n_tile;
  p.cout = c->cout;
  p.d_bufs = (4 * n_tile + 2 * STAGE_K <= TMEM_COLS) ? 2 : 1;
  p.a_col0 = p.d_bufs * 2 * n_tile;
  const int tps = STAGE_K / kc;
  p.slots = (TMEM_COLS - p.a_col0) / STAGE_K;
  if (p.slots > MAX_SLOTS) p.slots = MAX_SLOTS;
  p.halo_tx = halo_w * halo_h * row_bytes;
  p.halo_rows = halo_w * halo_h;
  p.halo_bytes = (p.halo_tx + 1023) & ~1023;
  p.b_bytes = tps * (2 * n_tile * kc * 2);  // n_tile % 16 == 0 keeps every tile 1024-byte aligned
  p.in_transform = c->in_transform;
  p.in_slope = c->in_slope;
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
  LSSVC_REQUIRE(opt(c->res1, &p.res1, &p.res1_pitch), "conv_h2: res1 shape mismatch");
  LSSVC_REQUIRE(opt(c->res2, &p.res2, &p.res2_pitch), "conv_h2: res2 shape mismatch");
  const float *o2 = nullptr;
  LSSVC_REQUIRE(opt(c->out2, &o2, &p.out2_pitch), "conv_h2: out2 shape mismatch");
  p.out2 = const_cast<float *>(o2);
  p.slope2 = c->slope2;
  if (c->epi != LSSVC_EPI_PLAIN) {
    LSSVC_REQUIRE(!c->pixel_shuffle && lssvc::view_ok(&c->gdn_x) && c->gdn_x.H == Ho && c->gdn_x.W == Wo &&
                      c->gdn_x.C == c->cout,
                  "conv_h2: GDN epilogue needs a matching gdn_x view");
    p.gdn_x = c->gdn_x.ptr;
    p.gdn_pitch = c->gdn_x.pitch;
    vec = vec && (c->gdn_x.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(c->gdn_x.ptr) & 15) == 0);
  }
  p.vec_ok = vec ? 1 : 0;

  // ---- TMA-store epilogue: the output tile is staged in shared memory as 128-byte (or 64-byte) swizzled rows
  // per slab of 32 (16) channels and written with cp.async.bulk.tensor stores (fully coalesced, clipped at the
  // image border and at the view's last channel).  Needs vectorisable views; with PixelShuffle a slab must not
  // straddle two sub-pixels.
  const int cq = c->cout / 4;
  int slab_w = 32;
  const bool no_tma_env = getenv("LSSVC_H2_NOTMA") != nullptr;  // A/B switch for the epilogue store path (tools/find_tma_bug.py)
  bool use_tma = vec && !no_tma_env && (!c->pixel_shuffle || cq % 16 == 0);
  // a slab must be fully owned by this tile's channels (n_tile % slab_w == 0) and, with PixelShuffle, by one sub-pixel
  if (n_tile % 32 != 0 || (c->pixel_shuffle && cq % 32 != 0)) slab_w = 16;
  int stage_bytes = 0;
  if (use_tma) {
    p.slab_w = slab_w;
    p.n_slabs = (n_tile + slab_w - 1) / slab_w;
    stage_bytes = p.n_slabs * 128 * slab_w * 4;
  }
  const int smem_budget = 224 * 1024;
  auto fits = [&](int halos, int slots, int stages) {
    return halos * p.halo_bytes + slots * p.b_bytes + stages * stage_bytes + 1024 <= smem_budget;
  };
  const int per_buf = use_tma ? (p.out2 ? 2 : 1) : 0;  // staging tiles per buffer (out, out2)
  // staging only when it leaves room for a healthy operand pipeline (>= 3 halo tiles, the TMEM-limited slots up to 4)
  const int want_slots = p.slots < 4 ? p.slots : 4;
  int n_stage = 0;
  p.stage_bufs = 0;
  if (per_buf && fits(4, want_slots, 2 * per_buf)) { n_stage = 2 * per_buf; p.stage_bufs = 2; }
  else if (per_buf && fits(3, want_slots, per_buf)) { n_stage = per_buf; p.stage_bufs = 1; }
  else use_tma = false;
  // as many halo tiles in flight as fit next to the weight ring: HBM latency x bandwidth wants > 64 KB per SM
  p.halo_bufs = MAX_HALO;
  while (p.halo_bufs > 3 && !fits(p.halo_bufs, want_slots, n_stage)) --p.halo_bufs;
  while (p.slots > 2 && !fits(p.halo_bufs, p.slots, n_stage)) --p.slots;
  while (p.halo_bufs > 2 && !fits(p.halo_bufs, p.slots, n_stage)) --p.halo_bufs;
  LSSVC_REQUIRE(p.slots >= 2 && fits(p.halo_bufs, p.slots, n_stage),
                "conv_h2: pipeline does not fit in shared memory (halo %d B x %d, weights %d B x %d, staging %d B x %d)",
                p.halo_bytes, p.halo_bufs, p.b_bytes, p.slots, stage_bytes, n_stage);
  p.stage_stride = per_buf * stage_bytes;
  p.use_tma = use_tma ? 1 : 0;
  p.stage_off = p.halo_bufs * p.halo_bytes + p.slots * p.b_bytes;
  p.stage2_off = p.stage_off + stage_bytes;
  if (use_tma) {
    auto make_out_map = [&](CUtensorMap *m, const lssvc_view &v) -> CUresult {
      const cuuint64_t px = static_cast<cuuint64_t>(v.pitch) * 4;
      const CUtensorMapSwizzle sw = slab_w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      if (c->pixel_shuffle) {
        // [2Ho][2Wo][cq] viewed as [Ho][2][Wo][2][cq]: one box per sub-pixel (i, j)
        const cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.C), 2, static_cast<cuuint64_t>(Wo), 2, static_cast<cuuint64_t>(Ho)};
        const cuuint64_t strides[4] = {px, px * 2, px * 2 * Wo, px * 2 * Wo * 2};
        const cuuint32_t box[5] = {static_cast<cuuint32_t>(slab_w), 1, TILE_W, 1, TILE_H};
        return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho)};
      const cuuint64_t strides[2] = {px, px * Wo};
      const cuuint32_t box[3] = {static_cast<cuuint32_t>(slab_w), TILE_W, TILE_H};
      return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = make_out_map(&p.out_map, c->out);
    if (r == CUDA_SUCCESS && p.out2) r = make_out_map(&p.out2_map, c->out2);
    if (r != CUDA_SUCCESS) {
      lssvc::set_error("conv_h2: cuTensorMapEncodeTiled(out) failed with %d (C=%d pitch=%d %dx%d ps=%d)", static_cast<int>(r),
                       c->out.C, c->out.pitch, Ho, Wo, c->pixel_shuffle);
      return LSSVC_ERR_CUDA;
    }
  }
  {
    // bottleneck isolation (tools/conv_bench.py): 1 no MMAs, 2 no halo reads, 4 no TMEM stores, 8 no global
    // stores, 16 no halo TMA, 32 no weight TMA.  Never set outside profiling: the output is garbage.
    static const int dbg = getenv("LSSVC_H2_DBG") ? atoi(getenv("LSSVC_H2_DBG")) : 0;
    p.dbg = dbg;
  }

  const int total_tiles = p.tiles_x * p.tiles_y * p.n_tiles;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  const size_t smem = static_cast<size_t>(p.halo_bufs) * p.halo_bytes + static_cast<size_t>(p.slots) * p.b_bytes +
                      static_cast<size_t>(n_stage) * stage_bytes + 1024;
  const int ki = kc == 32 ? 0 : 1;
  cudaStream_t s = lssvc::as_stream(stream);
  if (!g_attr_set[ki]) {
    const void *fn = kc == 32 ? reinterpret_cast<const void *>(conv_h2_kernel<32>)
                              : reinterpret_cast<const void *>(conv_h2_kernel<16>);
    LSSVC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    g_attr_set[ki] = true;
  }
  if (kc == 32) conv_h2_kernel<32><<<grid, NUM_THREADS, smem, s>>>(p);
  else conv_h2_kernel<16><<<grid, NUM_THREADS, smem, s>>>(p);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
