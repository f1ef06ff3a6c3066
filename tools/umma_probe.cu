// tcgen05.mma throughput probe (study tool, not part of the library): one thread per CTA issues a repeating pattern of up
// to 4 MMAs (kind::f16, M = 128, K = 16) and the cycles per pattern repetition are reported.  Answers, for the conv_hs
// operand scheme: what one MMA costs as a function of N, of the A source (shared memory vs TMEM), of the
// number of independent accumulators and of the issue order.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/umma_probe tools/umma_probe.cu && tools/_bin/umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../lssvc_b200/csrc/ptx.cuh"

struct Op {
  uint32_t d_col;   // accumulator column
  uint32_t a16;     // A start (bytes >> 4) inside the A region (SS) or TMEM column (TS)
  uint32_t b16;     // B start (bytes >> 4) inside the B region
  uint32_t n;       // N of the MMA
  uint32_t ts;      // 0: MMA with A from shared memory, 1: MMA with A from TMEM, 2: tcgen05.cp 128x256b smem -> TMEM column d_col
};
struct Cfg {
  Op op[8];
  int n_ops;
  int iters;
  uint32_t a_sbo, a_layout, a_lbo;  // A descriptor: SBO bytes, layout code, LBO bytes (no-swizzle only)
  int traffic;               // 1: the other warps stream shared memory (LDS.128 + STS.128) while the MMAs run
};

constexpr uint32_t A_BYTES = 48 * 1024, B_BYTES = 16 * 1024, X_BYTES = 64 * 1024;

template <int NOPS>
__device__ __forceinline__ void issue_loop(const Cfg &c, uint32_t tmem, uint32_t a_base, uint32_t b_base) {
  uint32_t d[NOPS], a[NOPS], b[NOPS], id[NOPS], ts[NOPS];
  const uint32_t a_hi = (c.a_sbo >> 4) | (1u << 14) | (c.a_layout << 29);
  const uint32_t b_hi = (512u >> 4) | (1u << 14) | (4u << 29);  // SWIZZLE_64B rows of 64 bytes
#pragma unroll
  for (int i = 0; i < NOPS; ++i) {
    d[i] = tmem + c.op[i].d_col;
    ts[i] = c.op[i].ts;
    a[i] = ts[i] == 1 ? tmem + c.op[i].a16 : (((a_base >> 4) + c.op[i].a16) | ((c.a_lbo >> 4) << 16));
    b[i] = ((b_base >> 4) + c.op[i].b16) | (1u << 16);
    id[i] = ptx::make_idesc_f16_m128(c.op[i].n);
  }
  for (int it = 0; it < c.iters; ++it) {
#pragma unroll
    for (int i = 0; i < NOPS; ++i) {
      if (ts[i] == 2) {
        asm volatile("{\n.reg .b64 da;\nmov.b64 da, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], da;\n}\n" ::"r"(d[i]), "r"(a[i]), "r"(a_hi) : "memory");
      } else if (ts[i]) ptx::mma_f16_ts2(d[i], a[i], b[i], b_hi, id[i], 1u);
      else ptx::mma_f16_ss2(d[i], a[i], a_hi, b[i], b_hi, id[i], 1u);
    }
  }
}

__global__ void __launch_bounds__(256, 1) probe_kernel(const Cfg c, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  for (uint32_t i = threadIdx.x * 16; i < A_BYTES + B_BYTES + X_BYTES; i += blockDim.x * 16) ptx::sts_u4(base + i, 0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
    done = 0;
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp >= 4) {  // zero the TMEM columns used as TS operands / accumulators
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0;
    for (int col = 0; col < 512; col += 16) ptx::tmem_st16(tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + col, z);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) {
    if (ptx::elect_one()) {
      const long long t0 = clock64();
      switch (c.n_ops) {
        case 1: issue_loop<1>(c, tmem, base, base + A_BYTES); break;
        case 2: issue_loop<2>(c, tmem, base, base + A_BYTES); break;
        case 3: issue_loop<3>(c, tmem, base, base + A_BYTES); break;
        case 4: issue_loop<4>(c, tmem, base, base + A_BYTES); break;
        case 6: issue_loop<6>(c, tmem, base, base + A_BYTES); break;
        default: issue_loop<8>(c, tmem, base, base + A_BYTES); break;
      }
      const long long t1 = clock64();
      ptx::mma_commit(ptx::smem_u32(&bar));
      ptx::mbar_wait(ptx::smem_u32(&bar), 0);
      const long long t2 = clock64();
      done = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
    __syncwarp();
  } else if (warp >= 2 && c.traffic) {
    // background shared-memory traffic in a separate region: 6 warps of LDS.128 + STS.128
    const uint32_t x = base + A_BYTES + B_BYTES + (threadIdx.x - 64) * 16;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    while (!done) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = ptx::lds_f4(x + k * 3072 * 2);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      ptx::sts_u4(x, __float_as_uint(acc.x), __float_as_uint(acc.y), __float_as_uint(acc.z), __float_as_uint(acc.w));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

static long long *g_out;
static int g_sms;

static void run(const char *name, Cfg c) {
  c.iters = 4000;
  long long h[2];
  float ms = 0.f;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    probe_kernel<<<g_sms, 256, A_BYTES + B_BYTES + X_BYTES + 1024>>>(c, g_out);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-58s FAILED: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaMemcpy(h, g_out, sizeof(h), cudaMemcpyDeviceToHost);
  double floor_c = 0;
  for (int i = 0; i < c.n_ops; ++i) floor_c += c.op[i].ts == 2 ? 0.0 : c.op[i].n / 2.0;
  printf("%-58s issue %7.1f  done %7.1f cyc/rep   math floor %6.1f   (%.3f ms, %.2f GHz)\n", name, double(h[0]) / c.iters,
         double(h[1]) / c.iters, floor_c, ms, double(h[1]) / (ms * 1e6));
  fflush(stdout);
}

static Cfg base_cfg() {
  Cfg c;
  memset(&c, 0, sizeof(c));
  c.a_sbo = 1024;
  c.a_layout = 2;
  c.a_lbo = 16;
  return c;
}
static Op op(uint32_t d, uint32_t a16, uint32_t b16, uint32_t n, uint32_t ts = 0) { return Op{d, a16, b16, n, ts}; }

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&g_out, 64);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES + X_BYTES + 1024);
  char name[128];
  // (a) one MMA per repetition, one accumulator, N sweep, A from shared memory / from TMEM
  for (int ts = 0; ts < 2; ++ts) {
    for (uint32_t n : {16u, 64u, 128u, 256u}) {
      Cfg c = base_cfg();
      c.n_ops = 1;
      c.op[0] = op(0, ts ? 480 : 0, 0, n, ts);
      snprintf(name, sizeof(name), "%s  N=%-3u 1 acc", ts ? "TS" : "SS", n);
      run(name, c);
    }
  }
  // (b) tcgen05.cp alone: 1, 2, 4 copies of a 128-row x 32-byte slice per repetition
  for (int k : {1, 2, 4}) {
    Cfg c = base_cfg();
    c.a_sbo = 2304;
    c.n_ops = k;
    for (int i = 0; i < k; ++i) c.op[i] = op(448 + 8 * i, 9 + 2 * i, 0, 0, 2);
    snprintf(name, sizeof(name), "CP  128x256b x%d", k);
    run(name, c);
  }
  // (c) the conv pair with the A slices copied to TMEM by tcgen05.cp right before the MMAs that read them
  for (uint32_t n : {32u, 64u, 96u}) {
    for (int variant = 0; variant < 3; ++variant) {
      Cfg c = base_cfg();
      c.a_sbo = 2304;
      const Op cph0 = op(448, 9, 0, 0, 2), cpl0 = op(456, 13, 0, 0, 2), cph1 = op(464, 9 + 64, 0, 0, 2), cpl1 = op(472, 13 + 64, 0, 0, 2);
      const Op hi0 = op(0, 448, 0, 2 * n, 1), lo0 = op(n, 456, 0, n, 1), hi1 = op(2 * n, 464, 0, 2 * n, 1), lo1 = op(3 * n, 472, 0, n, 1);
      c.n_ops = 8;
      const char *vn;
      if (variant == 0) { Op o[8] = {cph0, hi0, cpl0, lo0, cph1, hi1, cpl1, lo1}; memcpy(c.op, o, sizeof(o)); vn = "cp mma cp mma .."; }
      else if (variant == 1) { Op o[8] = {cph0, cpl0, cph1, cpl1, hi0, lo0, hi1, lo1}; memcpy(c.op, o, sizeof(o)); vn = "4 cp then 4 mma (same slots)"; }
      else { Op o[8] = {hi0, lo0, hi1, lo1, cph0, cpl0, cph1, cpl1}; memcpy(c.op, o, sizeof(o)); vn = "4 mma then 4 cp (same slots)"; }
      snprintf(name, sizeof(name), "CP+TS pair n=%-3u x2: %s", n, vn);
      run(name, c);
    }
    {  // double-buffered operand slots: the copies of repetition i+1 do not touch the slots the MMAs of repetition i read
      Cfg c = base_cfg();
      c.a_sbo = 2304;
      c.n_ops = 8;
      Op o[8] = {op(448, 9, 0, 0, 2), op(456, 13, 0, 0, 2), op(0, 480, 0, 2 * n, 1), op(n, 488, 0, n, 1),
                 op(480, 9 + 64, 0, 0, 2), op(488, 13 + 64, 0, 0, 2), op(2 * n, 448, 0, 2 * n, 1), op(3 * n, 456, 0, n, 1)};
      memcpy(c.op, o, sizeof(o));
      snprintf(name, sizeof(name), "CP+TS pair n=%-3u x2: cp ahead, alternating slots", n);
      run(name, c);
    }
  }
  return 0;
}
