// Shared host-side plumbing of the C-ABI library: status/err string, launch counter, view checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lssvc_b200.h"

namespace lssvc {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// device word raised by the split-fp16 kernels when an operand leaves the fp16 range (range.cu); nullptr if unavailable
unsigned int *range_flag();
// operands at or beyond this magnitude round to inf in fp16
constexpr float kSplitRangeLimit = 65520.f;

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool view_ok(const lssvc_view *v) {
  return v && v->ptr && v->H > 0 && v->W > 0 && v->C > 0 && v->pitch >= v->C;
}
inline bool view_present(const lssvc_view *v) { return v && v->ptr != nullptr; }

#define LSSVC_REQUIRE(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      lssvc::set_error(__VA_ARGS__);    \
      return LSSVC_ERR_ARG;             \
    }                                   \
  } while (0)

#define LSSVC_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      lssvc::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return LSSVC_ERR_CUDA;                                                      \
    }                                                                             \
  } while (0)

// call after a <<<>>> launch
#define LSSVC_LAUNCHED()                                                          \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      lssvc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,   \
                       cudaGetErrorString(e__));                                  \
      return LSSVC_ERR_CUDA;                                                      \
    }                                                                             \
    lssvc::count_launch();                                                        \
  } while (0)

// Launch with programmatic stream serialization (see ptx::pdl_wait): the kernel may start its prologue while its predecessor
// in the stream drains.  Not used while the stream is being captured into a CUDA graph; LSSVC_NO_PDL=1 switches it off (A/B).
template <typename P>
inline cudaError_t launch_pdl(void (*kernel)(const P), int grid, int block, size_t smem, cudaStream_t s, const P &p) {
  static const bool off = getenv("LSSVC_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  bool capturing = true;
  if (cudaStreamIsCapturing(s, &cap) == cudaSuccess) capturing = cap != cudaStreamCaptureStatusNone;
  else cudaGetLastError();  // (the query itself failed: launch plainly and let the launch report what is wrong)
  cfg.attrs = attr;
  cfg.numAttrs = (off || capturing) ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// The tensor core adds every K = 16 MMA into the fp32 TMEM accumulator with truncation (round towards zero), so a sum of
// `steps` MMAs comes out SMALLER in magnitude than the exact sum by a nearly deterministic factor: measured on the 15 layer
// shapes of tools/conv_bench.py (CONV_BENCH_BIAS=1, profiles/r1_accumulation_bias.txt) the mean signed relative error is
// -(0.264 * steps + 0.6) * 2^-24 for steps = 4 .. 196, and it — not the split-fp16 operands — was ~85 % of the rms error
// (3x3 64->64: 7.2e-7 against 4.3e-7 for fp32 CUDA cores), coherent from layer to layer.  The kernels undo its expected
// value by folding (1 + that factor) into the accumulator scale they apply anyway.  LSSVC_ACC_COMP=0 disables it (A/B).
inline float acc_comp(int steps) {
  static const bool off = [] {
    const char *e = getenv("LSSVC_ACC_COMP");
    return e != nullptr && atoi(e) == 0;
  }();
  if (off || steps <= 0) return 1.f;
  return 1.f + (0.264f * static_cast<float>(steps) + 0.6f) * 5.9604645e-8f;
}

}  // namespace lssvc
