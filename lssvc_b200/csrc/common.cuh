// Shared host-side plumbing of the C-ABI library: status/err string, launch counter, view checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lssvc_b200.h"

namespace lssvc {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

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

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace lssvc
