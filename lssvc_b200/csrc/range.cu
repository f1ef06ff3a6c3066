// Range guard of the split-fp16 arithmetic (conv_hs / conv_pw / conv_ffn).
//
// Activations are split x = rn_f16(x) + rn_f16(x - rn_f16(x)) with no per-tensor scale (a scale would have to be common to
// the whole K accumulation of a layer, i.e. known before the layer runs).  fp16 tops out at 65504: an operand beyond it
// (|x| >= 65520, or x^2 >= 65520 under LSSVC_IN_SQUARE, i.e. a GDN input beyond 255.9) turns into inf - inf = NaN inside
// the tensor core.  The kernels therefore keep the running max |operand| of what they split and raise ONE device-side
// flag when it leaves the range; the host fetches (and clears) the flag with the frame's bit counters and re-codes the
// frame on the fp32 engine (lssvc_b200/models.py) — a defined behaviour instead of silent NaNs, at one FMNMX per element.
// At the low end nothing overflows: operands below 2^-14 lose their lo term to fp16 subnormals gradually (absolute error
// <= 2^-25 per operand), which the parity tests bound (tests/test_kernels_gpu.py::test_split_fp16_range).
#include "common.cuh"

namespace lssvc {

unsigned int *range_flag() {
  static unsigned int *flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(unsigned int));
  }
  return flag;
}

}  // namespace lssvc

namespace {
__global__ void range_flag_fetch_kernel(unsigned int *flag, double *dst) {
  dst[0] = static_cast<double>(*flag);
  *flag = 0u;
}
}  // namespace

extern "C" int32_t lssvc_range_flag_fetch(double *dst, void *stream) {
  LSSVC_REQUIRE(dst != nullptr, "range_flag_fetch: null destination");
  unsigned int *flag = lssvc::range_flag();
  if (flag == nullptr) {
    cudaGetLastError();
    lssvc::set_error("range_flag_fetch: no CUDA device (cannot allocate the device flag)");
    return LSSVC_ERR_NO_DEVICE;
  }
  range_flag_fetch_kernel<<<1, 1, 0, lssvc::as_stream(stream)>>>(flag, dst);
  LSSVC_LAUNCHED();
  return LSSVC_OK;
}
