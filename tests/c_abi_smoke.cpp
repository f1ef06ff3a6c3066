// Header-only consumer of the C-ABI (include/lssvc_b200.h): what a maintainer binding the library from C / C++ / cgo / JNI
// would write — no Python, no torch.  Built and run by tests/test_abi.py.
//
//   c_abi_smoke host    (no GPU)  MLCodec_CXX.pmf_to_quantized_cdf (src/cpp/ops/ops.cpp:24-82) and a BufferedRansEncoder /
//                                 RansDecoder round trip incl. bypass symbols (src/cpp/rans/rans_interface.cpp:85-244),
//                                 argument validation of a device entry point
//   c_abi_smoke device  (B200)    INTEGRATION.md §3: nn.Conv2d(64, 64, 3, padding=1) + LeakyReLU(0.01) through lssvc_conv_hs with
//                                 weights packed HERE as the header documents (split fp16 of w * 2^shift, [taps][hi|lo][n_pad][cin16]),
//                                 checked against a double-precision loop; the range guard; the launch counter
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/lssvc_b200.h"

#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      fprintf(stderr, __VA_ARGS__); \
      fprintf(stderr, "\n");        \
      return 1;                     \
    }                               \
  } while (0)

static unsigned lcg(unsigned &s) { return s = s * 1664525u + 1013904223u; }
static float frand(unsigned &s) { return (lcg(s) >> 8) * (1.0f / 16777216.0f) * 2.f - 1.f; }

static int host_checks() {
  CHECK(lssvc_abi_version() == 4, "abi version %d", lssvc_abi_version());
  // ---- pmf_to_quantized_cdf: sums to 2^16, strictly increasing (every symbol keeps a non-zero frequency)
  const int n = 40;
  std::vector<float> pmf(n);
  unsigned s = 1;
  float tot = 0;
  for (auto &p : pmf) tot += (p = fabsf(frand(s)) + 1e-3f);
  for (auto &p : pmf) p /= tot;
  pmf[5] = pmf[6] = 1e-9f;
  std::vector<uint32_t> cdf(n + 1);
  CHECK(lssvc_pmf_to_quantized_cdf(pmf.data(), n, 16, cdf.data()) == LSSVC_OK, "pmf_to_quantized_cdf: %s", lssvc_last_error());
  CHECK(cdf[0] == 0 && cdf[n] == 65536u, "cdf ends %u %u", cdf[0], cdf[n]);
  for (int i = 0; i < n; ++i) CHECK(cdf[i + 1] > cdf[i], "symbol %d has zero frequency", i);
  // ---- rANS round trip: 2 table rows, symbols inside and outside the table (bypass), two pushes into one stream
  const int rows = 2, stride = n + 2;
  std::vector<int32_t> table(rows * stride, 0), sizes = {n + 1 + 1, n + 1 + 1}, offsets = {-20, -10};
  for (int r = 0; r < rows; ++r)
    for (int i = 0; i <= n; ++i) table[r * stride + i] = static_cast<int32_t>(cdf[i]);
  // the last symbol of a row is the escape symbol: sizes = pmf length + 2 (CompressAI convention)
  for (int r = 0; r < rows; ++r) sizes[r] = n + 1;
  const int N = 5000;
  std::vector<int32_t> sym(N), idx(N), out(N);
  for (int i = 0; i < N; ++i) {
    idx[i] = lcg(s) & 1;
    sym[i] = static_cast<int32_t>(lcg(s) % 36) + offsets[idx[i]];
    if (i % 97 == 0) sym[i] = static_cast<int32_t>(lcg(s) % 200000) - 100000;
  }
  lssvc_rans_encoder *enc = lssvc_rans_encoder_new();
  CHECK(lssvc_rans_encode_with_indexes(enc, sym.data(), idx.data(), N, table.data(), rows, stride, sizes.data(), offsets.data()) == LSSVC_OK,
        "encode: %s", lssvc_last_error());
  CHECK(lssvc_rans_encode_with_indexes(enc, sym.data(), idx.data(), 100, table.data(), rows, stride, sizes.data(), offsets.data()) == LSSVC_OK,
        "encode 2: %s", lssvc_last_error());
  const uint8_t *bytes = nullptr;
  const int64_t nbytes = lssvc_rans_encoder_flush(enc, &bytes);
  CHECK(nbytes > 8 && nbytes % 4 == 0, "flush returned %lld", static_cast<long long>(nbytes));
  lssvc_rans_decoder *dec = lssvc_rans_decoder_new();
  CHECK(lssvc_rans_decoder_set_stream(dec, bytes, nbytes) == LSSVC_OK, "set_stream: %s", lssvc_last_error());
  CHECK(lssvc_rans_decode_stream(dec, idx.data(), N, table.data(), rows, stride, sizes.data(), offsets.data(), out.data()) == LSSVC_OK,
        "decode: %s", lssvc_last_error());
  CHECK(memcmp(out.data(), sym.data(), N * sizeof(int32_t)) == 0, "decoded symbols differ");
  CHECK(lssvc_rans_decode_stream(dec, idx.data(), 100, table.data(), rows, stride, sizes.data(), offsets.data(), out.data()) == LSSVC_OK,
        "decode 2: %s", lssvc_last_error());
  CHECK(memcmp(out.data(), sym.data(), 100 * sizeof(int32_t)) == 0, "second push decodes differently");
  // out-of-table row -> error, not a wild read
  int32_t bad = 7;
  CHECK(lssvc_rans_encode_with_indexes(enc, sym.data(), &bad, 1, table.data(), rows, stride, sizes.data(), offsets.data()) == LSSVC_ERR_ARG,
        "a row outside the table must be rejected");
  lssvc_rans_encoder_free(enc);
  lssvc_rans_decoder_free(dec);
  // ---- argument validation happens before any CUDA call
  CHECK(lssvc_conv_hs(nullptr, nullptr) == LSSVC_ERR_ARG && strstr(lssvc_last_error(), "null descriptor"), "conv_hs(NULL)");
  printf("host: pmf_to_quantized_cdf, rANS round trip (%d symbols -> %lld bytes), argument checks OK\n", N, static_cast<long long>(nbytes));
  return 0;
}

static int device_checks() {
  CHECK(lssvc_device_check(0) == LSSVC_OK, "device_check: %s", lssvc_last_error());
  const int H = 40, W = 56, C = 64, K = 3, T = K * K;
  unsigned s = 3;
  std::vector<float> x(H * W * C), w(C * C * T), b(C);  // x NHWC, w [cout][cin][kh][kw] as nn.Conv2d stores it
  for (auto &v : x) v = frand(s);
  for (auto &v : w) v = frand(s) / 24.f;
  for (auto &v : b) v = frand(s) * 0.1f;
  // ---- weight repacking documented at lssvc_conv.weight_h2: fp16 [taps][2 (hi, lo)][n_pad][cin_pad16] of w * 2^shift
  float wmax = 0;
  for (float v : w) wmax = fmaxf(wmax, fabsf(v));
  const int shift = 13 - static_cast<int>(floorf(log2f(wmax)));
  std::vector<__half> wh(T * 2 * C * C);
  for (int t = 0; t < T; ++t)
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < C; ++ci) {
        const float v = ldexpf(w[(co * C + ci) * T + t], shift);
        const __half hi = __float2half_rn(v);
        wh[((t * 2 + 0) * C + co) * C + ci] = hi;
        wh[((t * 2 + 1) * C + co) * C + ci] = __float2half_rn(v - __half2float(hi));
      }
  float *x_d, *y_d, *b_d;
  void *w_d;
  double *flag_d;
  cudaMalloc(&x_d, x.size() * 4);
  cudaMalloc(&y_d, x.size() * 4);
  cudaMalloc(&b_d, C * 4);
  cudaMalloc(&w_d, wh.size() * 2);
  cudaMalloc(&flag_d, 8);
  cudaMemcpy(x_d, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(b_d, b.data(), C * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(w_d, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
  cudaStream_t stream;
  cudaStreamCreate(&stream);

  lssvc_conv c;
  memset(&c, 0, sizeof(c));
  c.n_src = 1;
  c.src[0] = {x_d, H, W, C, C};
  c.kh = c.kw = K; c.stride = 1; c.pad = 1; c.cout = C; c.n_pad = C; c.cin_total = C;
  c.bias = b_d;
  c.weight_h2 = w_d; c.cin_pad16 = C; c.acc_scale = ldexpf(1.f, -shift);
  c.precision = LSSVC_PREC_H2; c.act = LSSVC_ACT_LRELU; c.slope = 0.01f; c.out_scale = 1.f;
  c.out = {y_d, H, W, C, C};
  const int64_t l0 = lssvc_launch_count();
  CHECK(lssvc_conv_hs(&c, stream) == LSSVC_OK, "conv_hs: %s", lssvc_last_error());
  CHECK(lssvc_range_flag_fetch(flag_d, stream) == LSSVC_OK, "range_flag_fetch: %s", lssvc_last_error());
  CHECK(cudaStreamSynchronize(stream) == cudaSuccess, "kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  CHECK(lssvc_launch_count() == l0 + 2, "launch counter");
  std::vector<float> y(x.size());
  double flag = -1;
  cudaMemcpy(y.data(), y_d, y.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&flag, flag_d, 8, cudaMemcpyDeviceToHost);
  CHECK(flag == 0.0, "range flag %g", flag);
  double worst = 0, scale = 0;
  for (int oy = 0; oy < H; ++oy)
    for (int ox = 0; ox < W; ++ox)
      for (int co = 0; co < C; ++co) {
        double acc = b[co];
        for (int r = 0; r < K; ++r)
          for (int q = 0; q < K; ++q) {
            const int iy = oy + r - 1, ix = ox + q - 1;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
            const float *px = &x[(iy * W + ix) * C];
            for (int ci = 0; ci < C; ++ci) acc += static_cast<double>(px[ci]) * w[(co * C + ci) * T + r * K + q];
          }
        const double ref = acc > 0 ? acc : 0.01 * acc;
        worst = fmax(worst, fabs(ref - y[(oy * W + ox) * C + co]));
        scale = fmax(scale, fabs(ref));
      }
  printf("device: conv3x3 64->64 + LeakyReLU through lssvc_conv_hs from C++: max error %.2e of the output scale\n", worst / scale);
  // packed here WITHOUT the accumulator-truncation compensation of ops.acc_comp: the bound is the uncompensated one
  CHECK(worst / scale < 5e-6, "conv_hs deviates: %.3e", worst / scale);
  return 0;
}

int main(int argc, char **argv) {
  const bool device = argc > 1 && strcmp(argv[1], "device") == 0;
  if (int rc = host_checks()) return rc;
  if (device)
    if (int rc = device_checks()) return rc;
  printf("c_abi_smoke OK\n");
  return 0;
}
