/* ORACLE — test infrastructure only.  Plain-C restatement of the reference's entropy coder:
 *   - pmf_to_quantized_cdf           /root/reference/src/cpp/ops/ops.cpp:24-82
 *   - BufferedRansEncoder            /root/reference/src/cpp/rans/rans_interface.cpp:85-172
 *   - RansDecoder                    /root/reference/src/cpp/rans/rans_interface.cpp:176-244
 * on top of the restated rans64.h (third-party ryg_rans, pinned commit, see third_party/rans64.h).
 * The reference has no known-answer vectors for the bitstream ("parity unpinned" by its own tests); this file
 * is pinned byte-for-byte against the reference's own sources compiled into oracle/_ref (tests/test_rans.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rans64.h"

#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)

typedef struct { uint16_t start, range; int bypass; } sym_t;

/* ops.cpp:24-82 */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  const int size = n + 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)(roundf(pmf[i] * (1 << precision)) + 0.5);
  int total_i = 0; /* std::accumulate with an int initial value */
  for (int i = 0; i < size; ++i) total_i = (int)((uint32_t)total_i + cdf[i]);
  const uint32_t total = (uint32_t)total_i;
  for (int i = 0; i < size; ++i) cdf[i] = (uint32_t)(((1ull << precision) * cdf[i]) / total);
  for (int i = 1; i < size; ++i) cdf[i] += cdf[i - 1];
  cdf[size - 1] = 1u << precision;
  for (int i = 0; i < size - 1; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u;
      int best_steal = -1;
      for (int j = 0; j < size - 1; ++j) {
        uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
      }
      if (best_steal < 0) return -1;
      if (best_steal < i) { for (int j = best_steal + 1; j <= i; ++j) cdf[j]--; }
      else { for (int j = i + 1; j <= best_steal; ++j) cdf[j]++; }
    }
  }
  return 0;
}

/* rans_interface.cpp:46-64 */
static void enc_put_bits(Rans64State *r, uint32_t **pptr, uint32_t val, uint32_t nbits) {
  uint64_t x = *r;
  uint32_t freq = 1u << (16 - nbits);
  uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
  if (x >= x_max) { *pptr -= 1; **pptr = (uint32_t)x; x >>= 32; }
  *r = (x << nbits) | val;
}

/* rans_interface.cpp:66-82 */
static uint32_t dec_get_bits(Rans64State *r, uint32_t **pptr, uint32_t n_bits) {
  uint64_t x = *r;
  uint32_t val = (uint32_t)(x & ((1u << n_bits) - 1));
  x = x >> n_bits;
  if (x < RANS64_L) { x = (x << 32) | **pptr; *pptr += 1; }
  *r = x;
  return val;
}

/* encode_with_indexes + flush in one call (rans_interface.cpp:85-172).
 * Returns the number of bytes written to out (capacity cap bytes), or -1. */
long oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, long n, const int32_t *cdfs, int cdf_stride,
                        const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out, long cap) {
  sym_t *syms = (sym_t *)malloc(sizeof(sym_t) * (size_t)(n * 12 + 4));
  long ns = 0;
  for (long i = 0; i < n; ++i) {
    const int32_t cdf_idx = indexes[i];
    const int32_t *cdf = cdfs + (long)cdf_idx * cdf_stride;
    const int32_t max_value = cdf_sizes[cdf_idx] - 2;
    int32_t value = symbols[i] - offsets[cdf_idx];
    uint32_t raw_val = 0;
    if (value < 0) { raw_val = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw_val = (uint32_t)(2 * (value - max_value)); value = max_value; }
    syms[ns++] = (sym_t){(uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), 0};
    if (value == max_value) {
      int32_t n_bypass = 0;
      while ((raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= MAX_BYPASS_VAL) { syms[ns++] = (sym_t){MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, 1}; val -= MAX_BYPASS_VAL; }
      syms[ns++] = (sym_t){(uint16_t)val, (uint16_t)(val + 1), 1};
      for (int32_t j = 0; j < n_bypass; ++j) {
        const int32_t v1 = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
        syms[ns++] = (sym_t){(uint16_t)v1, (uint16_t)(v1 + 1), 1};
      }
    }
  }
  uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ns + 2));
  uint32_t *end = buf + ns + 2, *ptr = end;
  Rans64State rans;
  Rans64EncInit(&rans);
  while (ns > 0) {
    const sym_t s = syms[--ns];
    if (!s.bypass) Rans64EncPut(&rans, &ptr, s.start, s.range, PRECISION);
    else enc_put_bits(&rans, &ptr, s.start, BYPASS_PRECISION);
  }
  Rans64EncFlush(&rans, &ptr);
  long nbytes = (long)(end - ptr) * 4;
  long rc = -1;
  if (nbytes <= cap) { memcpy(out, ptr, (size_t)nbytes); rc = nbytes; }
  free(buf);
  free(syms);
  return rc;
}

/* set_stream + decode_stream (rans_interface.cpp:176-244) */
int oracle_rans_decode(const uint8_t *stream, long nbytes, const int32_t *indexes, long n, const int32_t *cdfs,
                       int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int32_t *out) {
  uint32_t *words = (uint32_t *)malloc((size_t)nbytes + 64);
  memset(words, 0, (size_t)nbytes + 64);
  memcpy(words, stream, (size_t)nbytes);
  uint32_t *ptr = words;
  Rans64State rans;
  Rans64DecInit(&rans, &ptr);
  for (long i = 0; i < n; ++i) {
    const int32_t cdf_idx = indexes[i];
    const int32_t *cdf = cdfs + (long)cdf_idx * cdf_stride;
    const int32_t max_value = cdf_sizes[cdf_idx] - 2;
    const int32_t offset = offsets[cdf_idx];
    const uint32_t cum_freq = Rans64DecGet(&rans, PRECISION);
    int32_t s = 0;
    while (s < cdf_sizes[cdf_idx] && (uint32_t)cdf[s] <= cum_freq) ++s; /* find_if(v > cum_freq) */
    s -= 1;
    Rans64DecAdvance(&rans, &ptr, (uint32_t)cdf[s], (uint32_t)(cdf[s + 1] - cdf[s]), PRECISION);
    int32_t value = s;
    if (value == max_value) {
      int32_t val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
      int32_t n_bypass = val;
      while (val == MAX_BYPASS_VAL) { val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION); n_bypass += val; }
      int32_t raw_val = 0;
      for (int j = 0; j < n_bypass; ++j) {
        val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
        raw_val |= val << (j * BYPASS_PRECISION);
      }
      value = raw_val >> 1;
      if (raw_val & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offset;
  }
  free(words);
  return 0;
}
