// Host-side range-ANS entropy coder, format-identical to the reference's MLCodec_rans
// (src/cpp/rans/rans_interface.cpp) and MLCodec_CXX.pmf_to_quantized_cdf (src/cpp/ops/ops.cpp).
//
// Stream format: 64-bit rANS state (lower bound 2^31, 32-bit renormalisation words, stream is an
// array of little-endian uint32), 16-bit probabilities.  A symbol is coded as v = sym - offset[row];
// v outside [0, cdf_size - 2) is coded as the escape symbol cdf_size - 2 followed by 4-bit "bypass"
// digits: a unary-of-15s digit count, then the digits of raw = -2v - 1 (v < 0) or 2 (v - max) little
// end first.  The encoder buffers (start, range) pairs and codes them in reverse on flush, so the
// decoder reads forward.  The state transitions follow the public-domain rans64 scheme by F. Giesen
// (ryg_rans @ c9d162d996fd600315af9ae8eb89d832576cb32d, the reference's un-vendored dependency).
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <numeric>
#include <vector>

#include "../../include/lssvc_b200.h"

namespace lssvc {
void set_error(const char *fmt, ...);
}

namespace {

constexpr uint32_t kProbBits = 16;
constexpr uint32_t kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;
constexpr uint64_t kLower = 1ull << 31;

struct Tok {
  uint16_t start;
  uint16_t range;
  uint16_t bypass;
};

inline void enc_renorm(uint64_t &x, uint32_t *&out, uint64_t x_max) {
  if (x >= x_max) {
    *--out = static_cast<uint32_t>(x);
    x >>= 32;
  }
}

// Every CDF row an index names must exist and fit its stride (a bad row would read outside cdfs / cdf_sizes / offsets).
bool rows_ok(const char *who, const int32_t *indexes, int64_t n, int32_t n_rows, int32_t cdf_stride, const int32_t *cdf_sizes) {
  for (int32_t r = 0; r < n_rows; ++r) {
    if (cdf_sizes[r] < 2 || cdf_sizes[r] > cdf_stride) {
      lssvc::set_error("%s: cdf_sizes[%d] = %d outside [2, cdf_stride = %d]", who, r, cdf_sizes[r], cdf_stride);
      return false;
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    if (indexes[i] < 0 || indexes[i] >= n_rows) {
      lssvc::set_error("%s: indexes[%lld] = %d outside [0, n_rows = %d)", who, static_cast<long long>(i), indexes[i], n_rows);
      return false;
    }
  }
  return true;
}

}  // namespace

struct lssvc_rans_encoder {
  std::vector<Tok> toks;
  std::vector<uint32_t> words;
  const uint8_t *data = nullptr;
  int64_t nbytes = 0;
};

struct lssvc_rans_decoder {
  std::vector<uint32_t> words;
  size_t pos = 0;
  uint64_t state = 0;
  bool underflow = false;

  uint32_t next_word() {
    if (pos >= words.size()) {
      underflow = true;
      return 0;
    }
    return words[pos++];
  }
  uint32_t get_bits(uint32_t nbits) {
    uint64_t x = state;
    const uint32_t val = static_cast<uint32_t>(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < kLower) x = (x << 32) | next_word();
    state = x;
    return val;
  }
};

extern "C" {

int32_t lssvc_pmf_to_quantized_cdf(const float *pmf, int32_t n, int32_t precision, uint32_t *cdf) {
  if (!pmf || !cdf || n <= 0 || precision <= 0 || precision > 30) {
    lssvc::set_error("pmf_to_quantized_cdf: bad arguments");
    return LSSVC_ERR_ARG;
  }
  const int size = n + 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) {
    cdf[i + 1] = static_cast<uint32_t>(std::round(pmf[i] * static_cast<float>(1 << precision)) + 0.5);
  }
  // the reference accumulates with an int initial value: wrap-around 32-bit arithmetic
  int total_i = 0;
  for (int i = 0; i < size; ++i) total_i = static_cast<int>(static_cast<uint32_t>(total_i) + cdf[i]);
  const uint32_t total = static_cast<uint32_t>(total_i);
  if (total == 0) {
    lssvc::set_error("pmf_to_quantized_cdf: empty pmf");
    return LSSVC_ERR_ARG;
  }
  for (int i = 0; i < size; ++i) cdf[i] = static_cast<uint32_t>(((1ull << precision) * cdf[i]) / total);
  for (int i = 1; i < size; ++i) cdf[i] += cdf[i - 1];
  cdf[size - 1] = 1u << precision;

  for (int i = 0; i < size - 1; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    // zero-frequency symbol: take one count from the least frequent symbol that can spare it
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < size - 1; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) {
        best_freq = f;
        best = j;
      }
    }
    if (best < 0) {
      lssvc::set_error("pmf_to_quantized_cdf: cannot give every symbol a non-zero frequency");
      return LSSVC_ERR_ARG;
    }
    if (best < i) {
      for (int j = best + 1; j <= i; ++j) cdf[j]--;
    } else {
      for (int j = i + 1; j <= best; ++j) cdf[j]++;
    }
  }
  return LSSVC_OK;
}

lssvc_rans_encoder *lssvc_rans_encoder_new(void) { return new lssvc_rans_encoder(); }
void lssvc_rans_encoder_free(lssvc_rans_encoder *e) { delete e; }
void lssvc_rans_encoder_reset(lssvc_rans_encoder *e) {
  if (e) e->toks.clear();
}

int32_t lssvc_rans_encode_with_indexes(lssvc_rans_encoder *e, const int32_t *symbols, const int32_t *indexes, int64_t n,
                                       const int32_t *cdfs, int32_t n_rows, int32_t cdf_stride,
                                       const int32_t *cdf_sizes, const int32_t *offsets) {
  if (!e || (n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets || cdf_stride <= 0 || n_rows <= 0) {
    lssvc::set_error("rans_encode_with_indexes: bad arguments");
    return LSSVC_ERR_ARG;
  }
  if (!rows_ok("rans_encode_with_indexes", indexes, n, n_rows, cdf_stride, cdf_sizes)) return LSSVC_ERR_ARG;
  e->toks.reserve(e->toks.size() + static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const int32_t row = indexes[i];
    const int32_t *cdf = cdfs + static_cast<int64_t>(row) * cdf_stride;
    const int32_t max_value = cdf_sizes[row] - 2;
    int32_t v = symbols[i] - offsets[row];
    uint32_t raw = 0;
    if (v < 0) {
      raw = static_cast<uint32_t>(-2 * v - 1);
      v = max_value;
    } else if (v >= max_value) {
      raw = static_cast<uint32_t>(2 * (v - max_value));
      v = max_value;
    }
    e->toks.push_back({static_cast<uint16_t>(cdf[v]), static_cast<uint16_t>(cdf[v + 1] - cdf[v]), 0});
    if (v == max_value) {
      int32_t digits = 0;
      while ((raw >> (digits * kBypassBits)) != 0) ++digits;
      int32_t left = digits;
      while (left >= static_cast<int32_t>(kBypassMax)) {
        e->toks.push_back({static_cast<uint16_t>(kBypassMax), static_cast<uint16_t>(kBypassMax + 1), 1});
        left -= kBypassMax;
      }
      e->toks.push_back({static_cast<uint16_t>(left), static_cast<uint16_t>(left + 1), 1});
      for (int32_t j = 0; j < digits; ++j) {
        const uint32_t d = (raw >> (j * kBypassBits)) & kBypassMax;
        e->toks.push_back({static_cast<uint16_t>(d), static_cast<uint16_t>(d + 1), 1});
      }
    }
  }
  return LSSVC_OK;
}

int64_t lssvc_rans_encoder_flush(lssvc_rans_encoder *e, const uint8_t **data) {
  if (!e) return LSSVC_ERR_ARG;
  // one word per token is an upper bound on the payload (16 bits of information at most each), + 2 for the state
  e->words.assign(e->toks.size() + 2, 0xCCu);
  uint32_t *end = e->words.data() + e->words.size();
  uint32_t *out = end;
  uint64_t x = kLower;
  for (size_t i = e->toks.size(); i-- > 0;) {
    const Tok &t = e->toks[i];
    if (!t.bypass) {
      const uint64_t freq = t.range;
      enc_renorm(x, out, ((kLower >> kProbBits) << 32) * freq);
      x = ((x / freq) << kProbBits) + (x % freq) + t.start;
    } else {
      const uint64_t freq = 1u << (16 - kBypassBits);
      enc_renorm(x, out, ((kLower >> 16) << 32) * freq);
      x = (x << kBypassBits) | t.start;
    }
  }
  e->toks.clear();
  out -= 2;
  out[0] = static_cast<uint32_t>(x);
  out[1] = static_cast<uint32_t>(x >> 32);
  e->data = reinterpret_cast<const uint8_t *>(out);
  e->nbytes = static_cast<int64_t>(end - out) * 4;
  if (data) *data = e->data;
  return e->nbytes;
}

lssvc_rans_decoder *lssvc_rans_decoder_new(void) { return new lssvc_rans_decoder(); }
void lssvc_rans_decoder_free(lssvc_rans_decoder *d) { delete d; }

int32_t lssvc_rans_decoder_set_stream(lssvc_rans_decoder *d, const uint8_t *data, int64_t nbytes) {
  if (!d || !data || nbytes < 8 || (nbytes & 3)) {
    lssvc::set_error("rans_decoder_set_stream: stream must be a non-empty multiple of 4 bytes (got %lld)",
                     static_cast<long long>(nbytes));
    return LSSVC_ERR_STREAM;
  }
  d->words.resize(static_cast<size_t>(nbytes / 4));
  memcpy(d->words.data(), data, static_cast<size_t>(nbytes));
  d->state = static_cast<uint64_t>(d->words[0]) | (static_cast<uint64_t>(d->words[1]) << 32);
  d->pos = 2;
  d->underflow = false;
  return LSSVC_OK;
}

int32_t lssvc_rans_decode_stream(lssvc_rans_decoder *d, const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                 int32_t n_rows, int32_t cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                 int32_t *out) {
  if (!d || (n > 0 && (!indexes || !out)) || !cdfs || !cdf_sizes || !offsets || cdf_stride <= 0 || n_rows <= 0) {
    lssvc::set_error("rans_decode_stream: bad arguments");
    return LSSVC_ERR_ARG;
  }
  if (!rows_ok("rans_decode_stream", indexes, n, n_rows, cdf_stride, cdf_sizes)) return LSSVC_ERR_ARG;
  const uint64_t mask = (1ull << kProbBits) - 1;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t row = indexes[i];
    const int32_t *cdf = cdfs + static_cast<int64_t>(row) * cdf_stride;
    const int32_t size = cdf_sizes[row];
    const int32_t max_value = size - 2;
    const uint32_t cum = static_cast<uint32_t>(d->state & mask);
    // first entry strictly above the cumulative frequency, minus one
    int32_t s = 0;
    while (s + 1 < size && static_cast<uint32_t>(cdf[s + 1]) <= cum) ++s;
    const uint64_t start = static_cast<uint32_t>(cdf[s]);
    const uint64_t freq = static_cast<uint32_t>(cdf[s + 1] - cdf[s]);
    uint64_t x = d->state;
    x = freq * (x >> kProbBits) + (x & mask) - start;
    if (x < kLower) x = (x << 32) | d->next_word();
    d->state = x;

    int32_t value = s;
    if (value == max_value) {
      int32_t val = static_cast<int32_t>(d->get_bits(kBypassBits));
      int32_t digits = val;
      while (val == static_cast<int32_t>(kBypassMax)) {
        val = static_cast<int32_t>(d->get_bits(kBypassBits));
        digits += val;
        if (d->underflow) break;
      }
      if (digits > 8 && !d->underflow) {   // a 32-bit symbol has at most 8 nibbles: the stream is corrupt or out of step
        lssvc::set_error("rans_decode_stream: %d bypass digits at symbol %lld of %lld (corrupt stream or wrong CDF indexes)",
                         digits, static_cast<long long>(i), static_cast<long long>(n));
        return LSSVC_ERR_STREAM;
      }
      int32_t raw = 0;
      for (int32_t j = 0; j < digits && j < 8; ++j) {
        val = static_cast<int32_t>(d->get_bits(kBypassBits));
        raw |= val << (j * kBypassBits);
      }
      value = raw >> 1;
      if (raw & 1) value = -value - 1;
      else value += max_value;
    }
    out[i] = value + offsets[row];
    if (d->underflow) {
      lssvc::set_error("rans_decode_stream: stream exhausted at symbol %lld of %lld", static_cast<long long>(i),
                       static_cast<long long>(n));
      return LSSVC_ERR_STREAM;
    }
  }
  return LSSVC_OK;
}

}  // extern "C"
