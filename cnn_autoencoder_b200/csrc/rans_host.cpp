// Host entropy coder of the product path: range-ANS with 64-bit state, 16-bit
// frequencies, 4-bit bypass escapes, 32-bit output words -- the stream format of
// compressai.ans.RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes
// that the reference reaches through EntropyBottleneck.compress / decompress
// (src/models/tasks/_autoencoders.py:549-551, 568-572, 645-647, 662-665;
// SURVEY.md Appendix A.3), and compressai._CXX.pmf_to_quantized_cdf (A.2) reached
// from EntropyBottleneck.update (R:502, R:615).
//
// Written for the EntropyBottleneck index pattern (symbol i of a C x hw raster
// uses table i / hw), as one backward pass over the symbols with no staging
// list: a symbol's escape nibbles are emitted (in reverse) right before its
// table code, which yields the same bytes as staging everything and flushing it
// back to front.  Thread-safe: no globals, caller-owned buffers.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/cae_b200.h"

void cae_set_error(const char *fmt, ...);

namespace {

constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr int32_t kMaxBypass = (1 << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31;

struct Writer {
  uint32_t *begin, *ptr;  // words are written downwards from the end of the buffer
  bool overflow = false;
  inline void put(uint32_t w) {
    if (ptr == begin) { overflow = true; return; }
    *--ptr = w;
  }
};

inline void enc_symbol(uint64_t &x, Writer &wr, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((kRansL >> kPrecision) << 32) * (uint64_t)freq;
  if (x >= x_max) { wr.put((uint32_t)x); x >>= 32; }
  x = ((x / freq) << kPrecision) + (x % freq) + start;
}

inline void enc_nibble(uint64_t &x, Writer &wr, uint32_t val) {
  const uint64_t x_max = ((kRansL >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits));
  if (x >= x_max) { wr.put((uint32_t)x); x >>= 32; }
  x = (x << kBypassBits) | val;
}

inline uint32_t dec_nibble(uint64_t &x, const uint32_t *&ptr, const uint32_t *end) {
  const uint32_t val = (uint32_t)(x & ((1u << kBypassBits) - 1));
  x >>= kBypassBits;
  if (x < kRansL) { x = (x << 32) | (ptr < end ? *ptr : 0u); ++ptr; }
  return val;
}

}  // namespace

extern "C" int cae_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  if (!pmf || !cdf || n <= 0 || precision <= 0 || precision > 16) {
    cae_set_error("cae_pmf_to_quantized_cdf: bad arguments");
    return 2;
  }
  const uint32_t one = 1u << precision;
  uint64_t total = 0;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) {
    if (!(pmf[i] >= 0.0f) || !std::isfinite(pmf[i])) {
      cae_set_error("cae_pmf_to_quantized_cdf: pmf[%d] is negative or not finite", i);
      return 4;
    }
    cdf[i + 1] = (uint32_t)std::round(pmf[i] * (float)one);
    total += cdf[i + 1];
  }
  total &= 0xffffffffu;  // the published routine accumulates in 32 bits
  if (total == 0) {
    cae_set_error("cae_pmf_to_quantized_cdf: pmf sums to zero");
    return 4;
  }
  uint32_t run = 0;
  for (int i = 0; i <= n; ++i) {
    run += (uint32_t)(((uint64_t)one * cdf[i]) / total);
    cdf[i] = run;
  }
  cdf[n] = one;
  // every symbol needs a non-zero frequency: take one count from the cheapest donor
  for (int i = 0; i < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best = ~0u;
    int donor = -1;
    for (int j = 0; j < n; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best) { best = f; donor = j; }
    }
    if (donor < 0) {
      cae_set_error("cae_pmf_to_quantized_cdf: no frequency left to steal");
      return 4;
    }
    if (donor < i) for (int j = donor + 1; j <= i; ++j) --cdf[j];
    else for (int j = i + 1; j <= donor; ++j) ++cdf[j];
  }
  return 0;
}

extern "C" int cae_rans_encode(const int32_t *symbols, int c, int hw, const int32_t *cdfs,
                               int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                               uint8_t *out, size_t out_cap, size_t *nbytes) {
  if (!symbols || !cdfs || !cdf_sizes || !offsets || !out || !nbytes || c <= 0 || hw <= 0) {
    cae_set_error("cae_rans_encode: bad arguments");
    return 2;
  }
  // worst case words: one per symbol plus escapes; grow on demand
  size_t cap_words = (size_t)c * hw + 16;
  std::vector<uint32_t> buf;
  for (int attempt = 0; attempt < 8; ++attempt, cap_words *= 2) {
    buf.assign(cap_words, 0);
    Writer wr{buf.data(), buf.data() + cap_words};
    uint64_t x = kRansL;
    for (int ch = c - 1; ch >= 0; --ch) {
      const int32_t *cdf = cdfs + (size_t)ch * cdf_stride;
      const int32_t max_value = cdf_sizes[ch] - 2;
      const int32_t offset = offsets[ch];
      if (max_value < 0 || cdf_sizes[ch] > cdf_stride) {
        cae_set_error("cae_rans_encode: bad cdf size for channel %d", ch);
        return 2;
      }
      const int32_t *sym = symbols + (size_t)ch * hw;
      for (int i = hw - 1; i >= 0; --i) {
        int64_t value = (int64_t)sym[i] - offset;
        if (value >= 0 && value < max_value) {
          enc_symbol(x, wr, (uint32_t)cdf[value], (uint32_t)(cdf[value + 1] - cdf[value]));
          continue;
        }
        // escape: sign-folded raw value in 4-bit groups, count in unary-of-15, then the
        // sentinel code; all emitted last-to-first
        uint32_t raw = value < 0 ? (uint32_t)(-2 * value - 1) : (uint32_t)(2 * (value - max_value));
        int32_t n_groups = 0;
        while (n_groups < 8 && (raw >> (n_groups * kBypassBits)) != 0) ++n_groups;
        for (int32_t g = n_groups - 1; g >= 0; --g)
          enc_nibble(x, wr, (raw >> (g * kBypassBits)) & (uint32_t)kMaxBypass);
        const int32_t fifteens = n_groups / kMaxBypass, rest = n_groups % kMaxBypass;
        enc_nibble(x, wr, (uint32_t)rest);
        for (int32_t k = 0; k < fifteens; ++k) enc_nibble(x, wr, (uint32_t)kMaxBypass);
        enc_symbol(x, wr, (uint32_t)cdf[max_value],
                   (uint32_t)(cdf[max_value + 1] - cdf[max_value]));
      }
    }
    wr.put((uint32_t)(x >> 32));
    wr.put((uint32_t)x);
    if (wr.overflow) continue;
    const size_t bytes = (size_t)((buf.data() + cap_words) - wr.ptr) * 4;
    if (bytes > out_cap) {
      cae_set_error("cae_rans_encode: output buffer too small (%zu > %zu)", bytes, out_cap);
      *nbytes = bytes;
      return 5;
    }
    std::memcpy(out, wr.ptr, bytes);
    *nbytes = bytes;
    return 0;
  }
  cae_set_error("cae_rans_encode: stream does not fit the staging buffer");
  return 5;
}

extern "C" int cae_rans_decode(const uint8_t *enc, size_t nbytes, int c, int hw,
                               const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                               const int32_t *offsets, int32_t *symbols) {
  if (!enc || !cdfs || !cdf_sizes || !offsets || !symbols || c <= 0 || hw <= 0) {
    cae_set_error("cae_rans_decode: bad arguments");
    return 2;
  }
  if (nbytes < 8 || (nbytes & 3)) {
    cae_set_error("cae_rans_decode: stream of %zu bytes is not a whole number of words >= 2", nbytes);
    return 6;
  }
  std::vector<uint32_t> words(nbytes / 4);
  std::memcpy(words.data(), enc, nbytes);
  const uint32_t *ptr = words.data(), *end = words.data() + words.size();
  uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
  ptr += 2;
  for (int ch = 0; ch < c; ++ch) {
    const int32_t *cdf = cdfs + (size_t)ch * cdf_stride;
    const int32_t size = cdf_sizes[ch];
    const int32_t max_value = size - 2;
    const int32_t offset = offsets[ch];
    if (max_value < 0 || size > cdf_stride) {
      cae_set_error("cae_rans_decode: bad cdf size for channel %d", ch);
      return 2;
    }
    int32_t *dst = symbols + (size_t)ch * hw;
    for (int i = 0; i < hw; ++i) {
      const uint32_t cf = (uint32_t)(x & 0xffffu);
      // last entry with cdf[s] <= cf  (tables are short: binary search over `size` entries)
      int32_t lo = 0, hi = size - 1;
      while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if ((uint32_t)cdf[mid] <= cf) lo = mid; else hi = mid;
      }
      const int32_t s = lo;
      const uint64_t start = (uint64_t)cdf[s], freq = (uint64_t)(cdf[s + 1] - cdf[s]);
      x = freq * (x >> kPrecision) + (x & 0xffffu) - start;
      if (x < kRansL) { x = (x << 32) | (ptr < end ? *ptr : 0u); ++ptr; }
      int32_t value = s;
      if (s == max_value) {
        int32_t v = (int32_t)dec_nibble(x, ptr, end), n_groups = v;
        while (v == kMaxBypass) { v = (int32_t)dec_nibble(x, ptr, end); n_groups += v; }
        uint32_t raw = 0;
        for (int32_t g = 0; g < n_groups; ++g) {
          const uint32_t nib = dec_nibble(x, ptr, end);
          if (g < 8) raw |= nib << (g * kBypassBits);
        }
        value = (int32_t)(raw >> 1);
        value = (raw & 1u) ? -value - 1 : value + max_value;
      }
      dst[i] = value + offset;
    }
  }
  if (ptr > end) {
    cae_set_error("cae_rans_decode: stream truncated");
    return 6;
  }
  return 0;
}
