// Batched range-ANS coder on the device: one thread per tile stream, thousands of streams in
// flight.  Same stream format as the host coder (rans_host.cpp) and as
// compressai.ans.RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes, which the
// reference reaches through EntropyBottleneck.compress / decompress
// (src/models/tasks/_autoencoders.py:549-551, 568-572): 64-bit state,
// 16-bit frequencies, 4-bit bypass escapes, 32-bit words, symbol i of a C x hw raster uses
// table i / hw.  A stream is inherently sequential, so the parallelism is ACROSS the tiles of
// a slide: the integer symbols never leave the device until they are bytes.
#include "cae_common.cuh"

namespace {

constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr int kMaxBypass = 15;
constexpr uint64_t kRansL = 1ull << 31;

// One entry per (channel, symbol): the state update x -> (x / freq << 16) + x % freq + start
// as x + bias + mulhi(x, rcp_freq) >> rcp_shift * (2^16 - freq), the division-free form of
// ryg_rans' rans64.h (Alverson's exact reciprocal: rcp_freq = ceil(2^(shift+63) / freq) with
// shift = ceil(log2 freq); freq == 1 uses rcp = 2^64 - 1, bias = start + 2^16 - 1).  The
// 64-bit division is most of the dependent chain of a symbol; this is ~10x shorter.
struct __align__(16) EncEntry {
  uint64_t rcp_freq;
  uint32_t bias;
  uint16_t cmpl_freq;   // 2^16 - freq
  uint16_t rcp_shift;
};

__global__ void rans_build_table_kernel(const int32_t *cdfs, int c, int stride,
                                        const int32_t *sizes, EncEntry *table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c * stride) return;
  const int ch = i / stride, v = i - ch * stride;
  EncEntry e{0ull, 0u, 0, 0};
  if (v + 1 < sizes[ch]) {
    const uint32_t start = (uint32_t)cdfs[(size_t)ch * stride + v];
    const uint32_t freq = (uint32_t)cdfs[(size_t)ch * stride + v + 1] - start;
    e.cmpl_freq = (uint16_t)((1u << kPrecision) - freq);
    if (freq < 2) {
      e.rcp_freq = ~0ull;
      e.rcp_shift = 0;
      e.bias = start + (1u << kPrecision) - 1;
    } else {
      uint32_t shift = 0;
      while (freq > (1u << shift)) ++shift;
      const unsigned __int128 num = ((unsigned __int128)1 << (shift + 63)) + freq - 1;
      e.rcp_freq = (uint64_t)(num / freq);
      e.rcp_shift = (uint16_t)(shift - 1);
      e.bias = start;
    }
  }
  table[i] = e;
}

struct RansEncParams {
  const EncEntry *table;   // [c][stride] or nullptr (plain division)
  int table_in_smem;
  const int32_t *symbols;  // [n][c][hw]
  int n, c, hw;
  const int32_t *cdfs;     // [c][stride]
  int stride;
  const int32_t *sizes, *offsets;
  uint32_t *words;         // [n][cap]; stream k occupies words[k][cap - nwords[k] .. cap)
  int cap;
  int32_t *nwords;         // [n]
  int32_t *status;         // bit 0: a stream overflowed its buffer
};

struct Emitter {
  uint32_t *base;
  int pos;        // next free slot is base[pos - 1]
  bool overflow;
  __device__ __forceinline__ void put(uint32_t w) {
    if (pos <= 0) { overflow = true; return; }
    base[--pos] = w;
  }
};

__device__ __forceinline__ void enc_symbol(uint64_t &x, Emitter &e, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((kRansL >> kPrecision) << 32) * (uint64_t)freq;
  if (x >= x_max) { e.put((uint32_t)x); x >>= 32; }
  const uint64_t q = x / freq;
  x = (q << kPrecision) + (x - q * freq) + start;
}

__device__ __forceinline__ void enc_nibble(uint64_t &x, Emitter &e, uint32_t val) {
  const uint64_t x_max = ((kRansL >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits));
  if (x >= x_max) { e.put((uint32_t)x); x >>= 32; }
  x = (x << kBypassBits) | val;
}

__device__ __forceinline__ void enc_symbol_fast(uint64_t &x, Emitter &e, const EncEntry &t) {
  const uint64_t freq = (1u << kPrecision) - (uint32_t)t.cmpl_freq;
  if (x >= (freq << 47)) { e.put((uint32_t)x); x >>= 32; }   // ((L >> 16) << 32) * freq
  const uint64_t q = __umul64hi(x, t.rcp_freq) >> t.rcp_shift;
  x = x + t.bias + q * t.cmpl_freq;
}

__global__ void __launch_bounds__(64) rans_encode_table_kernel(const RansEncParams p) {
  extern __shared__ __align__(16) uint8_t enc_smem[];
  const EncEntry *tab = p.table;
  if (p.table_in_smem) {
    EncEntry *st = reinterpret_cast<EncEntry *>(enc_smem);
    for (int i = threadIdx.x; i < p.c * p.stride; i += blockDim.x) st[i] = p.table[i];
    __syncthreads();
    tab = st;
  }
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  Emitter e{p.words + (size_t)k * p.cap, p.cap, false};
  uint64_t x = kRansL;
  const int32_t *sym = p.symbols + (size_t)k * p.c * p.hw;
  for (int ch = p.c - 1; ch >= 0; --ch) {
    const EncEntry *row = tab + (size_t)ch * p.stride;
    const int max_value = p.sizes[ch] - 2;
    const int offset = p.offsets[ch];
    const int32_t *s = sym + (size_t)ch * p.hw;
    for (int i = p.hw - 1; i >= 0; --i) {
      const long long value = (long long)__ldg(s + i) - offset;
      if (value >= 0 && value < max_value) {
        enc_symbol_fast(x, e, row[value]);
        continue;
      }
      const uint32_t raw = value < 0 ? (uint32_t)(-2 * value - 1) : (uint32_t)(2 * (value - max_value));
      int groups = 0;
      while (groups < 8 && (raw >> (groups * kBypassBits)) != 0) ++groups;
      for (int g = groups - 1; g >= 0; --g) enc_nibble(x, e, (raw >> (g * kBypassBits)) & 15u);
      enc_nibble(x, e, (uint32_t)(groups % kMaxBypass));
      for (int q = 0; q < groups / kMaxBypass; ++q) enc_nibble(x, e, (uint32_t)kMaxBypass);
      enc_symbol_fast(x, e, row[max_value]);
    }
  }
  e.put((uint32_t)(x >> 32));
  e.put((uint32_t)x);
  p.nwords[k] = p.cap - e.pos;
  if (e.overflow) atomicOr(p.status, 1);
}

__global__ void __launch_bounds__(64) rans_encode_kernel(const RansEncParams p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  Emitter e{p.words + (size_t)k * p.cap, p.cap, false};
  uint64_t x = kRansL;
  const int32_t *sym = p.symbols + (size_t)k * p.c * p.hw;
  for (int ch = p.c - 1; ch >= 0; --ch) {
    const int32_t *cdf = p.cdfs + (size_t)ch * p.stride;
    const int max_value = p.sizes[ch] - 2;
    const int offset = p.offsets[ch];
    const uint32_t esc_start = (uint32_t)cdf[max_value];
    const uint32_t esc_freq = (uint32_t)(cdf[max_value + 1] - cdf[max_value]);
    const int32_t *s = sym + (size_t)ch * p.hw;
    for (int i = p.hw - 1; i >= 0; --i) {
      const long long value = (long long)s[i] - offset;
      if (value >= 0 && value < max_value) {
        const uint32_t st = (uint32_t)cdf[value];
        enc_symbol(x, e, st, (uint32_t)cdf[value + 1] - st);
        continue;
      }
      // escape: sign-folded raw value as 4-bit groups, their count in unary-of-15, then the
      // sentinel code -- emitted last to first (the decoder reads them in the forward order)
      const uint32_t raw = value < 0 ? (uint32_t)(-2 * value - 1) : (uint32_t)(2 * (value - max_value));
      int groups = 0;
      while (groups < 8 && (raw >> (groups * kBypassBits)) != 0) ++groups;
      for (int g = groups - 1; g >= 0; --g) enc_nibble(x, e, (raw >> (g * kBypassBits)) & 15u);
      enc_nibble(x, e, (uint32_t)(groups % kMaxBypass));
      for (int q = 0; q < groups / kMaxBypass; ++q) enc_nibble(x, e, (uint32_t)kMaxBypass);
      enc_symbol(x, e, esc_start, esc_freq);
    }
  }
  e.put((uint32_t)(x >> 32));
  e.put((uint32_t)x);
  p.nwords[k] = p.cap - e.pos;
  if (e.overflow) atomicOr(p.status, 1);
}

// pack the used tails of the per-stream buffers back to back: block per stream
__global__ void rans_compact_kernel(const uint32_t *__restrict__ words, int cap,
                                    const int32_t *__restrict__ nwords,
                                    const int64_t *__restrict__ out_off, uint32_t *__restrict__ out) {
  const int k = blockIdx.x;
  const int nw = nwords[k];
  const uint32_t *src = words + (size_t)k * cap + (cap - nw);
  uint32_t *dst = out + out_off[k];
  for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
}

struct RansDecParams {
  const uint32_t *words;   // all streams back to back
  const int64_t *off;      // [n + 1] word offsets
  int n, c, hw;
  const int32_t *cdfs;
  int stride;
  const int32_t *sizes, *offsets;
  int32_t *symbols;        // [n][c][hw]
  int32_t *status;         // bit 1: a stream ran past its end
  int cdf_in_smem;
};

__device__ __forceinline__ uint32_t next_word(const uint32_t *&ptr, const uint32_t *end, bool &bad) {
  if (ptr >= end) { bad = true; ++ptr; return 0u; }
  return *ptr++;
}

__device__ __forceinline__ uint32_t dec_nibble(uint64_t &x, const uint32_t *&ptr,
                                               const uint32_t *end, bool &bad) {
  const uint32_t val = (uint32_t)(x & 15u);
  x >>= kBypassBits;
  if (x < kRansL) x = (x << 32) | next_word(ptr, end, bad);
  return val;
}

__global__ void __launch_bounds__(64) rans_decode_kernel(const RansDecParams p) {
  extern __shared__ __align__(16) uint8_t dec_smem[];
  const int32_t *cdfs = p.cdfs;
  if (p.cdf_in_smem) {
    int32_t *sc = reinterpret_cast<int32_t *>(dec_smem);
    for (int i = threadIdx.x; i < p.c * p.stride; i += blockDim.x) sc[i] = p.cdfs[i];
    __syncthreads();
    cdfs = sc;
  }
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  const uint32_t *ptr = p.words + p.off[k], *end = p.words + p.off[k + 1];
  bool bad = end - ptr < 2;
  uint64_t x = bad ? kRansL : ((uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32));
  ptr += 2;
  int32_t *out = p.symbols + (size_t)k * p.c * p.hw;
  for (int ch = 0; ch < p.c; ++ch) {
    const int32_t *cdf = cdfs + (size_t)ch * p.stride;
    const int size = p.sizes[ch], max_value = size - 2, offset = p.offsets[ch];
    int32_t *dst = out + (size_t)ch * p.hw;
    for (int i = 0; i < p.hw; ++i) {
      const uint32_t cf = (uint32_t)(x & 0xffffu);
      int lo = 0, hi = size - 1;          // last entry with cdf[s] <= cf
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)cdf[mid] <= cf) lo = mid; else hi = mid;
      }
      const uint32_t start = (uint32_t)cdf[lo], freq = (uint32_t)cdf[lo + 1] - start;
      x = (uint64_t)freq * (x >> kPrecision) + cf - start;
      if (x < kRansL) x = (x << 32) | next_word(ptr, end, bad);
      int value = lo;
      if (lo == max_value) {
        int v = (int)dec_nibble(x, ptr, end, bad), groups = v;
        while (v == kMaxBypass) { v = (int)dec_nibble(x, ptr, end, bad); groups += v; }
        uint32_t raw = 0;
        for (int g = 0; g < groups; ++g) {
          const uint32_t nib = dec_nibble(x, ptr, end, bad);
          if (g < 8) raw |= nib << (g * kBypassBits);
        }
        value = (int)(raw >> 1);
        value = (raw & 1u) ? -value - 1 : value + max_value;
      }
      dst[i] = value + offset;
    }
  }
  if (bad) atomicOr(p.status, 2);
}

}  // namespace

extern "C" size_t cae_rans_enc_table_bytes(int c, int cdf_stride) {
  return (size_t)c * cdf_stride * sizeof(EncEntry);
}

extern "C" int cae_rans_build_enc_table(const int32_t *cdfs, int c, int cdf_stride,
                                        const int32_t *cdf_sizes, void *table, void *stream) {
  CAE_CHECK(cdfs && cdf_sizes && table && c > 0 && cdf_stride > 1, 2,
            "cae_rans_build_enc_table: bad argument");
  const int total = c * cdf_stride;
  rans_build_table_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      cdfs, c, cdf_stride, cdf_sizes, reinterpret_cast<EncEntry *>(table));
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_rans_encode_batch(const int32_t *symbols, int n, int c, int hw,
                                     const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                     const int32_t *offsets, const void *enc_table, uint32_t *words,
                                     int cap_words, int32_t *nwords, int32_t *status, void *stream) {
  CAE_CHECK(symbols && cdfs && cdf_sizes && offsets && words && nwords && status, 2,
            "cae_rans_encode_batch: null argument");
  CAE_CHECK(n > 0 && c > 0 && hw > 0 && cap_words >= 4, 2, "cae_rans_encode_batch: bad shape");
  RansEncParams p{reinterpret_cast<const EncEntry *>(enc_table), 0, symbols, n, c, hw, cdfs,
                  cdf_stride, cdf_sizes, offsets, words, cap_words, nwords, status};
  if (enc_table) {
    const size_t tbytes = cae_rans_enc_table_bytes(c, cdf_stride);
    p.table_in_smem = tbytes <= 96 * 1024;
    const size_t smem = p.table_in_smem ? tbytes : 0;
    if (smem > 48 * 1024)
      CAE_CUDA(cudaFuncSetAttribute(rans_encode_table_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rans_encode_table_kernel<<<(n + 63) / 64, 64, smem, (cudaStream_t)stream>>>(p);
    cae_count_launch();
    CAE_CUDA(cudaGetLastError());
    return 0;
  }
  rans_encode_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_rans_compact(const uint32_t *words, int n, int cap_words, const int32_t *nwords,
                                const int64_t *out_offsets, uint32_t *out, void *stream) {
  CAE_CHECK(words && nwords && out_offsets && out && n > 0, 2, "cae_rans_compact: bad argument");
  rans_compact_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(words, cap_words, nwords, out_offsets, out);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_rans_decode_batch(const uint32_t *words, const int64_t *word_offsets, int n, int c,
                                     int hw, const int32_t *cdfs, int cdf_stride,
                                     const int32_t *cdf_sizes, const int32_t *offsets,
                                     int32_t *symbols, int32_t *status, void *stream) {
  CAE_CHECK(words && word_offsets && cdfs && cdf_sizes && offsets && symbols && status, 2,
            "cae_rans_decode_batch: null argument");
  CAE_CHECK(n > 0 && c > 0 && hw > 0, 2, "cae_rans_decode_batch: bad shape");
  RansDecParams p{words, word_offsets, n, c, hw, cdfs, cdf_stride, cdf_sizes, offsets, symbols, status, 0};
  const size_t cbytes = (size_t)c * cdf_stride * sizeof(int32_t);
  p.cdf_in_smem = cbytes <= 96 * 1024;
  const size_t smem = p.cdf_in_smem ? cbytes : 0;
  if (smem > 48 * 1024)
    CAE_CUDA(cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  rans_decode_kernel<<<(n + 63) / 64, 64, smem, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
