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


// ------------------------------------------------------------------------------------------
// Warp-staged coders (the default).  One thread still owns one stream -- a range-coder state is
// a dependent chain -- but the 32 streams of a warp move their symbols through shared memory in
// 32 x 32 chunks: the warp reads (writes) 32 symbols of each of its streams as one coalesced
// 128-byte access, software-pipelined one chunk ahead, and every lane then walks its own row of
// the chunk (row pitch 33 words: conflict free both ways).  The one-thread-per-stream kernels
// above issue one dependent, uncoalesced 4-byte global access per symbol (~330 ns each: 65 ms
// per call whatever the stream count); here the per-symbol chain is shared memory only.
// ------------------------------------------------------------------------------------------
constexpr int kChunk = 32;
constexpr int kPitch = kChunk + 1;

// Hot-path pieces of the warp-staged encoder.  The 32 lanes of a warp code 32 different streams,
// so any branch on the state (renormalise or not) diverges at almost every symbol; both the word
// emission and the state update are therefore predicated, and only the escape path (rare by
// construction of the tables) is a real, out-of-line branch.
__device__ __forceinline__ void st_global_if(uint32_t *ptr, uint32_t w, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.b32 [%0], %1;\n\t}"
               ::"l"(ptr), "r"(w), "r"((uint32_t)p)
               : "memory");
}

struct PredEmitter {
  uint32_t *base;
  int pos;        // words still free below the written tail; may go negative (= overflow)
  __device__ __forceinline__ void put_if(bool p, uint32_t w) {
    st_global_if(base + (pos - 1), w, p && pos > 0);
    pos -= p ? 1 : 0;
  }
};

__device__ __forceinline__ void enc_step(uint64_t &x, PredEmitter &e, const EncEntry &t) {
  const uint32_t freq = (1u << kPrecision) - (uint32_t)t.cmpl_freq;
  const bool p = (uint32_t)(x >> 47) >= freq;          // x >= ((L >> 16) << 32) * freq
  e.put_if(p, (uint32_t)x);
  x = p ? (x >> 32) : x;
  const uint64_t q = __umul64hi(x, t.rcp_freq) >> t.rcp_shift;
  x = x + t.bias + q * t.cmpl_freq;
}

struct EncState {
  uint64_t x;
  int pos;
};

// by value in, by value out: nothing of the hot loop's state has its address taken
static __device__ __noinline__ EncState enc_escape(uint64_t x, uint32_t *base, int pos, long long vv,
                                                   int max_value) {
  // sign-folded raw value as 4-bit groups, their count in unary-of-15 -- emitted last to first
  // (the decoder reads them in the forward order); the sentinel code follows in the caller
  PredEmitter e{base, pos};
  const uint32_t raw = vv < 0 ? (uint32_t)(-2 * vv - 1) : (uint32_t)(2 * (vv - max_value));
  int groups = 0;
  while (groups < 8 && (raw >> (groups * kBypassBits)) != 0) ++groups;
  auto nibble = [&](uint32_t val) {
    const bool p = (uint32_t)(x >> 47) >= (1u << (16 - kBypassBits));
    e.put_if(p, (uint32_t)x);
    x = p ? (x >> 32) : x;
    x = (x << kBypassBits) | val;
  };
  for (int g = groups - 1; g >= 0; --g) nibble((raw >> (g * kBypassBits)) & 15u);
  nibble((uint32_t)(groups % kMaxBypass));
  for (int q = 0; q < groups / kMaxBypass; ++q) nibble((uint32_t)kMaxBypass);
  return EncState{x, e.pos};
}

template <bool SMEM_TABLE>
__global__ void __launch_bounds__(32) rans_encode_warp_kernel(const RansEncParams p) {
  extern __shared__ __align__(16) uint8_t enc_smem[];
  const int lane = threadIdx.x;
  int32_t *stage = reinterpret_cast<int32_t *>(enc_smem);                 // [32][33]
  EncEntry *tab = reinterpret_cast<EncEntry *>(enc_smem + kChunk * kPitch * 4 + 32);
  // (template parameter rather than a runtime pointer choice: the compiler then knows the
  // address space and reads a table entry with one 16-byte shared-memory load)
  if (SMEM_TABLE)
    for (int i = lane; i < p.c * p.stride; i += 32) tab[i] = p.table[i];
  const EncEntry *table = SMEM_TABLE ? tab : p.table;
  __syncwarp();
  const int k0 = blockIdx.x * 32;
  const int k = k0 + lane;
  const bool live = k < p.n;
  const int n_here = min(32, p.n - k0);
  PredEmitter e{p.words + (size_t)(live ? k : k0) * p.cap, p.cap};
  uint64_t x = kRansL;
  const size_t per_stream = (size_t)p.c * p.hw;
  const int32_t *sym0 = p.symbols + (size_t)k0 * per_stream;
  const int chunks = (p.hw + kChunk - 1) / kChunk;
  const int total = p.c * chunks;

  int32_t pre[kChunk];
  auto fetch = [&](int t) {       // chunk t of every stream of the warp: 32 coalesced loads
    const int ch = t / chunks, j = t - ch * chunks;
    const int i = j * kChunk + lane;
    const int32_t *src = sym0 + (size_t)ch * p.hw + i;
    const bool ok = i < p.hw;
#pragma unroll
    for (int s = 0; s < kChunk; ++s)
      pre[s] = (ok && s < n_here) ? __ldg(src + (size_t)s * per_stream) : 0;
  };
  fetch(total - 1);
  for (int t = total - 1; t >= 0; --t) {
    const int ch = t / chunks, j = t - ch * chunks;
    const int max_value = p.sizes[ch] - 2;
    const int offset = p.offsets[ch];
    __syncwarp();
    // does any of the 32 x 32 symbols of this chunk need the escape path?  (rare by construction
    // of the tables: the quantiles put 1e-9 of the mass in each tail)
    bool esc_any = false;
#pragma unroll
    for (int s = 0; s < kChunk; ++s) {
      stage[s * kPitch + lane] = pre[s];
      esc_any |= (unsigned)(pre[s] - offset) >= (unsigned)max_value;
    }
    const int m = min(kChunk, p.hw - j * kChunk);
    const bool plain = !__any_sync(0xffffffffu, esc_any) && m == kChunk;
    __syncwarp();
    if (t > 0) fetch(t - 1);
    if (!live) continue;
    const EncEntry *row = table + (size_t)ch * p.stride;
    const int32_t *mine = stage + lane * kPitch;
    if (plain) {
      // no escape anywhere in the chunk: one straight-line block of 32 state updates, so the
      // scheduler is free to run the loads and the frequency arithmetic of the next symbols in
      // the stall slots of the current symbol's dependent chain
#pragma unroll
      for (int i = kChunk - 1; i >= 0; --i) enc_step(x, e, row[mine[i] - offset]);
      continue;
    }
    // The state update is a dependent chain (compare, 64-bit multiply-high, shift, multiply,
    // add); everything else of a symbol -- its load, the range test, the 16-byte table entry --
    // does not depend on the state, so it is fetched one symbol ahead and overlaps the chain.
    int value = mine[m - 1] - offset;
    bool esc = (unsigned)value >= (unsigned)max_value;
    EncEntry ent = row[esc ? max_value : value];
#pragma unroll 4
    for (int i = m - 1; i >= 0; --i) {
      const int v_cur = value;
      const bool esc_cur = esc;
      const EncEntry cur = ent;
      if (i > 0) {
        value = mine[i - 1] - offset;
        esc = (unsigned)value >= (unsigned)max_value;
        ent = row[esc ? max_value : value];
      }
      if (esc_cur) {
        const EncState st = enc_escape(x, e.base, e.pos, (long long)v_cur, max_value);
        x = st.x;
        e.pos = st.pos;
      }
      enc_step(x, e, cur);
    }
  }
  if (!live) return;
  e.put_if(true, (uint32_t)(x >> 32));
  e.put_if(true, (uint32_t)x);
  const bool overflow = e.pos < 0;
  p.nwords[k] = overflow ? 0 : p.cap - e.pos;
  if (overflow) atomicOr(p.status, 1);
}

// exclusive prefix sum of the stream lengths (one block; n is a few thousand at most)
__global__ void __launch_bounds__(1024) rans_scan_kernel(const int32_t *__restrict__ nwords, int n,
                                                         int64_t *__restrict__ off) {
  __shared__ long long part[1024];
  const int per = (n + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(n, b + per);
  long long s = 0;
  for (int i = b; i < e; ++i) s += nwords[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) { const long long v = part[i]; part[i] = run; run += v; }
    off[n] = run;
  }
  __syncthreads();
  long long run = part[threadIdx.x];
  for (int i = b; i < e; ++i) { off[i] = run; run += nwords[i]; }
}

struct WordReader {          // forward reader with one word of look-ahead
  const uint32_t *ptr, *end;
  uint32_t nxt;
  bool bad;
  __device__ __forceinline__ void init(const uint32_t *b, const uint32_t *e) {
    ptr = b; end = e; bad = false;
    nxt = ptr < end ? __ldg(ptr) : 0u;
  }
  __device__ __forceinline__ uint32_t take() {
    if (ptr >= end) { bad = true; return 0u; }
    const uint32_t w = nxt;
    ++ptr;
    if (ptr < end) nxt = __ldg(ptr);
    return w;
  }
};

__device__ __forceinline__ uint32_t dec_nibble_w(uint64_t &x, WordReader &r) {
  const uint32_t val = (uint32_t)(x & 15u);
  x >>= kBypassBits;
  if (x < kRansL) x = (x << 32) | r.take();
  return val;
}

// Decoder with a direct look-up: for the channel being decoded a 4096-entry table maps the top
// 12 bits of the cumulative frequency to (first symbol whose range reaches into that bucket, its
// start and frequency), so the symbol search on the state's dependent chain is ONE 8-byte
// shared-memory load (plus a rare step forward for buckets that straddle a boundary).  All
// streams of a warp move through the channels together, so one table (32 KB) serves the warp
// and is rebuilt at every channel switch (128 entries per lane, ~0.3 % of the channel's work).
constexpr int kFineBits = 12;
constexpr int kFine = 1 << kFineBits;

// Stream words reach the decoding lane through a per-lane ring in shared memory filled with
// 4-byte cp.async copies issued a whole chunk ahead.  A register look-ahead does not work here:
// the 32 lanes consume words at data-dependent moments, so almost every symbol step has SOME lane
// loading, and the next step's read of the look-ahead register then stalls the whole warp on
// that load's latency (measured: 470 cycles per symbol).  cp.async has no destination register,
// hence no scoreboard dependency; the word is picked up later with a shared-memory load.
constexpr int kRing = 128;              // words per lane; a chunk consumes at most 32 * 4

struct RingReader {
  const uint32_t *gbase;    // first word of this lane's stream
  uint32_t ring;            // shared-memory byte address of this lane's ring
  int n_words;              // stream length
  int rd, fl, safe;         // words consumed / copies issued / copies known to have landed

  __device__ __forceinline__ void issue() {      // top the ring up to rd + kRing, one group
    const int lim = min(rd + kRing, n_words);
    while (fl < lim) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring + (uint32_t)((fl & (kRing - 1)) << 2)),
                   "l"(gbase + fl)
                   : "memory");
      ++fl;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // chunk boundary: refill, then make sure at least `need` unread words have landed
  __device__ __forceinline__ void chunk_begin(int need) {
    const int fl_prev = fl;
    issue();
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // all but the group just issued
    safe = fl_prev;
    if (safe - rd < need && safe < n_words) {                  // rare (escape-heavy chunk)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      safe = fl;
    }
  }
  __device__ __forceinline__ uint32_t peek() const {
    uint32_t w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(ring + (uint32_t)((rd & (kRing - 1)) << 2)));
    return w;
  }
};

__device__ __forceinline__ void dec_renorm(uint64_t &x, RingReader &r) {
  const uint32_t w = r.peek();              // address known early: off the state's chain
  const bool need = x < kRansL;
  x = need ? ((x << 32) | w) : x;
  r.rd += need ? 1 : 0;
}

struct DecState {
  uint64_t x;
  int rd;
  int value;
};

static __device__ __noinline__ DecState dec_escape(uint64_t x, RingReader r, int max_value) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");        // everything issued has landed
  auto nibble = [&]() {
    const uint32_t val = (uint32_t)(x & 15u);
    x >>= kBypassBits;
    dec_renorm(x, r);
    return val;
  };
  int v = (int)nibble(), groups = v;
  while (v == kMaxBypass && groups < 64) { v = (int)nibble(); groups += v; }
  uint32_t raw = 0;
  for (int g = 0; g < groups; ++g) {
    const uint32_t nib = nibble();
    if (g < 8) raw |= nib << (g * kBypassBits);
  }
  const int value = (int)(raw >> 1);
  return DecState{x, r.rd, (raw & 1u) ? -value - 1 : value + max_value};
}

// the bucket of the fine table straddles a symbol boundary and cf lies beyond the first symbol:
// returns (lo << 32) | start of the symbol that holds cf
static __device__ __noinline__ uint64_t dec_walk(const int32_t *cdf, uint32_t cf, int lo) {
  uint32_t start;
  do {
    ++lo;
    start = (uint32_t)cdf[lo];
  } while (cf >= (uint32_t)cdf[lo + 1]);
  return ((uint64_t)(uint32_t)lo << 32) | start;
}

__global__ void __launch_bounds__(32) rans_decode_fine_kernel(const RansDecParams p) {
  extern __shared__ __align__(16) uint8_t dec_smem[];
  const int lane = threadIdx.x;
  uint2 *fine = reinterpret_cast<uint2 *>(dec_smem);                       // [4096]
  uint32_t *rings = reinterpret_cast<uint32_t *>(dec_smem + kFine * 8);    // [32][kRing]
  int32_t *stage = reinterpret_cast<int32_t *>(rings + 32 * kRing);        // [32][33]
  int32_t *scdf = stage + kChunk * kPitch + 8;                             // [c][stride]
  for (int i = lane; i < p.c * p.stride; i += 32) scdf[i] = p.cdfs[i];
  __syncwarp();
  const int k0 = blockIdx.x * 32;
  const int k = k0 + lane;
  const bool live = k < p.n;
  const int n_here = min(32, p.n - k0);
  RingReader r;
  r.ring = smem_u32(rings + lane * kRing);
  r.rd = r.fl = r.safe = 0;
  r.gbase = p.words;
  r.n_words = 0;
  uint64_t x = kRansL;
  if (live) {
    r.gbase = p.words + p.off[k];
    r.n_words = (int)(p.off[k + 1] - p.off[k]);
    r.issue();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    r.safe = r.fl;
    const uint32_t w0 = r.peek();
    r.rd = 1;
    const uint32_t w1 = r.peek();
    r.rd = 2;
    x = (uint64_t)w0 | ((uint64_t)w1 << 32);
  }
  const size_t per_stream = (size_t)p.c * p.hw;
  int32_t *out0 = p.symbols + (size_t)k0 * per_stream;
  const int chunks = (p.hw + kChunk - 1) / kChunk;
  for (int ch = 0; ch < p.c; ++ch) {
    const int32_t *cdf = scdf + (size_t)ch * p.stride;
    const int size = p.sizes[ch], max_value = size - 2, offset = p.offsets[ch];
    __syncwarp();
    {
      // buckets [128 lane, 128 lane + 128): s = last symbol with cdf[s] <= 16 b
      const int b0 = lane * (kFine / 32);
      int s = 0, hi = size - 1;
      const uint32_t c0 = (uint32_t)b0 << (kPrecision - kFineBits);
      while (hi - s > 1) {
        const int mid = (s + hi) >> 1;
        if ((uint32_t)cdf[mid] <= c0) s = mid; else hi = mid;
      }
      for (int b = b0; b < b0 + kFine / 32; ++b) {
        const uint32_t cb = (uint32_t)b << (kPrecision - kFineBits);
        while ((uint32_t)cdf[s + 1] <= cb) ++s;
        const uint32_t start = (uint32_t)cdf[s];
        fine[b] = make_uint2(start | ((uint32_t)s << 16), (uint32_t)cdf[s + 1] - start);
      }
    }
    __syncwarp();
    for (int j = 0; j < chunks; ++j) {
      const int m = min(kChunk, p.hw - j * kChunk);
      if (live) {
        r.chunk_begin(kChunk);
        int32_t *mine = stage + lane * kPitch;
        // one symbol; the two rare cases (bucket straddling a boundary, escape) are out of line.
        // (Running groups of eight symbols speculatively as one straight-line block with a
        // roll-back on the rare cases was measured and is no faster: 31.8 vs 31.0 ms.)
        auto careful = [&](int i) {
          const uint32_t cf = (uint32_t)x & 0xffffu;
          const uint2 fe = fine[cf >> (kPrecision - kFineBits)];
          uint32_t start = fe.x & 0xffffu, freq = fe.y;
          int lo = (int)(fe.x >> 16);
          if (cf - start >= freq) {
            const uint64_t w = dec_walk(cdf, cf, lo);
            lo = (int)(w >> 32);
            start = (uint32_t)w;
            freq = (uint32_t)cdf[lo + 1] - start;
          }
          x = (uint64_t)freq * (x >> kPrecision) + (cf - start);
          dec_renorm(x, r);
          int value = lo;
          if (lo == max_value) {
            const DecState st = dec_escape(x, r, max_value);
            x = st.x;
            r.rd = st.rd;
            value = st.value;
          }
          mine[i] = value + offset;
        };
#pragma unroll 4
        for (int i0 = 0; i0 < m; ++i0) careful(i0);
      }
      __syncwarp();
      const int i = j * kChunk + lane;
      if (i < p.hw) {
        int32_t *dst = out0 + (size_t)ch * p.hw + i;
#pragma unroll 8
        for (int s = 0; s < n_here; ++s) dst[(size_t)s * per_stream] = stage[s * kPitch + lane];
      }
      __syncwarp();
    }
  }
  if (live && r.rd > r.n_words) atomicOr(p.status, 2);
}

__global__ void __launch_bounds__(32) rans_decode_warp_kernel(const RansDecParams p, int coarse) {
  extern __shared__ __align__(16) uint8_t dec_smem[];
  const int lane = threadIdx.x;
  int32_t *stage = reinterpret_cast<int32_t *>(dec_smem);                 // [32][33]
  int32_t *scdf = stage + kChunk * kPitch + 8;
  uint8_t *first = reinterpret_cast<uint8_t *>(scdf + (p.cdf_in_smem ? p.c * p.stride : 0));
  const int32_t *cdfs = p.cdfs;
  if (p.cdf_in_smem) {
    for (int i = lane; i < p.c * p.stride; i += 32) scdf[i] = p.cdfs[i];
    cdfs = scdf;
  }
  __syncwarp();
  if (coarse) {
    // first[ch][b] = last symbol whose cumulative count is <= 256 b: the search for a
    // cumulative frequency cf starts there and walks forward (peaked densities: 0-2 steps)
    for (int ch = lane; ch < p.c; ch += 32) {
      const int32_t *cdf = cdfs + (size_t)ch * p.stride;
      const int size = p.sizes[ch];
      for (int s = 0; s + 1 < size; ++s) {
        const int b0 = (cdf[s] + 255) >> 8, b1 = (cdf[s + 1] + 255) >> 8;
        for (int b = b0; b < b1 && b < 256; ++b) first[ch * 256 + b] = (uint8_t)s;
      }
    }
  }
  __syncwarp();
  const int k0 = blockIdx.x * 32;
  const int k = k0 + lane;
  const bool live = k < p.n;
  const int n_here = min(32, p.n - k0);
  WordReader r;
  uint64_t x = kRansL;
  if (live) {
    const uint32_t *b = p.words + p.off[k], *e = p.words + p.off[k + 1];
    r.init(b, e);
    const uint32_t w0 = r.take(), w1 = r.take();
    x = (uint64_t)w0 | ((uint64_t)w1 << 32);
  } else {
    r.init(nullptr, nullptr);
  }
  const size_t per_stream = (size_t)p.c * p.hw;
  int32_t *out0 = p.symbols + (size_t)k0 * per_stream;
  const int chunks = (p.hw + kChunk - 1) / kChunk;
  for (int ch = 0; ch < p.c; ++ch) {
    const int32_t *cdf = cdfs + (size_t)ch * p.stride;
    const uint8_t *fst = first + ch * 256;
    const int size = p.sizes[ch], max_value = size - 2, offset = p.offsets[ch];
    for (int j = 0; j < chunks; ++j) {
      const int m = min(kChunk, p.hw - j * kChunk);
      if (live) {
        int32_t *mine = stage + lane * kPitch;
        for (int i = 0; i < m; ++i) {
          const uint32_t cf = (uint32_t)(x & 0xffffu);
          int lo;
          if (coarse) {
            lo = fst[cf >> 8];
            while ((uint32_t)cdf[lo + 1] <= cf) ++lo;
          } else {
            lo = 0;
            int hi = size - 1;          // last entry with cdf[s] <= cf
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if ((uint32_t)cdf[mid] <= cf) lo = mid; else hi = mid;
            }
          }
          const uint32_t start = (uint32_t)cdf[lo], freq = (uint32_t)cdf[lo + 1] - start;
          x = (uint64_t)freq * (x >> kPrecision) + cf - start;
          if (x < kRansL) x = (x << 32) | r.take();
          int value = lo;
          if (lo == max_value) {
            int v = (int)dec_nibble_w(x, r), groups = v;
            while (v == kMaxBypass && !r.bad) { v = (int)dec_nibble_w(x, r); groups += v; }
            uint32_t raw = 0;
            for (int g = 0; g < groups; ++g) {
              const uint32_t nib = dec_nibble_w(x, r);
              if (g < 8) raw |= nib << (g * kBypassBits);
            }
            value = (int)(raw >> 1);
            value = (raw & 1u) ? -value - 1 : value + max_value;
          }
          mine[i] = value + offset;
        }
      }
      __syncwarp();
      const int i = j * kChunk + lane;
      if (i < p.hw) {
        int32_t *dst = out0 + (size_t)ch * p.hw + i;
#pragma unroll 8
        for (int s = 0; s < n_here; ++s) dst[(size_t)s * per_stream] = stage[s * kPitch + lane];
      }
      __syncwarp();
    }
  }
  if (live && r.bad) atomicOr(p.status, 2);
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
  if (enc_table && !cae_knob(CAE_KNOB_RANS_V1)) {
    const size_t tbytes = cae_rans_enc_table_bytes(c, cdf_stride);
    p.table_in_smem = tbytes <= 160 * 1024;
    const size_t smem = kChunk * kPitch * 4 + 32 + (p.table_in_smem ? tbytes : 0);
    if (smem > 48 * 1024)
      CAE_CUDA(cudaFuncSetAttribute(rans_encode_warp_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (p.table_in_smem)
      rans_encode_warp_kernel<true><<<(n + 31) / 32, 32, smem, (cudaStream_t)stream>>>(p);
    else
      rans_encode_warp_kernel<false><<<(n + 31) / 32, 32, smem, (cudaStream_t)stream>>>(p);
    cae_count_launch();
    CAE_CUDA(cudaGetLastError());
    return 0;
  }
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
  if (!cae_knob(CAE_KNOB_RANS_V1) && !cae_knob(CAE_KNOB_RANS_V2) && cdf_stride <= 65535 &&
      cbytes <= 160 * 1024) {
    const size_t smem = kFine * 8 + 32 * kRing * 4 + (kChunk * kPitch + 8) * 4 + cbytes + 16;
    if (smem > 48 * 1024)
      CAE_CUDA(cudaFuncSetAttribute(rans_decode_fine_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rans_decode_fine_kernel<<<(n + 31) / 32, 32, smem, (cudaStream_t)stream>>>(p);
    cae_count_launch();
    CAE_CUDA(cudaGetLastError());
    return 0;
  }
  if (!cae_knob(CAE_KNOB_RANS_V1)) {
    // coarse start table (uint8 symbol indices): every table must have <= 256 entries
    const int coarse = cdf_stride <= 256 && cbytes + (size_t)c * 256 <= 160 * 1024;
    p.cdf_in_smem = coarse || cbytes <= 160 * 1024;
    const size_t smem = (kChunk * kPitch + 8) * 4 + (p.cdf_in_smem ? cbytes : 0) +
                        (coarse ? (size_t)c * 256 : 0) + 16;
    if (smem > 48 * 1024)
      CAE_CUDA(cudaFuncSetAttribute(rans_decode_warp_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rans_decode_warp_kernel<<<(n + 31) / 32, 32, smem, (cudaStream_t)stream>>>(p, coarse);
    cae_count_launch();
    CAE_CUDA(cudaGetLastError());
    return 0;
  }
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

extern "C" int cae_rans_scan(const int32_t *nwords, int n, int64_t *out_offsets, void *stream) {
  CAE_CHECK(nwords && out_offsets && n > 0, 2, "cae_rans_scan: bad argument");
  rans_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(nwords, n, out_offsets);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
