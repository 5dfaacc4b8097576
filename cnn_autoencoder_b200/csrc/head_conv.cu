// Fused image-side head of an analysis track: the channels_org -> channels_org stride-1
// stem and the channels_org -> channels_net stride-2 convolution of the reference's first
// DownsamplingUnit (src/models/tasks/_autoencoders.py:53-77: Conv2d k3 s1 reflect ->
// activation -> Conv2d k3 s2 reflect [-> activation]) as ONE kernel:
//
//   out = act2( conv_s2( act1( conv_s1(image) + b1 ) ) + b2 )
//
// Unfused, the stem writes a 16-channel-padded fp16 tensor at full image resolution that
// the tensor-core layer reads straight back (2 x 268 MB at 128 x 256^2, for 3 real
// channels) and the stride-2 layer spends 9 K=16 MMAs per tile on 3 channels.  Here the
// stem result never leaves the SM: per 8 x 32 output tile a group of six CUDA-core warps
// computes the 17 x 65 stem patch from the uint8 window into shared memory, gathers it as
// an im2col A operand with K = 9 * c_in taps + 2 bias columns (29 -> 32) in the UMMA K-major
// core-matrix layout, and issues K/16 tcgen05 MMAs per 128 pixels into its TMEM accumulator;
// two such groups work on alternating tiles, eight epilogue warps drain the accumulators
// (activation, fp16 planar stores with the consumer's reflect halo).  Measured per tile and SM:
// ~2.2 us of stem arithmetic, ~1.2 us window staging, ~1.2 us gather + MMA issue per group
// (two tiles in flight), ~2 us epilogue; 221 us for 128 x 256^2 (504 us unfused).
#include <stdlib.h>

#include "cae_common.cuh"

namespace {

constexpr int TH = 8, TW = 32;                 // output tile (two M=128 accumulators)
constexpr int SR = 2 * TH + 1, SC = 2 * TW + 1;  // stem patch
constexpr int SPITCH = SC + 1;
// image window: one halo ring for the plain unit, two for the residual unit (two stacked
// stride-1 convolutions); rows are 16-byte aligned and long enough for the widest work unit
template <bool RES>
struct Win {
  static constexpr int HALO = RES ? 2 : 1;
  static constexpr int WR = SR + 2 * HALO, WC = SC + 2 * HALO;
  static constexpr int WPITCH = RES ? 76 : 72;
  // residual unit: first convolution's output (fp32), one ring larger than the stem patch
  static constexpr int R1 = SR + 2, C1 = SC + 2, P1 = 72;
  static constexpr int PX1 = 8, G1 = (C1 + PX1 - 1) / PX1;     // 9 x 19 = 171 work units
};
constexpr int kLut = 260;                        // x/255 for a byte; entry 256 = 0 (zero padding)
constexpr int PXT = 6;                           // stem pixels per thread: 11 x 17 = 187 work units
constexpr int SG = (SC + PXT - 1) / PXT;         // pixel groups per patch row
// The CUDA-core work is done by two independent groups of warps on alternating tiles (group g
// fills A buffer g and accumulator g, and issues its own MMAs): two tiles are in flight per SM,
// so one group's arithmetic overlaps the other's latency-bound phases and barriers.
constexpr int kGroupWarps = 6, kGroupThreads = 32 * kGroupWarps;   // >= SR * SG: one pass per tile
constexpr int kStemWarps = 2 * kGroupWarps, kStemThreads = 32 * kStemWarps;
constexpr int kEpiWarps = 8;
constexpr int kHeadThreads = kStemThreads + 32 * kEpiWarps;   // 12 + 8 warps = 640
static_assert(SR * SG <= kGroupThreads && Win<true>::R1 * Win<true>::G1 <= kGroupThreads,
              "each stem stage is one pass of a group");
// (any 8 consecutive warps cover the four TMEM lane quadrants twice: (m, quadrant) below)

struct HeadParams {
  int n, h_in, w_in, h_out, w_out, c_out, N;
  int tiles_x, tiles_y, n_tiles;
  int in_fmt, pad_mode;
  const void *in;
  ActView out;
  const float *w1, *b1, *w2, *b2;
  const float *w1b, *b1b;      // residual unit: second stride-1 convolution
  float s1, s2, sm;            // activation slopes: stem, down, (residual) after the add
  uint32_t idesc;
  uint32_t tmem_cols;
  unsigned long long *trace;   // CAE_HEAD_TRACE: per-tile phase timestamps of CTA 0 (ns)
  int debug;   // CAE_HEAD_DEBUG: 1 = no output stores, 2 = no stem compute / gather (timing only)
};

__device__ __forceinline__ int reflect_i(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

__device__ __forceinline__ void stem_bar(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kGroupThreads));
}

template <int CI, bool RES>
struct HeadSmem {
  using W = Win<RES>;
  // K columns: 9 * CI taps, then two columns of ones that carry the bias as an fp16 hi + lo
  // pair (exact to ~2^-22 relative), so the epilogue has no bias add
  static constexpr int K = 9 * CI, KB = K + 2, KP = (KB + 15) / 16 * 16, KG = KP / 8;
  static constexpr int A_BYTES = 2 * KG * 2048;              // one buffer: two M tiles
  static constexpr int WIN_ELEMS = CI * W::WR * W::WPITCH;
  static constexpr int S1_ELEMS = RES ? CI * W::R1 * W::P1 : 0;
  static constexpr int S_ELEMS = CI * SR * SPITCH;
  static constexpr int W1N = (RES ? 2 : 1) * (CI * 3 * CI * 4 + 4);   // [ci][kh][co][4] weights + bias, per stem conv

};

template <int CI, bool U8, bool RES>
__global__ void __launch_bounds__(kHeadThreads, 1) head_conv_kernel(const HeadParams p) {
  using L = HeadSmem<CI, RES>;
  using W = Win<RES>;
  constexpr int WR = W::WR, WC = W::WC, WPITCH = W::WPITCH, HALO = W::HALO;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sA = smem;                                   // [2][2][KG][128][16 B]
  uint8_t *sB = sA + 2 * L::A_BYTES;                    // [KG][N][16 B]
  float *win0 = reinterpret_cast<float *>(sB + L::KG * p.N * 16);          // one window per group
  float *s1_0 = win0 + 2 * L::WIN_ELEMS;                                     // residual: first conv's output
  __half *S0 = reinterpret_cast<__half *>(s1_0 + 2 * L::S1_ELEMS);          // one stem patch per group
  float *lut = reinterpret_cast<float *>(S0 + 2 * ((L::S_ELEMS + 7) & ~7));
  float *w1s = lut + kLut;                               // [ci][kh][co][4] + bias
  uint64_t *bars = reinterpret_cast<uint64_t *>(w1s + L::W1N);
  uint64_t *a_empty = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ------------------------------------------------------------ prologue
  for (int i = tid; i < kLut; i += kHeadThreads) lut[i] = i < 256 ? (float)i / 255.0f : 0.f;
  for (int i = tid; i < CI * 3 * CI * 4; i += kHeadThreads) {
    // smem order [ci][kh][co][kw padded to 4]: one 16-byte load per (ci, kh, co)
    const int kw = i & 3, co = (i >> 2) % CI, kh = ((i >> 2) / CI) % 3, ci = (i >> 2) / (3 * CI);
    w1s[i] = kw < 3 ? p.w1[((co * CI + ci) * 3 + kh) * 3 + kw] : 0.f;
  }
  if (tid < 4) w1s[CI * 3 * CI * 4 + tid] = (p.b1 && tid < CI) ? p.b1[tid] : 0.f;
  constexpr int WSET = CI * 3 * CI * 4 + 4;
  if (RES) {
    for (int i = tid; i < CI * 3 * CI * 4; i += kHeadThreads) {
      const int kw = i & 3, co = (i >> 2) % CI, kh = ((i >> 2) / CI) % 3, ci = (i >> 2) / (3 * CI);
      w1s[WSET + i] = kw < 3 ? p.w1b[((co * CI + ci) * 3 + kh) * 3 + kw] : 0.f;
    }
    if (tid < 4) w1s[WSET + CI * 3 * CI * 4 + tid] = (p.b1b && tid < CI) ? p.b1b[tid] : 0.f;
  }
  for (int i = tid; i < p.N * L::KP; i += kHeadThreads) {
    const int nn = i / L::KP, k = i - nn * L::KP;
    __half hv = __float2half_rn(0.f);
    if (nn < p.c_out && k < L::K) {
      hv = __float2half_rn(p.w2[(size_t)nn * L::K + k]);
    } else if (nn < p.c_out && k < L::KB && p.b2) {
      const float b = p.b2[nn];
      const __half hi = __float2half_rn(b);
      hv = k == L::K ? hi : __float2half_rn(b - __half2float(hi));
    }
    *reinterpret_cast<__half *>(sB + ((size_t)(k >> 3) * p.N + nn) * 16 + (k & 7) * 2) = hv;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  fence_proxy_async();            // B operand written with generic stores, read by the MMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < 2 * kGroupWarps) {
    // ================================================= stem + im2col groups
    const int grp = warp / kGroupWarps, gwarp = warp - grp * kGroupWarps;
    const int gtid = tid - grp * kGroupThreads;
    float *win = win0 + grp * L::WIN_ELEMS;
    float *s1 = s1_0 + grp * L::S1_ELEMS;
    __half *S = S0 + grp * ((L::S_ELEMS + 7) & ~7);
    // The window is read line by line: a line is one image row of the window (uint8 HWC:
    // WC * CI contiguous bytes) or one (channel, row) of an fp32 NCHW image (WC floats).
    // Warp w of a group owns lines w, w + 5, ...; a lane owns elements lane, lane + 32, ... of
    // a line, so the element -> (column, channel) split is the same for every line and tile.
    constexpr int LINES = U8 ? WR : WR * CI, ELEMS = U8 ? WC * CI : WC;
    constexpr int LPW = (LINES + kGroupWarps - 1) / kGroupWarps, EPL = (ELEMS + 31) / 32;
    float pf[LPW][EPL];
    int soff[EPL];                  // window offset of element lane + 32 i within its line
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int j = lane + 32 * i;
      soff[i] = U8 ? (j % CI) * WR * WPITCH + j / CI : j;
    }
    auto prefetch = [&](int tile) {
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int wy0 = 2 * tyi * TH - 1 - HALO, wx0 = 2 * txi * TW - 1 - HALO;
      if (wx0 >= 0 && wx0 + WC <= p.w_in) {
        // no column of the window leaves the image (the common case): a line is one
        // contiguous run at constant offsets; its row resolves the padding by itself
#pragma unroll
        for (int l = 0; l < LPW; ++l) {
          const int line = gwarp + l * kGroupWarps;
          if (line < LINES) {
            const int r = U8 ? line : line % WR, c = U8 ? 0 : line / WR;
            const int gy = wy0 + r;
            const bool yok = p.pad_mode == CAE_PAD_REFLECT || (gy >= 0 && gy < p.h_in);
            const int ry = reflect_i(gy, p.h_in);
            if (U8) {
              const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in) +
                                   (((size_t)n * p.h_in + ry) * p.w_in + wx0) * CI + lane;
#pragma unroll
              for (int i = 0; i < EPL; ++i)
                if (lane + 32 * i < ELEMS)
                  pf[l][i] = __uint_as_float(yok ? (uint32_t)__ldg(src + 32 * i) : 256u);
            } else {
              const float *src = reinterpret_cast<const float *>(p.in) +
                                 (((size_t)n * CI + c) * p.h_in + ry) * p.w_in + wx0 + lane;
#pragma unroll
              for (int i = 0; i < EPL; ++i)
                if (lane + 32 * i < ELEMS) pf[l][i] = yok ? __ldg(src + 32 * i) : 0.f;
            }
          }
        }
        return;
      }
      int xoff[EPL];
      bool xok[EPL];
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const int j = lane + 32 * i;
        const int col = U8 ? j / CI : j, c = U8 ? j - col * CI : 0;
        const int gx = wx0 + col;
        xok[i] = j < ELEMS && (p.pad_mode == CAE_PAD_REFLECT || (gx >= 0 && gx < p.w_in));
        xoff[i] = U8 ? reflect_i(gx, p.w_in) * CI + c : reflect_i(gx, p.w_in);
      }
#pragma unroll
      for (int l = 0; l < LPW; ++l) {
        const int line = gwarp + l * kGroupWarps;
        const int r = U8 ? line : line % WR, c = U8 ? 0 : line / WR;
        const int gy = wy0 + r;
        const bool yok = line < LINES && (p.pad_mode == CAE_PAD_REFLECT || (gy >= 0 && gy < p.h_in));
        const size_t base = U8 ? ((size_t)n * p.h_in + reflect_i(gy, p.h_in)) * p.w_in * CI
                               : (((size_t)n * CI + c) * p.h_in + reflect_i(gy, p.h_in)) * p.w_in;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          if (U8) {
            const uint32_t b = (yok && xok[i])
                                   ? __ldg(reinterpret_cast<const uint8_t *>(p.in) + base + xoff[i])
                                   : 256u;
            pf[l][i] = __uint_as_float(b);
          } else {
            pf[l][i] = (yok && xok[i]) ? __ldg(reinterpret_cast<const float *>(p.in) + base + xoff[i])
                                       : 0.f;
          }
        }
      }
    };
    // group g takes the CTA's tiles number g, g + 2, ... and always fills A buffer g
    const int tstep = 2 * (int)gridDim.x;
    int tile = blockIdx.x + grp * (int)gridDim.x;
    if (tile < p.n_tiles) prefetch(tile);
    for (int k = 0; tile < p.n_tiles; tile += tstep, ++k) {
      const int it = 2 * k + grp;
      const int rem = tile % tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int oy0 = tyi * TH, ox0 = txi * TW;
      const int lo_y = 2 * oy0 - 1, lo_x = 2 * ox0 - 1;
      const bool tr = p.trace && blockIdx.x == 0 && gtid == 0 && it < 64;
      if (tr) p.trace[it * 8 + 0] = global_timer_ns();
      // 1. staged window -> shared memory (fp32, /255 through the exact table)
#pragma unroll
      for (int l = 0; l < LPW; ++l) {
        const int line = gwarp + l * kGroupWarps;
        if (line >= LINES) break;
        float *wl = win + line * WPITCH;       // fp32 NCHW: line = c * WR + r
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          if (lane + 32 * i >= ELEMS) break;
          wl[soff[i]] = U8 ? lut[__float_as_uint(pf[l][i])] : pf[l][i];
        }
      }
      stem_bar(grp);
      if (tr) p.trace[it * 8 + 1] = global_timer_ns();
      // 2. next tile's window loads fly while this tile is computed
      if (tile + tstep < p.n_tiles) prefetch(tile + tstep);
      // 3. stem arithmetic.  A work unit = (patch row, PX adjacent patch columns), one per
      //    thread; a source row segment is a few 8-byte shared loads, the 3 x CI weights of a
      //    (ci, kh) are CI 16-byte broadcast loads.
      const float *src2 = win;             // what the convolution feeding the patch reads
      int src2_rows = WR, src2_pitch = WPITCH;
      const float *wset2 = w1s;
      if (RES) {
        // 3a. residual unit, first convolution (R:114-124): S1 = act(conv_a(x) + b_a) in fp32 on
        //     the patch grown by one ring
        constexpr int PX1 = W::PX1;
        if (gtid < W::R1 * W::G1 && !(p.debug & 2)) {
          const int r = gtid / W::G1, x0 = (gtid - r * W::G1) * PX1;
          float acc[PX1][CI];
#pragma unroll
          for (int q = 0; q < PX1; ++q)
#pragma unroll
            for (int co = 0; co < CI; ++co) acc[q][co] = w1s[CI * 3 * CI * 4 + co];
#pragma unroll
          for (int ci = 0; ci < CI; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const float *wrow = win + (ci * WR + r + kh) * WPITCH + x0;     // 16-byte aligned
              float w10[PX1 + 2];
#pragma unroll
              for (int j = 0; j < PX1 + 2; j += 2) {
                const float2 a2 = *reinterpret_cast<const float2 *>(wrow + j);
                w10[j] = a2.x;
                w10[j + 1] = a2.y;
              }
#pragma unroll
              for (int co = 0; co < CI; ++co) {
                const float4 wv = *reinterpret_cast<const float4 *>(w1s + ((ci * 3 + kh) * CI + co) * 4);
#pragma unroll
                for (int q = 0; q < PX1; ++q) {
                  acc[q][co] = fmaf(w10[q], wv.x, acc[q][co]);
                  acc[q][co] = fmaf(w10[q + 1], wv.y, acc[q][co]);
                  acc[q][co] = fmaf(w10[q + 2], wv.z, acc[q][co]);
                }
              }
            }
#pragma unroll
          for (int q = 0; q < PX1; ++q)
            if (x0 + q < W::C1) {
#pragma unroll
              for (int co = 0; co < CI; ++co) {
                const float v = acc[q][co];
                s1[(co * W::R1 + r) * W::P1 + x0 + q] = fmaxf(v, v * p.s1);
              }
            }
        }
        stem_bar(grp);
        // 3b. the second convolution pads S1 by itself (R:125-137): positions of the ring that lie
        //     outside the image take the mirrored (or zero) value -- border tiles only
        const int g1y = lo_y - 1, g1x = lo_x - 1;
        if (g1y < 0 || g1y + W::R1 > p.h_in || g1x < 0 || g1x + W::C1 > p.w_in) {
          for (int e = gtid; e < W::R1 * W::C1; e += kGroupThreads) {
            const int r = e / W::C1, c = e - r * W::C1;
            const int gy = g1y + r, gx = g1x + c;
            if (gy >= 0 && gy < p.h_in && gx >= 0 && gx < p.w_in) continue;
            int sr = reflect_i(gy, p.h_in) - g1y, sc = reflect_i(gx, p.w_in) - g1x;
            const bool inside = sr >= 0 && sr < W::R1 && sc >= 0 && sc < W::C1 &&
                                p.pad_mode == CAE_PAD_REFLECT;
#pragma unroll
            for (int co = 0; co < CI; ++co)
              s1[(co * W::R1 + r) * W::P1 + c] = inside ? s1[(co * W::R1 + sr) * W::P1 + sc] : 0.f;
          }
          stem_bar(grp);
        }
        src2 = s1;
        src2_rows = W::R1;
        src2_pitch = W::P1;
        wset2 = w1s + WSET;
      }
      // 3c. the convolution that produces the stem patch (plain unit: the only one)
      if (gtid < SR * SG && !(p.debug & 2)) {
        const int r = gtid / SG, x0 = (gtid - r * SG) * PXT;
        float acc[PXT][CI];
#pragma unroll
        for (int q = 0; q < PXT; ++q)
#pragma unroll
          for (int co = 0; co < CI; ++co) acc[q][co] = wset2[CI * 3 * CI * 4 + co];
#pragma unroll
        for (int ci = 0; ci < CI; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float *wrow = src2 + (ci * src2_rows + r + kh) * src2_pitch + x0;     // 8-byte aligned
            float w6[PXT + 2];
#pragma unroll
            for (int j = 0; j < PXT + 2; j += 2) {
              const float2 a2 = *reinterpret_cast<const float2 *>(wrow + j);
              w6[j] = a2.x;
              w6[j + 1] = a2.y;
            }
#pragma unroll
            for (int co = 0; co < CI; ++co) {
              const float4 wv = *reinterpret_cast<const float4 *>(wset2 + ((ci * 3 + kh) * CI + co) * 4);
#pragma unroll
              for (int q = 0; q < PXT; ++q) {
                acc[q][co] = fmaf(w6[q], wv.x, acc[q][co]);
                acc[q][co] = fmaf(w6[q + 1], wv.y, acc[q][co]);
                acc[q][co] = fmaf(w6[q + 2], wv.z, acc[q][co]);
              }
            }
          }
        // a zero-padded consumer sees zeros outside the stem's own output domain
        const int gy = lo_y + r;
        const bool row_in = gy >= 0 && gy < p.h_in;
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
          const int gx = lo_x + x0 + q;
          const bool in = row_in && gx >= 0 && gx < p.w_in;
          if (x0 + q < SC) {
#pragma unroll
            for (int co = 0; co < CI; ++co) {
              float v = acc[q][co];
              if (RES) {
                // fx = conv_b(S1) + b_b + x, then the unit's activation (R:172 and :139-141)
                v += win[(co * WR + r + HALO) * WPITCH + x0 + q + HALO];
                v = fmaxf(v, v * p.sm);
              } else {
                v = fmaxf(v, v * p.s1);
              }
              S[(co * SR + r) * SPITCH + x0 + q] = __float2half_rn(in ? v : 0.f);
            }
          }
        }
      }
      stem_bar(grp);
      if (tr) p.trace[it * 8 + 2] = global_timer_ns();
      // 4. im2col gather into the A buffer: one output pixel (one M row) at a time
      mbar_wait(&a_empty[grp], (k & 1) ^ 1);
      if (tr) p.trace[it * 8 + 3] = global_timer_ns();
#pragma unroll 1
      for (int px = gtid; px < 2 * 128 && !(p.debug & 2); px += kGroupThreads) {
        const int m = px >> 7, row = px & 127;
        const int ty = m * 4 + (row >> 5), tx = row & 31;
        int ry[3], rx[3];
        bool vy[3], vx[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int gy = 2 * (oy0 + ty) - 1 + kk, gx = 2 * (ox0 + tx) - 1 + kk;
          vy[kk] = vx[kk] = true;
          if (p.pad_mode != CAE_PAD_REFLECT) {
            vy[kk] = gy >= 0 && gy < p.h_in;
            vx[kk] = gx >= 0 && gx < p.w_in;
          }
          int ly = reflect_i(gy, p.h_in) - lo_y, lx = reflect_i(gx, p.w_in) - lo_x;
          ry[kk] = ly < 0 ? 0 : (ly >= SR ? SR - 1 : ly);   // rows of partial tiles: unused
          rx[kk] = lx < 0 ? 0 : (lx >= SC ? SC - 1 : lx);
        }
        // k = (ci * 3 + kh) * 3 + kw, packed two halves per register
        uint32_t pk[L::KP / 2];
#pragma unroll
        for (int kk = 0; kk < L::KP / 2; ++kk) pk[kk] = 0u;
        pk[L::K >> 1] |= 0x3C00u << ((L::K & 1) * 16);              // the two bias columns = 1.0
        pk[(L::K + 1) >> 1] |= 0x3C00u << (((L::K + 1) & 1) * 16);
        const unsigned short *S16 = reinterpret_cast<const unsigned short *>(S);
#pragma unroll
        for (int ci = 0; ci < CI; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int kk = (ci * 3 + kh) * 3 + kw;
              uint32_t v = S16[(ci * SR + ry[kh]) * SPITCH + rx[kw]];
              if (!(vy[kh] && vx[kw])) v = 0u;
              pk[kk >> 1] |= v << ((kk & 1) * 16);
            }
        uint8_t *dst = sA + (size_t)grp * L::A_BYTES + (size_t)m * L::KG * 2048 + row * 16;
#pragma unroll
        for (int kg = 0; kg < L::KG; ++kg)
          *reinterpret_cast<uint4 *>(dst + kg * 2048) =
              make_uint4(pk[4 * kg], pk[4 * kg + 1], pk[4 * kg + 2], pk[4 * kg + 3]);
      }
      fence_proxy_async();
      stem_bar(grp);
      // 5. the group's first warp issues the tile's MMAs itself (K / 16 per 128 pixels: there is
      //    nothing to gain from a dedicated issuing warp, and the warp slot buys registers)
      if (gwarp == 0) {
        mbar_wait(&acc_empty[grp], (k & 1) ^ 1);       // the epilogue has drained accumulator g
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa_base = smem_u32(sA), sb_base = smem_u32(sB);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const uint32_t d = tmem_base + (uint32_t)((grp * 2 + m) * p.N);
#pragma unroll
            for (int ks = 0; ks < L::KP / 16; ++ks) {
              const uint64_t da = make_smem_desc(
                  sa_base + grp * L::A_BYTES + m * L::KG * 2048 + ks * 2 * 2048, 2048, 128);
              const uint64_t db = make_smem_desc(sb_base + ks * 2 * p.N * 16, p.N * 16, 128);
              umma_f16(d, da, db, p.idesc, ks > 0 ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[grp]);
          umma_commit(&acc_full[grp]);
        }
        __syncwarp();
      }
      if (tr) p.trace[it * 8 + 4] = global_timer_ns();
    }
  } else {
    // ============================================================ epilogue
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int m_first = kEpiWarps == 8 ? (warp - kStemWarps) >> 2 : 0;   // 8 warps: one M tile each
    const int m_step = kEpiWarps == 8 ? 2 : 1;
    const int tx = lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const __half2 post2 = __float2half2_rn(p.s2);
    uint4 *obase = reinterpret_cast<uint4 *>(p.out.ptr);
    const int PW = cae_row_units(p.out.W), PH = p.out.H + 2;
    const size_t plane_stride = (size_t)PH * PW;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int ox = txi * TW + tx;
      mbar_wait_backoff(&acc_full[buf], (it >> 1) & 1);
      if (p.trace && blockIdx.x == 0 && warp == kStemWarps && lane == 0 && it < 64) p.trace[it * 8 + 5] = global_timer_ns();
      tc_fence_after();
#pragma unroll 1
      for (int m = m_first; m < 2; m += m_step) {
        const int oy = tyi * TH + m * 4 + q;
        const bool valid = oy < p.h_out && ox < p.w_out;
        // reflect halo of the consumer: a border pixel is stored again in the halo row, the
        // halo column and the corner (one each unless the image is under 4 pixels wide / high)
        int y2 = -1, y3 = -1, x2 = -1, x3 = -1;
        if (p.out.halo == CAE_HALO_REFLECT) {
          if (oy == 1) y2 = 0;
          if (oy == p.h_out - 2) y3 = p.h_out + 1;
          if (ox == 1) x2 = 0;
          if (ox == p.w_out - 2) x3 = p.w_out + 1;
        }
        const bool border = valid && (y2 & y3 & x2 & x3) >= 0;   // any of them set
        const bool both = (y2 | y3) >= 0 || (x2 | x3) >= 0;      // tiny image: two halos per axis
        const int Yh = y2 >= 0 ? y2 : y3, Xh = x2 >= 0 ? x2 : x3;
        const uint32_t t0 = lane_base + (uint32_t)((buf * 2 + m) * p.N);
        uint4 *px = obase + (size_t)n * p.out.planes * plane_stride + (size_t)(oy + 1) * PW + ox + 1 +
                    CAE_COL_PAD;
        const ptrdiff_t d_x = Xh - (ox + 1), d_y = (ptrdiff_t)(Yh - (oy + 1)) * PW;
        for (int c0 = 0; c0 < p.N; c0 += 32, px += 4 * plane_stride) {
          uint32_t r0[16], r1[16];
          const bool second = c0 + 16 < p.N;
          __syncwarp();
          tmem_ld16(t0 + c0, r0);
          if (second) tmem_ld16(t0 + c0 + 16, r1);
          tmem_ld_wait();
          if (!valid || (p.debug & 1)) continue;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            if (part == 1 && !second) break;
            __align__(16) __half2 h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              h[i] = __floats2half2_rn(__uint_as_float(part ? r1[2 * i] : r0[2 * i]),
                                       __uint_as_float(part ? r1[2 * i + 1] : r0[2 * i + 1]));
              h[i] = __hmax2(h[i], __hmul2(h[i], post2));
            }
            const uint4 lo = *reinterpret_cast<const uint4 *>(&h[0]);
            const uint4 hi = *reinterpret_cast<const uint4 *>(&h[4]);
            uint4 *pl = px + part * 2 * plane_stride;
            pl[0] = lo;
            pl[plane_stride] = hi;
            if (border) {
              auto put = [&](ptrdiff_t d) {
                pl[d] = lo;
                pl[plane_stride + d] = hi;
              };
              if (!both) {
                if (Xh >= 0) put(d_x);
                if (Yh >= 0) {
                  put(d_y);
                  if (Xh >= 0) put(d_y + d_x);
                }
              } else {
                const int ys[3] = {oy + 1, y2, y3}, xs[3] = {ox + 1, x2, x3};
                for (int a = 0; a < 3; ++a)
                  for (int b = 0; b < 3; ++b)
                    if ((a | b) && ys[a] >= 0 && xs[b] >= 0)
                      put((ptrdiff_t)(ys[a] - (oy + 1)) * PW + xs[b] - (ox + 1));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (p.trace && blockIdx.x == 0 && warp == kStemWarps && lane == 0 && it < 64) p.trace[it * 8 + 6] = global_timer_ns();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

template <int CI, bool RES>
size_t head_smem_bytes(int N) {
  using L = HeadSmem<CI, RES>;
  size_t b = 2 * (size_t)L::A_BYTES + (size_t)L::KG * N * 16 + 2 * (size_t)L::WIN_ELEMS * 4 +
             2 * (size_t)L::S1_ELEMS * 4 + 2 * (size_t)((L::S_ELEMS + 7) & ~7) * 2 + kLut * 4 +
             L::W1N * 4 + 8 * 8 + 16;
  return b;
}

template <int CI, bool U8, bool RES = false>
int launch_head(const HeadParams &p, cudaStream_t stream) {
  // one CTA per SM by construction (the TMEM allocation must never wait on a neighbour)
  size_t smem = head_smem_bytes<CI, RES>(p.N);
  if (smem < 120 * 1024) smem = 120 * 1024;
  auto kern = head_conv_kernel<CI, U8, RES>;
  CAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  CAE_CUDA(cudaGetDevice(&dev));
  CAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  if (cae_knob(CAE_KNOB_HEAD_TRACE)) {
    // bring-up aid: phase timestamps of the first 64 tiles of CTA 0 (synchronous, prints to stderr)
    HeadParams q = p;
    unsigned long long *buf = nullptr, host[64 * 8];
    CAE_CUDA(cudaMalloc(&buf, sizeof(host)));
    CAE_CUDA(cudaMemset(buf, 0, sizeof(host)));
    q.trace = buf;
    kern<<<grid, kHeadThreads, smem, stream>>>(q);
    CAE_CUDA(cudaStreamSynchronize(stream));
    CAE_CUDA(cudaMemcpy(host, buf, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(buf);
    fprintf(stderr, "head trace (ns, relative to tile start of the stem group):\n"
                    " tile    start  win->smem  computed  a_empty  mma_issued | acc_full  epi_done  (abs since tile 0)\n");
    for (int t = 0; t < 64 && host[t * 8]; ++t) {
      const unsigned long long *h = host + t * 8, t0 = host[0];
      fprintf(stderr, " %3d %9llu %9llu %9llu %8llu %7llu | %8llu %9llu\n", t, h[0] - t0, h[1] - h[0],
              h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - t0, h[6] - t0);
    }
    cae_count_launch();
    return 0;
  }
  kern<<<grid, kHeadThreads, smem, stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int cae_conv_head(const cae_head_desc *d, void *stream) {
  CAE_CHECK(d, 2, "cae_conv_head: null descriptor");
  CAE_CHECK(d->n > 0 && d->h_in >= 2 && d->w_in >= 2, 2, "cae_conv_head: bad shape %dx%dx%d",
            d->n, d->h_in, d->w_in);
  CAE_CHECK(d->c_in >= 1 && d->c_in <= 4, 2, "cae_conv_head: c_in=%d not in 1..4", d->c_in);
  CAE_CHECK(d->c_out >= 1 && d->c_out <= 128, 2, "cae_conv_head: c_out=%d not in 1..128", d->c_out);
  CAE_CHECK(d->in.ptr && (d->in.fmt == CAE_FMT_U8_HWC || d->in.fmt == CAE_FMT_F32_NCHW), 2,
            "cae_conv_head: input must be U8_HWC or F32_NCHW");
  CAE_CHECK(d->out.ptr && d->out.fmt == CAE_FMT_F16_PLANAR, 2,
            "cae_conv_head: output must be F16_PLANAR");
  CAE_CHECK(d->w_stem && d->w_down, 2, "cae_conv_head: null weights");
  HeadParams p;
  memset(&p, 0, sizeof(p));
  p.n = d->n;
  p.h_in = d->h_in;
  p.w_in = d->w_in;
  p.h_out = (d->h_in - 1) / 2 + 1;
  p.w_out = (d->w_in - 1) / 2 + 1;
  p.c_out = d->c_out;
  p.N = (d->c_out + 15) / 16 * 16;
  CAE_CHECK(d->out.planes * 8 == p.N, 2, "cae_conv_head: output has %d planes, expected %d",
            d->out.planes, p.N / 8);
  p.tiles_x = (p.w_out + TW - 1) / TW;
  p.tiles_y = (p.h_out + TH - 1) / TH;
  const long long nt = (long long)p.n * p.tiles_x * p.tiles_y;
  CAE_CHECK(nt < (1ll << 31), 2, "cae_conv_head: too many tiles");
  p.n_tiles = (int)nt;
  p.in_fmt = d->in.fmt;
  p.pad_mode = d->pad_mode;
  p.in = d->in.ptr;
  p.out = ActView{d->out.ptr, d->out.fmt, d->out.planes, d->out.halo, p.h_out, p.w_out};
  p.w1 = d->w_stem;
  p.b1 = d->b_stem;
  p.w2 = d->w_down;
  p.b2 = d->b_down;
  auto slope = [](int act) { return act == CAE_ACT_LEAKY_RELU ? 0.01f : (act == CAE_ACT_RELU ? 0.f : 1.f); };
  CAE_CHECK(d->act_stem >= CAE_ACT_NONE && d->act_stem <= CAE_ACT_RELU && d->act_down >= CAE_ACT_NONE &&
                d->act_down <= CAE_ACT_RELU, 2, "cae_conv_head: bad activation");
  p.s1 = slope(d->act_stem);
  p.s2 = slope(d->act_down);
  if (d->residual) {
    CAE_CHECK(d->w_stem2, 2, "cae_conv_head: residual unit without its second convolution");
    CAE_CHECK(d->act_mid >= CAE_ACT_NONE && d->act_mid <= CAE_ACT_RELU, 2,
              "cae_conv_head: bad activation");
    p.w1b = d->w_stem2;
    p.b1b = d->b_stem2;
    p.sm = slope(d->act_mid);
  }
  p.idesc = make_idesc_f16(128, p.N);
  uint32_t cols = 32;
  while ((int)cols < 4 * p.N) cols <<= 1;
  p.tmem_cols = cols;
  if (const char *e = cae_knob(CAE_KNOB_HEAD_DEBUG)) p.debug = atoi(e);
  cudaStream_t s = (cudaStream_t)stream;
  const bool u8 = d->in.fmt == CAE_FMT_U8_HWC;
  if (d->residual) {
    switch (d->c_in) {
      case 1: return u8 ? launch_head<1, true, true>(p, s) : launch_head<1, false, true>(p, s);
      case 2: return u8 ? launch_head<2, true, true>(p, s) : launch_head<2, false, true>(p, s);
      case 3: return u8 ? launch_head<3, true, true>(p, s) : launch_head<3, false, true>(p, s);
      default: return u8 ? launch_head<4, true, true>(p, s) : launch_head<4, false, true>(p, s);
    }
  }
  switch (d->c_in) {
    case 1: return u8 ? launch_head<1, true>(p, s) : launch_head<1, false>(p, s);
    case 2: return u8 ? launch_head<2, true>(p, s) : launch_head<2, false>(p, s);
    case 3: return u8 ? launch_head<3, true>(p, s) : launch_head<3, false>(p, s);
    default: return u8 ? launch_head<4, true>(p, s) : launch_head<4, false>(p, s);
  }
}
