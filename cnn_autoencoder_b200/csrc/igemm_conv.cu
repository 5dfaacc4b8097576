// 3x3 (transposed) convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces nn.Conv2d / nn.ConvTranspose2d forward inside the reference's
// Analyzer.forward / Synthesizer.forward (src/models/tasks/_autoencoders.py:
// 53-304, 359-361, 442-455).
//
// Design (see DESIGN.md "igemm"):
//  * activations live in HBM as [N][C/8][H+2][W+2][8] fp16 ("planar", one-pixel
//    halo written by the producing layer: zeros or reflected copies), or the
//    parity-split variant of it for the stride-2 convolutions;
//  * one CTA tile = 16 x (8*mt) output pixels.  A TMA box load brings the input
//    patch (tile + halo, CK channels) into shared memory ONCE; in that layout
//    every 3x3 tap is the same bytes seen through a different start address, so
//    the nine taps are nine UMMA shared-memory descriptors (K-major, no swizzle:
//    core matrix = 8 consecutive pixels x 8 channels = 128 contiguous bytes) and
//    no im2col copy is ever materialised;
//  * weights are pre-packed per (channel chunk, tap) in exactly the shared-memory
//    image and streamed with 1-D bulk copies through a second mbarrier ring;
//  * tcgen05.mma (M=128, N=C_out, K=16, fp16 in / fp32 accumulate) issued by one
//    thread; accumulators in TMEM, double buffered across tiles when they fit;
//  * four epilogue warps read TMEM (tcgen05.ld), apply bias / activation /
//    residual add and write the next layer's input layout directly (halo copies
//    included), or the fp32 latent, or the uint8 image.
// Transposed stride-2 layers run as four output-phase accumulators (pixel
// shuffle in the epilogue); the final C_out<=4 layer merges the four phases into
// one N=16 GEMM over the 2x2 input neighbourhood.
#include <stdlib.h>
#include <string.h>

#include "cae_common.cuh"
#include "eb_device.cuh"

namespace {

constexpr int kMaxTaps = 9;
constexpr int kMaxSA = 6;
constexpr int kMaxSB = 12;
constexpr int kMaxEpiWarps = 12;  // 3 per TMEM lane quadrant
constexpr int kThreads = 128 + 32 * kMaxEpiWarps;

enum { EPI_ACT = 0, EPI_LATENT = 1, EPI_IMAGE = 2, EPI_PROJ = 3 };

// Projection fusion (EPI_PROJ): the last wide transposed layer does not write its 128-channel
// output U at all.  The image layer that follows it (ConvTranspose 128 -> c, c <= 3) is linear
// in U: I[2a+qy][2b+qx] = sum over (dy <= qy, dx <= qx) of V[qy+1-2dy][qx+1-2dx]^T U[a+dy][b+dx],
// so the nine per-pixel products P_t = V_t^T U[a][b] (t = kernel tap; 9 * c <= 27 numbers) are
// all the image layer ever needs from U.  The epilogue of this kernel stages the activated fp16
// tile of a pass (256 pixels x 128 channels) in shared memory as a K-major operand, a dedicated
// warp issues a second GEMM (M = 2 x 128 pixels, N = 32, K = 128) into the columns of the
// accumulator buffer that has just been drained, and P (64 bytes per pixel of U instead of 256,
// never read back as a 3x3 neighbourhood of 128 channels) goes to HBM.  image_from_proj_kernel
// then sums the up to four taps of every output pixel.  No halo, no recomputation: the spatial
// part of the image layer moved behind its channel reduction.
//
// Column order of P (and of the packed projection weights), three slots per tap, chosen so that
// what a neighbour needs from a record is one 16-byte unit:
//   unit 0-1: own taps (1,1) (1,2) (2,1) (2,2) at 0, 3, 6, 9; tap (0,0) at 12 (diagonal neighbour)
//   unit 2:   taps (1,0) (2,0) at 16, 19 (read by the left neighbour)
//   unit 3:   taps (0,1) (0,2) at 24, 27 (read by the upper neighbour)
constexpr int kProjN = 32;
constexpr int kProjK = 128;
constexpr int kProjStageBytes = (kProjK / 8) * 256 * 16;   // 16 planes x 256 pixels x 8 channels fp16
constexpr int kProjWBytes = (kProjK / 8) * kProjN * 16;    // [kplane][n][8] fp16
__host__ __device__ inline int proj_slot(int kh, int kw) {
  const int slot[3][3] = {{12, 24, 27}, {16, 0, 3}, {19, 6, 9}};
  return slot[kh][kw];
}

struct IgTap {
  uint32_t a_off;  // byte offset of the tap's view inside an A stage
  uint32_t acc;    // accumulator (output phase) this tap adds into
};

constexpr int kMaxItems = 18;
constexpr int kMaxStages = 9;

// One MMA group of the per-chunk work list of an issuing warp: tap view offset into the A
// stage (16-byte units, M-tile offset included), weight offset inside the B stage, TMEM
// column offset of the accumulator, and whether this is the first group writing it.
struct IgItem {
  uint32_t a_off16, b_off16, d_off, first;
};

struct IgParams {
  int n_img, dom_h, dom_w;
  int tiles_x, tiles_per_img, n_tiles;
  int mt, n_taps, n_chunks, ck;
  int N, n_acc, n_buf, tmem_cols;
  int PH, PW, n_par, par_stride;
  int a_stage_bytes, a_box_bytes, b_stage_bytes, sa, sb;
  int tpb, b_tap_bytes;  // taps per weight stage, bytes of one tap's weights
  // Passes: a transposed stride-2 layer is run as two passes per tile (output rows 2y and
  // 2y+1), each with two phase accumulators, so the accumulators stay double buffered and the
  // epilogue of one pass overlaps the MMAs of the next.  Every other kind has one pass.
  int n_pass;
  int pass_tap0[2];        // first tap of the pass (taps are ordered by output row phase)
  int n_stages[2];         // weight stages per channel chunk, per pass
  IgItem items[2][2][kMaxItems];              // [pass][issuer][item]
  int item_start[2][2][kMaxStages + 1];
  int epi_warps;  // 4, 8, 12 (or 16 on the fast-epilogue kernel)
  int debug;      // bring-up bit mask (CAE_IGEMM_DEBUG): 1 no A loads, 2 no B loads, 4 no MMA, 8 no stores
  int org_y, org_x;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b, idesc;
  IgTap taps[kMaxTaps];
  const uint8_t *wpack;
  size_t wpack_half_bytes;   // PAIR: bytes of one rank's half of the packed weights
  // resident: the CTA's whole share of the weights is loaded once and stays in shared memory
  // (one "stage" of all chunks; b_chunk16 = 16-byte units per channel chunk)
  int resident;
  uint32_t b_chunk16;
  // epilogue
  int up;  // 1, or 2 for the pixel-shuffling transposed stride-2 layers
  int c_out, pre_act, post_act;
  const float *bias;
  ActView out, skip;
  float *aux;  // optional fp32 NCHW copy (final image layer only)
  int out_h, out_w;
  // 16-byte-unit strides of out / skip (32-bit: buffers are < 2^31 units)
  uint32_t out_pitch, out_ps, out_is, skip_pitch, skip_ps, skip_is;
  int pair_store;   // up == 2, planar output without reflect halo: 32-byte stores of pixel pairs
  // quantizer fused into the latent layer's epilogue (EPI_LATENT only)
  int quant;
  int q_smem;        // 1: medians and the core of the likelihood table / histogram in shared memory
  int q_core_min, q_core_len;   // symbols [q_core_min, q_core_min + q_core_len) of every channel
  cae_eb_tables qt;
  float *q_yq;
  int32_t *q_sym, *q_hist, *q_status;
  double *q_rate;
  ActView q_pl;
  uint32_t q_pitch, q_ps, q_is;
  // projection fusion (EPI_PROJ only)
  const uint8_t *proj_w;   // packed [16][32][8] fp16
  __half *proj_out;        // [n][out_h][out_w][32] fp16
  uint32_t idesc2;
};

struct TapDef {
  int par, dy, dx, acc, kh, kw;
};

// Tap tables shared by the weight packer and the launcher.  kh/kw index the
// torch weight tensor; (par, dy, dx) say where the tap reads inside the patch.
int build_taps(int kind, bool merged, TapDef *t) {
  int n = 0;
  if (kind == CAE_CONV_S1) {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) t[n++] = {0, kh, kw, 0, kh, kw};
  } else if (kind == CAE_CONV_S2) {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        // padded column X = 2 ox + kw is physical column X + CAE_COL_PAD (odd shift): parity
        // (kw + 1) & 1 at half index ox + ((kw + 3) >> 1); the box origin is ox + 1 (org_x)
        t[n++] = {((kh & 1) << 1) | ((kw + CAE_COL_PAD) & 1), kh >> 1,
                  ((kw + CAE_COL_PAD) >> 1) - 1, 0, kh, kw};
  } else if (kind == CAE_CONVT_S1) {
    // conv_transpose(k3,s1,p1) == correlation with the flipped kernel, zero halo
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) t[n++] = {0, dy, dx, 0, 2 - dy, 2 - dx};
  } else if (merged) {
    // one tap per 2x2 input shift; kh/kw depend on the output phase (row of B)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) t[n++] = {0, dy, dx, 0, -1, -1};
  } else {
    // out[2a+py, 2b+px] += W[py+1-2dy, px+1-2dx]^T in[a+dy, b+dx]
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        for (int dy = 0; dy <= py; ++dy)
          for (int dx = 0; dx <= px; ++dx)
            t[n++] = {0, dy, dx, py * 2 + px, py + 1 - 2 * dy, px + 1 - 2 * dx};
  }
  return n;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

inline bool is_merged(int kind, int c_out) { return kind == CAE_CONVT_S2 && 4 * c_out <= 16; }

inline int mma_n(int kind, int c_out) {
  return is_merged(kind, c_out) ? 16 : round_up(c_out, 16);
}

// CTA-pair (cta_group::2) form: the 128-output-channel layers, whose weight stream is what
// bounds the single-CTA kernel.  Decided from the layer alone, so that the packer (which lays
// the weights out as two halves of output channels) and the launcher agree.
//
// Measured on B200 (net A, 128 x 256^2, round 2, gpurun_out/r2_layer_pair*.log): correct, but not
// faster -- conv s1 128->128 447 us either way (that layer is bound by the 64-cycle N = 128
// instruction and the power-limited clock, not by L2 or shared-memory bandwidth), the stride-2
// layers lose 12-40 % (their activation ring next to 144 KB of resident weights is too shallow
// for the L2 latency, and every accumulator hand-over crosses the cluster).  Hence opt-in:
// CAE_DEBUG=1 CAE_IGEMM_PAIR_MMA=1.
inline bool use_pair(int kind, int c_out) {
  return !is_merged(kind, c_out) && mma_n(kind, c_out) == 128 && cae_knob(CAE_KNOB_IGEMM_PAIR_MMA);
}

inline int auto_ck(int kind, int c_in_p, bool merged = false, bool pair = false) {
  if (pair && c_in_p % 32 == 0) {
    // CTA-pair layers keep their share of the weights resident (up to 144 KB); the chunk size
    // then only sets the size of an activation stage: three or more must fit beside the weights
    if (const char *e = cae_knob(CAE_KNOB_IGEMM_CK_PAIR)) {
      const int v = atoi(e);
      if ((v == 16 || v == 32 || v == 64) && c_in_p % v == 0) return v;
    }
    return kind == CAE_CONV_S2 ? 16 : (kind == CAE_CONVT_S2 ? 64 : 32);
  }
  // final image layer (N = 16): the MMAs are tiny and the layer is bound by the barrier round
  // trips per stage, so take the whole K = 128 in one stage
  if (merged && c_in_p % 128 == 0 && !cae_knob(CAE_KNOB_IGEMM_MERGED_CK64)) return 128;
  if (kind == CAE_CONV_S2) {
    if (const char *e = cae_knob(CAE_KNOB_IGEMM_CK_S2)) {      // experiment knob (pack and launch agree)
      const int v = atoi(e);
      if ((v == 16 || v == 32) && c_in_p % v == 0) return v;
    }
    return c_in_p % 32 == 0 ? 32 : 16;
  }
  if (c_in_p % 64 == 0) return 64;
  if (c_in_p % 48 == 0) return 48;
  if (c_in_p % 32 == 0) return 32;
  return 16;
}

// ---------------------------------------------------------------- packing
struct PackParams {
  int kind, merged, c_in, c_out, ck, N, n_taps, n_chunks, pair;
  TapDef taps[kMaxTaps];
};

// packed[chunk][tap][kplane][n][8]: the shared-memory image of operand B.  pair: two such
// images back to back, [rank][chunk][tap][kplane][n - rank * N / 2][8], one per CTA of a pair.
__global__ void pack_weights_kernel(PackParams q, const float *__restrict__ w,
                                    const float *__restrict__ scale, __half *__restrict__ out,
                                    size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k8 = (int)(i & 7);
  size_t r = i >> 3;
  const int n = (int)(r % q.N);
  r /= q.N;
  const int kplane = (int)(r % (q.ck / 8));
  r /= (q.ck / 8);
  const int t = (int)(r % q.n_taps);
  const int chunk = (int)(r / q.n_taps);
  const int ci = chunk * q.ck + kplane * 8 + k8;
  int co = n, kh = q.taps[t].kh, kw = q.taps[t].kw;
  if (q.merged) {
    const int phase = n / q.c_out;
    co = n % q.c_out;
    kh = (phase >> 1) + 1 - 2 * q.taps[t].dy;
    kw = (phase & 1) + 1 - 2 * q.taps[t].dx;
    if (phase >= 4) co = q.c_out;  // padding rows
  }
  float v = 0.f;
  if (ci < q.c_in && co < q.c_out && kh >= 0 && kh < 3 && kw >= 0 && kw < 3) {
    const bool transposed = q.kind == CAE_CONVT_S1 || q.kind == CAE_CONVT_S2;
    const size_t idx = transposed ? (((size_t)ci * q.c_out + co) * 3 + kh) * 3 + kw
                                  : (((size_t)co * q.c_in + ci) * 3 + kh) * 3 + kw;
    v = w[idx];
    if (scale) v *= scale[co];
  }
  size_t o = i;
  if (q.pair) {
    const int half = q.N >> 1, r = n / half, nl = n - r * half;
    o = (size_t)r * (total >> 1) +
        ((((size_t)chunk * q.n_taps + t) * (q.ck / 8) + kplane) * half + nl) * 8 + k8;
  }
  out[o] = __float2half_rn(v);
}

// --------------------------------------------------------------- epilogues
// The epilogue is kept deliberately compact (one copy of the math per kernel, rare paths
// out of line): eight epilogue warps share a ~6 KB L0 instruction cache with the other
// roles, and a large unrolled body stalls them on instruction fetch.
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint4 pack8(const float *v) {
  uint4 u;
  u.x = pack2(v[0], v[1]);
  u.y = pack2(v[2], v[3]);
  u.z = pack2(v[4], v[5]);
  u.w = pack2(v[6], v[7]);
  return u;
}

// max(v, v*slope): slope 1 = identity, 0.01 = LeakyReLU, 0 = ReLU (branch-free activations)
__device__ __forceinline__ float act_slope(int act) {
  return act == CAE_ACT_LEAKY_RELU ? 0.01f : (act == CAE_ACT_RELU ? 0.f : 1.f);
}

// 16-byte-unit offset of plane 0 of padded pixel (Y,X); pitch / ps / is are the row, plane and
// image strides precomputed on the host (is = planes * ps, times 4 parities for SPLIT).
__device__ __forceinline__ uint32_t pixel_unit(int fmt, uint32_t pitch, uint32_t ps, uint32_t is,
                                               int planes, int n, int Y, int X) {
  const uint32_t Xc = (uint32_t)(X + CAE_COL_PAD);
  if (fmt == CAE_FMT_F16_PLANAR) return (uint32_t)n * is + (uint32_t)Y * pitch + Xc;
  const uint32_t par = (uint32_t)(((Y & 1) << 1) | (Xc & 1));
  return (uint32_t)n * is + par * (uint32_t)planes * ps + (uint32_t)(Y >> 1) * pitch + (Xc >> 1);
}

// Mirrored halo copies of two consecutive planes of a border pixel (padding_mode='reflect').
__device__ __noinline__ void store_halo2(const IgParams &p, int n, int plane, int oy, int ox,
                                         uint4 v0, uint4 v1) {
  uint4 *base = reinterpret_cast<uint4 *>(p.out.ptr);
  const int ys[3] = {oy + 1, oy == 1 ? 0 : -1, oy == p.out.H - 2 ? p.out.H + 1 : -1};
  const int xs[3] = {ox + 1, ox == 1 ? 0 : -1, ox == p.out.W - 2 ? p.out.W + 1 : -1};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      if ((a | b) != 0 && ys[a] >= 0 && xs[b] >= 0) {
        const uint32_t off = pixel_unit(p.out.fmt, p.out_pitch, p.out_ps, p.out_is, p.out.planes,
                                        n, ys[a], xs[b]) + (uint32_t)plane * p.out_ps;
        base[off] = v0;
        base[off + p.out_ps] = v1;
      }
}

// Fast path of the planar epilogue: bias in fp32, then convert to half2 and run the
// activations as packed fp16 max(v, v*slope) -- half the instructions of the fp32 form; the
// result only differs by fp16 rounding of already-rounded negatives.
template <bool SKIP>
__device__ __forceinline__ void fast_half16(const IgParams &p, const uint32_t (&r)[16], int c0,
                                            float pre_s, __half2 pre2, __half2 post2,
                                            const uint4 *skip2, __half2 (&h)[8]) {
  if (SKIP) {
    // residual layers: bias, pre-activation and the skip add stay in fp32 (the residual path is
    // the precision-critical one), only the post-activation is packed
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = __uint_as_float(r[i]);
      if (p.bias && c0 + i < p.c_out) v[i] += __ldg(p.bias + c0 + i);
    }
    if (p.pre_act != CAE_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * pre_s);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const uint4 sv = skip2[hh];
      const __half2 *sh = reinterpret_cast<const __half2 *>(&sv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(sh[k]);
        h[hh * 4 + k] = __floats2half2_rn(v[hh * 8 + 2 * k] + f.x, v[hh * 8 + 2 * k + 1] + f.y);
      }
    }
  } else {
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + 2 * i;
        const float b0 = c < p.c_out ? __ldg(p.bias + c) : 0.f;
        const float b1 = c + 1 < p.c_out ? __ldg(p.bias + c + 1) : 0.f;
        h[i] = __floats2half2_rn(__uint_as_float(r[2 * i]) + b0, __uint_as_float(r[2 * i + 1]) + b1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        h[i] = __floats2half2_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    }
    if (p.pre_act != CAE_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __hmax2(h[i], __hmul2(h[i], pre2));
    }
  }
  if (p.post_act != CAE_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __hmax2(h[i], __hmul2(h[i], post2));
  }
}

__device__ __forceinline__ uint32_t h2u(const __half2 &h) {
  return *reinterpret_cast<const uint32_t *>(&h);
}

// Transposed stride-2 layers: one thread owns the two horizontally adjacent output pixels
// 2x, 2x+1 of a plane = 32 contiguous, sector-aligned bytes -> one 256-bit store per plane
// (two 16-byte stores from separate instructions would each write half a sector).
__device__ __forceinline__ void store16_pair(const IgParams &p, const __half2 (&a)[8],
                                             const __half2 (&b)[8], int c0, int n, int oy, int ox0) {
  const uint32_t off = pixel_unit(p.out.fmt, p.out_pitch, p.out_ps, p.out_is, p.out.planes, n,
                                  oy + 1, ox0 + 1) + (uint32_t)(c0 >> 3) * p.out_ps;
  uint4 *dst = reinterpret_cast<uint4 *>(p.out.ptr) + off;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(h2u(a[0])),
               "r"(h2u(a[1])), "r"(h2u(a[2])), "r"(h2u(a[3])), "r"(h2u(b[0])), "r"(h2u(b[1])),
               "r"(h2u(b[2])), "r"(h2u(b[3]))
               : "memory");
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + p.out_ps),
               "r"(h2u(a[4])), "r"(h2u(a[5])), "r"(h2u(a[6])), "r"(h2u(a[7])), "r"(h2u(b[4])),
               "r"(h2u(b[5])), "r"(h2u(b[6])), "r"(h2u(b[7]))
               : "memory");
}

__device__ __forceinline__ void store16(const IgParams &p, __half2 (&h)[8], int c0, int n, int oy,
                                        int ox) {
  uint4 lo, hi;
  lo.x = *reinterpret_cast<uint32_t *>(&h[0]);
  lo.y = *reinterpret_cast<uint32_t *>(&h[1]);
  lo.z = *reinterpret_cast<uint32_t *>(&h[2]);
  lo.w = *reinterpret_cast<uint32_t *>(&h[3]);
  hi.x = *reinterpret_cast<uint32_t *>(&h[4]);
  hi.y = *reinterpret_cast<uint32_t *>(&h[5]);
  hi.z = *reinterpret_cast<uint32_t *>(&h[6]);
  hi.w = *reinterpret_cast<uint32_t *>(&h[7]);
  const uint32_t off = pixel_unit(p.out.fmt, p.out_pitch, p.out_ps, p.out_is, p.out.planes, n,
                                  oy + 1, ox + 1) + (uint32_t)(c0 >> 3) * p.out_ps;
  uint4 *dst = reinterpret_cast<uint4 *>(p.out.ptr) + off;
  if (p.debug & 16) {          // experiment: streaming (evict-first) stores
    __stcs(dst, lo);
    __stcs(dst + p.out_ps, hi);
  } else {
    dst[0] = lo;
    dst[p.out_ps] = hi;
  }
  if (p.out.halo == CAE_HALO_REFLECT &&
      (oy == 1 || oy == p.out.H - 2 || ox == 1 || ox == p.out.W - 2))
    store_halo2(p, n, c0 >> 3, oy, ox, lo, hi);
}

template <bool SKIP>
__device__ __forceinline__ void emit16_fast(const IgParams &p, const uint32_t (&r)[16], int c0,
                                            int n, int oy, int ox, float pre_s, __half2 pre2,
                                            __half2 post2, const uint4 *skip2) {
  __half2 h[8];
  fast_half16<SKIP>(p, r, c0, pre_s, pre2, post2, skip2, h);
  store16(p, h, c0, n, oy, ox);
}

// Final image layer (merged transposed conv): accumulator columns j = phase * CO + c with
// phase = py * 2 + px.  Thread = input pixel (y, x); per output row 2y+py it owns the two
// adjacent pixels 2x, 2x+1 = 2*CO contiguous uint8 in HWC (always 2-byte aligned: out_w even).
template <int CO>
__device__ __forceinline__ void image_store(const IgParams &p, const uint32_t (&r)[16], int n, int y,
                                            int x, float pre_s, float post_s) {
  float t[4 * CO];
#pragma unroll
  for (int j = 0; j < 4 * CO; ++j) {
    float u = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + (j % CO)) : 0.f);
    u = fmaxf(u, u * pre_s);
    t[j] = fmaxf(u, u * post_s);
  }
  if (p.aux) {
#pragma unroll
    for (int j = 0; j < 4 * CO; ++j) {
      const int ph = j / CO, c = j % CO;
      p.aux[(((size_t)n * CO + c) * p.out_h + (y * 2 + (ph >> 1))) * p.out_w + x * 2 + (ph & 1)] = t[j];
    }
  }
  if (p.out.ptr) {
    uint8_t *o = reinterpret_cast<uint8_t *>(p.out.ptr);
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      uint16_t *row = reinterpret_cast<uint16_t *>(
          o + (((size_t)n * p.out_h + (y * 2 + py)) * p.out_w + x * 2) * CO);
#pragma unroll
      for (int q = 0; q < CO; ++q)
        row[q] = (uint16_t)(to_u8_trunc(t[py * 2 * CO + 2 * q]) |
                            ((uint16_t)to_u8_trunc(t[py * 2 * CO + 2 * q + 1]) << 8));
    }
  }
}

// Quantizer fused behind the latent layer (cae_quant_fuse): all 32 lanes call this with 16
// consecutive channels of their pixel (valid = the pixel exists).  Same arithmetic as
// eb_quantize_kernel.  The medians, the likelihood table and the histogram live in shared
// memory when they fit (s_med != nullptr; the histogram is flushed once per CTA at the end),
// the rate is accumulated in a register per thread and reduced once per warp at the end.
struct QuantSmem {
  const float *med, *lut;   // [c_out], [c_out][q_core_len]: the symbols around the median
  int *hist;                // [c_out][q_core_len]
};

__device__ __forceinline__ void quant_emit16(const IgParams &p, const QuantSmem &qs,
                                             const float (&v)[16], int c0, int n, int oy, int ox,
                                             bool valid, float &bits) {
  const size_t cs = (size_t)p.out_h * p.out_w;
  const int bins = p.q_hist ? p.qt.hist_bins : 0;
  const int lane = (int)(threadIdx.x & 31);
  uint4 *pl = nullptr;
  if (p.q_pl.ptr && valid)
    pl = reinterpret_cast<uint4 *>(p.q_pl.ptr) +
         pixel_unit(p.q_pl.fmt, p.q_pitch, p.q_ps, p.q_is, p.q_pl.planes, n, oy + 1, ox + 1) +
         (uint32_t)(c0 >> 3) * p.q_ps;
  // eight channels (one planar unit) at a time keeps the live registers low
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int cb = c0 + 8 * h;
    const size_t o0 = (((size_t)n * p.c_out + cb) * p.out_h + oy) * p.out_w + ox;
    float yq[8];
    int sym[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool on = valid && cb + i < p.c_out;
      const float med = !on ? 0.f : (qs.med ? qs.med[cb + i] : __ldg(p.qt.medians + cb + i));
      const float r = rintf(v[8 * h + i] - med);   // torch.round: half to even
      yq[i] = on ? r + med : 0.f;
      sym[i] = eb_symbol(r);
    }
    if (p.q_rate) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (valid && cb + i < p.c_out) {
          const int li = sym[i] - p.q_core_min;
          float lik;
          if (qs.lut && li >= 0 && li < p.q_core_len) lik = qs.lut[(cb + i) * p.q_core_len + li];
          else lik = eb_lookup(p.qt, cb + i, sym[i], yq[i], p.q_status);   // rare: global table
          bits -= log2f(lik);
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (cb + i < p.c_out) {
          if (p.q_yq) p.q_yq[o0 + i * cs] = yq[i];
          if (p.q_sym) p.q_sym[o0 + i * cs] = sym[i];
        }
      }
    }
    if (bins) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (cb + i < p.c_out) {                  // warp-uniform
          int b = sym[i] - p.qt.hist_min;
          b = b < 0 ? 0 : (b >= bins ? bins - 1 : b);
          if (!valid) b = -1 - lane;             // lanes that do not count never match each other
          const uint32_t peers = __match_any_sync(0xffffffffu, b);
          if (b >= 0 && lane == __ffs(peers) - 1) {
            const int lc = sym[i] - p.q_core_min;
            if (qs.hist && lc >= 0 && lc < p.q_core_len)
              atomicAdd(qs.hist + (cb + i) * p.q_core_len + lc, __popc(peers));
            else
              atomicAdd(p.q_hist + (size_t)(cb + i) * bins + b, __popc(peers));
          }
        }
      }
    }
    if (pl) pl[h * p.q_ps] = pack8(yq);
  }
}

// One epilogue job = two 16-column TMEM loads in flight, then the math and stores.
//  up == 1: columns [c0, c0+32) of accumulator m
//  up == 2: columns [c0, c0+16) of the two horizontal output phases (py,0) and (py,1), i.e.
//           two adjacent output pixels -> 32 contiguous bytes per plane
template <int EPI, int FAST>
__device__ __forceinline__ void epilogue_job(const IgParams &p, uint32_t tmem_lane_base,
                                             int acc_base, int n, int y, int x, int job,
                                             bool valid, float pre_s, float post_s, int pass,
                                             const QuantSmem &qs, float &q_bits) {
  uint32_t r0[16], r1[16];
  if (EPI == EPI_IMAGE) {
    // merged final transposed layer: one 16-column accumulator, columns j = phase * c_out + c
    __syncwarp();
    tmem_ld16(tmem_lane_base + (uint32_t)(acc_base * p.N), r0);
    tmem_ld_wait();
    if (!valid) return;
    switch (p.c_out) {
      case 1: image_store<1>(p, r0, n, y, x, pre_s, post_s); break;
      case 2: image_store<2>(p, r0, n, y, x, pre_s, post_s); break;
      case 3: image_store<3>(p, r0, n, y, x, pre_s, post_s); break;
      default: image_store<4>(p, r0, n, y, x, pre_s, post_s); break;
    }
    return;
  }

  int c_first, c_second, oy, ox0;
  bool second = true;
  if (p.up == 1) {
    c_first = job * 32;
    c_second = c_first + 16;
    second = c_second < p.N;
    oy = y;
    ox0 = x;
    const uint32_t t = tmem_lane_base + (uint32_t)(acc_base * p.N + c_first);
    __syncwarp();
    tmem_ld16(t, r0);
    if (second) tmem_ld16(t + 16, r1);
  } else {
    // two-pass layers hold one output row phase (two accumulators) per pass
    const int per_row = p.N >> 4;
    const int prow = p.n_pass == 2 ? 0 : job / per_row;
    const int py = p.n_pass == 2 ? pass : prow;
    c_first = c_second = (job - prow * per_row) * 16;
    oy = y * 2 + py;
    ox0 = x * 2;
    const uint32_t t = tmem_lane_base + (uint32_t)((acc_base + prow * 2) * p.N + c_first);
    __syncwarp();
    tmem_ld16(t, r0);
    tmem_ld16(t + (uint32_t)p.N, r1);
  }
  uint4 sk[4] = {};
  if (FAST == 2 && valid && !(p.debug & 32)) {
    // residual input of this job: in flight together with the TMEM loads
    const int ox1 = ox0 + (p.up == 2 ? 1 : 0);
    const uint4 *s0 = reinterpret_cast<const uint4 *>(p.skip.ptr) +
                      pixel_unit(p.skip.fmt, p.skip_pitch, p.skip_ps, p.skip_is, p.skip.planes, n,
                                 oy + 1, ox0 + 1) + (uint32_t)(c_first >> 3) * p.skip_ps;
    const uint4 *s1 = reinterpret_cast<const uint4 *>(p.skip.ptr) +
                      pixel_unit(p.skip.fmt, p.skip_pitch, p.skip_ps, p.skip_is, p.skip.planes, n,
                                 oy + 1, ox1 + 1) + (uint32_t)(c_second >> 3) * p.skip_ps;
    sk[0] = __ldg(s0);
    sk[1] = __ldg(s0 + p.skip_ps);
    if (second) {
      sk[2] = __ldg(s1);
      sk[3] = __ldg(s1 + p.skip_ps);
    }
  }
  tmem_ld_wait();
  const bool fused_q = EPI == EPI_LATENT && p.quant;   // its warp collectives need every lane
  if ((!valid && !fused_q) || (p.debug & 8)) return;

  if (FAST) {
    const __half2 pre2 = __float2half2_rn(pre_s), post2 = __float2half2_rn(post_s);
    if (FAST == 1 && p.pair_store) {
      __half2 ha[8], hb[8];
      fast_half16<false>(p, r0, c_first, pre_s, pre2, post2, sk, ha);
      fast_half16<false>(p, r1, c_second, pre_s, pre2, post2, sk, hb);
      store16_pair(p, ha, hb, c_first, n, oy, ox0);
      return;
    }
    emit16_fast<FAST == 2>(p, r0, c_first, n, oy, ox0, pre_s, pre2, post2, sk);
    if (second)
      emit16_fast<FAST == 2>(p, r1, c_second, n, oy, ox0 + (p.up == 2 ? 1 : 0), pre_s, pre2, post2,
                             sk + 2);
    return;
  }

#pragma unroll 1
  for (int part = 0; part < 2; ++part) {
    if (part == 1 && !second) break;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(part ? r1[i] : r0[i]);
    const int c0 = part ? c_second : c_first;
    const int ox = ox0 + (p.up == 2 ? part : 0);
    if (p.bias) {
      if (c0 + 16 <= p.c_out) {
        const float4 *bp = reinterpret_cast<const float4 *>(p.bias + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bv = __ldg(bp + q);
          v[4 * q] += bv.x;
          v[4 * q + 1] += bv.y;
          v[4 * q + 2] += bv.z;
          v[4 * q + 3] += bv.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.c_out) v[i] += __ldg(p.bias + c0 + i);
      }
    }
    if (pre_s != 1.f) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * pre_s);
    }
    if (p.skip.ptr) {
      const uint32_t off = pixel_unit(p.skip.fmt, p.skip_pitch, p.skip_ps, p.skip_is,
                                      p.skip.planes, n, oy + 1, ox + 1) +
                           (uint32_t)(c0 >> 3) * p.skip_ps;
      const uint4 *sp = reinterpret_cast<const uint4 *>(p.skip.ptr) + off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 sv = __ldg(sp + h * p.skip_ps);
        const __half2 *sh = reinterpret_cast<const __half2 *>(&sv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(sh[k]);
          v[h * 8 + 2 * k] += f.x;
          v[h * 8 + 2 * k + 1] += f.y;
        }
      }
    }
    if (post_s != 1.f) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * post_s);
    }
    if (EPI == EPI_LATENT) {
      if (valid) {
        float *o = reinterpret_cast<float *>(p.out.ptr) +
                   (((size_t)n * p.c_out + c0) * p.out_h + oy) * p.out_w + ox;
        const size_t cs = (size_t)p.out_h * p.out_w;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.c_out) o[i * cs] = v[i];
      }
      if (fused_q) quant_emit16(p, qs, v, c0, n, oy, ox, valid, q_bits);
    } else {
      const uint4 lo = pack8(v), hi = pack8(v + 8);
      const uint32_t off = pixel_unit(p.out.fmt, p.out_pitch, p.out_ps, p.out_is, p.out.planes,
                                      n, oy + 1, ox + 1) + (uint32_t)(c0 >> 3) * p.out_ps;
      uint4 *dst = reinterpret_cast<uint4 *>(p.out.ptr) + off;
      dst[0] = lo;
      dst[p.out_ps] = hi;
      if (p.out.halo == CAE_HALO_REFLECT &&
          (oy == 1 || oy == p.out.H - 2 || ox == 1 || ox == p.out.W - 2))
        store_halo2(p, n, c0 >> 3, oy, ox, lo, hi);
    }
  }
}

// ------------------------------------------------------------------ kernel
template <int EPI, int FAST, int PAIR>
__global__ void __launch_bounds__(EPI == EPI_PROJ ? 128 + 32 * 17 : (FAST == 1 ? 128 + 32 * 16 : kThreads), 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ IgParams p) {
  // PAIR: a cluster of two CTAs works on two tiles at a time with cta_group::2 MMAs (M = 256):
  // each CTA loads its own activation patch and only HALF of every weight stage (N / 2 rows of
  // B), which halves the L2 -> SM weight traffic per MMA -- the bound of the single-CTA form
  // (~2.6 KB per 128x128x16 MMA against ~43 B/clk/SM of L2 bandwidth) -- and the issue count.
  // The leader (rank 0) issues; the peer's idle issuing warps relay its full barriers.
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxSA], a_empty[kMaxSA];
  __shared__ __align__(8) uint64_t b_full[kMaxSB], b_empty[kMaxSB];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t b_peer[kMaxSB];   // PAIR: the peer's weight stages landed
  __shared__ __align__(8) uint64_t u_full, u_free, d2_full;   // EPI_PROJ
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // work units (tile, pass) are walked by clusters; rank r of a pair takes tile 2 * pairtile + r
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_units = (PAIR ? (p.n_tiles + 1) >> 1 : p.n_tiles) * p.n_pass;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + (size_t)p.sa * p.a_stage_bytes;
  // fused quantizer: medians | likelihood table | histogram behind the rings
  QuantSmem qs{nullptr, nullptr, nullptr};
  if (EPI == EPI_LATENT && p.quant && p.q_smem) {
    float *q0 = reinterpret_cast<float *>(smem_b + (size_t)p.sb * p.b_stage_bytes);
    const int nl = p.c_out * p.q_core_len, nh = p.q_hist ? nl : 0;
    for (int i = threadIdx.x; i < p.c_out; i += blockDim.x) q0[i] = p.qt.medians[i];
    for (int i = threadIdx.x; i < nl; i += blockDim.x) {
      const int c = i / p.q_core_len, k = i - c * p.q_core_len;
      q0[p.c_out + i] = p.qt.lut[(size_t)c * p.qt.lut_len + (p.q_core_min - p.qt.lut_min) + k];
    }
    int *h0 = reinterpret_cast<int *>(q0 + p.c_out + nl);
    for (int i = threadIdx.x; i < nh; i += blockDim.x) h0[i] = 0;
    qs.med = q0;
    qs.lut = q0 + p.c_out;
    qs.hist = nh ? h0 : nullptr;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 2);
    }
    for (int i = 0; i < p.sb; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 2);
      mbar_init(&acc_empty[i], (uint32_t)p.epi_warps * (PAIR ? 2u : 1u));
    }
    if (PAIR) {
      for (int i = 0; i < p.sb; ++i) mbar_init(&b_peer[i], 1);
    }
    if (EPI == EPI_PROJ) {
      mbar_init(&u_full, (uint32_t)p.epi_warps * (PAIR ? 2u : 1u));
      mbar_init(&u_free, 1);
      mbar_init(&d2_full, 1);
    }
    fence_barrier_init();
  }
  // EPI_PROJ: the staged output tile and the projection weights behind the rings
  uint8_t *u_stage = smem_b + (size_t)p.sb * p.b_stage_bytes;
  uint8_t *proj_w_s = u_stage + kProjStageBytes;
  if (EPI == EPI_PROJ) {
    // PAIR: rows [16 r, 16 r + 16) of V, from the pair-ordered copy behind the full one
    const uint4 *src = reinterpret_cast<const uint4 *>(
        p.proj_w + (PAIR ? kProjWBytes + rank * (kProjWBytes / 2) : 0));
    uint4 *dst = reinterpret_cast<uint4 *>(proj_w_s);
    for (int i = threadIdx.x; i < kProjWBytes / 16 / (PAIR ? 2 : 1); i += blockDim.x)
      dst[i] = __ldg(src + i);
    fence_proxy_async();       // read by the tensor cores (async proxy) below
  }
  if (warp == 3) tmem_alloc_g<PAIR>(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmA);
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== activation patches: TMA box loads =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int vt = unit0; vt < n_units; vt += unit_step) {
        // (a pair's tile beyond the last one reads outside the tensor: zero fill, full byte count)
        const int tile = PAIR ? (vt / p.n_pass) * 2 + (int)rank : vt / p.n_pass;
        const int n = tile / p.tiles_per_img, rem = tile - n * p.tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * 16 + p.org_y, x0 = tx * 8 * p.mt + p.org_x;
        for (int ch = 0; ch < p.n_chunks; ++ch, ++it) {
          const int s = it % p.sa;
          mbar_wait(&a_empty[s], ((it / p.sa) & 1) ^ 1);
          if (PAIR) {
            // both patches complete on the leader's barrier, which expects the bytes of both
            if (rank == 0) mbar_expect_tx(&a_full[s], 2u * (uint32_t)(p.n_par * p.a_box_bytes));
            const uint32_t bar = mapa_u32(smem_u32(&a_full[s]), 0);
            for (int par = 0; par < p.n_par; ++par)
              tma_load_4d_pair(&tmA, bar,
                               smem_a + (size_t)s * p.a_stage_bytes + (size_t)par * p.par_stride,
                               x0 * 8, y0, ch * (p.ck >> 3), n * p.n_par + par);
            continue;
          }
          if (p.debug & 1) { mbar_arrive(&a_full[s]); continue; }
          mbar_expect_tx(&a_full[s], (uint32_t)(p.n_par * p.a_box_bytes));
          for (int par = 0; par < p.n_par; ++par)
            tma_load_4d(&tmA, &a_full[s],
                        smem_a + (size_t)s * p.a_stage_bytes + (size_t)par * p.par_stride, x0 * 8,
                        y0, ch * (p.ck >> 3), n * p.n_par + par);
        }
      }
    }
  } else if (warp == 1) {
    // ===== packed weights: 1-D bulk copies =====
    if (lane == 0) {
      uint32_t it = 0;
      // PAIR: rank r streams rows [r * N / 2, (r + 1) * N / 2) of B (the second half of wpack)
      const uint8_t *wsrc = p.wpack + (size_t)rank * p.wpack_half_bytes;
      if (p.resident) {
        mbar_expect_tx(&b_full[0], (uint32_t)p.b_stage_bytes);
        for (int k = 0; k < p.n_chunks * p.n_taps; ++k)
          bulk_load_1d(smem_b + (size_t)k * p.b_tap_bytes, wsrc + (size_t)k * p.b_tap_bytes,
                       (uint32_t)p.b_tap_bytes, &b_full[0]);
      }
      for (int vt = unit0; vt < n_units && !p.resident; vt += unit_step) {
        const int pass = vt % p.n_pass;
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          for (int st = 0; st < p.n_stages[pass]; ++st, ++it) {
            const int s = it % p.sb;
            mbar_wait(&b_empty[s], ((it / p.sb) & 1) ^ 1);
            if (p.debug & 2) { mbar_arrive(&b_full[s]); continue; }
            mbar_expect_tx(&b_full[s], (uint32_t)p.b_stage_bytes);
            bulk_load_1d(smem_b + (size_t)s * p.b_stage_bytes,
                         wsrc + (size_t)(ch * p.n_taps + p.pass_tap0[pass] + st * p.tpb) *
                                    p.b_tap_bytes,
                         (uint32_t)p.b_stage_bytes, &b_full[s]);
          }
        }
      }
    }
  }
  if (PAIR && rank != 0 && (warp == 2 || warp == 3)) {
    // ===== peer CTA: its issuing warps are idle; warp 3 relays "weights landed" to the leader
    // (the activation patches complete on the leader's barrier by themselves) =====
    if (warp == 3 && lane == 0) {
      if (p.resident) {
        mbar_wait(&b_full[0], 0);
        mbar_arrive_cluster(mapa_u32(smem_u32(&b_peer[0]), 0));
      }
      uint32_t it = 0;
      for (int vt = unit0; vt < n_units && !p.resident; vt += unit_step) {
        const int pass = vt % p.n_pass;
        for (int k = 0; k < p.n_chunks * p.n_stages[pass]; ++k, ++it) {
          const int s = it % p.sb;
          mbar_wait(&b_full[s], (it / p.sb) & 1);
          mbar_arrive_cluster(mapa_u32(smem_u32(&b_peer[s]), 0));
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===== MMA issue: two issuing warps =====
    // tcgen05.mma is issued by one thread, and with M=128,N<=128 tiles an instruction only
    // covers 64 tensor-pipe cycles, so issue cost bounds the pipe.  Hence: (1) warps 2 and 3
    // each issue the MMAs of the accumulators with (index & 1) == issuer and each commit to
    // the empty / full barriers (which expect two arrivals); (2) the per-chunk work of an
    // issuer is a precomputed list in the kernel parameters (constant bank), walked with
    // warp-uniform control flow so every descriptor lives in uniform registers, and only the
    // MMA / commit instructions themselves are predicated on the elected lane; (3) descriptors
    // are (lo, hi) words: hi (SBO, version) constant, lo = address | LBO advanced by 32-bit
    // adds in 16-byte units.
    const int issuer = warp - 2;
    const bool leader = elect_one();
    const uint32_t a_hi = ((p.sbo_a >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t b_hi = ((p.sbo_b >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t a_lbo = ((p.lbo_a >> 4) & 0x3FFFu) << 16;
    const uint32_t b_lbo = ((p.lbo_b >> 4) & 0x3FFFu) << 16;
    const uint32_t a_kstep = (2 * p.lbo_a) >> 4, b_kstep = (2 * p.lbo_b) >> 4;
    const uint32_t sa_base = smem_base >> 4;
    const uint32_t sb_base = (smem_base + (uint32_t)(p.sa * p.a_stage_bytes)) >> 4;
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
    const uint32_t idesc = p.idesc;
    const int ksteps = p.ck >> 4;
    const uint32_t buf_cols = (uint32_t)(p.mt * p.n_acc * p.N);
    const bool no_mma = (p.debug & 4) != 0;
    uint32_t j = 0;
    int sA = 0, sB = 0;
    uint32_t phA = 0, phB = 0;
    const bool resident = p.resident != 0;
    if (resident) {
      mbar_wait(&b_full[0], 0);
      if (PAIR) mbar_wait_cluster(&b_peer[0], 0);
    }
    for (int vt = unit0; vt < n_units; vt += unit_step, ++j) {
      const int pass = vt % p.n_pass;
      const int n_stages = p.n_stages[pass];
      const int buf = j % p.n_buf;
      if (PAIR) mbar_wait_cluster(&acc_empty[buf], ((j / p.n_buf) & 1) ^ 1);
      else mbar_wait(&acc_empty[buf], ((j / p.n_buf) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_buf = tmem_base + (uint32_t)buf * buf_cols;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        if (PAIR) mbar_wait_cluster(&a_full[sA], phA); else mbar_wait(&a_full[sA], phA);
        const uint32_t a_stage = ((sa_base + (uint32_t)sA * a_stage16) & 0x3FFFu) | a_lbo;
        const uint32_t later = ch > 0 ? 1u : 0u;
        for (int st = 0; st < n_stages; ++st) {
          if (!resident) {
            mbar_wait(&b_full[sB], phB);
            if (PAIR) mbar_wait_cluster(&b_peer[sB], phB);
          }
          tc_fence_after();
          const uint32_t b_stage =
              ((sb_base + (resident ? (uint32_t)ch * p.b_chunk16 : (uint32_t)sB * b_stage16)) & 0x3FFFu) | b_lbo;
          const int i1 = p.item_start[pass][issuer][st + 1];
          for (int i = p.item_start[pass][issuer][st]; i < i1; ++i) {
            const IgItem it = p.items[pass][issuer][i];
            const uint32_t d = d_buf + it.d_off;
            uint32_t a_lo = a_stage + it.a_off16, b_lo = b_stage + it.b_off16;
            uint32_t flag = later | (it.first ^ 1u);
            if (leader && !no_mma) {
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k) {
                umma_f16_lohi_g<PAIR>(d, a_lo, a_hi, b_lo, b_hi, idesc, flag);
                a_lo += a_kstep;
                b_lo += b_kstep;
                flag = 1u;
              }
            }
          }
          if (leader) {
            if (!resident) umma_commit_g<PAIR>(&b_empty[sB]);
            if (st == n_stages - 1) umma_commit_g<PAIR>(&a_empty[sA]);
            if (st == n_stages - 1 && ch == p.n_chunks - 1) umma_commit_g<PAIR>(&acc_full[buf]);
          }
          if (!resident && ++sB == p.sb) { sB = 0; phB ^= 1u; }
        }
        if (++sA == p.sa) { sA = 0; phA ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 4 + p.epi_warps) {
    // ===== epilogue: TMEM -> registers -> HBM (epi_warps / 4 warps per TMEM lane quadrant) =====
    const int quad = warp & 3, half = (warp - 4) >> 2, n_halves = p.epi_warps >> 2;
    const int row = quad * 32 + lane;
    const int ty = row >> 3, txl = row & 7;
    const int acc_per_buf = p.mt * p.n_acc;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    int jobs_per_m;
    if (EPI == EPI_IMAGE) jobs_per_m = 1;
    else if (p.up == 1) jobs_per_m = (p.N + 31) >> 5;
    else jobs_per_m = (p.n_pass == 2 ? 1 : 2) * (p.N >> 4);
    const int n_jobs = p.mt * jobs_per_m;
    const float pre_s = act_slope(p.pre_act), post_s = act_slope(p.post_act);
    float q_bits = 0.f;
    uint32_t j = 0;
    const uint32_t u_full_addr = PAIR ? mapa_u32(smem_u32(&u_full), 0) : smem_u32(&u_full);
    const uint32_t acc_empty_addr0 = PAIR ? mapa_u32(smem_u32(&acc_empty[0]), 0) : 0u;
    const uint32_t acc_empty_addr1 = PAIR ? mapa_u32(smem_u32(&acc_empty[1]), 0) : 0u;
    for (int vt = unit0; vt < n_units; vt += unit_step, ++j) {
      const int pass = vt % p.n_pass;
      const int tile = PAIR ? (vt / p.n_pass) * 2 + (int)rank : vt / p.n_pass;
      const int n = tile / p.tiles_per_img, rem = tile - n * p.tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int buf = j % p.n_buf;
      // rows past the image (and the whole tile past the last one of an odd count) are not stored
      const int y = tile < p.n_tiles ? tyi * 16 + ty : p.dom_h;
      if (FAST == 2 && p.up == 1 && (txl == 0 || txl == 7) && y < p.dom_h && !(p.debug & 64)) {
        // pull this warp's share of the residual tile into L2 while the MMAs still run: 8
        // pixels x 16 B per plane row, so the first and last lane of a row cover its lines
        for (int job = half; job < n_jobs; job += n_halves) {
          const int m = job / jobs_per_m, jj = job - m * jobs_per_m;
          const int x = (txi * p.mt + m) * 8 + txl;
          if (x >= p.dom_w) continue;
          const uint4 *sp = reinterpret_cast<const uint4 *>(p.skip.ptr) +
                            pixel_unit(p.skip.fmt, p.skip_pitch, p.skip_ps, p.skip_is,
                                       p.skip.planes, n, y + 1, x + 1) +
                            (uint32_t)(jj * 4) * p.skip_ps;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + q * p.skip_ps));
        }
      }
      mbar_wait(&acc_full[buf], (j / p.n_buf) & 1);
      tc_fence_after();
      if (EPI == EPI_PROJ) {
        // E1: accumulators -> bias / activation -> fp16 -> the staged tile.  Stage row = px * 128
        // + TMEM lane (the order of the rows of the second GEMM is free), plane stride 4096 bytes:
        // a warp's 16-byte stores are contiguous, so the staging costs no bank conflicts.
        mbar_wait(&u_free, (j & 1u) ^ 1u);            // the previous pass's GEMM has read the stage
        const __half2 pre2 = __float2half2_rn(pre_s), post2 = __float2half2_rn(post_s);
        for (int job = half; job < (kProjK >> 4); job += n_halves) {
          const int c_first = job * 16;
          uint32_t r0[16], r1[16];
          const uint32_t t = lane_base + (uint32_t)(buf * acc_per_buf * p.N + c_first);
          __syncwarp();
          tmem_ld16(t, r0);
          tmem_ld16(t + (uint32_t)p.N, r1);
          tmem_ld_wait();
          __half2 ha[8], hb[8];
          fast_half16<false>(p, r0, c_first, pre_s, pre2, post2, nullptr, ha);
          fast_half16<false>(p, r1, c_first, pre_s, pre2, post2, nullptr, hb);
          uint4 *dst = reinterpret_cast<uint4 *>(u_stage + (size_t)(c_first >> 3) * 4096) + row;
          dst[0] = make_uint4(h2u(ha[0]), h2u(ha[1]), h2u(ha[2]), h2u(ha[3]));
          dst[128] = make_uint4(h2u(hb[0]), h2u(hb[1]), h2u(hb[2]), h2u(hb[3]));
          dst[256] = make_uint4(h2u(ha[4]), h2u(ha[5]), h2u(ha[6]), h2u(ha[7]));
          dst[384] = make_uint4(h2u(hb[4]), h2u(hb[5]), h2u(hb[6]), h2u(hb[7]));
        }
        fence_proxy_async();                           // generic-proxy writes -> UMMA reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(u_full_addr); else mbar_arrive(&u_full);
        }
        // E2: the projected tile (block = px, TMEM lane = input pixel) from the first 64 columns
        // of the drained buffer -> HBM, one 32-byte sector per thread and unit
        mbar_wait(&d2_full, j & 1u);
        tc_fence_after();
        const uint32_t d2 = lane_base + (uint32_t)(buf * acc_per_buf * p.N);
        const int x = txi * 8 + txl;
        for (int unit = half; unit < 4; unit += n_halves) {
          const int blk = unit >> 1, hh = unit & 1;
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(d2 + (uint32_t)(blk * kProjN + hh * 16), r);
          tmem_ld_wait();
          if (y < p.dom_h && x < p.dom_w && !(p.debug & 8)) {
            __half *o = p.proj_out +
                        (((size_t)n * p.out_h + (y * 2 + pass)) * p.out_w + (x * 2 + blk)) * kProjN +
                        hh * 16;
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o),
                         "r"(pack2(__uint_as_float(r[0]), __uint_as_float(r[1]))),
                         "r"(pack2(__uint_as_float(r[2]), __uint_as_float(r[3]))),
                         "r"(pack2(__uint_as_float(r[4]), __uint_as_float(r[5]))),
                         "r"(pack2(__uint_as_float(r[6]), __uint_as_float(r[7]))),
                         "r"(pack2(__uint_as_float(r[8]), __uint_as_float(r[9]))),
                         "r"(pack2(__uint_as_float(r[10]), __uint_as_float(r[11]))),
                         "r"(pack2(__uint_as_float(r[12]), __uint_as_float(r[13]))),
                         "r"(pack2(__uint_as_float(r[14]), __uint_as_float(r[15])))
                         : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
        if (PAIR) mbar_arrive_cluster_relaxed(buf ? acc_empty_addr1 : acc_empty_addr0); else mbar_arrive(&acc_empty[buf]);
      }
        continue;
      }
      for (int job = half; job < n_jobs; job += n_halves) {
        const int m = job / jobs_per_m, jj = job - m * jobs_per_m;
        const int x = (txi * p.mt + m) * 8 + txl;
        const bool valid = y < p.dom_h && x < p.dom_w;
        epilogue_job<EPI, FAST>(p, lane_base, buf * acc_per_buf + m * p.n_acc, n, y, x, jj, valid,
                                pre_s, post_s, pass, qs, q_bits);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster_relaxed(buf ? acc_empty_addr1 : acc_empty_addr0); else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (EPI == EPI_LATENT && p.quant && p.q_rate) {
      for (int o = 16; o > 0; o >>= 1) q_bits += __shfl_xor_sync(0xffffffffu, q_bits, o);
      if (lane == 0 && q_bits != 0.f) atomicAdd(p.q_rate, (double)q_bits);
    }
  } else if (EPI == EPI_PROJ && warp == 4 + p.epi_warps && rank == 0) {
    // ===== projection GEMM: P[2 x 128 pixels][32] = U_stage[256 x 128] * V[128 x 32] =====
    // A warp of its own, so the two main issuers never wait for the epilogue; the result goes
    // into the first 64 columns of the accumulator buffer the epilogue has just drained.
    const bool leader = elect_one();
    const uint32_t us = smem_u32(u_stage), ws = smem_u32(proj_w_s);
    const uint32_t buf_cols = (uint32_t)(p.mt * p.n_acc * p.N);
    uint32_t j = 0;
    constexpr uint32_t kRowsB = kProjN / (PAIR ? 2 : 1);     // rows of V held by this CTA
    for (int vt = unit0; vt < n_units; vt += unit_step, ++j) {
      if (PAIR) mbar_wait_cluster(&u_full, j & 1u); else mbar_wait(&u_full, j & 1u);
      tc_fence_after();
      if (leader) {
        const uint32_t d2 = tmem_base + (j % (uint32_t)p.n_buf) * buf_cols;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
          for (int k = 0; k < kProjK / 16; ++k) {
            const uint64_t da = make_smem_desc(us + blk * 2048 + k * 8192, 4096, 128);
            const uint64_t db = make_smem_desc(ws + k * (2 * kRowsB * 16), kRowsB * 16, 128);
            umma_f16_g<PAIR>(d2 + blk * kProjN, da, db, p.idesc2, k > 0 ? 1u : 0u);
          }
        }
        umma_commit_g<PAIR>(&u_free);
        umma_commit_g<PAIR>(&d2_full);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (EPI == EPI_LATENT && qs.hist) {
    // core bins of this CTA -> the global histogram (same clamping as the direct path)
    const int nh = p.c_out * p.q_core_len, bins = p.qt.hist_bins;
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      const int v = qs.hist[i];
      if (v) {
        const int c = i / p.q_core_len;
        int b = p.q_core_min + (i - c * p.q_core_len) - p.qt.hist_min;
        b = b < 0 ? 0 : (b >= bins ? bins - 1 : b);
        atomicAdd(p.q_hist + (size_t)c * bins + b, v);
      }
    }
  }
  if (warp == 3) tmem_dealloc_g<PAIR>(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------- projection fusion: helpers
// packed[kplane][n][8] fp16 from the image layer's ConvTranspose2d weight (128, c_out, 3, 3)
__global__ void pack_proj_kernel(const float *__restrict__ w, const float *__restrict__ scale,
                                 int c_out, __half *__restrict__ out) {
  // first kProjWBytes: [kplane][n][8]; then the CTA-pair order [rank][kplane][n - 16 rank][8]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kProjK * kProjN) return;
  const int k8 = i & 7, n = (i >> 3) % kProjN, kplane = i / (8 * kProjN);
  const int ci = kplane * 8 + k8;
  float v = 0.f;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int c = n - proj_slot(kh, kw);
      if (c >= 0 && c < c_out && c < 3)
        v = w[(((size_t)ci * c_out + c) * 3 + kh) * 3 + kw] * (scale ? scale[c] : 1.f);
    }
  out[i] = __float2half_rn(v);
  const int half = kProjN / 2, r = n / half;
  out[kProjK * kProjN + r * (kProjK * half) + (kplane * half + (n - r * half)) * 8 + k8] =
      __float2half_rn(v);
}

struct ProjGatherParams {
  const uint4 *proj;     // [n][h][w] records of four 16-byte units
  int n, h, w, c_out;    // h, w: size of U (the projected layer's output)
  const float *bias;
  float pre_s, post_s;
  uint8_t *out;          // [n][2h][2w][c_out] or nullptr
  float *aux;            // [n][c_out][2h][2w] or nullptr
};

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// One thread per pixel (a, b) of U = the 2x2 output pixels (2a + qy, 2b + qx):
//   o00 = R(a,b)[0..]   o01 = R(a,b)[3..] + R(a,b+1)[16..]   o10 = R(a,b)[6..] + R(a+1,b)[24..]
//   o11 = R(a,b)[9..] + R(a,b+1)[19..] + R(a+1,b)[27..] + R(a+1,b+1)[12..]
// HBM bound: 64 bytes read (neighbour units come from L1 / L2), 4 * c_out bytes written per thread.
template <int CO>
__global__ void __launch_bounds__(256) image_from_proj_kernel(const ProjGatherParams q) {
  const int b = blockIdx.x * 32 + (threadIdx.x & 31);
  const int a = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (a >= q.h || b >= q.w) return;
  const uint4 *r = q.proj + (((size_t)n * q.h + a) * q.w + b) * 4;
  const bool right = b + 1 < q.w, down = a + 1 < q.h;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const uint4 u0 = __ldg(r), u1 = __ldg(r + 1);
  const uint4 ur = right ? __ldg(r + 4 + 2) : zero;
  const uint4 ud = down ? __ldg(r + (size_t)q.w * 4 + 3) : zero;
  const uint4 ux = (right && down) ? __ldg(r + (size_t)q.w * 4 + 4 + 1) : zero;
  float f0[8], f1[8], fr[8], fd[8], fx[8];
  unpack8(u0, f0);
  unpack8(u1, f1);
  unpack8(ur, fr);
  unpack8(ud, fd);
  unpack8(ux, fx);
  float own[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    own[i] = f0[i];
    own[8 + i] = f1[i];
  }
  float t[4][CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    t[0][c] = own[c];
    t[1][c] = own[3 + c] + fr[c];
    t[2][c] = own[6 + c] + fd[c];
    t[3][c] = own[9 + c] + fr[3 + c] + fd[3 + c] + fx[4 + c];
  }
#pragma unroll
  for (int ph = 0; ph < 4; ++ph)
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      float u = t[ph][c] + (q.bias ? __ldg(q.bias + c) : 0.f);
      u = fmaxf(u, u * q.pre_s);
      t[ph][c] = fmaxf(u, u * q.post_s);
    }
  const int H2 = 2 * q.h, W2 = 2 * q.w;
  if (q.aux) {
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int c = 0; c < CO; ++c)
        q.aux[(((size_t)n * CO + c) * H2 + (2 * a + (ph >> 1))) * W2 + 2 * b + (ph & 1)] = t[ph][c];
  }
  if (q.out) {
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      // two adjacent pixels = 2 * CO contiguous bytes, 2-byte aligned (2 b * CO is even)
      uint16_t *row = reinterpret_cast<uint16_t *>(
          q.out + (((size_t)n * H2 + (2 * a + py)) * W2 + 2 * b) * CO);
      uint8_t v[2 * CO];
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        v[c] = to_u8_trunc(t[py * 2][c]);
        v[CO + c] = to_u8_trunc(t[py * 2 + 1][c]);
      }
#pragma unroll
      for (int k = 0; k < CO; ++k) row[k] = (uint16_t)(v[2 * k] | ((uint16_t)v[2 * k + 1] << 8));
    }
  }
}

// ------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int pow2_cols(int c) {
  int r = 32;
  while (r < c) r <<= 1;
  return r;
}

}  // namespace

// ------------------------------------------------------------------- C ABI
extern "C" size_t cae_packed_weight_bytes(int kind, int c_in, int c_out, int ck) {
  const int c_in_p = round_up(c_in, 16);
  if (ck <= 0) ck = auto_ck(kind, c_in_p, is_merged(kind, c_out), use_pair(kind, c_out));
  TapDef taps[kMaxTaps];
  const int n_taps = build_taps(kind, is_merged(kind, c_out), taps);
  return (size_t)(c_in_p / ck) * n_taps * mma_n(kind, c_out) * ck * sizeof(__half);
}

extern "C" int cae_pack_weights(int kind, int c_in, int c_out, int ck, const float *w,
                                const float *scale, void *packed, void *stream) {
  CAE_CHECK(kind >= CAE_CONV_S1 && kind <= CAE_CONVT_S2, 2, "cae_pack_weights: bad kind %d", kind);
  CAE_CHECK(w && packed, 2, "cae_pack_weights: null pointer");
  const int c_in_p = round_up(c_in, 16);
  if (ck <= 0) ck = auto_ck(kind, c_in_p, is_merged(kind, c_out), use_pair(kind, c_out));
  CAE_CHECK(ck % 16 == 0 && ck <= 128 && c_in_p % ck == 0, 2,
            "cae_pack_weights: ck=%d does not divide padded c_in=%d", ck, c_in_p);
  PackParams q;
  memset(&q, 0, sizeof(q));
  q.kind = kind;
  q.merged = is_merged(kind, c_out);
  q.c_in = c_in;
  q.c_out = c_out;
  q.ck = ck;
  q.N = mma_n(kind, c_out);
  q.n_taps = build_taps(kind, q.merged, q.taps);
  q.n_chunks = c_in_p / ck;
  q.pair = use_pair(kind, c_out) ? 1 : 0;
  const size_t total = (size_t)q.n_chunks * q.n_taps * q.N * ck;
  const int threads = 256;
  pack_weights_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0,
                        (cudaStream_t)stream>>>(q, w, scale, (__half *)packed, total);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_conv_igemm(const cae_conv_desc *d, void *stream) {
  CAE_CHECK(d, 2, "cae_conv_igemm: null descriptor");
  const int kind = d->kind;
  CAE_CHECK(kind >= CAE_CONV_S1 && kind <= CAE_CONVT_S2, 2, "cae_conv_igemm: bad kind %d", kind);
  CAE_CHECK(d->in.ptr && d->weights, 2, "cae_conv_igemm: null input or weights");
  CAE_CHECK(d->groups <= 1, 2, "cae_conv_igemm: grouped convolutions run on cae_conv_direct");
  CAE_CHECK(d->n > 0 && d->h_in > 0 && d->w_in > 0, 2, "cae_conv_igemm: bad shape");
  const int c_in_p = round_up(d->c_in, 16);
  CAE_CHECK(d->in.planes * 8 == c_in_p, 2,
            "cae_conv_igemm: input has %d planes, need %d (c_in padded to 16)", d->in.planes,
            c_in_p / 8);
  const bool merged = is_merged(kind, d->c_out);
  const int want_fmt = kind == CAE_CONV_S2 ? CAE_FMT_F16_SPLIT : CAE_FMT_F16_PLANAR;
  CAE_CHECK(d->in.fmt == want_fmt, 2, "cae_conv_igemm: input format %d, kind %d needs %d",
            d->in.fmt, kind, want_fmt);
  if (kind == CAE_CONV_S2)
    CAE_CHECK(d->h_in % 2 == 0 && d->w_in % 2 == 0, 2,
              "cae_conv_igemm: stride-2 conv needs even input size, got %dx%d", d->h_in, d->w_in);

  IgParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.N = mma_n(kind, d->c_out);
  CAE_CHECK(p.N <= 256, 2, "cae_conv_igemm: c_out=%d too large", d->c_out);
  const bool pair = use_pair(kind, d->c_out);
  p.ck = d->ck > 0 ? d->ck : auto_ck(kind, c_in_p, merged, pair);
  CAE_CHECK(p.ck % 16 == 0 && p.ck <= 128 && c_in_p % p.ck == 0, 2, "cae_conv_igemm: bad ck=%d",
            p.ck);
  p.n_chunks = c_in_p / p.ck;
  CAE_CHECK(!(pair && d->out.fmt == CAE_FMT_F32_NCHW), 2,
            "cae_conv_igemm: a 128-channel fp32 latent layer is not covered by the CTA-pair form "
            "(unset CAE_IGEMM_PAIR_MMA)");
  TapDef taps[kMaxTaps];
  p.n_taps = build_taps(kind, merged, taps);
  // transposed stride-2 (not merged): four phase accumulators of N columns.  When they do not
  // fit TMEM twice, the tile is done in two passes (output rows 2y, then 2y+1) of two
  // accumulators each, so that the epilogue of one pass overlaps the MMAs of the next
  // (measured with the 256-bit pair stores: 128->128 @64x64 209 -> 199 us, 48->128 43 -> 37 us)
  const bool convt2 = kind == CAE_CONVT_S2 && !merged;
  p.n_pass = 1;
  if (convt2 && (cae_knob(CAE_KNOB_IGEMM_TWO_PASS) || (4 * p.N > 256 && !cae_knob(CAE_KNOB_IGEMM_ONE_PASS))))
    p.n_pass = 2;
  p.n_acc = (kind == CAE_CONVT_S2 && !merged) ? (p.n_pass == 2 ? 2 : 4) : 1;
  p.up = (kind == CAE_CONVT_S2) ? 2 : 1;

  // M domain: output pixels, except transposed stride-2 where it is input pixels
  if (kind == CAE_CONV_S2) {
    p.dom_h = d->h_in / 2;
    p.dom_w = d->w_in / 2;
  } else {
    p.dom_h = d->h_in;
    p.dom_w = d->w_in;
  }
  p.out_h = p.dom_h * p.up;
  p.out_w = p.dom_w * p.up;

  int mt = d->mt > 0 ? d->mt : 2;
  if (d->mt <= 0)
    if (const char *e = cae_knob(CAE_KNOB_IGEMM_MT)) mt = atoi(e) == 1 ? 1 : 2;
  if (p.dom_w <= 8) mt = 1;
  if (p.n_acc * p.N * mt > 512) mt = 1;
  if (p.n_pass == 2 && p.n_acc * p.N * mt > 256) mt = 1;   // keep the accumulators double buffered
  CAE_CHECK(p.n_acc * p.N * mt <= 512, 2, "cae_conv_igemm: accumulators exceed TMEM");
  p.mt = mt;
  p.n_buf = (512 / (p.n_acc * p.N * mt)) >= 2 ? 2 : 1;
  p.tmem_cols = pow2_cols(p.n_buf * p.n_acc * p.N * mt);

  if (kind == CAE_CONV_S1 || kind == CAE_CONVT_S1) {
    p.PH = 18;
    p.PW = 8 * mt + 2;
    p.n_par = 1;
    p.org_y = 0;
    p.org_x = CAE_COL_PAD;
  } else if (kind == CAE_CONV_S2) {
    p.PH = 17;
    p.PW = 8 * mt + 1;
    p.n_par = 4;
    p.org_y = 0;
    p.org_x = 1;             // half index of physical column 2 ox + CAE_COL_PAD, rounded down
  } else {
    p.PH = 17;
    p.PW = 8 * mt + 1;
    p.n_par = 1;
    p.org_y = 1;
    p.org_x = 1 + CAE_COL_PAD;
  }
  const int kplanes = p.ck / 8;
  p.a_box_bytes = kplanes * p.PH * p.PW * 16;
  p.par_stride = round_up(p.a_box_bytes, 128);
  p.a_stage_bytes = round_up(p.par_stride * p.n_par, 128);
  p.lbo_a = (uint32_t)(p.PH * p.PW * 16);
  p.sbo_a = (uint32_t)(p.PW * 16);
  p.lbo_b = (uint32_t)((pair ? p.N / 2 : p.N) * 16);
  p.sbo_b = 128;
  if (cae_knob(CAE_KNOB_IGEMM_SWAP_LBO_SBO)) {  // bring-up knob, see DESIGN.md
    uint32_t t = p.lbo_a; p.lbo_a = p.sbo_a; p.sbo_a = t;
    t = p.lbo_b; p.lbo_b = p.sbo_b; p.sbo_b = t;
  }
  p.idesc = make_idesc_f16(pair ? 256 : 128, p.N);
  for (int t = 0; t < p.n_taps; ++t) {
    p.taps[t].a_off = (uint32_t)(taps[t].par * p.par_stride + (taps[t].dy * p.PW + taps[t].dx) * 16);
    p.taps[t].acc = (uint32_t)taps[t].acc;
  }

  // shared-memory rings.  A weight stage holds `tpb` taps: every stage costs the issuing
  // warps one barrier round trip (~700 cycles), so a stage must carry well over that much
  // tensor work; tpb is the largest divisor of the tap count that leaves room for two stages
  // of each ring.
  int budget = 227 * 1024 - 2048;
  // fused quantizer tables (medians, likelihoods, histogram) behind the rings when they fit
  int q_bytes = 0;
  if (d->quant && d->out.fmt == CAE_FMT_F32_NCHW && !merged && d->quant->tables.lut &&
      !cae_knob(CAE_KNOB_QUANT_NO_SMEM)) {
    // as many symbols around the median as fit 40 KB (likelihoods + histogram); the rest of
    // the table stays in global memory (rare symbols)
    const cae_eb_tables &t = d->quant->tables;
    int core = (40 * 1024 / (4 * d->c_out) - 1) / 2;
    if (core > t.lut_len) core = t.lut_len;
    if (core >= 16) {
      int cmin = -(core / 2);
      if (cmin < t.lut_min) cmin = t.lut_min;
      if (cmin + core > t.lut_min + t.lut_len) cmin = t.lut_min + t.lut_len - core;
      p.q_core_min = cmin;
      p.q_core_len = core;
      q_bytes = round_up(4 * d->c_out * (1 + 2 * core), 128);
    }
  }
  budget -= q_bytes;
  // projection fusion: the staged tile and the projection weights live behind the rings
  const bool proj = d->proj != nullptr;
  if (proj) {
    CAE_CHECK(convt2 && p.n_pass == 2 && p.mt == 1 && p.N == kProjK && d->c_out == kProjK, 2,
              "cae_conv_igemm(proj): needs ConvTranspose2d stride 2 with %d output channels", kProjK);
    CAE_CHECK(d->proj->weights && d->proj->proj, 2, "cae_conv_igemm(proj): null pointer");
    CAE_CHECK(!d->skip.ptr && !d->quant && !d->aux_out, 2,
              "cae_conv_igemm(proj): no skip / quantizer / aux output on a projected layer");
    budget -= kProjStageBytes + kProjWBytes;
  }
  p.b_tap_bytes = (pair ? p.N / 2 : p.N) * p.ck * 2;     // what ONE CTA holds of a tap
  p.wpack_half_bytes = (size_t)p.n_chunks * p.n_taps * p.b_tap_bytes;
  int pass_ntaps[2] = {p.n_taps, 0};
  p.pass_tap0[0] = p.pass_tap0[1] = 0;
  if (p.n_pass == 2) {
    pass_ntaps[0] = 0;
    for (int t = 0; t < p.n_taps; ++t) pass_ntaps[taps[t].acc >> 1]++;   // taps are sorted by row phase
    p.pass_tap0[1] = pass_ntaps[0];
  }
  auto divides = [&](int v) {
    return pass_ntaps[0] % v == 0 && (p.n_pass == 1 || pass_ntaps[1] % v == 0);
  };
  // CTA-pair layers: this CTA's share of ALL weights stays in shared memory when two activation
  // stages still fit beside it; then there is no weight ring at all
  const int resident_bytes = p.n_chunks * p.n_taps * p.b_tap_bytes;
  const bool resident = pair && !cae_knob(CAE_KNOB_IGEMM_NO_RESIDENT) &&
                        resident_bytes + 2 * p.a_stage_bytes <= budget;
  p.resident = resident ? 1 : 0;
  p.b_chunk16 = (uint32_t)((p.n_taps * p.b_tap_bytes) >> 4);
  int tpb = 1;
  for (int cand = p.n_taps; cand >= 1 && !resident; --cand) {
    if (!divides(cand)) continue;
    if (2 * p.a_stage_bytes + 2 * cand * p.b_tap_bytes <= budget && cand * p.b_tap_bytes <= 73728) {
      tpb = cand;
      break;
    }
  }
  if (const char *e = resident ? nullptr : cae_knob(CAE_KNOB_IGEMM_TPB)) {
    const int v = atoi(e);
    if (v >= 1 && divides(v) && 2 * p.a_stage_bytes + 2 * v * p.b_tap_bytes <= budget) tpb = v;
  }
  p.tpb = tpb;
  p.b_stage_bytes = resident ? resident_bytes : tpb * p.b_tap_bytes;
  for (int ps = 0; ps < p.n_pass; ++ps) {
    // resident: one "stage" per chunk holding every tap of the pass, offsets inside the chunk
    if (resident) tpb = pass_ntaps[ps];
    const int b_tap0 = resident ? p.pass_tap0[ps] : 0;
    p.n_stages[ps] = pass_ntaps[ps] / tpb;
    CAE_CHECK(p.n_stages[ps] <= kMaxStages, 2, "cae_conv_igemm: too many weight stages");
    int cnt[2] = {0, 0};
    uint32_t seen = 0;
    for (int st = 0; st < p.n_stages[ps]; ++st) {
      p.item_start[ps][0][st] = cnt[0];
      p.item_start[ps][1][st] = cnt[1];
      for (int tt = 0; tt < tpb; ++tt) {
        const int t = p.pass_tap0[ps] + st * tpb + tt;
        for (int m = 0; m < p.mt; ++m) {
          const int acc = m * p.n_acc + (p.n_pass == 2 ? (taps[t].acc & 1) : taps[t].acc);
          const int who = acc & 1;
          CAE_CHECK(cnt[who] < kMaxItems, 2, "cae_conv_igemm: work list overflow");
          IgItem &it = p.items[ps][who][cnt[who]++];
          it.a_off16 = (uint32_t)((taps[t].par * p.par_stride + (taps[t].dy * p.PW + taps[t].dx) * 16 +
                                   m * 128) >> 4);
          it.b_off16 = (uint32_t)(((b_tap0 + tt) * p.b_tap_bytes) >> 4);
          it.d_off = (uint32_t)(acc * p.N);
          it.first = (seen >> acc) & 1u ? 0u : 1u;
          seen |= 1u << acc;
        }
      }
    }
    p.item_start[ps][0][p.n_stages[ps]] = cnt[0];
    p.item_start[ps][1][p.n_stages[ps]] = cnt[1];
  }
  int sa = 2, sb = resident ? 1 : 2;
  CAE_CHECK(sa * p.a_stage_bytes + sb * p.b_stage_bytes <= budget, 2,
            "cae_conv_igemm: tile does not fit shared memory (A %d B %d)", p.a_stage_bytes,
            p.b_stage_bytes);
  for (;;) {
    bool grew = false;
    const int sb_max = resident ? 1 : kMaxSB;
    if (sb < 3 && sb < sb_max && sa * p.a_stage_bytes + (sb + 1) * p.b_stage_bytes <= budget) { ++sb; grew = true; }
    else if (sa < 3 && (sa + 1) * p.a_stage_bytes + sb * p.b_stage_bytes <= budget) { ++sa; grew = true; }
    else if (sb < 4 && sb < sb_max && sa * p.a_stage_bytes + (sb + 1) * p.b_stage_bytes <= budget) { ++sb; grew = true; }
    else if (sa < 4 && (sa + 1) * p.a_stage_bytes + sb * p.b_stage_bytes <= budget) { ++sa; grew = true; }
    else if (sa < kMaxSA && (sa + 1) * p.a_stage_bytes + sb * p.b_stage_bytes <= budget) { ++sa; grew = true; }
    else if (sb < sb_max && sa * p.a_stage_bytes + (sb + 1) * p.b_stage_bytes <= budget) { ++sb; grew = true; }
    if (!grew) break;
  }
  p.sa = sa;
  p.sb = sb;
  const int smem_bytes = sa * p.a_stage_bytes + sb * p.b_stage_bytes + 1024 + q_bytes +
                         (proj ? kProjStageBytes + kProjWBytes : 0);
  p.q_smem = q_bytes > 0;
  p.tiles_x = (p.dom_w + 8 * mt - 1) / (8 * mt);
  const int tiles_y = (p.dom_h + 15) / 16;
  p.tiles_per_img = p.tiles_x * tiles_y;
  p.n_tiles = p.tiles_per_img * d->n;
  if (cae_knob(CAE_KNOB_IGEMM_VERBOSE))
    fprintf(stderr, "cae_conv_igemm: kind %d %d->%d @%dx%d N=%d ck=%d mt=%d n_pass=%d n_acc=%d n_buf=%d "
            "tpb=%d a_stage=%d x%d b_stage=%d x%d smem=%d tiles=%d pair=%d resident=%d\n", kind, d->c_in, d->c_out, d->h_in,
            d->w_in, p.N, p.ck, p.mt, p.n_pass, p.n_acc, p.n_buf, p.tpb, p.a_stage_bytes, sa,
            p.b_stage_bytes, sb, smem_bytes, p.n_tiles, (int)pair, p.resident);

  p.wpack = (const uint8_t *)d->weights;

  // epilogue
  p.c_out = d->c_out;
  p.pre_act = d->pre_act;
  p.post_act = d->post_act;
  p.bias = d->bias;
  p.aux = (float *)d->aux_out;
  int epi;
  if (proj) {
    epi = EPI_PROJ;
    p.proj_w = (const uint8_t *)d->proj->weights;
    p.proj_out = (__half *)d->proj->proj;
    p.idesc2 = make_idesc_f16(pair ? 256 : 128, kProjN);
  } else if (merged) {
    epi = EPI_IMAGE;
    CAE_CHECK(d->out.fmt == CAE_FMT_U8_HWC || d->out.fmt == CAE_FMT_NONE, 2,
              "cae_conv_igemm: final layer writes U8_HWC (and/or aux fp32)");
    CAE_CHECK(d->out.ptr || d->aux_out, 2, "cae_conv_igemm: no output");
    p.out.ptr = d->out.fmt == CAE_FMT_U8_HWC ? d->out.ptr : nullptr;
  } else if (d->out.fmt == CAE_FMT_F32_NCHW) {
    epi = EPI_LATENT;
    CAE_CHECK(p.up == 1 && d->out.ptr, 2, "cae_conv_igemm: fp32 NCHW output needs up==1");
    p.out.ptr = d->out.ptr;
    if (d->quant) {
      const cae_quant_fuse *q = d->quant;
      CAE_CHECK(!d->skip.ptr, 2, "cae_conv_igemm: fused quantizer on a residual layer");
      if (int rc = eb_check_tables(&q->tables, "cae_conv_igemm(quant)")) return rc;
      CAE_CHECK(q->tables.hist_bins >= 0 && q->tables.hist_bins <= 8192, 2,
                "cae_conv_igemm(quant): hist_bins out of range");
      p.quant = 1;
      p.qt = q->tables;
      p.q_yq = q->y_q;
      p.q_sym = q->symbols;
      p.q_hist = q->hist;
      p.q_rate = q->rate_bits;
      p.q_status = q->status;
      if (cae_knob(CAE_KNOB_QUANT_NO_HIST)) p.q_hist = nullptr;     // bring-up timing knobs
      if (cae_knob(CAE_KNOB_QUANT_NO_RATE)) p.q_rate = nullptr;
      if (cae_knob(CAE_KNOB_QUANT_NO_YQ)) p.q_yq = nullptr;
      if (q->y_q_planar.fmt != CAE_FMT_NONE && q->y_q_planar.ptr) {
        CAE_CHECK(q->y_q_planar.fmt == CAE_FMT_F16_PLANAR && q->y_q_planar.planes * 8 == p.N, 2,
                  "cae_conv_igemm(quant): planar y_q needs %d planes", p.N / 8);
        p.q_pl = ActView{q->y_q_planar.ptr, q->y_q_planar.fmt, q->y_q_planar.planes,
                         q->y_q_planar.halo, p.out_h, p.out_w};
      }
    }
  } else {
    epi = EPI_ACT;
    CAE_CHECK(d->out.ptr && (d->out.fmt == CAE_FMT_F16_PLANAR || d->out.fmt == CAE_FMT_F16_SPLIT),
              2, "cae_conv_igemm: bad output format %d", d->out.fmt);
    CAE_CHECK(d->out.planes * 8 == p.N, 2, "cae_conv_igemm: output has %d planes, need %d",
              d->out.planes, p.N / 8);
    if (d->out.fmt == CAE_FMT_F16_SPLIT)
      CAE_CHECK(p.out_h % 2 == 0 && p.out_w % 2 == 0, 2, "cae_conv_igemm: split output needs even size");
    p.out.ptr = d->out.ptr;
  }
  p.out.fmt = proj ? CAE_FMT_NONE : d->out.fmt;
  p.out.planes = d->out.planes;
  p.out.halo = d->out.halo;
  p.out.H = p.out_h;
  p.out.W = p.out_w;
  auto strides = [](int fmt, int planes, int H, int W, uint32_t &pitch, uint32_t &ps,
                    uint32_t &is) {
    if (fmt == CAE_FMT_F16_SPLIT) {
      pitch = (uint32_t)(cae_row_units(W) / 2);
      ps = pitch * (uint32_t)((H + 2) / 2);
      is = 4u * (uint32_t)planes * ps;
    } else {
      pitch = (uint32_t)cae_row_units(W);
      ps = pitch * (uint32_t)(H + 2);
      is = (uint32_t)planes * ps;
    }
  };
  if (epi == EPI_ACT) {
    CAE_CHECK(!d->aux_out, 2, "cae_conv_igemm: aux_out is only available on the final image layer");
    CAE_CHECK((double)d->n * d->out.planes * (p.out_h + 2) * cae_row_units(p.out_w) < 2147483648.0, 2,
              "cae_conv_igemm: output tensor too large for 32-bit unit offsets; split the batch");
    strides(p.out.fmt, p.out.planes, p.out_h, p.out_w, p.out_pitch, p.out_ps, p.out_is);
  }
  if (p.q_pl.ptr) strides(p.q_pl.fmt, p.q_pl.planes, p.out_h, p.out_w, p.q_pitch, p.q_ps, p.q_is);
  if (d->skip.fmt != CAE_FMT_NONE && d->skip.ptr) {
    CAE_CHECK(epi != EPI_IMAGE, 2, "cae_conv_igemm: skip unsupported on the final layer");
    CAE_CHECK((d->skip.fmt == CAE_FMT_F16_PLANAR || d->skip.fmt == CAE_FMT_F16_SPLIT) &&
                  d->skip.planes * 8 == p.N,
              2, "cae_conv_igemm: skip must be planar fp16 with %d planes", p.N / 8);
    p.skip.ptr = d->skip.ptr;
    p.skip.fmt = d->skip.fmt;
    p.skip.planes = d->skip.planes;
    p.skip.H = p.out_h;
    p.skip.W = p.out_w;
    strides(p.skip.fmt, p.skip.planes, p.out_h, p.out_w, p.skip_pitch, p.skip_ps, p.skip_is);
  }

  // tensor map over the input
  EncodeTiledFn encode = get_encode_fn();
  CAE_CHECK(encode, 3, "cae_conv_igemm: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tm;
  const int Hp = d->h_in + 2, Wp = cae_row_units(d->w_in);
  cuuint64_t gdim[4], gstr[3];
  if (kind == CAE_CONV_S2) {
    const int Hh = Hp / 2, Wh = Wp / 2;
    gdim[0] = (cuuint64_t)Wh * 8; gdim[1] = Hh; gdim[2] = d->in.planes; gdim[3] = (cuuint64_t)d->n * 4;
    gstr[0] = (cuuint64_t)Wh * 16; gstr[1] = gstr[0] * Hh; gstr[2] = gstr[1] * d->in.planes;
  } else {
    gdim[0] = (cuuint64_t)Wp * 8; gdim[1] = Hp; gdim[2] = d->in.planes; gdim[3] = d->n;
    gstr[0] = (cuuint64_t)Wp * 16; gstr[1] = gstr[0] * Hp; gstr[2] = gstr[1] * d->in.planes;
  }
  cuuint32_t box[4] = {(cuuint32_t)(p.PW * 8), (cuuint32_t)p.PH, (cuuint32_t)kplanes, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->in.ptr, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CAE_CHECK(cr == CUDA_SUCCESS, 3, "cae_conv_igemm: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const int sm_count = cae_sm_count();
  int grid = d->grid > 0 ? d->grid : sm_count;
  if (pair) {
    const int units = ((p.n_tiles + 1) / 2) * p.n_pass;
    grid &= ~1;
    if (grid > 2 * units) grid = 2 * units;
    CAE_CHECK(grid >= 2, 2, "cae_conv_igemm: the CTA-pair form needs at least two CTAs");
  } else if (grid > p.n_tiles * p.n_pass) grid = p.n_tiles * p.n_pass;

  p.debug = 0;
  if (const char *e = cae_knob(CAE_KNOB_IGEMM_DEBUG)) p.debug = atoi(e);
  p.pair_store = epi == EPI_ACT && p.up == 2 && p.out.fmt == CAE_FMT_F16_PLANAR &&
                 p.out.halo != CAE_HALO_REFLECT && !(p.debug & 16) && !cae_knob(CAE_KNOB_IGEMM_NO_PAIR_STORE);
  p.epi_warps = 8;
  if (const char *e = cae_knob(CAE_KNOB_IGEMM_EPI_WARPS)) {
    const int v = atoi(e);
    if (v == 4 || v == 8 || v == 12 || v == 16) p.epi_warps = v;
  }
  // epilogue variant: 1 = packed half2 (16 epilogue warps), 2 = the same after an fp32 residual
  // add (12 warps: more registers), 0 = plain fp32 (bring-up / latent / image layers)
  int fast = 0;
  if (epi == EPI_ACT && !cae_knob(CAE_KNOB_IGEMM_NO_FAST_EPILOGUE)) fast = p.skip.ptr ? 2 : 1;
  if (!cae_knob(CAE_KNOB_IGEMM_EPI_WARPS)) p.epi_warps = fast == 1 ? 16 : ((fast == 2 || p.quant) ? 12 : 8);
  if (fast != 1 && p.epi_warps > kMaxEpiWarps) p.epi_warps = kMaxEpiWarps;
  if (proj) p.epi_warps = 16;
  const int threads = 128 + 32 * p.epi_warps + (proj ? 32 : 0);    // + the projection-GEMM warp
  typedef void (*KernFn)(const CUtensorMap, const IgParams);
  KernFn kern;
  if (pair) {
    kern = epi == EPI_PROJ ? (KernFn)igemm_conv_kernel<EPI_PROJ, 1, 1>
           : (fast == 1 ? (KernFn)igemm_conv_kernel<EPI_ACT, 1, 1>
                        : (fast == 2 ? (KernFn)igemm_conv_kernel<EPI_ACT, 2, 1>
                                     : (KernFn)igemm_conv_kernel<EPI_ACT, 0, 1>));
    CAE_CHECK(epi == EPI_PROJ || epi == EPI_ACT, 2, "cae_conv_igemm: internal (pair form of epilogue %d)", epi);
  } else {
    kern = epi == EPI_PROJ ? (KernFn)igemm_conv_kernel<EPI_PROJ, 1, 0>
           : epi == EPI_ACT ? (fast == 1 ? (KernFn)igemm_conv_kernel<EPI_ACT, 1, 0>
                                         : (fast == 2 ? (KernFn)igemm_conv_kernel<EPI_ACT, 2, 0>
                                                      : (KernFn)igemm_conv_kernel<EPI_ACT, 0, 0>))
                            : (epi == EPI_LATENT ? (KernFn)igemm_conv_kernel<EPI_LATENT, 0, 0>
                                                 : (KernFn)igemm_conv_kernel<EPI_IMAGE, 0, 0>);
  }
  static_assert(sizeof(IgParams) + sizeof(CUtensorMap) <= 4096, "kernel parameters exceed 4 KB");
  CAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CAE_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, p));
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------- projection fusion
extern "C" size_t cae_proj_weight_bytes(void) { return (size_t)2 * kProjWBytes; }

extern "C" size_t cae_proj_bytes(int n, int h, int w) {
  return (size_t)n * h * w * kProjN * sizeof(__half);
}

extern "C" int cae_pack_proj_weights(int c_in, int c_out, const float *w, const float *scale,
                                     void *packed, void *stream) {
  CAE_CHECK(c_in == kProjK && c_out >= 1 && c_out <= 3, 2,
            "cae_pack_proj_weights: needs c_in = %d and c_out <= 3 (got %d -> %d)", kProjK, c_in, c_out);
  CAE_CHECK(w && packed, 2, "cae_pack_proj_weights: null pointer");
  pack_proj_kernel<<<(kProjK * kProjN + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w, scale, c_out, (__half *)packed);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_image_from_proj(const void *proj, int n, int h, int w, int c_out,
                                   const float *bias, int pre_act, int post_act, void *out_u8,
                                   float *aux, void *stream) {
  CAE_CHECK(proj && (out_u8 || aux), 2, "cae_image_from_proj: null pointer");
  CAE_CHECK(n > 0 && h > 0 && w > 0 && c_out >= 1 && c_out <= 3, 2, "cae_image_from_proj: bad shape");
  CAE_CHECK(n <= 65535 && (h + 7) / 8 <= 65535, 2, "cae_image_from_proj: batch too large for one grid");
  ProjGatherParams q;
  q.proj = (const uint4 *)proj;
  q.n = n; q.h = h; q.w = w; q.c_out = c_out;
  q.bias = bias;
  q.pre_s = pre_act == CAE_ACT_LEAKY_RELU ? 0.01f : (pre_act == CAE_ACT_RELU ? 0.f : 1.f);
  q.post_s = post_act == CAE_ACT_LEAKY_RELU ? 0.01f : (post_act == CAE_ACT_RELU ? 0.f : 1.f);
  q.out = (uint8_t *)out_u8;
  q.aux = aux;
  const dim3 grid((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8), (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  if (c_out == 1) image_from_proj_kernel<1><<<grid, 256, 0, st>>>(q);
  else if (c_out == 2) image_from_proj_kernel<2><<<grid, 256, 0, st>>>(q);
  else image_from_proj_kernel<3><<<grid, 256, 0, st>>>(q);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
