// Weight gradient of a 3x3 (transposed) convolution on the sm_100a tensor cores.
//
// Replaces the wgrad half of `loss.backward()` for nn.Conv2d / nn.ConvTranspose2d inside the
// reference's training step (src/train_cae_ms.py:209-219 through the units of
// src/models/tasks/_autoencoders.py:53-304).
//
//   dW[m][n][tap] = sum over images and pixels of  P[pixel][m] * Q[pixel * stride + tap][n]
//
// P is the tensor indexed without a shift (the gradient of the layer's output for Conv2d and
// ConvTranspose2d stride 1, the layer's INPUT for ConvTranspose2d stride 2), Q the one read
// through the 3x3 window (the layer's input, resp. the output gradient in the parity-split
// layout) -- exactly the activation patch the forward kernel reads, so the TMA box, the tap
// offsets and the split layout are shared with igemm_conv.cu.
//
// GEMM view: M = channels of P (<= 128, TMEM lanes), N = a chunk of <= 48 channels of Q,
// K = pixels.  In the planar layout eight channels of a pixel are one 16-byte unit and eight
// consecutive pixels 128 contiguous bytes: that IS the MN-major (transposed) core matrix of a
// UMMA operand, so both operands are read in place (instruction descriptor a_major = b_major =
// MN, LBO = pitch between rows of eight pixels, SBO = plane stride).  A CTA owns one channel
// chunk of Q and a strided share of the 16x8-pixel tiles, accumulates all nine taps (9 x 48
// TMEM columns) over its whole share without ever draining, and stores its partial result
// ([slot][chunk][tap][m][48] fp32, coalesced) at the end; wgrad_reduce_kernel sums the slots,
// applies the loss scale and accumulates into dW in torch layout (one writer per element: 8 M
// contended atomics per layer were 2-3 ms, the partials are ~30 MB of streaming traffic).
// Persistent, one CTA per SM; warp 0 = TMA, warp 1 = MMA issue, warps 2-5 = the final drain.
#include <string.h>

#include "cae_common.cuh"

namespace {

constexpr int kWgNc = 48;          // channels of Q per CTA (6 planes): 9 taps x 48 columns <= 512
constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 4;

struct WgParams {
  int n_img, tiles_x, tiles_per_img, n_tiles;
  int n_chunks;                    // channel chunks of Q
  int planes_p;                    // planes of P (channels padded to 16) actually loaded
  int PH, PW, n_par, par_stride;   // Q patch geometry (as the forward kernel's A stage)
  int q_org_y, q_org_x;
  int p_off;                       // P is embedded at (p_off, p_off) of a larger zero-ringed buffer
  int p_plane0, m0;                // first plane / channel of P handled by this launch (M chunks of 128)
  int p_bytes, q_box_bytes, stage_bytes, stages;
  uint32_t tap_off[9];             // byte offset of a tap's view inside the Q part of a stage
  uint32_t idesc;
  // output
  int kind, c_in, c_out, c_m, c_n; // c_m / c_n: real channels of P / Q
  float *dw;
  const float *scale;              // device scalar multiplied into the result, or nullptr
  float *partial;                  // [n_slots][n_chunks][9][128][kWgNc]
  int n_slots;
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return make_smem_desc(saddr, lbo, sbo);   // same fields; the major-ness lives in the idesc
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
             const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kWgMaxStages], empty[kWgMaxStages], done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int chunk = (int)blockIdx.x % p.n_chunks;
  const int slot = (int)blockIdx.x / p.n_chunks;
  const int n_slots = (int)gridDim.x / p.n_chunks;
  const bool active = slot < n_slots;          // (CTAs beyond a whole number of chunk groups idle)
  int my_tiles = 0;
  if (active && slot < p.n_tiles) my_tiles = (p.n_tiles - slot + n_slots - 1) / n_slots;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmP);
    prefetch_tensormap(&tmQ);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = slot + i * n_slots;
        const int n = tile / p.tiles_per_img, rem = tile - n * p.tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int s = i % p.stages;
        mbar_wait(&empty[s], ((i / p.stages) & 1) ^ 1);
        uint8_t *st = smem + (size_t)s * p.stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)(p.p_bytes + p.n_par * p.q_box_bytes));
        // P: the 16 x 8 pixel tile itself (interior pixel (y, x) = row y + 1, column x + 4)
        tma_load_4d(&tmP, &full[s], st, (tx * 8 + 1 + CAE_COL_PAD + p.p_off) * 8,
                    ty * 16 + 1 + p.p_off, p.p_plane0, n);
        // Q: the patch the forward kernel reads for this tile, channels of this CTA's chunk
        for (int par = 0; par < p.n_par; ++par)
          tma_load_4d(&tmQ, &full[s], st + p.p_bytes + (size_t)par * p.par_stride,
                      (tx * 8 + p.q_org_x) * 8, ty * 16 + p.q_org_y, chunk * (kWgNc / 8),
                      n * p.n_par + par);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i % p.stages;
      mbar_wait(&full[s], (i / p.stages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t pa = smem_base + (uint32_t)(s * p.stage_bytes);
        const uint32_t qa = pa + (uint32_t)p.p_bytes;
        const uint32_t q_lbo = (uint32_t)(p.PW * 16), q_sbo = (uint32_t)(p.PH * p.PW * 16);
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks) {            // K = 16 pixels = two rows of the tile
          const uint64_t da = make_desc_mn(pa + (uint32_t)(ks * 256), 128, 2048);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint64_t db = make_desc_mn(qa + p.tap_off[t] + (uint32_t)(ks * 2) * q_lbo, q_lbo, q_sbo);
            umma_f16(tmem_base + (uint32_t)(t * kWgNc), da, db, p.idesc, (i | ks) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (i == my_tiles - 1) umma_commit(&done);
      }
      __syncwarp();
    }
  } else if (active) {
    // ===== drain: lane = channel m of P, columns = (tap, channel of Q) =====
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(&done, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    float *dst = p.partial + ((size_t)(slot * p.n_chunks + chunk) * 9 * 128 + m) * kWgNc;
    for (int t = 0; t < 9; ++t) {
      for (int c0 = 0; c0 < kWgNc; c0 += 16) {
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(lane_base + (uint32_t)(t * kWgNc + c0), r);
        tmem_ld_wait();
        float4 *o = reinterpret_cast<float4 *>(dst + (size_t)t * 128 * kWgNc + c0);
        if (my_tiles == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 16; ++q4) r[q4] = 0u;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_float4(__uint_as_float(r[4 * q4]), __uint_as_float(r[4 * q4 + 1]),
                              __uint_as_float(r[4 * q4 + 2]), __uint_as_float(r[4 * q4 + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW[torch index of (m, n, tap)] += scale * sum over slots of partial[slot][chunk][tap][m][n']
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over (tap, m, n), n fastest
  const int total = 9 * p.c_m * p.c_n;
  if (i >= total) return;
  const int n = i % p.c_n, m = (i / p.c_n) % p.c_m, t = i / (p.c_n * p.c_m);
  const int chunk = n / kWgNc, nl = n - chunk * kWgNc;
  const float *src = p.partial + ((size_t)chunk * 9 * 128 + (size_t)t * 128 + m) * kWgNc + nl;
  const size_t slot_stride = (size_t)p.n_chunks * 9 * 128 * kWgNc;
  float acc = 0.f;
  for (int s = 0; s < p.n_slots; ++s) acc += __ldg(src + (size_t)s * slot_stride);
  const int dy = t / 3, dx = t - dy * 3;
  const int mg = m + p.m0;                                          // channel of P in the whole layer
  size_t idx;
  if (p.kind == CAE_CONV_S1 || p.kind == CAE_CONV_S2)
    idx = ((size_t)mg * p.c_in + n) * 9 + t;                        // P = dz (c_out), Q = x (c_in)
  else if (p.kind == CAE_CONVT_S1)
    idx = ((size_t)n * p.c_out + mg) * 9 + (2 - dy) * 3 + (2 - dx); // flipped correlation
  else
    idx = ((size_t)mg * p.c_out + n) * 9 + t;                       // P = x (c_in), Q = dz (c_out)
  p.dw[idx] += acc * (p.scale ? __ldg(p.scale) : 1.f);
}

// ---------------------------------------------------------------- cae_act_grad
// One 16-byte unit (8 channels of a pixel) per thread: gather the incoming gradient (with the
// reflect fold), multiply by the activation's derivative taken from the saved forward output,
// accumulate the bias gradient, scale, and write in the layout the next kernel wants.
struct AgView {
  void *ptr;
  int fmt, planes;
  int H, W;          // logical size of the buffer
  int oy, ox;        // where pixel (0, 0) of the layer sits in it
};

struct AgParams {
  AgView g, out, dz;
  AgView g2, skip, gsum;   // residual add: extra incoming gradient, the tensor added, where g_sum goes
  int n, h, w, c;
  int fold, fold_shift;
  float slope;       // activation: d/dv max(v, v * slope); 1 = none
  float post_slope;  // activation after the residual add; 1 = none
  const float *scale;
  float *db;
};

__device__ __forceinline__ size_t ag_unit(const AgView &v, int n, int plane, int y, int x) {
  // 16-byte unit of plane `plane` of logical pixel (y, x) of the BUFFER (planar / split)
  const int Y = y + 1, Xc = x + 1 + CAE_COL_PAD;
  if (v.fmt == CAE_FMT_F16_PLANAR)
    return (((size_t)n * v.planes + plane) * (v.H + 2) + Y) * cae_row_units(v.W) + Xc;
  const int Hh = (v.H + 2) >> 1, Wh = cae_row_units(v.W) >> 1;
  const int par = ((Y & 1) << 1) | (Xc & 1);
  return ((((size_t)n * 4 + par) * v.planes + plane) * Hh + (Y >> 1)) * Wh + (Xc >> 1);
}

__device__ __forceinline__ void ag_load8(const AgView &v, int n, int plane, int y, int x, int c,
                                         float (&f)[8]) {
  if (v.fmt == CAE_FMT_F32_NCHW) {
    const float *src = reinterpret_cast<const float *>(v.ptr);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = plane * 8 + k;
      f[k] = ch < c ? __ldg(src + (((size_t)n * c + ch) * v.H + y) * v.W + x) : 0.f;
    }
    return;
  }
  const uint4 u = __ldg(reinterpret_cast<const uint4 *>(v.ptr) + ag_unit(v, n, plane, y, x));
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

__device__ __forceinline__ void ag_store8(const AgView &v, int n, int plane, int y, int x, int c,
                                          const float (&g)[8], float sc) {
  if (v.fmt == CAE_FMT_F32_NCHW) {
    float *dst = reinterpret_cast<float *>(v.ptr);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = plane * 8 + k;
      if (ch < c) dst[(((size_t)n * c + ch) * v.H + y + v.oy) * v.W + x + v.ox] = g[k] * sc;
    }
  } else if (plane < v.planes) {
    uint4 u;
    __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = plane * 8 + 2 * k < c ? g[2 * k] * sc : 0.f;
      const float b = plane * 8 + 2 * k + 1 < c ? g[2 * k + 1] * sc : 0.f;
      h[k] = __floats2half2_rn(a, b);
    }
    reinterpret_cast<uint4 *>(v.ptr)[ag_unit(v, n, plane, y + v.oy, x + v.ox)] = u;
  }
}

__global__ void __launch_bounds__(256) act_grad_kernel(const AgParams q) {
  // grid (w / 32, h / 8, n * planes), block 32 x 8: a warp = 32 consecutive pixels of one row
  // and plane (coalesced 16-byte units, no index divisions, one shuffle reduction per channel)
  const int planes = (q.c + 7) >> 3;
  const int x = (int)(blockIdx.x * 32 + (threadIdx.x & 31));
  const int y = (int)(blockIdx.y * 8 + (threadIdx.x >> 5));
  const int plane = (int)(blockIdx.z % (unsigned)planes), n = (int)(blockIdx.z / (unsigned)planes);
  const bool inside = x < q.w && y < q.h;
  __shared__ float s_db[8];
  if (q.db) {
    if (threadIdx.x < 8) s_db[threadIdx.x] = 0.f;
    __syncthreads();
  }
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (inside) {
    if (!q.fold) {
      ag_load8(q.g, n, plane, y + q.g.oy, x + q.g.ox, q.c, g);
    } else {
      // padded index P of the reflect-padded input: P = y + 1, plus the mirrored ring (P = 0
      // mirrors pixel 1, P = h + 1 mirrors pixel h - 2); buffer index = P + fold_shift
      const int fs = q.fold_shift;
      const int y2 = y == 1 ? fs : (y == q.h - 2 ? q.h + 1 + fs : -1);
      const int x2 = x == 1 ? fs : (x == q.w - 2 ? q.w + 1 + fs : -1);
      const bool ry = y2 >= 0 && y2 < q.g.H, rx = x2 >= 0 && x2 < q.g.W;
      ag_load8(q.g, n, plane, y + 1 + fs, x + 1 + fs, q.c, g);
      if (ry | rx) {               // border pixels only
        float t[8];
        if (ry) {
          ag_load8(q.g, n, plane, y2, x + 1 + fs, q.c, t);
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] += t[k];
        }
        if (rx) {
          ag_load8(q.g, n, plane, y + 1 + fs, x2, q.c, t);
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] += t[k];
        }
        if (ry & rx) {
          ag_load8(q.g, n, plane, y2, x2, q.c, t);
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] += t[k];
        }
      }
    }
    if (q.g2.ptr) {             // gradient arriving through a residual connection further up
      float t[8];
      ag_load8(q.g2, n, plane, y + q.g2.oy, x + q.g2.ox, q.c, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
    }
    if (!q.skip.ptr) {
      if (q.out.ptr && q.slope != 1.f) {
        float o[8];
        ag_load8(q.out, n, plane, y + q.out.oy, x + q.out.ox, q.c, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = o[k] > 0.f ? g[k] : g[k] * q.slope;
      }
    } else {
      // out = post(pre(z) + skip): g_sum = g * post'(out) goes on to the skip source as it is;
      // pre'(z) needs the sign of the branch = post^-1(out) - skip
      float o[8], sk[8];
      ag_load8(q.out, n, plane, y + q.out.oy, x + q.out.ox, q.c, o);
      ag_load8(q.skip, n, plane, y + q.skip.oy, x + q.skip.ox, q.c, sk);
      const float inv_post = q.post_slope != 0.f ? 1.f / q.post_slope : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (q.post_slope != 1.f && !(o[k] > 0.f)) g[k] *= q.post_slope;
      }
      if (q.gsum.ptr) ag_store8(q.gsum, n, plane, y, x, q.c, g, q.scale ? __ldg(q.scale) : 1.f);
      if (q.slope != 1.f) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float sum = o[k] > 0.f ? o[k] : o[k] * inv_post;
          if (!(sum - sk[k] > 0.f)) g[k] *= q.slope;
        }
      }
    }
  }
  if (q.db) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = g[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&s_db[k], v);
    }
  }
  if (inside) ag_store8(q.dz, n, plane, y, x, q.c, g, q.scale ? __ldg(q.scale) : 1.f);
  if (q.db) {
    __syncthreads();
    // (the sums are taken before scaling)
    if (threadIdx.x < 8 && plane * 8 + (int)threadIdx.x < q.c && s_db[threadIdx.x] != 0.f)
      atomicAdd(q.db + plane * 8 + threadIdx.x, s_db[threadIdx.x]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn wg_encode_fn() {
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

inline int wg_round_up(int a, int b) { return (a + b - 1) / b * b; }

// tensor map over a planar (or split) fp16 tensor, box = [pw pixels][ph rows][planes][1]
int wg_tensor_map(CUtensorMap *tm, const cae_tensor &t, int n, int H, int W, bool split, int pw,
                  int ph, int planes_box) {
  EncodeTiledFn encode = wg_encode_fn();
  CAE_CHECK(encode, 3, "cae_conv_wgrad: cuTensorMapEncodeTiled unavailable");
  const int Hp = H + 2, Wp = cae_row_units(W);
  cuuint64_t gdim[4], gstr[3];
  if (split) {
    const int Hh = Hp / 2, Wh = Wp / 2;
    gdim[0] = (cuuint64_t)Wh * 8; gdim[1] = Hh; gdim[2] = t.planes; gdim[3] = (cuuint64_t)n * 4;
    gstr[0] = (cuuint64_t)Wh * 16; gstr[1] = gstr[0] * Hh; gstr[2] = gstr[1] * t.planes;
  } else {
    gdim[0] = (cuuint64_t)Wp * 8; gdim[1] = Hp; gdim[2] = t.planes; gdim[3] = n;
    gstr[0] = (cuuint64_t)Wp * 16; gstr[1] = gstr[0] * Hp; gstr[2] = gstr[1] * t.planes;
  }
  cuuint32_t box[4] = {(cuuint32_t)(pw * 8), (cuuint32_t)ph, (cuuint32_t)planes_box, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t.ptr, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CAE_CHECK(cr == CUDA_SUCCESS, 3, "cae_conv_wgrad: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  return 0;
}

}  // namespace

extern "C" size_t cae_conv_wgrad_workspace_bytes(void) {
  return (size_t)cae_sm_count() * 9 * 128 * kWgNc * sizeof(float);
}

extern "C" int cae_conv_wgrad(int kind, int n, int h_in, int w_in, int c_in, int c_out,
                              cae_tensor x, cae_tensor dz, int dz_embed, float *dw,
                              const float *scale, void *workspace, size_t workspace_bytes,
                              void *stream) {
  CAE_CHECK(kind >= CAE_CONV_S1 && kind <= CAE_CONVT_S2, 2, "cae_conv_wgrad: bad kind %d", kind);
  CAE_CHECK(x.ptr && dz.ptr && dw && workspace, 2, "cae_conv_wgrad: null pointer");
  CAE_CHECK(workspace_bytes >= cae_conv_wgrad_workspace_bytes(), 2,
            "cae_conv_wgrad: workspace smaller than cae_conv_wgrad_workspace_bytes()");
  CAE_CHECK(n > 0 && h_in > 0 && w_in > 0 && c_in > 0 && c_out > 0, 2, "cae_conv_wgrad: bad shape");
  CAE_CHECK(c_in <= 256 && c_out <= 256, 2, "cae_conv_wgrad: at most 256 channels per side");
  const bool down = kind == CAE_CONV_S2, up = kind == CAE_CONVT_S2;
  if (down)
    CAE_CHECK(h_in % 2 == 0 && w_in % 2 == 0, 2, "cae_conv_wgrad: stride-2 convolution needs an even input size");
  CAE_CHECK(dz_embed == 0 || kind == CAE_CONV_S1 || kind == CAE_CONV_S2, 2,
            "cae_conv_wgrad: only Conv2d layers take an embedded output gradient");
  const int h_out = down ? h_in / 2 : (up ? h_in * 2 : h_in);
  const int w_out = down ? w_in / 2 : (up ? w_in * 2 : w_in);
  // P (unshifted) / Q (through the window) and their sizes
  const cae_tensor &P = up ? x : dz;
  const cae_tensor &Q = up ? dz : x;
  const int c_m = up ? c_in : c_out, c_n = up ? c_out : c_in;
  const int hP = up ? h_in : h_out, wP = up ? w_in : w_out;     // tile domain
  const int hQ = up ? h_out : h_in, wQ = up ? w_out : w_in;
  const bool q_split = down || up;
  CAE_CHECK(P.fmt == CAE_FMT_F16_PLANAR, 2, "cae_conv_wgrad: %s must be planar fp16", up ? "x" : "dz");
  CAE_CHECK(Q.fmt == (q_split ? CAE_FMT_F16_SPLIT : CAE_FMT_F16_PLANAR), 2,
            "cae_conv_wgrad: %s must be %s fp16", up ? "dz" : "x", q_split ? "split" : "planar");
  CAE_CHECK(P.planes * 8 == wg_round_up(c_m, 16) && Q.planes * 8 == wg_round_up(c_n, 16), 2,
            "cae_conv_wgrad: plane counts do not match the channel counts");

  WgParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n;
  p.tiles_x = (wP + 7) / 8;
  p.tiles_per_img = p.tiles_x * ((hP + 15) / 16);
  p.n_tiles = p.tiles_per_img * n;
  p.planes_p = P.planes > 16 ? 16 : P.planes;     // one 128-channel slice of P per launch
  p.n_chunks = (Q.planes * 8 + kWgNc - 1) / kWgNc;
  if (q_split) {                 // the stride-2 pattern of igemm_conv.cu (CAE_CONV_S2)
    p.PH = 17; p.PW = 9; p.n_par = 4; p.q_org_y = 0; p.q_org_x = 1;
  } else {
    p.PH = 18; p.PW = 10; p.n_par = 1; p.q_org_y = 0; p.q_org_x = CAE_COL_PAD;
  }
  p.q_box_bytes = (kWgNc / 8) * p.PH * p.PW * 16;
  p.par_stride = wg_round_up(p.q_box_bytes, 128);
  p.p_bytes = p.planes_p * 16 * 8 * 16;
  p.stage_bytes = wg_round_up(p.p_bytes + p.par_stride * p.n_par, 1024);
  // the M = 128 instruction always reads 16 planes of P: with fewer real planes the rest are
  // rows nobody drains, but the reads must stay inside the allocation (slack behind the ring)
  const int slack = 16 * 2048 - p.p_bytes;
  p.stages = (227 * 1024 - 2048 - slack) / p.stage_bytes;
  if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  CAE_CHECK(p.stages >= 1, 2, "cae_conv_wgrad: a stage does not fit shared memory");
  for (int t = 0; t < 9; ++t) {
    const int kh = t / 3, kw = t % 3;
    int par = 0, dy = kh, dx = kw;
    if (q_split) {
      par = ((kh & 1) << 1) | ((kw + CAE_COL_PAD) & 1);
      dy = kh >> 1;
      dx = ((kw + CAE_COL_PAD) >> 1) - 1;
    }
    p.tap_off[t] = (uint32_t)(par * p.par_stride + (dy * p.PW + dx) * 16);
  }
  p.idesc = make_idesc_f16(128, kWgNc) | (1u << 15) | (1u << 16);     // A and B MN-major
  p.kind = kind;
  p.c_in = c_in;
  p.c_out = c_out;
  p.c_m = c_m;
  p.c_n = c_n;
  p.dw = dw;
  p.scale = scale;

  CUtensorMap tmP, tmQ;
  // dz_embed: the gradient sits at offset (1, 1) of a zero-ringed buffer whose logical size is
  // (h + 2) x (w + 2) for stride 1 and (h + 1) x (w + 1) for stride 2 -- the form the data
  // gradient of a reflect-padded Conv2d wants (cae_act_grad)
  const int grow = dz_embed ? (down ? 1 : 2) : 0;
  p.p_off = dz_embed ? 1 : 0;
  if (int rc = wg_tensor_map(&tmP, P, n, hP + grow, wP + grow, false, 8, 16, p.planes_p)) return rc;
  if (int rc = wg_tensor_map(&tmQ, Q, n, hQ, wQ, q_split, p.PW, p.PH, kWgNc / 8)) return rc;

  int grid = cae_sm_count();
  grid -= grid % p.n_chunks;
  if (grid > p.n_tiles * p.n_chunks) grid = p.n_tiles * p.n_chunks;
  const int smem_bytes = p.stages * p.stage_bytes + 1024 + slack;
  CAE_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  p.partial = (float *)workspace;
  p.n_slots = grid / p.n_chunks;
  // P has up to 256 channels, the accumulators 128 lanes: one launch per 128-channel slice of P
  // (its planes are a coordinate of the same tensor map)
  const int c_m_total = c_m;
  for (int m0 = 0; m0 < c_m_total; m0 += 128) {
    p.m0 = m0;
    p.p_plane0 = m0 / 8;
    p.c_m = c_m_total - m0 < 128 ? c_m_total - m0 : 128;
    wgrad_kernel<<<grid, kWgThreads, smem_bytes, (cudaStream_t)stream>>>(tmP, tmQ, p);
    const int total = 9 * p.c_m * c_n;
    wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    cae_count_launch(2);
  }
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_act_grad(const cae_act_grad_desc *d, void *stream) {
  CAE_CHECK(d && d->g.ptr && d->dz.ptr, 2, "cae_act_grad: null pointer");
  CAE_CHECK(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0, 2, "cae_act_grad: bad shape");
  auto ok = [](int fmt) {
    return fmt == CAE_FMT_F32_NCHW || fmt == CAE_FMT_F16_PLANAR || fmt == CAE_FMT_F16_SPLIT;
  };
  CAE_CHECK(ok(d->g.fmt) && ok(d->dz.fmt) && (!d->out.ptr || ok(d->out.fmt)), 2, "cae_act_grad: bad format");
  CAE_CHECK(!d->g2.ptr || ok(d->g2.fmt), 2, "cae_act_grad: bad format of g2");
  if (d->skip.ptr) {
    CAE_CHECK(d->out.ptr && ok(d->skip.fmt) && (!d->gsum.ptr || ok(d->gsum.fmt)), 2,
              "cae_act_grad: a residual layer needs its saved output and the tensor that was added");
    CAE_CHECK(d->post_act != CAE_ACT_RELU, 2,
              "cae_act_grad: ReLU after a residual add is not invertible (LeakyReLU or none)");
  }
  auto slope = [](int act) { return act == CAE_ACT_LEAKY_RELU ? 0.01f : (act == CAE_ACT_RELU ? 0.f : 1.f); };
  AgParams q;
  memset(&q, 0, sizeof(q));
  q.g = AgView{d->g.ptr, d->g.fmt, d->g.planes, d->g_h, d->g_w, d->g_oy, d->g_ox};
  q.out = AgView{d->out.ptr, d->out.fmt, d->out.planes, d->out_h, d->out_w, 0, 0};
  q.dz = AgView{d->dz.ptr, d->dz.fmt, d->dz.planes, d->dz_h, d->dz_w, d->dz_oy, d->dz_ox};
  q.g2 = AgView{d->g2.ptr, d->g2.fmt, d->g2.planes, d->h, d->w, 0, 0};
  q.skip = AgView{d->skip.ptr, d->skip.fmt, d->skip.planes, d->h, d->w, 0, 0};
  q.gsum = AgView{d->gsum.ptr, d->gsum.fmt, d->gsum.planes, d->h, d->w, 0, 0};
  q.n = d->n; q.h = d->h; q.w = d->w; q.c = d->c;
  q.fold = d->fold;
  q.fold_shift = d->fold_shift;
  q.slope = slope(d->act);
  q.post_slope = slope(d->post_act);
  q.scale = d->scale;
  q.db = d->db;
  const int planes = (d->c + 7) / 8;
  CAE_CHECK((long long)d->n * planes <= 65535 && (d->h + 7) / 8 <= 65535, 2,
            "cae_act_grad: tensor too large for one launch; split the batch");
  const dim3 grid((unsigned)((d->w + 31) / 32), (unsigned)((d->h + 7) / 8), (unsigned)(d->n * planes));
  act_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
