// CUDA-core direct 3x3 (transposed) convolution for the thin image-side layers
// (c_in < 16: the channels_org -> channels_org stem of the reference's
// DownsamplingUnit / ResidualDownsamplingUnit, _autoencoders.py:63-70, 114-137)
// and for nets too narrow for the tensor-core tile (the MNIST-size net).  It
// accepts every tensor format of the ABI, resolves reflect / zero padding by
// index arithmetic on the interior (it never reads a halo), and shares the
// epilogue definition of cae_conv_desc with the implicit-GEMM kernel:
//   out = post_act( pre_act(conv(in) + bias) + skip ).
#include "cae_common.cuh"

namespace {

struct DcParams {
  int kind, n, h_in, w_in, h_out, w_out, c_in, c_out;
  int stride, transposed, pad_mode;
  int in_fmt;
  ActView in, out, skip;
  int out_fmt;
  const float *w;
  const float *bias;
  int pre_act, post_act;
  float *aux;
  int out_channels_padded;  // planar outputs: planes * 8
  int groups, cin_pg, cout_pg;   // grouped convolution (groups=channels_in in the reference)
};

__device__ __forceinline__ float load_in(const DcParams &p, int n, int c, int y, int x) {
  if (p.in_fmt == CAE_FMT_U8_HWC) {
    const uint8_t *q = reinterpret_cast<const uint8_t *>(p.in.ptr);
    return (float)q[(((size_t)n * p.h_in + y) * p.w_in + x) * p.c_in + c] / 255.0f;
  }
  if (p.in_fmt == CAE_FMT_F32_NCHW) {
    const float *q = reinterpret_cast<const float *>(p.in.ptr);
    return q[(((size_t)n * p.c_in + c) * p.h_in + y) * p.w_in + x];
  }
  const __half *q = reinterpret_cast<const __half *>(p.in.ptr);
  return __half2float(q[act_unit_offset(p.in, n, c >> 3, y + 1, x + 1) * 8 + (c & 7)]);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

constexpr int CO_BLK = 8;

__global__ void __launch_bounds__(128) direct_conv_kernel(const DcParams p) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)p.n * p.h_out * p.w_out;
  if (pix >= total) return;
  const int ox = (int)(pix % p.w_out);
  const int oy = (int)((pix / p.w_out) % p.h_out);
  const int n = (int)(pix / ((size_t)p.w_out * p.h_out));
  const int co0 = blockIdx.y * CO_BLK;

  float acc[CO_BLK];
#pragma unroll
  for (int i = 0; i < CO_BLK; ++i) acc[i] = 0.f;

  for (int kh = 0; kh < 3; ++kh) {
    for (int kw = 0; kw < 3; ++kw) {
      int iy, ix;
      if (!p.transposed) {
        iy = oy * p.stride + kh - 1;
        ix = ox * p.stride + kw - 1;
        if (p.pad_mode == CAE_PAD_REFLECT) {
          iy = reflect_idx(iy, p.h_in);
          ix = reflect_idx(ix, p.w_in);
        } else if (iy < 0 || iy >= p.h_in || ix < 0 || ix >= p.w_in) {
          continue;
        }
      } else {
        // out[o] += in[i] * w[kh] with o = i*stride - 1 + kh
        const int ty = oy + 1 - kh, tx = ox + 1 - kw;
        if (ty < 0 || tx < 0 || ty % p.stride || tx % p.stride) continue;
        iy = ty / p.stride;
        ix = tx / p.stride;
        if (iy >= p.h_in || ix >= p.w_in) continue;
      }
      if (p.groups > 1) {
        // nn.Conv2d(groups=G): weight (c_out, c_in/G, 3, 3); nn.ConvTranspose2d(groups=G): weight
        // (c_in, c_out/G, 3, 3); output channel co belongs to group co / (c_out/G) and only sees
        // that group's c_in/G input channels (R:68, 83, 119, 135, 153: groups=channels_in)
#pragma unroll
        for (int i = 0; i < CO_BLK; ++i) {
          const int co = co0 + i;
          if (co >= p.c_out) continue;
          const int g = co / p.cout_pg, col = co - g * p.cout_pg;
          for (int cl = 0; cl < p.cin_pg; ++cl) {
            const int ci = g * p.cin_pg + cl;
            const size_t wi = p.transposed ? (((size_t)ci * p.cout_pg + col) * 3 + kh) * 3 + kw
                                           : (((size_t)co * p.cin_pg + cl) * 3 + kh) * 3 + kw;
            acc[i] = fmaf(load_in(p, n, ci, iy, ix), __ldg(p.w + wi), acc[i]);
          }
        }
        continue;
      }
      for (int ci = 0; ci < p.c_in; ++ci) {
        const float v = load_in(p, n, ci, iy, ix);
#pragma unroll
        for (int i = 0; i < CO_BLK; ++i) {
          const int co = co0 + i;
          if (co < p.c_out) {
            const size_t wi = p.transposed ? (((size_t)ci * p.c_out + co) * 3 + kh) * 3 + kw
                                           : (((size_t)co * p.c_in + ci) * 3 + kh) * 3 + kw;
            acc[i] = fmaf(v, __ldg(p.w + wi), acc[i]);
          }
        }
      }
    }
  }

  float v[CO_BLK];
#pragma unroll
  for (int i = 0; i < CO_BLK; ++i) {
    const int co = co0 + i;
    float t = acc[i];
    if (co < p.c_out) {
      if (p.bias) t += p.bias[co];
      t = apply_act(t, p.pre_act);
      if (p.skip.ptr) {
        if (p.skip.fmt == CAE_FMT_U8_HWC) {
          const uint8_t *q = reinterpret_cast<const uint8_t *>(p.skip.ptr);
          t += (float)q[(((size_t)n * p.h_out + oy) * p.w_out + ox) * p.c_out + co] / 255.0f;
        } else if (p.skip.fmt == CAE_FMT_F32_NCHW) {
          const float *q = reinterpret_cast<const float *>(p.skip.ptr);
          t += q[(((size_t)n * p.c_out + co) * p.h_out + oy) * p.w_out + ox];
        } else {
          const __half *q = reinterpret_cast<const __half *>(p.skip.ptr);
          t += __half2float(
              q[act_unit_offset(p.skip, n, co >> 3, oy + 1, ox + 1) * 8 + (co & 7)]);
        }
      }
      t = apply_act(t, p.post_act);
    } else {
      t = 0.f;
    }
    v[i] = t;
  }

  if (p.aux) {
#pragma unroll
    for (int i = 0; i < CO_BLK; ++i) {
      const int co = co0 + i;
      if (co < p.c_out) p.aux[(((size_t)n * p.c_out + co) * p.h_out + oy) * p.w_out + ox] = v[i];
    }
  }
  if (!p.out.ptr) return;
  if (p.out_fmt == CAE_FMT_U8_HWC) {
    uint8_t *q = reinterpret_cast<uint8_t *>(p.out.ptr);
#pragma unroll
    for (int i = 0; i < CO_BLK; ++i) {
      const int co = co0 + i;
      if (co < p.c_out)
        q[(((size_t)n * p.h_out + oy) * p.w_out + ox) * p.c_out + co] = to_u8_trunc(v[i]);
    }
  } else if (p.out_fmt == CAE_FMT_F32_NCHW) {
    float *q = reinterpret_cast<float *>(p.out.ptr);
#pragma unroll
    for (int i = 0; i < CO_BLK; ++i) {
      const int co = co0 + i;
      if (co < p.c_out) q[(((size_t)n * p.c_out + co) * p.h_out + oy) * p.w_out + ox] = v[i];
    }
  } else {
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const uint4 u = *reinterpret_cast<uint4 *>(h);
    uint4 *base = reinterpret_cast<uint4 *>(p.out.ptr);
    const int plane = blockIdx.y;
    int ys[3], xs[3], ny = 1, nx = 1;
    ys[0] = oy + 1;
    xs[0] = ox + 1;
    if (p.out.halo == CAE_HALO_REFLECT) {
      if (oy == 1) ys[ny++] = 0;
      if (oy == p.h_out - 2) ys[ny++] = p.h_out + 1;
      if (ox == 1) xs[nx++] = 0;
      if (ox == p.w_out - 2) xs[nx++] = p.w_out + 1;
    }
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) base[act_unit_offset(p.out, n, plane, ys[a], xs[b])] = u;
  }
}

// Tiled fast path for the image-side stem (Conv2d k3 s1, c_in <= 4, c_out <= 4): the input
// tile with its halo is staged in shared memory once (coalesced uint8 / fp32 reads), every
// thread produces ST_PX horizontally adjacent pixels from a register window (weights are
// read once per thread, not once per pixel), and a planar output is one 16-byte store per
// pixel (channels 4..7 of plane 0 are zero; the other planes stay zero from allocation).
constexpr int ST_W = 64, ST_H = 16, ST_PX = 4;
constexpr int ST_TX = ST_W / ST_PX;

template <int CI, int CO>
__global__ void __launch_bounds__(ST_TX * ST_H) stem_conv_kernel(const DcParams p) {
  __shared__ float tile[CI][ST_H + 2][ST_W + 2];
  __shared__ float wsm[CO * CI * 9];
  const int tid = threadIdx.y * ST_TX + threadIdx.x;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * ST_W, y0 = blockIdx.y * ST_H;
  for (int i = tid; i < CO * CI * 9; i += ST_TX * ST_H) wsm[i] = p.w[i];
  constexpr int cells = (ST_H + 2) * (ST_W + 2);
  if (p.in_fmt == CAE_FMT_U8_HWC) {
    // uint8 HWC rows are contiguous bytes: one row of the tile (+halo) = (ST_W+2)*CI bytes,
    // loaded by consecutive threads; (float)b / 255.0f is precomputed per byte value (exact).
    __shared__ float lut[256];
    lut[tid & 255] = (float)(tid & 255) / 255.0f;
    __syncthreads();
    const uint8_t *src = reinterpret_cast<const uint8_t *>(p.in.ptr) +
                         (size_t)n * p.h_in * p.w_in * CI;
    constexpr int row_bytes = (ST_W + 2) * CI;
    constexpr int passes = (row_bytes + ST_TX * ST_H - 1) / (ST_TX * ST_H);
#pragma unroll
    for (int pass = 0; pass < passes; ++pass) {
      const int t = tid + pass * ST_TX * ST_H;
      if (t >= row_bytes) break;
      const int col = t / CI, c = t - col * CI;
      int gx = x0 - 1 + col;
      bool col_ok = true;
      if (p.pad_mode == CAE_PAD_REFLECT) {
        gx = reflect_idx(gx, p.w_in);
        gx = gx < 0 ? 0 : (gx >= p.w_in ? p.w_in - 1 : gx);
      } else {
        col_ok = gx >= 0 && gx < p.w_in;
      }
      // all row loads are issued before any is consumed (18 independent global loads in flight)
      uint8_t b[ST_H + 2];
      bool ok[ST_H + 2];
#pragma unroll
      for (int r = 0; r < ST_H + 2; ++r) {
        int gy = y0 - 1 + r;
        ok[r] = col_ok;
        if (p.pad_mode == CAE_PAD_REFLECT) {
          gy = reflect_idx(gy, p.h_in);
          gy = gy < 0 ? 0 : (gy >= p.h_in ? p.h_in - 1 : gy);
        } else {
          ok[r] = ok[r] && gy >= 0 && gy < p.h_in;
          gy = gy < 0 ? 0 : (gy >= p.h_in ? p.h_in - 1 : gy);
        }
        b[r] = ok[r] ? __ldg(src + ((size_t)gy * p.w_in + gx) * CI + c) : (uint8_t)0;
      }
#pragma unroll
      for (int r = 0; r < ST_H + 2; ++r) tile[c][r][col] = ok[r] ? lut[b[r]] : 0.f;
    }
  } else {
    for (int i = tid; i < cells * CI; i += ST_TX * ST_H) {
      const int cell = i % cells, c = i / cells;
      const int r = cell / (ST_W + 2), col = cell - r * (ST_W + 2);
      int gy = y0 - 1 + r, gx = x0 - 1 + col;
      float v = 0.f;
      bool inside = true;
      if (p.pad_mode == CAE_PAD_REFLECT) {
        gy = reflect_idx(gy, p.h_in);
        gx = reflect_idx(gx, p.w_in);
        // tiles past the image edge (partial tiles): clamp, the result is never stored
        gy = gy < 0 ? 0 : (gy >= p.h_in ? p.h_in - 1 : gy);
        gx = gx < 0 ? 0 : (gx >= p.w_in ? p.w_in - 1 : gx);
      } else {
        inside = gy >= 0 && gy < p.h_in && gx >= 0 && gx < p.w_in;
      }
      if (inside) v = load_in(p, n, c, gy, gx);
      tile[c][r][col] = v;
    }
  }
  __syncthreads();

  float acc[ST_PX][CO];
#pragma unroll
  for (int q = 0; q < ST_PX; ++q)
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[q][co] = 0.f;
  const int lx = threadIdx.x * ST_PX;
#pragma unroll
  for (int ci = 0; ci < CI; ++ci)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      float win[ST_PX + 2];
#pragma unroll
      for (int j = 0; j < ST_PX + 2; ++j) win[j] = tile[ci][threadIdx.y + kh][lx + j];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float wv = wsm[((co * CI + ci) * 3 + kh) * 3 + kw];
#pragma unroll
          for (int q = 0; q < ST_PX; ++q) acc[q][co] = fmaf(win[q + kw], wv, acc[q][co]);
        }
    }

  const int oy = y0 + threadIdx.y;
  if (oy >= p.h_out) return;
  // max(v, v*slope): slope 1 identity, 0.01 LeakyReLU, 0 ReLU
  const float pre_s = p.pre_act == CAE_ACT_LEAKY_RELU ? 0.01f : (p.pre_act == CAE_ACT_RELU ? 0.f : 1.f);
  const float post_s = p.post_act == CAE_ACT_LEAKY_RELU ? 0.01f : (p.post_act == CAE_ACT_RELU ? 0.f : 1.f);
#pragma unroll 1
  for (int q = 0; q < ST_PX; ++q) {
    const int ox = x0 + lx + q;
    if (ox >= p.w_out) break;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      float a = acc[0][co];
#pragma unroll
      for (int qq = 1; qq < ST_PX; ++qq) a = q == qq ? acc[qq][co] : a;
      if (p.bias) a += p.bias[co];
      a = fmaxf(a, a * pre_s);
      if (p.skip.ptr) {
        if (p.skip.fmt == CAE_FMT_U8_HWC) {
          const uint8_t *s8 = reinterpret_cast<const uint8_t *>(p.skip.ptr);
          a += (float)s8[(((size_t)n * p.h_out + oy) * p.w_out + ox) * CO + co] / 255.0f;
        } else if (p.skip.fmt == CAE_FMT_F32_NCHW) {
          const float *sf = reinterpret_cast<const float *>(p.skip.ptr);
          a += sf[(((size_t)n * CO + co) * p.h_out + oy) * p.w_out + ox];
        } else {
          const __half *sh = reinterpret_cast<const __half *>(p.skip.ptr);
          a += __half2float(sh[act_unit_offset(p.skip, n, 0, oy + 1, ox + 1) * 8 + co]);
        }
      }
      v[co] = fmaxf(a, a * post_s);
    }
    if (p.aux)
      for (int co = 0; co < CO; ++co)
        p.aux[(((size_t)n * CO + co) * p.h_out + oy) * p.w_out + ox] = v[co];
    if (!p.out.ptr) continue;
    if (p.out_fmt == CAE_FMT_U8_HWC) {
      uint8_t *o8 = reinterpret_cast<uint8_t *>(p.out.ptr);
      for (int co = 0; co < CO; ++co)
        o8[(((size_t)n * p.h_out + oy) * p.w_out + ox) * CO + co] = to_u8_trunc(v[co]);
    } else if (p.out_fmt == CAE_FMT_F32_NCHW) {
      float *of = reinterpret_cast<float *>(p.out.ptr);
      for (int co = 0; co < CO; ++co)
        of[(((size_t)n * CO + co) * p.h_out + oy) * p.w_out + ox] = v[co];
    } else {
      __half2 h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      const uint4 u = *reinterpret_cast<uint4 *>(h);
      uint4 *base = reinterpret_cast<uint4 *>(p.out.ptr);
      base[act_unit_offset(p.out, n, 0, oy + 1, ox + 1)] = u;
      if (p.out.halo == CAE_HALO_REFLECT) {
        const int y2 = oy == 1 ? 0 : -1, y3 = oy == p.h_out - 2 ? p.h_out + 1 : -1;
        const int x2 = ox == 1 ? 0 : -1, x3 = ox == p.w_out - 2 ? p.w_out + 1 : -1;
        if (y2 < 0 && y3 < 0 && x2 < 0 && x3 < 0) continue;
        const int ys[3] = {oy + 1, y2, y3}, xs[3] = {ox + 1, x2, x3};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            if ((a | b) != 0 && ys[a] >= 0 && xs[b] >= 0)
              base[act_unit_offset(p.out, n, 0, ys[a], xs[b])] = u;
      }
    }
  }
}

// fp32 NCHW -> planar fp16 (zero halo untouched, reflect halo optional)
__global__ void nchw_to_planar_kernel(const float *__restrict__ src, int n_img, int c, int h, int w,
                                      ActView dst) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n_img * dst.planes * h * w;
  if (idx >= total) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const int plane = (int)((idx / ((size_t)w * h)) % dst.planes);
  const int n = (int)(idx / ((size_t)w * h * dst.planes));
  __half2 hv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = 0.f, b = 0.f;
    const int c0 = plane * 8 + 2 * i;
    if (c0 < c) a = src[(((size_t)n * c + c0) * h + y) * w + x];
    if (c0 + 1 < c) b = src[(((size_t)n * c + c0 + 1) * h + y) * w + x];
    hv[i] = __floats2half2_rn(a, b);
  }
  const uint4 u = *reinterpret_cast<uint4 *>(hv);
  uint4 *base = reinterpret_cast<uint4 *>(dst.ptr);
  int ys[3], xs[3], ny = 1, nx = 1;
  ys[0] = y + 1;
  xs[0] = x + 1;
  if (dst.halo == CAE_HALO_REFLECT) {
    if (y == 1) ys[ny++] = 0;
    if (y == h - 2) ys[ny++] = h + 1;
    if (x == 1) xs[nx++] = 0;
    if (x == w - 2) xs[nx++] = w + 1;
  }
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) base[act_unit_offset(dst, n, plane, ys[a], xs[b])] = u;
}

__global__ void planar_to_nchw_kernel(ActView src, int n_img, int c, int h, int w,
                                      float *__restrict__ dst) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n_img * c * h * w;
  if (idx >= total) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const int ch = (int)((idx / ((size_t)w * h)) % c);
  const int n = (int)(idx / ((size_t)w * h * c));
  const __half *q = reinterpret_cast<const __half *>(src.ptr);
  dst[idx] = __half2float(q[act_unit_offset(src, n, ch >> 3, y + 1, x + 1) * 8 + (ch & 7)]);
}

bool planar_fmt(int f) { return f == CAE_FMT_F16_PLANAR || f == CAE_FMT_F16_SPLIT; }

}  // namespace

extern "C" int cae_conv_direct(const cae_conv_desc *d, void *stream) {
  CAE_CHECK(d, 2, "cae_conv_direct: null descriptor");
  CAE_CHECK(d->kind >= CAE_CONV_S1 && d->kind <= CAE_CONVT_S2, 2, "cae_conv_direct: bad kind %d",
            d->kind);
  CAE_CHECK(d->in.ptr && d->weights, 2, "cae_conv_direct: null input or weights");
  CAE_CHECK(d->out.ptr || d->aux_out, 2, "cae_conv_direct: no output");
  DcParams p;
  memset(&p, 0, sizeof(p));
  p.kind = d->kind;
  p.n = d->n;
  p.h_in = d->h_in;
  p.w_in = d->w_in;
  p.c_in = d->c_in;
  p.c_out = d->c_out;
  p.transposed = d->kind == CAE_CONVT_S1 || d->kind == CAE_CONVT_S2;
  p.stride = (d->kind == CAE_CONV_S2 || d->kind == CAE_CONVT_S2) ? 2 : 1;
  p.pad_mode = p.transposed ? CAE_PAD_ZERO : d->pad_mode;
  if (d->kind == CAE_CONV_S2) {
    p.h_out = (d->h_in - 1) / 2 + 1;
    p.w_out = (d->w_in - 1) / 2 + 1;
  } else if (d->kind == CAE_CONVT_S2) {
    p.h_out = d->h_in * 2;
    p.w_out = d->w_in * 2;
  } else {
    p.h_out = d->h_in;
    p.w_out = d->w_in;
  }
  if (p.pad_mode == CAE_PAD_REFLECT)
    CAE_CHECK(d->h_in >= 2 && d->w_in >= 2, 2, "cae_conv_direct: reflect padding needs size >= 2");
  p.in_fmt = d->in.fmt;
  p.in.ptr = d->in.ptr;
  p.in.fmt = d->in.fmt;
  p.in.planes = d->in.planes;
  p.in.H = d->h_in;
  p.in.W = d->w_in;
  if (planar_fmt(d->in.fmt))
    CAE_CHECK(d->in.planes * 8 >= d->c_in, 2, "cae_conv_direct: input planes too few");
  p.out_fmt = d->out.fmt;
  p.out.ptr = d->out.ptr;
  p.out.fmt = d->out.fmt;
  p.out.planes = d->out.planes;
  p.out.halo = d->out.halo;
  p.out.H = p.h_out;
  p.out.W = p.w_out;
  int co_blocks = (d->c_out + CO_BLK - 1) / CO_BLK;
  if (planar_fmt(d->out.fmt) && d->out.ptr) {
    CAE_CHECK(d->out.planes * 8 >= d->c_out, 2, "cae_conv_direct: output planes too few");
    if (d->out.fmt == CAE_FMT_F16_SPLIT)
      CAE_CHECK(p.h_out % 2 == 0 && p.w_out % 2 == 0, 2, "cae_conv_direct: split output needs even size");
    co_blocks = d->out.planes;  // also zero-fills the padding channels
  }
  if (d->skip.fmt != CAE_FMT_NONE && d->skip.ptr) {
    p.skip.ptr = d->skip.ptr;
    p.skip.fmt = d->skip.fmt;
    p.skip.planes = d->skip.planes;
    p.skip.H = p.h_out;
    p.skip.W = p.w_out;
  }
  p.w = (const float *)d->weights;
  p.bias = d->bias;
  p.pre_act = d->pre_act;
  p.post_act = d->post_act;
  p.aux = (float *)d->aux_out;
  p.groups = d->groups > 1 ? d->groups : 1;
  CAE_CHECK(d->c_in % p.groups == 0 && d->c_out % p.groups == 0, 2,
            "cae_conv_direct: groups=%d does not divide c_in=%d / c_out=%d", p.groups, d->c_in, d->c_out);
  p.cin_pg = d->c_in / p.groups;
  p.cout_pg = d->c_out / p.groups;
  if (p.groups == 1 && d->kind == CAE_CONV_S1 && d->c_in <= 4 && d->c_out <= 4 && p.n <= 65535) {
    dim3 grid((unsigned)((p.w_out + ST_W - 1) / ST_W), (unsigned)((p.h_out + ST_H - 1) / ST_H),
              (unsigned)p.n);
    const dim3 blk(ST_TX, ST_H);
    cudaStream_t st = (cudaStream_t)stream;
#define CAE_STEM(CI, CO) \
  if (d->c_in == CI && d->c_out == CO) stem_conv_kernel<CI, CO><<<grid, blk, 0, st>>>(p)
    CAE_STEM(1, 1); CAE_STEM(1, 2); CAE_STEM(1, 3); CAE_STEM(1, 4);
    CAE_STEM(2, 1); CAE_STEM(2, 2); CAE_STEM(2, 3); CAE_STEM(2, 4);
    CAE_STEM(3, 1); CAE_STEM(3, 2); CAE_STEM(3, 3); CAE_STEM(3, 4);
    CAE_STEM(4, 1); CAE_STEM(4, 2); CAE_STEM(4, 3); CAE_STEM(4, 4);
#undef CAE_STEM
    cae_count_launch();
    CAE_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t total = (size_t)p.n * p.h_out * p.w_out;
  dim3 grid((unsigned)((total + 127) / 128), (unsigned)co_blocks);
  direct_conv_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_nchw_to_planar(const float *src, int n, int c, int h, int w, cae_tensor dst,
                                  void *stream) {
  CAE_CHECK(src && dst.ptr && planar_fmt(dst.fmt), 2, "cae_nchw_to_planar: bad arguments");
  CAE_CHECK(dst.planes * 8 >= c, 2, "cae_nchw_to_planar: planes too few");
  if (dst.fmt == CAE_FMT_F16_SPLIT)
    CAE_CHECK(h % 2 == 0 && w % 2 == 0, 2, "cae_nchw_to_planar: split layout needs even size");
  ActView v;
  v.ptr = dst.ptr; v.fmt = dst.fmt; v.planes = dst.planes; v.halo = dst.halo; v.H = h; v.W = w;
  const size_t total = (size_t)n * dst.planes * h * w;
  nchw_to_planar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, n, c, h, w, v);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_planar_to_nchw(cae_tensor src, int n, int c, int h, int w, float *dst,
                                  void *stream) {
  CAE_CHECK(dst && src.ptr && planar_fmt(src.fmt), 2, "cae_planar_to_nchw: bad arguments");
  CAE_CHECK(src.planes * 8 >= c, 2, "cae_planar_to_nchw: planes too few");
  ActView v;
  v.ptr = src.ptr; v.fmt = src.fmt; v.planes = src.planes; v.halo = src.halo; v.H = h; v.W = w;
  const size_t total = (size_t)n * c * h * w;
  planar_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      v, n, c, h, w, dst);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
