// Generalised divisive normalisation (GDN / IGDN), the optional activation of the
// reference's units (_define_act_layer, src/models/tasks/_autoencoders.py:29-30, from
// compressai.layers.GDN):
//     norm_i = beta_i + sum_j gamma_ij * x_j^2 ;  out_i = x_i * rsqrt(norm_i)   (GDN, analysis)
//                                                  out_i = x_i * sqrt(norm_i)    (IGDN, synthesis)
// computed in fp32 on CUDA cores from / to the internal fp16 layouts, with the optional
// residual add that follows it in the residual units (R:172, R:302) and the halo the consumer
// needs.  beta / gamma arrive already re-parametrised (max(p, bound)^2 - pedestal) from the
// host.  First correct version: C^2 MACs per pixel on the FMA pipe (a tensor-core 1x1 form is
// the planned replacement).
#include "cae_common.cuh"

namespace {

constexpr int GP = 64;  // pixels per block

struct GdnParams {
  ActView in, out, skip;
  int n, h, w, c;
  const float *beta, *gamma;
  int inverse;
};

__device__ __forceinline__ size_t unit_off(const ActView &v, int n, int p, int Y, int X) {
  return act_unit_offset(v, n, p, Y, X);
}

__global__ void __launch_bounds__(256) gdn_kernel(const GdnParams p) {
  extern __shared__ float sm[];
  const int C = p.c, Cp = (C + 7) & ~7;
  float *gam = sm;                 // [C][C]
  float *xsq = gam + C * C;        // [Cp][GP]
  float *xv = xsq + Cp * GP;       // [Cp][GP]
  const int tid = threadIdx.x;
  for (int i = tid; i < C * C; i += blockDim.x) gam[i] = p.gamma[i];
  const size_t total = (size_t)p.n * p.h * p.w;
  const size_t pix0 = (size_t)blockIdx.x * GP;
  // stage x and x^2 of GP pixels: one 16-byte unit (8 channels) per (plane, pixel)
  for (int u = tid; u < (Cp / 8) * GP; u += blockDim.x) {
    const int plane = u / GP, px = u - plane * GP;
    const size_t pix = pix0 + px;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pix < total) {
      const int x = (int)(pix % p.w), y = (int)((pix / p.w) % p.h), n = (int)(pix / ((size_t)p.w * p.h));
      const uint4 raw = reinterpret_cast<const uint4 *>(p.in.ptr)[unit_off(p.in, n, plane, y + 1, x + 1)];
      const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h2[k]);
        v[2 * k] = f.x;
        v[2 * k + 1] = f.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xv[(plane * 8 + k) * GP + px] = v[k];
      xsq[(plane * 8 + k) * GP + px] = v[k] * v[k];
    }
  }
  __syncthreads();
  const int px = tid % GP, grp = tid / GP;            // 4 channel groups of Cp/4 channels
  const size_t pix = pix0 + px;
  if (pix >= total) return;
  const int x = (int)(pix % p.w), y = (int)((pix / p.w) % p.h), n = (int)(pix / ((size_t)p.w * p.h));
  const int planes = Cp / 8;
  for (int plane = grp; plane < planes; plane += 4) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = plane * 8 + k;
      float val = 0.f;
      if (i < C) {
        float norm = p.beta[i];
        const float *g = gam + (size_t)i * C;
        for (int j = 0; j < C; ++j) norm = fmaf(g[j], xsq[j * GP + px], norm);
        const float xi = xv[i * GP + px];
        val = p.inverse ? xi * sqrtf(norm) : xi * rsqrtf(norm);
      }
      o[k] = val;
    }
    if (p.skip.ptr && p.skip.fmt == CAE_FMT_U8_HWC) {      // unit input is the raw image
      const uint8_t *q = reinterpret_cast<const uint8_t *>(p.skip.ptr);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (plane * 8 + k < C)
          o[k] += (float)q[(((size_t)n * p.h + y) * p.w + x) * C + plane * 8 + k] / 255.0f;
    } else if (p.skip.ptr && p.skip.fmt == CAE_FMT_F32_NCHW) {
      const float *q = reinterpret_cast<const float *>(p.skip.ptr);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (plane * 8 + k < C) o[k] += q[(((size_t)n * C + plane * 8 + k) * p.h + y) * p.w + x];
    } else if (p.skip.ptr) {
      const uint4 raw = reinterpret_cast<const uint4 *>(p.skip.ptr)[unit_off(p.skip, n, plane, y + 1, x + 1)];
      const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h2[k]);
        o[2 * k] += f.x;
        o[2 * k + 1] += f.y;
      }
    }
    __half2 hh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) hh[k] = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
    const uint4 u = *reinterpret_cast<uint4 *>(hh);
    uint4 *base = reinterpret_cast<uint4 *>(p.out.ptr);
    const int ys[3] = {y + 1, (p.out.halo == CAE_HALO_REFLECT && y == 1) ? 0 : -1,
                       (p.out.halo == CAE_HALO_REFLECT && y == p.h - 2) ? p.h + 1 : -1};
    const int xs[3] = {x + 1, (p.out.halo == CAE_HALO_REFLECT && x == 1) ? 0 : -1,
                       (p.out.halo == CAE_HALO_REFLECT && x == p.w - 2) ? p.w + 1 : -1};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        if (ys[a] >= 0 && xs[b] >= 0) base[unit_off(p.out, n, plane, ys[a], xs[b])] = u;
  }
}

bool planar_like(int f) { return f == CAE_FMT_F16_PLANAR || f == CAE_FMT_F16_SPLIT; }

}  // namespace

extern "C" int cae_gdn(cae_tensor in, cae_tensor out, cae_tensor skip, int n, int h, int w, int c,
                       const float *beta, const float *gamma, int inverse, void *stream) {
  CAE_CHECK(in.ptr && out.ptr && beta && gamma, 2, "cae_gdn: null argument");
  CAE_CHECK(planar_like(in.fmt) && planar_like(out.fmt), 2, "cae_gdn: tensors must be planar fp16");
  CAE_CHECK(n > 0 && h > 0 && w > 0 && c > 0, 2, "cae_gdn: bad shape");
  const int Cp = (c + 7) & ~7;
  CAE_CHECK(in.planes * 8 >= Cp && out.planes * 8 >= Cp, 2, "cae_gdn: planes too few");
  const size_t smem = ((size_t)c * c + 2 * (size_t)Cp * GP) * sizeof(float);
  CAE_CHECK(smem <= 227 * 1024, 2, "cae_gdn: %d channels need %zu B of shared memory", c, smem);
  GdnParams p;
  memset(&p, 0, sizeof(p));
  p.in.ptr = in.ptr; p.in.fmt = in.fmt; p.in.planes = in.planes; p.in.H = h; p.in.W = w;
  p.out.ptr = out.ptr; p.out.fmt = out.fmt; p.out.planes = out.planes; p.out.halo = out.halo;
  p.out.H = h; p.out.W = w;
  if (skip.fmt != CAE_FMT_NONE && skip.ptr) {
    CAE_CHECK((planar_like(skip.fmt) && skip.planes * 8 >= Cp) || skip.fmt == CAE_FMT_F32_NCHW ||
                  skip.fmt == CAE_FMT_U8_HWC,
              2, "cae_gdn: bad skip tensor");
    p.skip.ptr = skip.ptr; p.skip.fmt = skip.fmt; p.skip.planes = skip.planes;
    p.skip.H = h; p.skip.W = w;
  }
  if (out.fmt == CAE_FMT_F16_SPLIT || in.fmt == CAE_FMT_F16_SPLIT)
    CAE_CHECK(h % 2 == 0 && w % 2 == 0, 2, "cae_gdn: split layout needs even size");
  p.n = n; p.h = h; p.w = w; p.c = c;
  p.beta = beta; p.gamma = gamma; p.inverse = inverse;
  CAE_CUDA(cudaFuncSetAttribute(gdn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t total = (size_t)n * h * w;
  gdn_kernel<<<(unsigned)((total + GP - 1) / GP), 256, smem, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
