// Rounding quantizer + factorized-prior likelihood + per-channel symbol
// histogram + rate, one pass over the latent (HBM-bound elementwise work).
//
// Replaces CompressAI's EntropyBottleneck.forward in eval mode and the symbol
// extraction of EntropyBottleneck.compress, which the reference reaches from
// src/models/tasks/_taskutils.py:97 and src/models/tasks/_autoencoders.py:549
// (SURVEY.md Appendix A.1), plus the rate term of
// src/models/criteria/_ratedist.py:49-54.  In eval mode y_q - median_c is an
// integer, so the likelihood takes one value per (channel, symbol): the host
// builds that table once with the model's own fp32 ops and the kernel looks it
// up; symbols outside the table fall back to evaluating the density MLP here.
#include <stdlib.h>

#include "cae_common.cuh"
#include "eb_device.cuh"

namespace {

struct EbParams {
  const float *y;
  int n, c, hw;
  cae_eb_tables t;
  float *y_q, *p_y;
  int32_t *symbols, *hist;
  double *rate_bits;
  int32_t *status;
};

// grid = (N*C, blocks over hw); every block works inside one channel
__global__ void __launch_bounds__(256) eb_quantize_kernel(const EbParams p) {
  extern __shared__ int32_t s_hist[];
  __shared__ float s_red[8];
  // block = (channel c, image group g, hw slice): it walks images n = g, g + groups, ... of one
  // channel, so the shared histogram is zeroed and flushed once per many thousand symbols
  const int c = blockIdx.x % p.c, g = blockIdx.x / p.c, groups = gridDim.x / p.c;
  const int bins = p.hist ? p.t.hist_bins : 0;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();

  const float med = p.t.medians[c];
  float bits = 0.f;
  const float *lut = p.t.lut ? p.t.lut + (size_t)c * p.t.lut_len : nullptr;
  // one symbol: everything eb_quantize produces for it; returns the histogram bin (or -1)
  auto one = [&](float y, float &yq, int &sym, float &lik) -> int {
    const float r = rintf(y - med);            // torch.round: half to even
    yq = r + med;
    sym = eb_symbol(r);                        // saturating float -> int32, escapes stay exact
    if (p.p_y || p.rate_bits) {
      const int li = sym - p.t.lut_min;
      lik = (lut && li >= 0 && li < p.t.lut_len) ? __ldg(lut + li)
                                                 : eb_lookup(p.t, c, sym, yq, p.status);
      bits -= __log2f(lik);
    }
    if (!bins) return -1;
    const int b = sym - p.t.hist_min;
    return b < 0 ? 0 : (b >= bins ? bins - 1 : b);
  };
  const bool vec4 = (p.hw & 3) == 0;           // rows of four symbols: 16-byte loads and stores
  for (int n = g; n < p.n; n += groups) {
    const size_t base = ((size_t)n * p.c + c) * p.hw;
    if (vec4) {
      const int hw4 = p.hw >> 2;
      for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < hw4; i += gridDim.y * blockDim.x) {
        const float4 y4 = *reinterpret_cast<const float4 *>(p.y + base + 4 * (size_t)i);
        float4 q4, l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 s4;
        const int b0 = one(y4.x, q4.x, s4.x, l4.x), b1 = one(y4.y, q4.y, s4.y, l4.y);
        const int b2 = one(y4.z, q4.z, s4.z, l4.z), b3 = one(y4.w, q4.w, s4.w, l4.w);
        if (p.y_q) *reinterpret_cast<float4 *>(p.y_q + base + 4 * (size_t)i) = q4;
        if (p.symbols) *reinterpret_cast<int4 *>(p.symbols + base + 4 * (size_t)i) = s4;
        if (p.p_y) *reinterpret_cast<float4 *>(p.p_y + base + 4 * (size_t)i) = l4;
        if (bins) {
          // neighbouring symbols often share a bin: merge equal bins before the atomics
          int cnt0 = 1, cnt1 = 1, cnt2 = 1, cnt3 = 1;
          if (b1 == b0) { cnt0 += 1; cnt1 = 0; }
          if (b2 == b0) { cnt0 += 1; cnt2 = 0; } else if (b2 == b1 && cnt1) { cnt1 += 1; cnt2 = 0; }
          if (b3 == b0) { cnt0 += 1; cnt3 = 0; } else if (b3 == b1 && cnt1) { cnt1 += 1; cnt3 = 0; }
          else if (b3 == b2 && cnt2) { cnt2 += 1; cnt3 = 0; }
          atomicAdd(&s_hist[b0], cnt0);
          if (cnt1) atomicAdd(&s_hist[b1], cnt1);
          if (cnt2) atomicAdd(&s_hist[b2], cnt2);
          if (cnt3) atomicAdd(&s_hist[b3], cnt3);
        }
      }
    } else {
      for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < p.hw; i += gridDim.y * blockDim.x) {
        float yq, lik = 0.f;
        int sym;
        const int bb = one(p.y[base + i], yq, sym, lik);
        if (p.y_q) p.y_q[base + i] = yq;
        if (p.symbols) p.symbols[base + i] = sym;
        if (p.p_y) p.p_y[base + i] = lik;
        if (bins) atomicAdd(&s_hist[bb], 1);
      }
    }
  }

  if (p.rate_bits) {
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bits;
  }
  __syncthreads();
  if (p.rate_bits && threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += (double)s_red[w];
    atomicAdd(p.rate_bits, tot);
  }
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const int v = s_hist[i];
    if (v) atomicAdd(&p.hist[(size_t)c * bins + i], v);
  }
}

// int32 symbols (N x C x h x w) -> y_q = sym + median_c in the synthesis track's input layout
// (planar fp16, zero halo untouched): the de-quantisation of EntropyBottleneck.decompress
// ("symbols + medians"; reached from _autoencoders.py:568-572) fused with the
// layout conversion, one pass.  Thread = one pixel x one 8-channel plane.
__global__ void __launch_bounds__(256) eb_dequantize_planar_kernel(
    const int32_t *__restrict__ sym, const float *__restrict__ medians, int n_img, int c, int h,
    int w, ActView dst) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t hw = (size_t)h * w, total = (size_t)n_img * dst.planes * hw;
  if (idx >= total) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const int plane = (int)((idx / hw) % dst.planes);
  const int n = (int)(idx / (hw * dst.planes));
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = plane * 8 + i;
    v[i] = ch < c ? (float)__ldg(sym + ((size_t)n * c + ch) * hw + (size_t)y * w + x) + __ldg(medians + ch)
                  : 0.f;
  }
  __half2 hv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) hv[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  reinterpret_cast<uint4 *>(dst.ptr)[act_unit_offset(dst, n, plane, y + 1, x + 1)] =
      *reinterpret_cast<uint4 *>(hv);
}

}  // namespace

extern "C" int cae_eb_dequantize_planar(const int32_t *symbols, const float *medians, int n, int c,
                                        int h, int w, cae_tensor dst, void *stream) {
  CAE_CHECK(symbols && medians && dst.ptr && dst.fmt == CAE_FMT_F16_PLANAR, 2,
            "cae_eb_dequantize_planar: bad argument");
  CAE_CHECK(n > 0 && c > 0 && h > 0 && w > 0 && dst.planes * 8 >= c, 2,
            "cae_eb_dequantize_planar: bad shape");
  ActView v;
  v.ptr = dst.ptr; v.fmt = dst.fmt; v.planes = dst.planes; v.halo = dst.halo; v.H = h; v.W = w;
  const size_t total = (size_t)n * dst.planes * h * w;
  eb_dequantize_planar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      symbols, medians, n, c, h, w, v);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_eb_quantize(const float *y, int n, int c, int hw, const cae_eb_tables *t,
                               float *y_q, float *p_y, int32_t *symbols, int32_t *hist,
                               double *rate_bits, int32_t *status, void *stream) {
  CAE_CHECK(y && t && t->medians, 2, "cae_eb_quantize: null argument");
  CAE_CHECK(n > 0 && c > 0 && hw > 0, 2, "cae_eb_quantize: bad shape");
  if (int rc = eb_check_tables(t, "cae_eb_quantize")) return rc;
  const int bins = hist ? t->hist_bins : 0;
  CAE_CHECK(bins >= 0 && bins <= 8192, 2, "cae_eb_quantize: hist_bins out of range");
  EbParams p;
  p.y = y; p.n = n; p.c = c; p.hw = hw; p.t = *t;
  p.y_q = y_q; p.p_y = p_y; p.symbols = symbols; p.hist = hist;
  p.rate_bits = rate_bits; p.status = status;
  int bx = (hw + 256 * 16 - 1) / (256 * 16);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  // about eight blocks per SM in total: images are grouped per channel to reach that
  int groups = (8 * cae_sm_count()) / (c * bx);
  if (groups < 1) groups = 1;
  if (groups > n) groups = n;
  dim3 grid((unsigned)(groups * c), (unsigned)bx);
  eb_quantize_kernel<<<grid, 256, bins * sizeof(int32_t), (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
