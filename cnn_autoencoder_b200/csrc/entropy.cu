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
  for (int n = g; n < p.n; n += groups) {
  const size_t base = ((size_t)n * p.c + c) * p.hw;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < p.hw; i += gridDim.y * blockDim.x) {
    const float y = p.y[base + i];
    const float r = rintf(y - med);  // torch.round: half to even
    const float yq = r + med;
    // saturate like a float->int32 cast of an in-range value; escapes stay exact up to 2^31
    const int sym = eb_symbol(r);
    if (p.y_q) p.y_q[base + i] = yq;
    if (p.symbols) p.symbols[base + i] = sym;
    if (p.p_y || p.rate_bits) {
      const float lik = eb_lookup(p.t, c, sym, yq, p.status);
      if (p.p_y) p.p_y[base + i] = lik;
      bits -= log2f(lik);
    }
    if (bins) {
      int b = sym - p.t.hist_min;
      b = b < 0 ? 0 : (b >= bins ? bins - 1 : b);
      atomicAdd(&s_hist[b], 1);
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

}  // namespace

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
  int bx = (hw + 256 * 4 - 1) / (256 * 4);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  // about eight blocks per SM in total: images are grouped per channel to reach that
  int groups = (8 * 148) / (c * bx);
  if (groups < 1) groups = 1;
  if (groups > n) groups = n;
  dim3 grid((unsigned)(groups * c), (unsigned)bx);
  eb_quantize_kernel<<<grid, 256, bins * sizeof(int32_t), (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
