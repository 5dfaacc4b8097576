// Training-mode EntropyBottleneck: additive-noise quantisation proxy + factorized-prior
// likelihood, forward and backward, one fused kernel each way.
//
// Replaces CompressAI's EntropyBottleneck.forward in training mode,
// which the reference reaches from forward_func (src/models/tasks/_taskutils.py:97) inside the
// train step (src/train_cae_ms.py:209-219) and which runs there as ~60 small ATen launches plus
// two permute copies per direction.  Arithmetic per element (channel c):
//   v   = y + noise
//   l,u = logits_c(v -+ 1/2)           logits_c: z_i = W_i a_i + b_i; a_{i+1} = z_i + phi_i tanh(z_i)
//   s   = -sign(l + u)   (detached);   p = |sigmoid(s u) - sigmoid(s l)|;   lik = max(p, bound)
// with the EFFECTIVE per-channel parameters W_i = softplus(_matrix_i), b_i = _bias_i,
// phi_i = tanh(_factor_i) handed in as one blob (the tiny softplus / tanh Jacobians stay with
// torch autograd on the host side).  Backward: LowerBound passes the gradient where p >= bound or
// the gradient is negative; the density network is differentiated by hand, the 58 parameter
// gradients of a channel are accumulated in registers over the thread's elements, reduced with
// warp shuffles and added to the per-channel result with one atomic per block and parameter.
// filters = (3, 3, 3, 3) (the reference's K = 4, r = 3, _taskargs.py) is the compiled shape.
#include "cae_common.cuh"

namespace {

constexpr int kL = 5;                          // layers of the density network
constexpr int kR = 3;                          // hidden width
constexpr int kBlob = 58;                      // 9 + 3 * 15 + 4 effective parameters per channel
// offsets of (W, b, phi) of every layer inside the blob
// (constexpr functions, not arrays: after unrolling every index into the register-resident
// accumulators is a compile-time constant)
__host__ __device__ constexpr int offW(int i) { return i == 0 ? 0 : 9 + 15 * (i - 1); }
__host__ __device__ constexpr int din_of(int i) { return i == 0 ? 1 : kR; }
__host__ __device__ constexpr int dout_of(int i) { return i == kL - 1 ? 1 : kR; }
__host__ __device__ constexpr int offB(int i) { return offW(i) + din_of(i) * dout_of(i); }
__host__ __device__ constexpr int offP(int i) { return offB(i) + dout_of(i); }
static_assert(offW(4) == 54 && offB(4) == 57 && offP(3) == 51 && offB(0) == 3, "blob layout");

__device__ __forceinline__ float sigmoidf_(float t) { return 1.0f / (1.0f + expf(-t)); }

struct Trace {               // activations of one evaluation, kept for the backward pass
  float a[kL][kR];           // layer inputs
  float t[kL - 1][kR];       // tanh(z_i)
};

__device__ __forceinline__ float logits_fwd(const float *w, float x, Trace *tr) {
  float a[kR] = {x, 0.f, 0.f};
  float out = 0.f;
#pragma unroll
  for (int i = 0; i < kL; ++i) {
    const int din = din_of(i), dout = dout_of(i);
    float z[kR];
#pragma unroll
    for (int o = 0; o < kR; ++o) {
      if (o < dout) {
        float s = w[offB(i) + o];
#pragma unroll
        for (int k = 0; k < kR; ++k)
          if (k < din) s = fmaf(w[offW(i) + o * din + k], a[k], s);
        z[o] = s;
      }
    }
    if (tr) {
#pragma unroll
      for (int k = 0; k < kR; ++k) tr->a[i][k] = k < din ? a[k] : 0.f;
    }
    if (i < kL - 1) {
#pragma unroll
      for (int o = 0; o < kR; ++o) {
        const float t = tanhf(z[o]);
        if (tr) tr->t[i][o] = t;
        a[o] = fmaf(w[offP(i) + o], t, z[o]);
      }
    } else {
      out = z[0];
    }
  }
  return out;
}

// accumulate d out / d params * g into acc[], return d out / d x * g
__device__ __forceinline__ float logits_bwd(const float *w, const Trace &tr, float g, float *acc) {
  float ga[kR] = {g, 0.f, 0.f};       // gradient w.r.t. the layer output (a_{i+1}; z for the last)
#pragma unroll
  for (int i = kL - 1; i >= 0; --i) {
    const int din = din_of(i), dout = dout_of(i);
    float gz[kR];
#pragma unroll
    for (int o = 0; o < kR; ++o) {
      if (o < dout) {
        if (i < kL - 1) {
          const float t = tr.t[i][o], phi = w[offP(i) + o];
          acc[offP(i) + o] = fmaf(ga[o], t, acc[offP(i) + o]);
          gz[o] = ga[o] * fmaf(phi, 1.0f - t * t, 1.0f);
        } else {
          gz[o] = ga[o];
        }
        acc[offB(i) + o] += gz[o];
#pragma unroll
        for (int k = 0; k < kR; ++k)
          if (k < din) acc[offW(i) + o * din + k] = fmaf(gz[o], tr.a[i][k], acc[offW(i) + o * din + k]);
      }
    }
    float gin[kR] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kR; ++k) {
      if (k < din) {
#pragma unroll
        for (int o = 0; o < kR; ++o)
          if (o < dout) gin[k] = fmaf(w[offW(i) + o * din + k], gz[o], gin[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kR; ++k) ga[k] = gin[k];
  }
  return ga[0];
}

struct EbTrainParams {
  const float *y, *noise, *blob;       // y, noise: N x C x hw; blob: C x 58
  int n, c, hw;
  float bound;
  float *y_hat, *lik;                  // forward outputs
  // backward
  const float *g_yhat, *g_lik;         // either may be null
  float *g_y, *g_blob;                 // g_blob: C x 58, accumulated into
};

__global__ void __launch_bounds__(256) eb_train_fwd_kernel(const EbTrainParams p) {
  __shared__ float w[kBlob];
  const int c = blockIdx.x;
  if (threadIdx.x < kBlob) w[threadIdx.x] = p.blob[(size_t)c * kBlob + threadIdx.x];
  __syncthreads();
  const size_t per = (size_t)p.n * p.hw;
  for (size_t e = (size_t)blockIdx.y * blockDim.x + threadIdx.x; e < per; e += (size_t)gridDim.y * blockDim.x) {
    const size_t n = e / p.hw, i = e - n * p.hw;
    const size_t idx = (n * p.c + c) * p.hw + i;
    const float v = p.y[idx] + (p.noise ? p.noise[idx] : 0.f);
    const float l = logits_fwd(w, v - 0.5f, nullptr), u = logits_fwd(w, v + 0.5f, nullptr);
    const float sum = l + u;
    const float s = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    const float pr = fabsf(sigmoidf_(s * u) - sigmoidf_(s * l));
    p.y_hat[idx] = v;
    p.lik[idx] = fmaxf(pr, p.bound);
  }
}

__global__ void __launch_bounds__(128) eb_train_bwd_kernel(const EbTrainParams p) {
  __shared__ float w[kBlob];
  __shared__ float red[4][kBlob];
  const int c = blockIdx.x;
  if (threadIdx.x < kBlob) w[threadIdx.x] = p.blob[(size_t)c * kBlob + threadIdx.x];
  __syncthreads();
  float acc[kBlob];
#pragma unroll
  for (int k = 0; k < kBlob; ++k) acc[k] = 0.f;
  const size_t per = (size_t)p.n * p.hw;
  for (size_t e = (size_t)blockIdx.y * blockDim.x + threadIdx.x; e < per; e += (size_t)gridDim.y * blockDim.x) {
    const size_t n = e / p.hw, i = e - n * p.hw;
    const size_t idx = (n * p.c + c) * p.hw + i;
    const float v = p.y_hat[idx];
    float gv = p.g_yhat ? p.g_yhat[idx] : 0.f;
    const float gl = p.g_lik ? p.g_lik[idx] : 0.f;
    if (gl != 0.f) {
      Trace tl, tu;
      const float l = logits_fwd(w, v - 0.5f, &tl), u = logits_fwd(w, v + 0.5f, &tu);
      const float sum = l + u;
      const float s = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
      const float su = sigmoidf_(s * u), sl = sigmoidf_(s * l);
      const float d = su - sl, pr = fabsf(d);
      const bool pass = pr >= p.bound || gl < 0.f;             // LowerBound's gradient rule
      if (pass && s != 0.f) {
        const float gd = gl * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        const float gu = gd * su * (1.f - su) * s;
        const float gll = -gd * sl * (1.f - sl) * s;
        gv += logits_bwd(w, tu, gu, acc);
        gv += logits_bwd(w, tl, gll, acc);
      }
    }
    p.g_y[idx] = gv;
  }
  // 58 sums: warp shuffles, then one shared-memory round and one atomic per parameter
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kBlob; ++k) {
    float v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kBlob) {
    const float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    if (v != 0.f) atomicAdd(p.g_blob + (size_t)c * kBlob + threadIdx.x, v);
  }
}

int blocks_y(int n, int c, int hw, int threads) {
  const size_t per = (size_t)n * hw;
  int by = (int)((per + threads - 1) / threads);
  const int want = (8 * cae_sm_count() + c - 1) / c;        // ~8 blocks per SM over all channels
  if (by > want) by = want;
  return by < 1 ? 1 : by;
}

}  // namespace

extern "C" int cae_eb_train_blob_size(void) { return kBlob; }

extern "C" int cae_eb_train_fwd(const float *y, const float *noise, const float *blob, int n, int c,
                                int hw, float bound, float *y_hat, float *lik, void *stream) {
  CAE_CHECK(y && blob && y_hat && lik && n > 0 && c > 0 && hw > 0 && c <= 65535, 2,
            "cae_eb_train_fwd: bad argument");
  EbTrainParams p{};
  p.y = y; p.noise = noise; p.blob = blob; p.n = n; p.c = c; p.hw = hw; p.bound = bound;
  p.y_hat = y_hat; p.lik = lik;
  eb_train_fwd_kernel<<<dim3((unsigned)c, (unsigned)blocks_y(n, c, hw, 256)), 256, 0,
                        (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_eb_train_bwd(const float *y_hat, const float *blob, int n, int c, int hw,
                                float bound, const float *g_yhat, const float *g_lik, float *g_y,
                                float *g_blob, void *stream) {
  CAE_CHECK(y_hat && blob && g_y && g_blob && n > 0 && c > 0 && hw > 0 && c <= 65535, 2,
            "cae_eb_train_bwd: bad argument");
  EbTrainParams p{};
  p.blob = blob; p.n = n; p.c = c; p.hw = hw; p.bound = bound;
  p.y_hat = const_cast<float *>(y_hat);
  p.g_yhat = g_yhat; p.g_lik = g_lik; p.g_y = g_y; p.g_blob = g_blob;
  eb_train_bwd_kernel<<<dim3((unsigned)c, (unsigned)blocks_y(n, c, hw, 128)), 128, 0,
                        (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
