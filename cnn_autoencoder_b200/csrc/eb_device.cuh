// Device functions of the factorized-prior entropy model shared by the stand-alone quantizer
// (entropy.cu) and the quantizer fused into the latent layer's epilogue (igemm_conv.cu).
// CompressAI EntropyBottleneck._logits_cumulative / _likelihood,
// reached from src/models/tasks/_taskutils.py:97 and _autoencoders.py:549.
#pragma once

#include "cae_common.cuh"

constexpr int kEbMaxDim = 8;

__device__ __forceinline__ float eb_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// logits_cumulative for one scalar input of one channel
static __device__ __noinline__ float eb_logits(const cae_eb_tables &t, const float *mlp, float x) {
  float v[kEbMaxDim], u[kEbMaxDim];
  v[0] = x;
  const float *q = mlp;
  for (int i = 0; i < t.n_layers; ++i) {
    const int din = t.dims[i], dout = t.dims[i + 1];
    for (int o = 0; o < dout; ++o) {
      float s = 0.f;
      for (int k = 0; k < din; ++k) s += q[o * din + k] * v[k];
      u[o] = s;
    }
    q += dout * din;
    for (int o = 0; o < dout; ++o) u[o] += q[o];
    q += dout;
    if (i < t.n_layers - 1) {
      for (int o = 0; o < dout; ++o) u[o] += q[o] * tanhf(u[o]);
      q += dout;
    }
    for (int o = 0; o < dout; ++o) v[o] = u[o];
  }
  return v[0];
}

static __device__ __noinline__ float eb_likelihood(const cae_eb_tables &t, int c, float v) {
  const float *mlp = t.mlp + (size_t)c * t.mlp_stride;
  const float lower = eb_logits(t, mlp, v - 0.5f);
  const float upper = eb_logits(t, mlp, v + 0.5f);
  const float s = lower + upper;
  const float sign = s > 0.f ? -1.f : (s < 0.f ? 1.f : 0.f);
  const float p = fabsf(eb_sigmoid(sign * upper) - eb_sigmoid(sign * lower));
  return fmaxf(p, t.lik_bound);
}

// sym = rint(y - median) saturated like an in-range float -> int32 cast
__device__ __forceinline__ int eb_symbol(float r) {
  return (int)fminf(fmaxf(r, -2147483520.f), 2147483520.f);
}

// likelihood of symbol `sym` (value yq) of channel c: table, else MLP, else the floor + status
__device__ __forceinline__ float eb_lookup(const cae_eb_tables &t, int c, int sym, float yq,
                                           int32_t *status) {
  const int li = sym - t.lut_min;
  if (t.lut && li >= 0 && li < t.lut_len) return __ldg(t.lut + (size_t)c * t.lut_len + li);
  if (t.lut && t.tail_lik > 0.f) return t.tail_lik;   // the table ends where the bound begins
  if (t.mlp) return eb_likelihood(t, c, yq);
  if (status) atomicOr(status, 1);
  return t.lik_bound;
}

inline int eb_check_tables(const cae_eb_tables *t, const char *who) {
  CAE_CHECK(t->medians, 2, "%s: null medians", who);
  if (t->mlp) {
    CAE_CHECK(t->n_layers >= 1 && t->n_layers <= 9, 2, "%s: bad n_layers", who);
    for (int i = 0; i <= t->n_layers; ++i)
      CAE_CHECK(t->dims[i] >= 1 && t->dims[i] <= kEbMaxDim, 2, "%s: filter dim %d > %d", who,
                t->dims[i], kEbMaxDim);
  }
  return 0;
}
