// Shared device helpers: sm_100a PTX wrappers (mbarrier, TMA, tcgen05, TMEM),
// internal tensor addressing and the error plumbing of the C ABI.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cae_b200.h"

// ------------------------------------------------------------------ errors
void cae_set_error(const char *fmt, ...);
void cae_count_launch(int n = 1);
int cae_sm_count();   // SMs of the current device (cached per device)

// Bring-up / experiment knobs (environment variables CAE_<NAME>).  The environment is read ONCE,
// at the first call, and a knob is only honoured when CAE_DEBUG=1 is set as well: several of
// them produce wrong output on purpose (loads / MMAs / stores switched off), so a stray variable
// in a production environment must not reach the kernels.  Returns the value or nullptr.
#define CAE_KNOB_LIST(X)                                                                      \
  X(HEAD_TRACE) X(HEAD_DEBUG) X(IGEMM_MERGED_CK64) X(IGEMM_CK_S2) X(IGEMM_TWO_PASS)           \
  X(IGEMM_ONE_PASS) X(IGEMM_MT) X(IGEMM_SWAP_LBO_SBO) X(QUANT_NO_SMEM) X(IGEMM_TPB)           \
  X(IGEMM_VERBOSE) X(QUANT_NO_HIST) X(QUANT_NO_RATE) X(QUANT_NO_YQ) X(IGEMM_DEBUG)            \
  X(IGEMM_NO_PAIR_STORE) X(IGEMM_EPI_WARPS) X(IGEMM_NO_FAST_EPILOGUE) X(IGEMM_NO_TMA_STORE)   \
  X(IGEMM_PAIR_MMA) X(IGEMM_CK_PAIR) X(IGEMM_NO_RESIDENT) X(RANS_V1) X(RANS_V2)
enum CaeKnob {
#define X(n) CAE_KNOB_##n,
  CAE_KNOB_LIST(X)
#undef X
  CAE_KNOB_COUNT
};
const char *cae_knob(CaeKnob k);

#define CAE_CHECK(cond, code, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      cae_set_error(__VA_ARGS__);             \
      return (code);                          \
    }                                         \
  } while (0)

#define CAE_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      cae_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                    __FILE__, __LINE__);                                      \
      return 100 + (int)_e;                                                   \
    }                                                                         \
  } while (0)

// ------------------------------------------------------- internal layouts
// PLANAR: [N][P][H+2][W+8][8] half.  SPLIT: [N][4][P][(H+2)/2][(W+8)/2][8] half.
// Rows carry CAE_COL_PAD unused units on each side of the 1-pixel halo: pixel x = 0 starts on a
// 32-byte sector, so a warp's run of consecutive pixels is written as whole sectors (a 16-byte
// shift costs ~35 % of the store bandwidth, tools/micro/storebench.cu).
__host__ __device__ inline int cae_row_units(int W) { return W + 2 + 2 * CAE_COL_PAD; }
struct ActView {
  void *ptr;
  int fmt, planes, halo;
  int H, W;  // logical (unpadded) spatial size
};

__host__ __device__ inline size_t cae_act_bytes(int fmt, int n, int planes, int H, int W) {
  return (size_t)n * planes * (H + 2) * cae_row_units(W) * 16;  // same for PLANAR and SPLIT
}

// Element offset (in 16-byte units) of padded pixel (Y,X), plane p, image n.
__device__ __forceinline__ size_t act_unit_offset(const ActView &v, int n, int p, int Y, int X) {
  if (v.fmt == CAE_FMT_F16_PLANAR) {
    return (((size_t)n * v.planes + p) * (v.H + 2) + Y) * cae_row_units(v.W) + X + CAE_COL_PAD;
  } else {
    const int Hh = (v.H + 2) >> 1, Wh = cae_row_units(v.W) >> 1, Xc = X + CAE_COL_PAD;
    const int par = ((Y & 1) << 1) | (Xc & 1);
    return ((((size_t)n * 4 + par) * v.planes + p) * Hh + (Y >> 1)) * Wh + (Xc >> 1);
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CAE_ACT_LEAKY_RELU) return v > 0.f ? v : v * 0.01f;
  if (act == CAE_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}

// (uint8) clip(v*255, 0, 255): truncation toward zero, NaN -> 0
__device__ __forceinline__ uint8_t to_u8_trunc(float v) {
  float s = v * 255.0f;
  s = fminf(fmaxf(s, 0.0f), 255.0f);
  return (uint8_t)(int)s;
}

// ------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Bounded wait: a pipeline bug must surface as a launch failure, not as a hung GPU.  The
// polling loop is out of line so that every wait site costs a handful of instructions.
static __device__ __noinline__ void mbar_wait_slow(uint64_t *bar, uint32_t parity) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("cae_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// Wait that yields its issue slots while it polls: for warps that wait a long time next to
// warps doing CUDA-core work on the same SM (the fused head kernel).
static __device__ __noinline__ void mbar_wait_backoff_slow(uint64_t *bar, uint32_t parity) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(100);
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("cae_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_backoff_slow(bar, parity);
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// TMA: 4-D tiled tensor load global -> shared, completes on an mbarrier.
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *dst,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

// The same load issued by either CTA of a pair, completing on an mbarrier of the LEADER CTA
// (`bar_cluster` = its shared::cluster address, mapa_u32): the leader waits on one barrier for
// the patches of both CTAs.
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap *map, uint32_t bar_cluster,
                                                 void *dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

// TMA: 1-D bulk copy global -> shared (bytes multiple of 16, 16-byte aligned).
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// One lane of a converged warp (warp-uniform control flow around it keeps the
// operands of the guarded instructions in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// tcgen05 ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, with each descriptor passed as (lo, hi) 32-bit words and assembled inside the asm
// block: the issuing loop then only advances the lo words (no 64-bit arithmetic per MMA).
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                              uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier when every tcgen05 op issued so far by this thread is done.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA pairs (cta_group::2): one MMA instruction drives the tensor cores of both SMs of a
// cluster of two; each CTA supplies its own 128 rows of A and HALF of the rows of B, the
// accumulator rows of a CTA's tile land in its own TMEM.  PAIR is a template parameter because
// every tcgen05 instruction of a kernel must name the same cta_group.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}

// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}

// The same without release semantics, for arrivals that only hand TMEM / shared-memory stages
// back (ordered by tcgen05.fence / the async proxy): a cluster-scope release is MEMBAR + ERRBAR
// and waits for the warp's outstanding GLOBAL stores -- on the accumulator hand-over that put the
// HBM write latency of the epilogue on the tensor pipe's critical path (ncu source page).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}

// wait on a local barrier whose arrivals come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

static __device__ __noinline__ void mbar_wait_cluster_slow(uint64_t *bar, uint32_t parity) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("cae_b200: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  if (!mbar_try_wait_cluster(bar, parity)) mbar_wait_cluster_slow(bar, parity);
}

template <int PAIR>
__device__ __forceinline__ void tmem_alloc_g(uint32_t *dst_smem, uint32_t ncols) {
  if (PAIR) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    tmem_alloc(dst_smem, ncols);
  }
}

template <int PAIR>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr, uint32_t ncols) {
  if (PAIR)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  else
    tmem_dealloc(taddr, ncols);
}

template <int PAIR>
__device__ __forceinline__ void umma_f16_lohi_g(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                                uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        :
        : "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_f16_lohi(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  }
}

template <int PAIR>
__device__ __forceinline__ void umma_f16_g(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_f16(tmem_d, desc_a, desc_b, idesc, accumulate);
  }
}

// PAIR: the arrival is multicast to the barrier at this offset in BOTH CTAs of the pair
template <int PAIR>
__device__ __forceinline__ void umma_commit_g(uint64_t *bar) {
  if (PAIR) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  } else {
    umma_commit(bar);
  }
}

// Shared-memory matrix descriptor, K-major, no swizzle ("interleave") layout:
// core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = byte distance between the two core matrices of one K=16 step,
// SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// Instruction descriptor for kind::f16: F16 x F16 -> F32, both operands K-major.
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                       // D format F32
         | (0u << 7) | (0u << 10)        // A, B format F16
         | ((uint32_t)(N >> 3) << 17)    // N
         | ((uint32_t)(M >> 4) << 24);   // M
}
