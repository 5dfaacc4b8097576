// Micro-benchmark: how fast can 148 persistent CTAs write a 128 x 16 planes x 130 x 130 x 16 B
// planar fp16 tensor (537 MB) with different store patterns?  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int NIMG = 128, PLANES = 16, H = 128, W = 128, PH = H + 2, PW = W + 2;

// mode 0: tile 8 rows x 32 px; warp = one row of 32 px; per (row, plane): 512 B at a 16 B shift
// mode 1: same but no +1 shift and pitch 128 (fully aligned 512 B runs)
// mode 2: tile 2 rows x 128 px: warp writes 4 consecutive 512 B runs of one plane row (2 KB)
// mode 3: linear fill
// mode 4: tile 8 x 32, but plane-major inside the warp: for plane: for row (8 rows of one plane)
template <int MODE>
__global__ void __launch_bounds__(512) store_kernel(uint4 *out, int warps_used) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_used) return;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  const size_t plane_stride = (size_t)PH * PW;
  if (MODE == 3) {
    const size_t total = (size_t)NIMG * PLANES * plane_stride;
    const size_t nw = (size_t)gridDim.x * warps_used;
    for (size_t i = ((size_t)blockIdx.x * warps_used + warp) * 32 + lane; i < total; i += nw * 32) out[i] = v;
    return;
  }
  const int tiles_x = MODE == 2 ? 1 : W / 32, tiles_y = MODE == 2 ? H / 2 : H / 8;
  const int n_tiles = NIMG * tiles_x * tiles_y;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y), rem = tile % (tiles_x * tiles_y);
    const int tyi = rem / tiles_x, txi = rem % tiles_x;
    if (MODE == 0 || MODE == 1 || MODE == 4) {
      // 8 rows x 16 planes = 128 (row, plane) runs per tile, spread over the warps
      for (int u = warp; u < 8 * PLANES; u += warps_used) {
        const int row = MODE == 4 ? (u & 7) : (u / PLANES), plane = MODE == 4 ? (u >> 3) : (u % PLANES);
        const int oy = tyi * 8 + row, ox = txi * 32 + lane;
        const size_t off = ((size_t)n * PLANES + plane) * plane_stride +
                           (MODE == 1 ? (size_t)oy * PW + ox : (size_t)(oy + 1) * PW + ox + 1);
        out[off] = v;
      }
    } else {
      // 2 rows x 16 planes = 32 runs of 2 KB
      for (int u = warp; u < 2 * PLANES; u += warps_used) {
        const int row = u / PLANES, plane = u % PLANES;
        const int oy = tyi * 2 + row;
        const size_t off = ((size_t)n * PLANES + plane) * plane_stride + (size_t)(oy + 1) * PW + 1 + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) out[off + 32 * q] = v;
      }
    }
  }
}

int main(int argc, char **argv) {
  const size_t bytes = (size_t)NIMG * PLANES * PH * PW * 16;
  uint4 *out;
  cudaMalloc(&out, bytes);
  char *flush;
  cudaMalloc(&flush, 256 << 20);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 5; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(flush, rep, 256 << 20);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        switch (mode) {
          case 0: store_kernel<0><<<148, 512>>>(out, warps); break;
          case 1: store_kernel<1><<<148, 512>>>(out, warps); break;
          case 2: store_kernel<2><<<148, 512>>>(out, warps); break;
          case 3: store_kernel<3><<<148, 512>>>(out, warps); break;
          default: store_kernel<4><<<148, 512>>>(out, warps); break;
        }
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      const double payload = (double)NIMG * PLANES * H * W * 16;
      printf("warps %2d mode %d: %.1f us  %.2f TB/s (%s)\n", warps, mode, best * 1e3, payload / best / 1e9,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
