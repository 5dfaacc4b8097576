// Evaluation sums of the reference's src/test_cae.py on the device, for uint8
// H x W x C images that are already in HBM (the reconstruction as the synthesis transform left it):
//
//  cae_ssim_u8     skimage.metrics.structural_similarity(x, x_r, channel_axis=2) as test_cae.py:52-54
//                  calls it (defaults: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance,
//                  data_range 255, mean over the map cropped by 3 pixels, mean over channels):
//                  per image the SUM of the SSIM map over the cropped region and the channels, in
//                  double; the caller divides by (H - 6)(W - 6)C.
//  cae_delta_e_u8  mean CIE76 colour difference as compute_deltaCIELAB (test_cae.py:21-44):
//                  skimage.color.rgb2lab (sRGB -> linear -> XYZ, D65 / 2 degree observer -> L*a*b*)
//                  of both images and deltaE_cie76: per image the SUM of the per-pixel distances.
//
// Both are HBM bound (two uint8 images in, a handful of doubles out): 16-byte-free scalar loads
// through shared-memory tiles for SSIM (each input byte is read once per 32 x 32 output tile plus
// its 6-pixel apron), one pixel per thread for the colour difference; warp-shuffle + one double
// atomic per block.
#include "cae_common.cuh"

namespace {

constexpr int kSsimWin = 7, kSsimPad = 3;
constexpr int kSsimTile = 32;                       // outputs per block side
constexpr int kSsimIn = kSsimTile + kSsimWin - 1;   // 38

struct SsimParams {
  const uint8_t *a, *b;
  int n, H, W, C;
  double *sum;    // [n]
};

// Block = one 32 x 32 tile of window positions (top-left corners) of one image; the window sums of
// every channel are built separably: horizontal 7-sums of the five moment images into shared
// memory, then vertical 7-sums per output.
__global__ void __launch_bounds__(256) ssim_u8_kernel(const SsimParams p) {
  __shared__ float tile_a[kSsimIn][kSsimIn + 1], tile_b[kSsimIn][kSsimIn + 1];
  __shared__ float h[5][kSsimIn][kSsimTile + 1];     // horizontal sums: a, b, aa, bb, ab
  __shared__ double s_part[8];
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * kSsimTile, ox0 = blockIdx.x * kSsimTile;
  const int OH = p.H - kSsimWin + 1, OW = p.W - kSsimWin + 1;     // number of window positions
  const uint8_t *A = p.a + (size_t)n * p.H * p.W * p.C, *B = p.b + (size_t)n * p.H * p.W * p.C;
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  const double inv_np = 1.0 / 49.0, cov_norm = 49.0 / 48.0;
  double acc = 0.0;
  for (int c = 0; c < p.C; ++c) {
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += blockDim.x) {
      const int r = i / kSsimIn, q = i - r * kSsimIn;
      const int y = oy0 + r, x = ox0 + q;
      float va = 0.f, vb = 0.f;
      if (y < p.H && x < p.W) {
        const size_t o = ((size_t)y * p.W + x) * p.C + c;
        va = (float)A[o];
        vb = (float)B[o];
      }
      tile_a[r][q] = va;
      tile_b[r][q] = vb;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += blockDim.x) {
      const int r = i / kSsimTile, q = i - r * kSsimTile;
      float sa = 0.f, sb = 0.f, saa = 0.f, sbb = 0.f, sab = 0.f;
#pragma unroll
      for (int k = 0; k < kSsimWin; ++k) {
        const float va = tile_a[r][q + k], vb = tile_b[r][q + k];
        sa += va;
        sb += vb;
        saa += va * va;
        sbb += vb * vb;
        sab += va * vb;
      }
      h[0][r][q] = sa;
      h[1][r][q] = sb;
      h[2][r][q] = saa;
      h[3][r][q] = sbb;
      h[4][r][q] = sab;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += blockDim.x) {
      const int r = i / kSsimTile, q = i - r * kSsimTile;
      if (oy0 + r >= OH || ox0 + q >= OW) continue;
      float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < kSsimWin; ++k)
#pragma unroll
        for (int m = 0; m < 5; ++m) s[m] += h[m][r + k][q];
      // the sums are integers < 49 * 255^2 < 2^24, exact in fp32; the moments cancel, so the
      // rest is done in double like skimage does
      const double ux = (double)s[0] * inv_np, uy = (double)s[1] * inv_np;
      const double vx = cov_norm * ((double)s[2] * inv_np - ux * ux);
      const double vy = cov_norm * ((double)s[3] * inv_np - uy * uy);
      const double vxy = cov_norm * ((double)s[4] * inv_np - ux * uy);
      const double a1 = 2.0 * ux * uy + C1, a2 = 2.0 * vxy + C2;
      const double b1 = ux * ux + uy * uy + C1, b2 = vx + vy + C2;
      acc += (a1 * a2) / (b1 * b2);
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(p.sum + n, t);
  }
}

// skimage.color.rgb2lab for one 8-bit sRGB pixel (D65, 2 degree observer)
__device__ __forceinline__ void rgb_to_lab(float r, float g, float b, float &L, float &A, float &Bv) {
  auto lin = [](float v) {
    v *= (1.f / 255.f);
    return v > 0.04045f ? powf((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
  };
  const float R = lin(r), G = lin(g), Bl = lin(b);
  float X = 0.412453f * R + 0.357580f * G + 0.180423f * Bl;
  float Y = 0.212671f * R + 0.715160f * G + 0.072169f * Bl;
  float Z = 0.019334f * R + 0.119193f * G + 0.950227f * Bl;
  X /= 0.95047f;
  Z /= 1.08883f;
  auto f = [](float t) { return t > 0.008856f ? cbrtf(t) : 7.787f * t + 16.f / 116.f; };
  const float fx = f(X), fy = f(Y), fz = f(Z);
  L = 116.f * fy - 16.f;
  A = 500.f * (fx - fy);
  Bv = 200.f * (fy - fz);
}

__global__ void __launch_bounds__(256) delta_e_u8_kernel(const uint8_t *__restrict__ a,
                                                         const uint8_t *__restrict__ b,
                                                         size_t pixels, double *__restrict__ sum) {
  __shared__ double s_part[8];
  const int n = blockIdx.y;
  const uint8_t *A = a + (size_t)n * pixels * 3, *B = b + (size_t)n * pixels * 3;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
       i += (size_t)gridDim.x * blockDim.x) {
    float L1, a1, b1, L2, a2, b2;
    rgb_to_lab((float)A[3 * i], (float)A[3 * i + 1], (float)A[3 * i + 2], L1, a1, b1);
    rgb_to_lab((float)B[3 * i], (float)B[3 * i + 1], (float)B[3 * i + 2], L2, a2, b2);
    const float dL = L1 - L2, da = a1 - a2, db = b1 - b2;
    acc += (double)sqrtf(dL * dL + da * da + db * db);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(sum + n, t);
  }
}

// ------------------------------------------------------------------------------------------
// MS-SSIM (pytorch_msssim.ms_ssim as compute_ms_ssim calls it, src/test_cae.py:46-50): five
// scales of Gaussian-window (11 taps, sigma 1.5, valid convolution) SSIM / contrast-structure
// means on fp32 planes, 2x2 average pooling (zero padding on odd sizes, pad counted) in between.
constexpr int kMsWin = 11;
constexpr int kMsIn = kSsimTile + kMsWin - 1;    // 42

struct MsParams {
  const float *a, *b;   // [planes][H][W]
  int planes, H, W;
  float win[kMsWin];
  float c1, c2;
  double *sum_ssim, *sum_cs;   // [planes]
};

__global__ void __launch_bounds__(256) msssim_level_kernel(const MsParams p) {
  __shared__ float tile_a[kMsIn][kMsIn + 1], tile_b[kMsIn][kMsIn + 1];
  __shared__ float h[5][kMsIn][kSsimTile + 1];
  __shared__ double s_part[2][8];
  const int pl = blockIdx.z;
  const int oy0 = blockIdx.y * kSsimTile, ox0 = blockIdx.x * kSsimTile;
  const int OH = p.H - kMsWin + 1, OW = p.W - kMsWin + 1;
  const float *A = p.a + (size_t)pl * p.H * p.W, *B = p.b + (size_t)pl * p.H * p.W;
  for (int i = threadIdx.x; i < kMsIn * kMsIn; i += blockDim.x) {
    const int r = i / kMsIn, q = i - r * kMsIn;
    const int y = oy0 + r, x = ox0 + q;
    const bool in = y < p.H && x < p.W;
    tile_a[r][q] = in ? __ldg(A + (size_t)y * p.W + x) : 0.f;
    tile_b[r][q] = in ? __ldg(B + (size_t)y * p.W + x) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMsIn * kSsimTile; i += blockDim.x) {
    const int r = i / kSsimTile, q = i - r * kSsimTile;
    float sa = 0.f, sb = 0.f, saa = 0.f, sbb = 0.f, sab = 0.f;
#pragma unroll
    for (int k = 0; k < kMsWin; ++k) {
      const float w = p.win[k], va = tile_a[r][q + k], vb = tile_b[r][q + k];
      sa += w * va;
      sb += w * vb;
      saa += w * va * va;
      sbb += w * vb * vb;
      sab += w * va * vb;
    }
    h[0][r][q] = sa;
    h[1][r][q] = sb;
    h[2][r][q] = saa;
    h[3][r][q] = sbb;
    h[4][r][q] = sab;
  }
  __syncthreads();
  double acc_s = 0.0, acc_c = 0.0;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += blockDim.x) {
    const int r = i / kSsimTile, q = i - r * kSsimTile;
    if (oy0 + r >= OH || ox0 + q >= OW) continue;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kMsWin; ++k)
#pragma unroll
      for (int m = 0; m < 5; ++m) s[m] += p.win[k] * h[m][r + k][q];
    const float mu1 = s[0], mu2 = s[1];
    const float s1 = s[2] - mu1 * mu1, s2 = s[3] - mu2 * mu2, s12 = s[4] - mu1 * mu2;
    const float cs = (2.f * s12 + p.c2) / (s1 + s2 + p.c2);
    const float ss = ((2.f * mu1 * mu2 + p.c1) / (mu1 * mu1 + mu2 * mu2 + p.c1)) * cs;
    acc_s += (double)ss;
    acc_c += (double)cs;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_part[0][threadIdx.x >> 5] = acc_s;
    s_part[1][threadIdx.x >> 5] = acc_c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t0 += s_part[0][w];
      t1 += s_part[1][w];
    }
    atomicAdd(p.sum_ssim + pl, t0);
    atomicAdd(p.sum_cs + pl, t1);
  }
}

// N x H x W x C uint8 -> [N * C][H][W] fp32
__global__ void __launch_bounds__(256) u8_to_planes_kernel(const uint8_t *__restrict__ src, int n,
                                                           int h, int w, int c, float *__restrict__ dst) {
  const size_t total = (size_t)n * h * w * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const size_t px = i / c;
    const size_t img = px / ((size_t)h * w), rem = px - img * (size_t)h * w;
    dst[(img * c + ch) * (size_t)h * w + rem] = (float)src[i];
  }
}

// F.avg_pool2d(x, 2, padding=(H % 2, W % 2)) with the padding counted: out[i][j] = mean over the
// 2 x 2 window of the zero-padded plane
__global__ void __launch_bounds__(256) avgpool2_kernel(const float *__restrict__ src, int planes,
                                                       int H, int W, float *__restrict__ dst, int OH, int OW) {
  const int py = H & 1, px = W & 1;
  const size_t total = (size_t)planes * OH * OW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const size_t r = i / OW;
    const int y = (int)(r % OH);
    const size_t pl = r / OH;
    const float *s = src + pl * (size_t)H * W;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * y + dy - py, xx = 2 * x + dx - px;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += __ldg(s + (size_t)yy * W + xx);
      }
    dst[i] = acc * 0.25f;
  }
}

}  // namespace

extern "C" int cae_u8_to_planes_f32(const uint8_t *src, int n, int h, int w, int c, float *dst,
                                    void *stream) {
  CAE_CHECK(src && dst && n > 0 && h > 0 && w > 0 && c > 0, 2, "cae_u8_to_planes_f32: bad argument");
  const size_t total = (size_t)n * h * w * c;
  const int blocks = (int)((total + 255) / 256 < (size_t)(16 * cae_sm_count()) ? (total + 255) / 256 : 16 * cae_sm_count());
  u8_to_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, n, h, w, c, dst);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_avgpool2_planes_f32(const float *src, int planes, int h, int w, float *dst,
                                       void *stream) {
  CAE_CHECK(src && dst && planes > 0 && h > 0 && w > 0, 2, "cae_avgpool2_planes_f32: bad argument");
  const int OH = (h + 2 * (h & 1) - 2) / 2 + 1, OW = (w + 2 * (w & 1) - 2) / 2 + 1;
  const size_t total = (size_t)planes * OH * OW;
  const int blocks = (int)((total + 255) / 256 < (size_t)(16 * cae_sm_count()) ? (total + 255) / 256 : 16 * cae_sm_count());
  avgpool2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, planes, h, w, dst, OH, OW);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_ssim_gauss_planes_f32(const float *a, const float *b, int planes, int h, int w,
                                         float data_range, double *sum_ssim, double *sum_cs,
                                         void *stream) {
  CAE_CHECK(a && b && sum_ssim && sum_cs && planes > 0, 2, "cae_ssim_gauss_planes_f32: bad argument");
  CAE_CHECK(h >= kMsWin && w >= kMsWin, 2, "cae_ssim_gauss_planes_f32: planes smaller than the 11-tap window");
  CAE_CHECK(planes <= 65535, 2, "cae_ssim_gauss_planes_f32: more than 65535 planes per call");
  MsParams p;
  p.a = a; p.b = b; p.planes = planes; p.H = h; p.W = w;
  float sum = 0.f;
  for (int k = 0; k < kMsWin; ++k) {
    const float x = (float)(k - kMsWin / 2);
    p.win[k] = expf(-(x * x) / (2.f * 1.5f * 1.5f));
    sum += p.win[k];
  }
  for (int k = 0; k < kMsWin; ++k) p.win[k] /= sum;
  p.c1 = (0.01f * data_range) * (0.01f * data_range);
  p.c2 = (0.03f * data_range) * (0.03f * data_range);
  p.sum_ssim = sum_ssim;
  p.sum_cs = sum_cs;
  const int OH = h - kMsWin + 1, OW = w - kMsWin + 1;
  const dim3 grid((unsigned)((OW + kSsimTile - 1) / kSsimTile), (unsigned)((OH + kSsimTile - 1) / kSsimTile),
                  (unsigned)planes);
  CAE_CHECK(grid.y <= 65535, 2, "cae_ssim_gauss_planes_f32: plane too tall for one launch");
  msssim_level_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_ssim_u8(const uint8_t *a, const uint8_t *b, int n_images, int h, int w, int c,
                           double *sum, void *stream) {
  CAE_CHECK(a && b && sum && n_images > 0 && c > 0, 2, "cae_ssim_u8: bad argument");
  CAE_CHECK(h >= kSsimWin && w >= kSsimWin, 2, "cae_ssim_u8: images smaller than the 7x7 window");
  CAE_CHECK(n_images <= 65535, 2, "cae_ssim_u8: more than 65535 images per call");
  SsimParams p{a, b, n_images, h, w, c, sum};
  const int OH = h - kSsimWin + 1, OW = w - kSsimWin + 1;
  const dim3 grid((unsigned)((OW + kSsimTile - 1) / kSsimTile), (unsigned)((OH + kSsimTile - 1) / kSsimTile),
                  (unsigned)n_images);
  CAE_CHECK(grid.y <= 65535, 2, "cae_ssim_u8: image too tall for one launch");
  ssim_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cae_delta_e_u8(const uint8_t *a, const uint8_t *b, int n_images, int64_t pixels,
                              double *sum, void *stream) {
  CAE_CHECK(a && b && sum && n_images > 0 && pixels > 0, 2, "cae_delta_e_u8: bad argument");
  CAE_CHECK(n_images <= 65535, 2, "cae_delta_e_u8: more than 65535 images per call");
  int bx = (int)((pixels + 255) / 256);
  const int want = (8 * cae_sm_count() + n_images - 1) / n_images;
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  delta_e_u8_kernel<<<dim3((unsigned)bx, (unsigned)n_images), 256, 0, (cudaStream_t)stream>>>(
      a, b, (size_t)pixels, sum);
  cae_count_launch();
  CAE_CUDA(cudaGetLastError());
  return 0;
}
