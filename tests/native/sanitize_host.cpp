// ASan / UBSan driver for the host-side native code of the hot path (rans_host.cpp, host_io.cpp):
// built and run by tests/test_host_sanitizers.py with -fsanitize=address,undefined.  Exercises the
// table builder, encode -> decode round trips with escapes on both sides and ragged sizes,
// truncated streams, the tile gather with ragged edges and the threaded file writer / reader.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cae_b200.h"

void cae_set_error(const char *fmt, ...) {      // api.cu is CUDA code: a local error sink
  va_list ap;
  va_start(ap, fmt);
  va_end(ap);
}

#define REQUIRE(c)                                                        \
  do {                                                                    \
    if (!(c)) {                                                           \
      fprintf(stderr, "sanitize_host: %s failed (line %d)\n", #c, __LINE__); \
      return 1;                                                           \
    }                                                                     \
  } while (0)

static uint32_t rng_state = 12345u;
static uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

int main(int argc, char **argv) {
  const char *dir = argc > 1 ? argv[1] : "/tmp";
  // ---- tables: skewed PMFs incl. zeros (steals in both directions)
  const int C = 5, L = 24;                     // symbols per channel incl. the escape slot
  std::vector<int32_t> cdfs(C * (L + 2)), sizes(C), offsets(C);
  for (int c = 0; c < C; ++c) {
    std::vector<float> pmf(L);
    float sum = 0.f;
    for (int i = 0; i < L; ++i) {
      pmf[i] = (i % (c + 2) == 0) ? 1e-7f : (float)(1 + rnd() % 1000) / (1.f + (float)abs(i - L / 2));
      sum += pmf[i];
    }
    for (int i = 0; i < L; ++i) pmf[i] /= sum;
    std::vector<uint32_t> cdf(L + 1);
    REQUIRE(cae_pmf_to_quantized_cdf(pmf.data(), L, 16, cdf.data()) == 0);
    REQUIRE(cdf[0] == 0 && cdf[L] == 65536);
    for (int i = 0; i <= L; ++i) cdfs[c * (L + 2) + i] = (int32_t)cdf[i];
    sizes[c] = L + 1;
    offsets[c] = -(L / 2);
  }
  float bad[3] = {0.5f, -1.f, 0.5f};
  uint32_t out3[4];
  REQUIRE(cae_pmf_to_quantized_cdf(bad, 3, 16, out3) != 0);
  // ---- round trips, escapes on both sides, odd sizes
  for (int hw : {1, 7, 64, 1000}) {
    std::vector<int32_t> sym((size_t)C * hw), dec((size_t)C * hw);
    for (size_t i = 0; i < sym.size(); ++i) {
      const uint32_t r = rnd() % 100;
      sym[i] = r < 3 ? -(int32_t)(rnd() % 5000) - 20 : (r < 6 ? (int32_t)(rnd() % 70000) + 20 : (int32_t)(rnd() % 20) - 10);
    }
    std::vector<uint8_t> enc((size_t)C * hw * 16 + 64);
    size_t nbytes = 0;
    REQUIRE(cae_rans_encode(sym.data(), C, hw, cdfs.data(), L + 2, sizes.data(), offsets.data(),
                            enc.data(), enc.size(), &nbytes) == 0);
    REQUIRE(nbytes > 0 && nbytes <= enc.size());
    std::vector<uint8_t> exact(enc.begin(), enc.begin() + nbytes);       // no slack behind the stream
    REQUIRE(cae_rans_decode(exact.data(), nbytes, C, hw, cdfs.data(), L + 2, sizes.data(),
                            offsets.data(), dec.data()) == 0);
    REQUIRE(memcmp(sym.data(), dec.data(), sym.size() * 4) == 0);
    // truncated / too small: an error, never an out-of-bounds access
    if (nbytes > 8)
      (void)cae_rans_decode(exact.data(), nbytes / 2, C, hw, cdfs.data(), L + 2, sizes.data(),
                            offsets.data(), dec.data());
    size_t nb2 = 0;
    REQUIRE(cae_rans_encode(sym.data(), C, hw, cdfs.data(), L + 2, sizes.data(), offsets.data(),
                            enc.data(), 4, &nb2) != 0);
  }
  // ---- tile gather with ragged edges
  {
    const int64_t H = 70, W = 45;
    const int c = 3, ps = 32;
    std::vector<uint8_t> img((size_t)H * W * c);
    for (auto &v : img) v = (uint8_t)rnd();
    std::vector<int32_t> yx;
    for (int ty = 0; ty * ps < H; ++ty)            // tile indices (row, column)
      for (int tx = 0; tx * ps < W; ++tx) { yx.push_back(ty); yx.push_back(tx); }
    const int n = (int)yx.size() / 2;
    std::vector<uint8_t> tiles((size_t)n * ps * ps * c, 0xAB);
    REQUIRE(cae_tiles_gather_u8(img.data(), H, W, c, ps, yx.data(), n, tiles.data(), 3) == 0);
    for (int k = 0; k < n; ++k)
      for (int y = 0; y < ps; ++y)
        for (int x = 0; x < ps; ++x)
          for (int ch = 0; ch < c; ++ch) {
            const int64_t Y = (int64_t)yx[2 * k] * ps + y, X = (int64_t)yx[2 * k + 1] * ps + x;
            const uint8_t want = (Y < H && X < W) ? img[(size_t)(Y * W + X) * c + ch] : 0;
            REQUIRE(tiles[(((size_t)k * ps + y) * ps + x) * c + ch] == want);
          }
  }
  // ---- threaded file write / stat / read
  {
    const int n = 37, hdr = 16;
    std::string paths;
    std::vector<int64_t> off(n + 1, 0);
    for (int k = 0; k < n; ++k) {
      paths += std::string(dir) + "/chunk." + std::to_string(k);
      paths.push_back('\0');
      off[k + 1] = off[k] + (int64_t)(rnd() % 5000);       // includes empty payloads
    }
    std::vector<uint8_t> headers((size_t)n * hdr), payload((size_t)off[n] + 1);
    for (auto &v : headers) v = (uint8_t)rnd();
    for (auto &v : payload) v = (uint8_t)rnd();
    REQUIRE(cae_files_write(paths.c_str(), n, headers.data(), hdr, payload.data(), off.data(), 4) == 0);
    std::vector<int64_t> sz(n);
    REQUIRE(cae_files_stat(paths.c_str(), n, sz.data(), 4) == 0);
    for (int k = 0; k < n; ++k) REQUIRE(sz[k] == hdr + off[k + 1] - off[k]);
    std::vector<uint8_t> h2((size_t)n * hdr), p2((size_t)off[n] + 1);
    REQUIRE(cae_files_read(paths.c_str(), n, h2.data(), hdr, p2.data(), off.data(), 4) == 0);
    REQUIRE(memcmp(h2.data(), headers.data(), h2.size()) == 0);
    REQUIRE(memcmp(p2.data(), payload.data(), (size_t)off[n]) == 0);
    std::string missing = std::string(dir) + "/does.not.exist";
    missing.push_back('\0');
    int64_t s1 = 0;
    (void)cae_files_stat(missing.c_str(), 1, &s1, 1);
    REQUIRE(s1 == -1);
    // header-less files (raw chunks), then the removal of everything incl. a name that is not there
    REQUIRE(cae_files_write(paths.c_str(), n, nullptr, 0, payload.data(), off.data(), 3) == 0);
    REQUIRE(cae_files_stat(paths.c_str(), n, sz.data(), 2) == 0);
    for (int k = 0; k < n; ++k) REQUIRE(sz[k] == off[k + 1] - off[k]);
    std::string all = paths + missing;
    REQUIRE(cae_files_remove(all.c_str(), n + 1, 4) == 0);
    REQUIRE(cae_files_stat(paths.c_str(), n, sz.data(), 4) == 0);
    for (int k = 0; k < n; ++k) REQUIRE(sz[k] == -1);
    REQUIRE(cae_files_remove(nullptr, 1, 1) != 0);
  }
  printf("sanitize_host: ok\n");
  return 0;
}
