// Host side of the whole-slide tile loops: the per-tile work that the reference leaves to
// dask's threaded scheduler and zarr's chunk store (src/compress.py:101-128,
// src/decompress.py:72-96) -- cutting tiles out of the slide, writing / reading one file per
// chunk -- done by native threads on whole batches, so that a Python loop (and its GIL) is not
// what bounds a GPU that transforms several gigapixels per second.  Plain C ABI, no globals.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cae_b200.h"

void cae_set_error(const char *fmt, ...);

namespace {

template <typename F>
void parallel_for(int n, int threads, F fn) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    });
  for (auto &th : pool) th.join();
}

bool write_all(int fd, const uint8_t *p, size_t n) {
  while (n) {
    const ssize_t w = ::write(fd, p, n);
    if (w <= 0) return false;
    p += w;
    n -= (size_t)w;
  }
  return true;
}

// header + payload of one chunk file as one system call (short writes finished by write_all)
bool write_two(int fd, const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
  if (na == 0) return write_all(fd, b, nb);
  iovec v[2] = {{const_cast<uint8_t *>(a), na}, {const_cast<uint8_t *>(b), nb}};
  const ssize_t w = ::writev(fd, v, 2);
  if (w < 0) return false;
  size_t done = (size_t)w;
  if (done < na) {
    if (!write_all(fd, a + done, na - done)) return false;
    done = na;
  }
  return write_all(fd, b + (done - na), nb - (done - na));
}

}  // namespace

// dst[k] = the ps x ps x c tile (ty, tx) = tile_yx[2k], tile_yx[2k+1] of the row-major
// H x W x c uint8 image `src`; the part of an edge tile beyond the image is zero filled (what
// zarr hands the chunk codec).  `threads` native threads.
extern "C" int cae_tiles_gather_u8(const uint8_t *src, int64_t H, int64_t W, int c, int ps,
                                   const int32_t *tile_yx, int n, uint8_t *dst, int threads) {
  if (!src || !dst || !tile_yx || H <= 0 || W <= 0 || c <= 0 || ps <= 0 || n < 0) {
    cae_set_error("cae_tiles_gather_u8: bad argument");
    return 2;
  }
  const size_t row = (size_t)ps * c, tile = row * ps;
  // rows of all tiles are independent copies: split the batch by (tile, row block)
  const int blocks = 4;
  parallel_for(n * blocks, threads, [&](int job) {
    const int k = job / blocks, b = job % blocks;
    const int64_t y0 = (int64_t)tile_yx[2 * k] * ps, x0 = (int64_t)tile_yx[2 * k + 1] * ps;
    const int64_t w_in = x0 >= W ? 0 : (x0 + ps <= W ? ps : W - x0);
    uint8_t *d = dst + (size_t)k * tile;
    for (int r = b * ps / blocks; r < (b + 1) * ps / blocks; ++r) {
      uint8_t *dr = d + (size_t)r * row;
      if (y0 + r >= H || w_in == 0) {
        memset(dr, 0, row);
        continue;
      }
      memcpy(dr, src + ((size_t)(y0 + r) * W + x0) * c, (size_t)w_in * c);
      if (w_in < ps) memset(dr + (size_t)w_in * c, 0, (size_t)(ps - w_in) * c);
    }
  });
  return 0;
}

// Write n files.  File k gets hdr_len bytes from headers + k * hdr_len followed by
// payload[payload_off[k] .. payload_off[k + 1]).  paths: n NUL-terminated strings back to back.
// Each file is written as <path>.partial and renamed (readers never see half a chunk).
extern "C" int cae_files_write(const char *paths, int n, const uint8_t *headers, int hdr_len,
                               const uint8_t *payload, const int64_t *payload_off, int threads) {
  if (!paths || n < 0 || (hdr_len > 0 && !headers) || !payload || !payload_off) {
    cae_set_error("cae_files_write: bad argument");
    return 2;
  }
  std::vector<const char *> name(n);
  const char *p = paths;
  for (int k = 0; k < n; ++k) {
    name[k] = p;
    p += strlen(p) + 1;
  }
  std::atomic<int> failed(-1);
  parallel_for(n, threads, [&](int k) {
    const std::string tmp = std::string(name[k]) + ".partial";
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    bool ok = fd >= 0;
    if (ok)
      ok = write_two(fd, hdr_len > 0 ? headers + (size_t)k * hdr_len : nullptr, (size_t)hdr_len,
                     payload + payload_off[k], (size_t)(payload_off[k + 1] - payload_off[k]));
    if (fd >= 0) ::close(fd);
    if (ok) ok = ::rename(tmp.c_str(), name[k]) == 0;
    if (!ok) failed.store(k);
  });
  if (failed.load() >= 0) {
    cae_set_error("cae_files_write: could not write %s", name[failed.load()]);
    return 4;
  }
  return 0;
}

// Remove n files (missing ones are not an error): what zarr's `overwrite=True` does to the chunks
// of an existing array (compress.py:123-128 of the reference), by native threads.
extern "C" int cae_files_remove(const char *paths, int n, int threads) {
  if (!paths || n < 0) {
    cae_set_error("cae_files_remove: bad argument");
    return 2;
  }
  std::vector<const char *> name(n);
  const char *p = paths;
  for (int k = 0; k < n; ++k) {
    name[k] = p;
    p += strlen(p) + 1;
  }
  parallel_for(n, threads, [&](int k) { ::unlink(name[k]); });
  return 0;
}

// Sizes of n files (bytes; -1 if missing).
extern "C" int cae_files_stat(const char *paths, int n, int64_t *sizes, int threads) {
  if (!paths || n < 0 || !sizes) {
    cae_set_error("cae_files_stat: bad argument");
    return 2;
  }
  std::vector<const char *> name(n);
  const char *p = paths;
  for (int k = 0; k < n; ++k) {
    name[k] = p;
    p += strlen(p) + 1;
  }
  parallel_for(n, threads, [&](int k) {
    struct stat st;
    sizes[k] = ::stat(name[k], &st) == 0 ? (int64_t)st.st_size : -1;
  });
  return 0;
}

// Read n files: the first hdr_len bytes of file k go to headers + k * hdr_len, the rest to
// payload + payload_off[k] (payload_off[k + 1] - payload_off[k] = size - hdr_len, from
// cae_files_stat).
extern "C" int cae_files_read(const char *paths, int n, uint8_t *headers, int hdr_len,
                              uint8_t *payload, const int64_t *payload_off, int threads) {
  if (!paths || n < 0 || (hdr_len > 0 && !headers) || !payload || !payload_off) {
    cae_set_error("cae_files_read: bad argument");
    return 2;
  }
  std::vector<const char *> name(n);
  const char *p = paths;
  for (int k = 0; k < n; ++k) {
    name[k] = p;
    p += strlen(p) + 1;
  }
  std::atomic<int> failed(-1);
  parallel_for(n, threads, [&](int k) {
    const int fd = ::open(name[k], O_RDONLY);
    bool ok = fd >= 0;
    auto read_all = [&](uint8_t *d, size_t len) {
      while (ok && len) {
        const ssize_t r = ::read(fd, d, len);
        if (r <= 0) { ok = false; break; }
        d += r;
        len -= (size_t)r;
      }
    };
    if (ok && hdr_len > 0) read_all(headers + (size_t)k * hdr_len, (size_t)hdr_len);
    if (ok) read_all(payload + payload_off[k], (size_t)(payload_off[k + 1] - payload_off[k]));
    if (fd >= 0) ::close(fd);
    if (!ok) failed.store(k);
  });
  if (failed.load() >= 0) {
    cae_set_error("cae_files_read: could not read %s", name[failed.load()]);
    return 4;
  }
  return 0;
}
