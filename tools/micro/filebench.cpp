// Where does a chunk-file write go?  Writes N files of SZ bytes into ONE tmpfs directory (the
// flat zarr v2 layout of a slide's chunks) with T threads, in the ways cae_files_write could:
//   0  <name>.partial: open(O_CREAT|O_TRUNC) + write(header) + write(payload) + rename
//   1  the directory operations alone (create + 16-byte write + rename)
//   2  overwrite in place, no truncate (pure copy into existing pages)
//   3  open(O_TRUNC) in place (page free + allocate, no directory operation)
//   4  like 0 with one writev
//   5  O_TMPFILE + writev + linkat (one directory operation; falls back to 4 when the name exists)
// build: g++ -O2 -pthread tools/micro/filebench.cpp -o /tmp/filebench; run: /tmp/filebench [threads] [n] [bytes]
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/uio.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

using clk = std::chrono::steady_clock;

template <typename F>
static void pfor(int n, int t, F fn) {
  std::atomic<int> nx(0);
  std::vector<std::thread> p;
  for (int i = 0; i < t; ++i)
    p.emplace_back([&] { for (int k = nx.fetch_add(1); k < n; k = nx.fetch_add(1)) fn(k); });
  for (auto &x : p) x.join();
}

int main(int argc, char **argv) {
  const int thr = argc > 1 ? atoi(argv[1]) : 16, n = argc > 2 ? atoi(argv[2]) : 4096;
  const int sz = argc > 3 ? atoi(argv[3]) : 136000;
  std::vector<char> buf((size_t)n * sz, 7);
  const std::string dir = "/dev/shm/cae_filebench";
  for (int mode = 0; mode <= 5; ++mode) {
    if (system(("rm -rf " + dir + " && mkdir -p " + dir).c_str())) return 1;
    std::vector<std::string> names;
    for (int i = 0; i < n; ++i)
      names.push_back(dir + "/" + std::to_string(i / 128) + "." + std::to_string(i % 128) + ".0");
    int dfd = open(dir.c_str(), O_RDONLY | O_DIRECTORY);
    for (int rep = 0; rep < 3; ++rep) {   // rep 0: fresh directory, rep 1-2: names exist
      std::atomic<int> fallbacks(0);
      const auto t0 = clk::now();
      pfor(n, thr, [&](int k) {
        const std::string tmp = names[k] + ".partial";
        iovec v[2] = {{buf.data(), 16}, {buf.data() + (size_t)k * sz, (size_t)sz}};
        if (mode == 0) {
          int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
          if (write(fd, v[0].iov_base, 16) < 0 || write(fd, v[1].iov_base, sz) < 0) abort();
          close(fd);
          rename(tmp.c_str(), names[k].c_str());
        } else if (mode == 1) {
          int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
          if (write(fd, v[0].iov_base, 16) < 0) abort();
          close(fd);
          rename(tmp.c_str(), names[k].c_str());
        } else if (mode == 2 || mode == 3) {
          int fd = open(names[k].c_str(), O_WRONLY | O_CREAT | (mode == 3 ? O_TRUNC : 0), 0644);
          if (writev(fd, v, 2) < 0) abort();
          close(fd);
        } else {
          int fd = mode == 5 ? open(dir.c_str(), O_TMPFILE | O_WRONLY, 0644) : -1;
          if (fd >= 0) {
            if (writev(fd, v, 2) < 0) abort();
            char proc[64];
            snprintf(proc, sizeof(proc), "/proc/self/fd/%d", fd);
            if (linkat(AT_FDCWD, proc, AT_FDCWD, names[k].c_str(), AT_SYMLINK_FOLLOW) == 0) {
              close(fd);
              return;
            }
            close(fd);
            fallbacks++;
          }
          fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
          if (writev(fd, v, 2) < 0) abort();
          close(fd);
          rename(tmp.c_str(), names[k].c_str());
        }
      });
      const double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
      printf("mode %d threads %d %s: %.1f ms for %d files (%.1f us/file, %.2f GB/s)%s\n", mode, thr,
             rep ? "overwrite" : "fresh    ", ms, n, ms * 1e3 / n, (double)n * sz / ms / 1e6,
             fallbacks ? "  [fell back to rename]" : "");
    }
    close(dfd);
  }
  return system(("rm -rf " + dir).c_str());
}
