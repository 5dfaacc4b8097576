/* CPU oracle: pmf_to_quantized_cdf and the rANS64 interface that CompressAI's
 * EntropyBottleneck.update/compress/decompress execute.  TEST INFRASTRUCTURE
 * ONLY (see oracle/__init__.py).
 *
 * The algorithm lives in the un-vendored dependency compressai>=1.2.4
 * (/root/reference/requirements.txt:26): compressai._CXX.pmf_to_quantized_cdf
 * and compressai.ans.RansEncoder/RansDecoder (ryg rans64, 16-bit precision,
 * 4-bit bypass escapes, 32-bit words).  It is restated from the published
 * algorithm; the reference reaches it through
 * /root/reference/src/models/tasks/_autoencoders.py:502,549-551,568-572.
 * PARITY UNPINNED: no upstream source, package or golden vector is available
 * in this container.
 *
 * Deliberately the "by the book" form: stage every (start, range, bypass)
 * item in a list, then flush the list back to front -- exactly the order of
 * operations Appendix A.3 states.  The product coder (cnn_autoencoder_b200/
 * csrc/rans_host.cpp) is a separate single-pass implementation; the tests
 * require the two to emit identical bytes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16u
#define BYPASS_BITS 4u
#define MAX_BYPASS 15
#define RANS_L (1ull << 31)

/* A.2.  Returns 0 on success, <0 on a domain error. cdf_out has n+1 entries. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    int i, j;
    for (i = 0; i < n; ++i)
        if (pmf[i] < 0 || !isfinite(pmf[i])) return -1;
    cdf[0] = 0;
    for (i = 0; i < n; ++i)
        cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    uint32_t total = 0;
    for (i = 0; i <= n; ++i) total += cdf[i];
    if (total == 0) return -2;
    for (i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
    for (i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;
    for (i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            uint32_t best_freq = ~0u;
            int best = -1;
            for (j = 0; j < n; ++j) {
                uint32_t f = cdf[j + 1] - cdf[j];
                if (f > 1 && f < best_freq) { best_freq = f; best = j; }
            }
            if (best < 0) return -3;
            if (best < i) { for (j = best + 1; j <= i; ++j) cdf[j]--; }
            else          { for (j = i + 1; j <= best; ++j) cdf[j]++; }
        }
    }
    return 0;
}

typedef struct { uint16_t start; uint16_t range; uint8_t bypass; } item_t;

typedef struct { item_t *v; size_t n, cap; } list_t;

static int push(list_t *l, uint32_t start, uint32_t range, int bypass)
{
    if (l->n == l->cap) {
        size_t nc = l->cap ? l->cap * 2 : 1024;
        item_t *nv = (item_t *)realloc(l->v, nc * sizeof(item_t));
        if (!nv) return -1;
        l->v = nv; l->cap = nc;
    }
    l->v[l->n].start = (uint16_t)start;
    l->v[l->n].range = (uint16_t)range;
    l->v[l->n].bypass = (uint8_t)bypass;
    l->n++;
    return 0;
}

/* A.3 encode.  cdfs is row-major [n_cdfs][cdf_stride] int32.  Returns the
 * number of bytes written to out (the used tail of the word buffer, host
 * little-endian), or <0 on error (-2: out too small). */
long oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, long n,
                        const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                        const int32_t *offsets, uint8_t *out, long out_cap)
{
    list_t l = {0, 0, 0};
    long i;
    for (i = 0; i < n; ++i) {
        const int32_t k = indexes[i];
        const int32_t *cdf = cdfs + (size_t)k * cdf_stride;
        const int32_t max_value = cdf_sizes[k] - 2;
        int32_t value = symbols[i] - offsets[k];
        uint32_t raw = 0;
        if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
        else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
        if (push(&l, (uint32_t)cdf[value], (uint32_t)(cdf[value + 1] - cdf[value]), 0)) goto oom;
        if (value == max_value) {
            int32_t nb = 0;
            while (nb < 8 && (raw >> (nb * BYPASS_BITS)) != 0) ++nb;
            int32_t val = nb;
            while (val >= MAX_BYPASS) {
                if (push(&l, MAX_BYPASS, MAX_BYPASS + 1, 1)) goto oom;
                val -= MAX_BYPASS;
            }
            if (push(&l, (uint32_t)val, (uint32_t)val + 1, 1)) goto oom;
            for (int32_t j = 0; j < nb; ++j) {
                uint32_t nib = (raw >> (j * BYPASS_BITS)) & MAX_BYPASS;
                if (push(&l, nib, nib + 1, 1)) goto oom;
            }
        }
    }
    {
        size_t nwords = l.n + 2;
        uint32_t *buf = (uint32_t *)malloc(nwords * sizeof(uint32_t));
        if (!buf) goto oom;
        uint32_t *ptr = buf + nwords;
        uint64_t x = RANS_L;
        while (l.n) {
            item_t s = l.v[--l.n];
            if (!s.bypass) {
                uint64_t freq = s.range;
                uint64_t x_max = ((RANS_L >> PRECISION) << 32) * freq;
                if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
                x = ((x / freq) << PRECISION) + (x % freq) + s.start;
            } else {
                uint64_t freq = 1u << (16 - BYPASS_BITS);
                uint64_t x_max = ((RANS_L >> 16) << 32) * freq;
                if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
                x = (x << BYPASS_BITS) | s.start;
            }
        }
        ptr -= 2;
        ptr[0] = (uint32_t)x;
        ptr[1] = (uint32_t)(x >> 32);
        long nbytes = (long)((buf + nwords) - ptr) * 4;
        if (nbytes > out_cap) { free(buf); free(l.v); return -2; }
        memcpy(out, ptr, (size_t)nbytes);
        free(buf);
        free(l.v);
        return nbytes;
    }
oom:
    free(l.v);
    return -1;
}

static uint32_t get_bits(uint64_t *r, const uint32_t **pp, uint32_t nbits)
{
    uint64_t x = *r;
    uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < RANS_L) { x = (x << 32) | **pp; *pp += 1; }
    *r = x;
    return val;
}

/* A.3 decode.  Returns 0 on success. */
int oracle_rans_decode(const uint8_t *enc, long nbytes, const int32_t *indexes, long n,
                       const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                       const int32_t *offsets, int32_t *out)
{
    if (nbytes < 8) return -1;
    uint32_t *words = (uint32_t *)malloc((size_t)nbytes + 8);
    if (!words) return -1;
    memcpy(words, enc, (size_t)nbytes);
    memset((uint8_t *)words + nbytes, 0, 8);
    const uint32_t *ptr = words;
    uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;
    for (long i = 0; i < n; ++i) {
        const int32_t k = indexes[i];
        const int32_t *cdf = cdfs + (size_t)k * cdf_stride;
        const int32_t size = cdf_sizes[k];
        const int32_t max_value = size - 2;
        const uint32_t cf = (uint32_t)(x & 0xffffu);
        int32_t j = 0;
        while (j < size && !((uint32_t)cdf[j] > cf)) ++j;
        const int32_t s = j - 1;
        const uint64_t start = (uint64_t)cdf[s], freq = (uint64_t)(cdf[s + 1] - cdf[s]);
        x = freq * (x >> PRECISION) + (x & 0xffffu) - start;
        if (x < RANS_L) { x = (x << 32) | *ptr++; }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)get_bits(&x, &ptr, BYPASS_BITS);
            int32_t nb = val;
            while (val == MAX_BYPASS) { val = (int32_t)get_bits(&x, &ptr, BYPASS_BITS); nb += val; }
            int32_t raw = 0;
            for (int32_t q = 0; q < nb; ++q) {
                val = (int32_t)get_bits(&x, &ptr, BYPASS_BITS);
                raw |= val << (q * BYPASS_BITS);
            }
            value = raw >> 1;
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + offsets[k];
    }
    free(words);
    return 0;
}
