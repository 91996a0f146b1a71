/* bra_gen.c -- deterministic synthetic workloads of SURVEY.md section 8(d) (C2 text-like, C3 uniform random,
 * C4a/C4b periodic), host memory. Measurement and test tooling only: built as tools/libbra_gen.so, separate from
 * the product library, so that the reference arm of bench.py never maps libbra_b200.so.
 * PRNG for all shapes: splitmix64. */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* C3: little-endian bytes of successive splitmix64(seed) outputs */
void bra_gen_random(uint8_t* out, uint64_t n, uint64_t seed)
{
    uint64_t s = seed, i = 0;
    for (; i + 8 <= n; i += 8)
    {
        const uint64_t v = splitmix64(&s);
        memcpy(out + i, &v, 8);
    }
    if (i < n)
    {
        const uint64_t v = splitmix64(&s);
        memcpy(out + i, &v, (size_t) (n - i));
    }
}

/* C2: vocab[splitmix64(seed) % nvocab] joined by single spaces, truncated to n */
void bra_gen_text(uint8_t* out, uint64_t n, uint64_t seed, const char* const* vocab, uint32_t nvocab)
{
    uint64_t s = seed, i = 0;
    if (nvocab == 0) return;
    uint32_t* wl = (uint32_t*) malloc(sizeof(uint32_t) * nvocab);
    if (wl == NULL) return;
    for (uint32_t k = 0; k < nvocab; ++k) wl[k] = (uint32_t) strlen(vocab[k]);
    int first = 1;
    while (i < n)
    {
        if (!first) out[i++] = ' ';
        first = 0;
        if (i >= n) break;
        const uint32_t k = (uint32_t) (splitmix64(&s) % nvocab);
        const uint64_t m = wl[k] < n - i ? wl[k] : n - i;
        memcpy(out + i, vocab[k], (size_t) m);
        i += m;
    }
    free(wl);
}

/* C4a / C4b: `pattern` repeated, truncated to n */
void bra_gen_periodic(uint8_t* out, uint64_t n, const uint8_t* pattern, uint32_t plen)
{
    if (plen == 0) return;
    for (uint64_t i = 0; i < n; i += plen) memcpy(out + i, pattern, (size_t) (plen < n - i ? plen : n - i));
}
