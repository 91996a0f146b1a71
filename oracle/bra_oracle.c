/*
 * oracle/bra_oracle.c -- CPU restatement of BR-Archive's block-compression chain
 * (CRC32C, BWT, MTF, PackBits-style RLE, canonical Huffman).
 *
 * TEST INFRASTRUCTURE ONLY -- see bra_oracle.h for who may use it and how it is
 * pinned to the reference (golden vectors + oracle/_ref/libbra_ref.so).
 * Plain scalar C, written for clarity; it is the checker, not the product.
 */
#include "bra_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* CRC-32C (Castagnoli, reflected polynomial 0x82F63B78)                      */
/* reference src/utils/lib_bra_crc32c.c:26 (polynomial), :102-114 (byte loop) */
/* ------------------------------------------------------------------------- */
#define ORA_POLY 0x82F63B78u

static uint32_t g_crc_tab[256];
static int      g_crc_tab_ready = 0;

static void crc_tab_init(void)
{
    for (uint32_t b = 0; b < 256; ++b)
    {
        uint32_t r = b;
        for (int k = 0; k < 8; ++k)
            r = (r & 1u) ? (r >> 1) ^ ORA_POLY : (r >> 1);
        g_crc_tab[b] = r;
    }
    g_crc_tab_ready = 1;
}

uint32_t ora_crc32c(const void* data, uint64_t len, uint32_t prev)
{
    if (!g_crc_tab_ready)
        crc_tab_init();
    const uint8_t* p   = (const uint8_t*) data;
    uint32_t       crc = ~prev; /* init/xorout 0xFFFFFFFF are applied inside, like the reference */
    for (uint64_t i = 0; i < len; ++i)
        crc = g_crc_tab[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

/* a(x)*b(x) mod P(x), reflected bit order (bit 31 is x^0). */
static uint32_t gf_mulmod(uint32_t a, uint32_t b)
{
    uint32_t acc = 0;
    for (uint32_t m = 0x80000000u; m != 0; m >>= 1)
    {
        if (a & m)
            acc ^= b;
        b = (b & 1u) ? (b >> 1) ^ ORA_POLY : (b >> 1);
    }
    return acc;
}

/* reference :181-231 advances crc_a over len_b zero bytes with GF(2) matrix squaring;
 * that operator is multiplication by x^(8*len_b) mod P, computed here by square-and-multiply. */
uint32_t ora_crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint32_t len_b)
{
    if (len_b == 0)
        return crc_a;
    uint32_t pw  = 0x00800000u; /* x^8 : one zero byte */
    uint32_t mul = 0x80000000u; /* x^0 */
    for (uint32_t e = len_b; e != 0; e >>= 1)
    {
        if (e & 1u)
            mul = gf_mulmod(mul, pw);
        pw = gf_mulmod(pw, pw);
    }
    return gf_mulmod(crc_a, mul) ^ crc_b;
}

/* ------------------------------------------------------------------------- */
/* BWT                                                                        */
/* ------------------------------------------------------------------------- */

/* reference src/encoders/bra_bwt.c:31-53: compare rotations a and b byte by byte, n bytes, 0 if equal. */
static int rot_cmp(const uint8_t* t, uint32_t n, uint32_t a, uint32_t b)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        uint32_t pa = a + i, pb = b + i;
        if (pa >= n) pa -= n;
        if (pb >= n) pb -= n;
        if (t[pa] != t[pb])
            return t[pa] < t[pb] ? -1 : 1;
    }
    return 0;
}

/* Stable merge sort: equal rotations keep ascending start index, which is what
 * glibc's qsort_r (a stable merge sort) gives the reference at bra_bwt.c:91. */
static void rot_msort(const uint8_t* t, uint32_t n, uint32_t* a, uint32_t* tmp, uint32_t lo, uint32_t hi)
{
    if (hi - lo < 2)
        return;
    uint32_t mid = lo + (hi - lo) / 2;
    rot_msort(t, n, a, tmp, lo, mid);
    rot_msort(t, n, a, tmp, mid, hi);
    uint32_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi)
        tmp[k++] = (rot_cmp(t, n, a[j], a[i]) < 0) ? a[j++] : a[i++];
    while (i < mid) tmp[k++] = a[i++];
    while (j < hi) tmp[k++] = a[j++];
    memcpy(a + lo, tmp + lo, (size_t) (hi - lo) * sizeof(uint32_t));
}

/* reference src/encoders/bra_bwt.c:94-104: last column + position of rotation 0. */
static void bwt_emit(const uint8_t* in, uint32_t n, const uint32_t* sa, uint32_t* primary, uint8_t* out)
{
    *primary = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        out[i] = in[(sa[i] + n - 1) % n];
        if (sa[i] == 0)
            *primary = i;
    }
}

int ora_bwt_encode_naive(const uint8_t* in, uint32_t n, uint32_t* primary, uint8_t* out)
{
    if (n == 0)
        return -1;
    uint32_t* sa  = (uint32_t*) malloc((size_t) n * sizeof(uint32_t));
    uint32_t* tmp = (uint32_t*) malloc((size_t) n * sizeof(uint32_t));
    if (!sa || !tmp) { free(sa); free(tmp); return -1; }
    for (uint32_t i = 0; i < n; ++i)
        sa[i] = i;
    rot_msort(in, n, sa, tmp, 0, n);
    bwt_emit(in, n, sa, primary, out);
    free(sa);
    free(tmp);
    return 0;
}

/* Same order as above, computed by prefix doubling. Every round rebuilds the order from
 * the identity permutation with two stable counting sorts (second key, then first key),
 * so the result is ordered by (rank[i], rank[i+h], i) and fully equal rotations end in
 * ascending index exactly like the stable comparison sort. */
int ora_bwt_encode(const uint8_t* in, uint32_t n, uint32_t* primary, uint8_t* out)
{
    if (n == 0)
        return -1;
    const size_t m    = (size_t) n;
    const size_t nb   = (m > 256 ? m : 256) + 1;
    uint32_t*    rank = (uint32_t*) malloc(m * sizeof(uint32_t));
    uint32_t*    rk2  = (uint32_t*) malloc(m * sizeof(uint32_t));
    uint32_t*    sa   = (uint32_t*) malloc(m * sizeof(uint32_t));
    uint32_t*    sb   = (uint32_t*) malloc(m * sizeof(uint32_t));
    uint32_t*    cnt  = (uint32_t*) malloc(nb * sizeof(uint32_t));
    if (!rank || !rk2 || !sa || !sb || !cnt) { free(rank); free(rk2); free(sa); free(sb); free(cnt); return -1; }

    /* round 0: order by first byte, stable */
    memset(cnt, 0, nb * sizeof(uint32_t));
    for (size_t i = 0; i < m; ++i) cnt[in[i] + 1]++;
    for (size_t c = 1; c <= 256; ++c) cnt[c] += cnt[c - 1];
    for (size_t i = 0; i < m; ++i) sa[cnt[in[i]]++] = (uint32_t) i;
    uint32_t groups = 0;
    for (size_t j = 0; j < m; ++j)
    {
        if (j == 0 || in[sa[j]] != in[sa[j - 1]])
            groups++;
        rank[sa[j]] = groups - 1;
    }

    for (uint64_t h = 1; groups < n && h < n; h *= 2)
    {
        /* pass 1: identity order -> by rank[i+h] */
        memset(cnt, 0, nb * sizeof(uint32_t));
        for (size_t i = 0; i < m; ++i) cnt[rank[(i + h) % m] + 1]++;
        for (size_t c = 1; c <= groups; ++c) cnt[c] += cnt[c - 1];
        for (size_t i = 0; i < m; ++i) sb[cnt[rank[(i + h) % m]]++] = (uint32_t) i;
        /* pass 2: stable by rank[i] */
        memset(cnt, 0, nb * sizeof(uint32_t));
        for (size_t i = 0; i < m; ++i) cnt[rank[i] + 1]++;
        for (size_t c = 1; c <= groups; ++c) cnt[c] += cnt[c - 1];
        for (size_t j = 0; j < m; ++j) sa[cnt[rank[sb[j]]]++] = sb[j];
        /* re-rank */
        uint32_t g = 0;
        for (size_t j = 0; j < m; ++j)
        {
            if (j == 0 || rank[sa[j]] != rank[sa[j - 1]] || rank[(sa[j] + h) % m] != rank[(sa[j - 1] + h) % m])
                g++;
            rk2[sa[j]] = g - 1;
        }
        uint32_t* t = rank; rank = rk2; rk2 = t;
        groups = g;
    }
    if (groups < n)
    {
        /* periodic input: remaining ties are fully equal rotations; one more stable
         * sort from identity order puts them in ascending index. */
        memset(cnt, 0, nb * sizeof(uint32_t));
        for (size_t i = 0; i < m; ++i) cnt[rank[i] + 1]++;
        for (size_t c = 1; c <= groups; ++c) cnt[c] += cnt[c - 1];
        for (size_t i = 0; i < m; ++i) sa[cnt[rank[i]]++] = (uint32_t) i;
    }
    bwt_emit(in, n, sa, primary, out);
    free(rank); free(rk2); free(sa); free(sb); free(cnt);
    return 0;
}

/* reference src/encoders/bra_bwt.c:133-168 */
int ora_bwt_decode(const uint8_t* in, uint32_t n, uint32_t primary, uint8_t* out)
{
    if (n == 0 || primary >= n)
        return -1;
    uint32_t* transform = (uint32_t*) malloc((size_t) n * sizeof(uint32_t));
    if (!transform)
        return -1;
    uint32_t count[256] = {0};
    for (uint32_t i = 0; i < n; ++i)
        count[in[i]]++;
    uint32_t first[256];
    uint32_t run = 0;
    for (int c = 0; c < 256; ++c) { first[c] = run; run += count[c]; }
    for (uint32_t i = 0; i < n; ++i)
        transform[first[in[i]]++] = i;
    uint32_t idx = primary;
    for (uint32_t i = 0; i < n; ++i)
    {
        idx    = transform[idx];
        out[i] = in[idx];
    }
    free(transform);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* MTF  (reference src/encoders/bra_mtf.c:9-13 identity table per call)       */
/* ------------------------------------------------------------------------- */
void ora_mtf_encode(const uint8_t* in, size_t n, uint8_t* out)
{
    uint8_t tab[256];
    for (int i = 0; i < 256; ++i) tab[i] = (uint8_t) i;
    for (size_t i = 0; i < n; ++i)
    {
        const uint8_t s = in[i];
        unsigned      p = 0;
        while (tab[p] != s) ++p;
        memmove(tab + 1, tab, p);
        tab[0] = s;
        out[i] = (uint8_t) p;
    }
}

void ora_mtf_decode(const uint8_t* in, size_t n, uint8_t* out)
{
    uint8_t tab[256];
    for (int i = 0; i < 256; ++i) tab[i] = (uint8_t) i;
    for (size_t i = 0; i < n; ++i)
    {
        const unsigned p = in[i];
        const uint8_t  s = tab[p];
        memmove(tab + 1, tab, p);
        tab[0] = s;
        out[i] = s;
    }
}

/* ------------------------------------------------------------------------- */
/* RLE (PackBits variant; MIN_RUN 3, MAX_RUN 128, MAX_LITERAL 128)            */
/* reference src/lib_bra_defs.h:95-98, src/encoders/bra_rle.c                  */
/* ------------------------------------------------------------------------- */
#define ORA_RLE_MAX 128u
#define ORA_RLE_MIN 3u

/* reference bra_rle.c:9-18 */
static size_t rle_run_at(const uint8_t* in, size_t n, size_t i)
{
    size_t run = 1;
    while (i + run < n && in[i + run] == in[i] && run < ORA_RLE_MAX)
        ++run;
    return run;
}

/* One greedy token starting at i. Returns bytes consumed; *is_run tells which kind. */
static size_t rle_token_at(const uint8_t* in, size_t n, size_t i, int* is_run)
{
    const size_t run = rle_run_at(in, n, i);
    if (run >= ORA_RLE_MIN) { *is_run = 1; return run; }
    *is_run    = 0;
    size_t lit = run;
    size_t j   = i + run;
    while (j < n)
    {
        if (rle_run_at(in, n, j) >= ORA_RLE_MIN)
            break;
        ++j;
        if (++lit == ORA_RLE_MAX)
            break;
    }
    return lit;
}

size_t ora_rle_encode_size(const uint8_t* in, size_t n)
{
    size_t size = 0;
    for (size_t i = 0; i < n;)
    {
        int          is_run;
        const size_t len = rle_token_at(in, n, i, &is_run);
        size += is_run ? 2 : 1 + len;
        i += len;
    }
    return size;
}

int ora_rle_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n)
{
    *out_n = 0;
    const size_t s = ora_rle_encode_size(in, n);
    if (s == 0 || s > cap)
        return -1;
    uint8_t* p = out;
    for (size_t i = 0; i < n;)
    {
        int          is_run;
        const size_t len = rle_token_at(in, n, i, &is_run);
        if (is_run)
        {
            *p++ = (uint8_t) (int8_t) (-(int) (len - 1));
            *p++ = in[i];
        }
        else
        {
            *p++ = (uint8_t) (len - 1);
            memcpy(p, in + i, len);
            p += len;
        }
        i += len;
    }
    *out_n = s;
    return 0;
}

/* reference bra_rle.c:122-160 */
size_t ora_rle_decode_size(const uint8_t* in, size_t n)
{
    size_t size = 0;
    for (size_t i = 0; i < n;)
    {
        const int8_t c = (int8_t) in[i++];
        if (c >= 0)
        {
            const size_t cnt = (size_t) c + 1;
            if (i + cnt > n) return 0;
            size += cnt;
            i += cnt;
        }
        else if (c >= -127)
        {
            if (i >= n) return 0;
            size += (size_t) (1 - c);
            ++i;
        }
        /* -128: no-op */
    }
    return size;
}

/* reference bra_rle.c:162-224 */
int ora_rle_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n)
{
    *out_n = 0;
    const size_t s = ora_rle_decode_size(in, n);
    if (s == 0 || s > cap)
        return -1;
    uint8_t* p = out;
    for (size_t i = 0; i < n;)
    {
        const int8_t c = (int8_t) in[i++];
        if (c >= 0)
        {
            const size_t cnt = (size_t) c + 1;
            memcpy(p, in + i, cnt);
            i += cnt;
            p += cnt;
        }
        else if (c >= -127)
        {
            const size_t cnt = (size_t) (1 - c);
            memset(p, in[i++], cnt);
            p += cnt;
        }
    }
    *out_n = s;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Huffman                                                                    */
/* ------------------------------------------------------------------------- */

/* Code lengths. The reference keeps a singly linked list ascending by frequency
 * (bra_huffman.c:90-118): a new node becomes the head only if head.freq > x.freq;
 * otherwise it walks while next.freq < x.freq and is linked after that node. Here
 * the list is an array of node ids; the insertion slot follows from the same rule. */
int ora_huffman_lengths(const uint32_t freq[256], uint8_t lengths[256])
{
    uint32_t nfreq[512];
    int      left[512], right[512], sym[512];
    int      list[512];
    int      nlist = 0, nnodes = 0;

    memset(lengths, 0, 256);
    for (int s = 0; s < 256; ++s) /* leaves in symbol order, bra_huffman.c:140-153 */
    {
        if (freq[s] == 0)
            continue;
        const int id = nnodes++;
        nfreq[id] = freq[s]; left[id] = right[id] = -1; sym[id] = s;
        int pos;
        if (nlist == 0 || nfreq[list[0]] > nfreq[id])
            pos = 0;
        else
        {
            int cur = 0;
            while (cur + 1 < nlist && nfreq[list[cur + 1]] < nfreq[id]) ++cur;
            pos = cur + 1;
        }
        memmove(list + pos + 1, list + pos, (size_t) (nlist - pos) * sizeof(int));
        list[pos] = id;
        nlist++;
    }
    if (nlist == 0)
        return -1; /* bra_huffman.c:155-156 */

    while (nlist > 1) /* bra_huffman.c:158-175 */
    {
        const int l = list[0], r = list[1];
        memmove(list, list + 2, (size_t) (nlist - 2) * sizeof(int));
        nlist -= 2;
        const int id = nnodes++;
        nfreq[id] = nfreq[l] + nfreq[r]; /* uint32 sum, as in the reference */
        left[id] = l; right[id] = r; sym[id] = 0;
        int pos;
        if (nlist == 0 || nfreq[list[0]] > nfreq[id])
            pos = 0;
        else
        {
            int cur = 0;
            while (cur + 1 < nlist && nfreq[list[cur + 1]] < nfreq[id]) ++cur;
            pos = cur + 1;
        }
        memmove(list + pos + 1, list + pos, (size_t) (nlist - pos) * sizeof(int));
        list[pos] = id;
        nlist++;
    }

    /* depths, bra_huffman.c:188-220; a lone leaf gets length 1 (:201-207) */
    int stack_n[512], stack_d[512], sp = 0;
    stack_n[sp] = list[0]; stack_d[sp] = 0; sp++;
    while (sp > 0)
    {
        sp--;
        const int nd = stack_n[sp], d = stack_d[sp];
        if (left[nd] < 0 && right[nd] < 0)
        {
            lengths[sym[nd]] = (uint8_t) (d == 0 ? 1 : d);
            continue;
        }
        stack_n[sp] = left[nd];  stack_d[sp] = d + 1; sp++;
        stack_n[sp] = right[nd]; stack_d[sp] = d + 1; sp++;
    }
    return 0;
}

/* reference bra_huffman.c:227-261. The running code is a uint32_t that is shifted for every
 * length 1..256, so it wraps exactly like the reference for lengths above 32. */
void ora_huffman_canonical(const uint8_t lengths[256], uint32_t codes[256])
{
    uint32_t count[257];
    memset(count, 0, sizeof(count));
    for (int i = 0; i < 256; ++i)
        if (lengths[i] > 0) ++count[lengths[i]];
    uint32_t code = 0;
    count[0]      = 0;
    for (int len = 1; len <= 256; ++len)
    {
        code <<= 1;
        const uint32_t c = count[len];
        count[len] = code;
        code += c;
    }
    for (int i = 0; i < 256; ++i)
        codes[i] = lengths[i] ? count[lengths[i]]++ : 0;
}

/* bit j (0 = first emitted) of symbol's code: bra_huffman.c:254-258 stores c's low bits
 * right-aligned in a len-long bit array, so positions above bit 31 read as 0. */
static inline unsigned code_bit(uint32_t code, unsigned len, unsigned j)
{
    const unsigned shift = len - 1 - j;
    return shift < 32 ? (code >> shift) & 1u : 0u;
}

int ora_huffman_encode(const uint8_t* in, uint32_t n, uint8_t lengths[256], uint8_t* out, size_t cap, uint32_t* encoded_size)
{
    uint32_t freq[256] = {0};
    uint32_t codes[256];
    *encoded_size = 0;
    for (uint32_t i = 0; i < n; ++i)
        ++freq[in[i]];
    if (ora_huffman_lengths(freq, lengths) != 0)
        return -1;
    ora_huffman_canonical(lengths, codes);

    uint32_t bits = 0; /* uint32 like bra_huffman.c:390-392 */
    for (uint32_t i = 0; i < n; ++i)
        bits += lengths[in[i]];
    const uint32_t nbytes = (bits + 7) / 8;
    if (nbytes > cap)
        return -1;

    uint8_t* p   = out;
    uint8_t  cur = 0;
    int      pos = 0;
    for (uint32_t i = 0; i < n; ++i) /* MSB-first, bra_huffman.c:406-425 */
    {
        const uint8_t  s   = in[i];
        const unsigned len = lengths[s];
        for (unsigned j = 0; j < len; ++j)
        {
            if (code_bit(codes[s], len, j))
                cur |= (uint8_t) (1u << (7 - pos));
            if (++pos == 8) { *p++ = cur; cur = 0; pos = 0; }
        }
    }
    if (pos > 0)
        *p = cur;
    *encoded_size = nbytes;
    return 0;
}

/* Decoder: rebuild the pointer tree from the lengths in symbol order and walk it bit by
 * bit, restating bra_huffman.c:263-348 and :434-498 including their error exits. Writes
 * past orig_size (possible in the reference only on corrupt input) are counted, not stored. */
int ora_huffman_decode(const uint8_t lengths[256], const uint8_t* data, uint32_t encoded_size, uint32_t orig_size, uint8_t* out)
{
    enum { MAXN = 1 + 256 * 255 + 8 };
    uint32_t codes[256];
    ora_huffman_canonical(lengths, codes);

    int32_t* left  = (int32_t*) malloc(sizeof(int32_t) * MAXN);
    int32_t* right = (int32_t*) malloc(sizeof(int32_t) * MAXN);
    uint8_t* symb  = (uint8_t*) malloc(MAXN);
    if (!left || !right || !symb) { free(left); free(right); free(symb); return -1; }
    int nn = 0, rc = -1;
    left[0] = right[0] = -1; symb[0] = 0; nn = 1;

    for (int i = 0; i < 256; ++i)
    {
        const unsigned len = lengths[i];
        if (len == 0)
            continue;
        int cur = 0;
        for (unsigned j = 0; j < len; ++j)
        {
            const unsigned bit   = code_bit(codes[i], len, j);
            int32_t*       child = bit ? &right[cur] : &left[cur];
            if (j == len - 1)
            {
                if (*child != -1)
                    goto done; /* collision, :294-303 */
                left[nn] = right[nn] = -1; symb[nn] = (uint8_t) i;
                *child = nn++;
            }
            else
            {
                if (*child == -1)
                {
                    left[nn] = right[nn] = -1; symb[nn] = 0;
                    *child = nn++;
                }
                cur = *child;
            }
        }
    }

    {
        uint32_t idx = 0;
        int      cur = 0;
        for (uint32_t i = 0; i < encoded_size; ++i)
        {
            const uint8_t byte = data[i];
            for (int bit = 7; bit >= 0; --bit)
            {
                cur = ((byte >> bit) & 1) ? right[cur] : left[cur];
                if (cur == -1)
                    goto done; /* invalid code sequence, :466-470 */
                if (left[cur] == -1 && right[cur] == -1)
                {
                    if (idx < orig_size)
                        out[idx] = symb[cur];
                    idx++;
                    cur = 0;
                    if (idx >= orig_size)
                        break; /* leaves only the bit loop, :476-479 */
                }
            }
        }
        if (idx == orig_size)
            rc = 0; /* :485-489 */
    }
done:
    free(left); free(right); free(symb);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Whole-block chain                                                          */
/* ------------------------------------------------------------------------- */
static void put_u32(uint8_t* p, uint32_t v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); p[2] = (uint8_t) (v >> 16); p[3] = (uint8_t) (v >> 24); }
static uint32_t get_u32(const uint8_t* p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }

/* reference src/io/lib_bra_io_file_chunks.c:214-246 */
int ora_encode_block(const uint8_t* in, uint32_t n, uint8_t hdr[ORA_HDR_BYTES], uint8_t* payload, size_t cap, uint32_t* crc_raw, int naive_bwt)
{
    if (n == 0)
        return -1;
    const size_t rcap = (size_t) n + n / 128 + 2;
    uint8_t*     a    = (uint8_t*) malloc(n);
    uint8_t*     b    = (uint8_t*) malloc(n);
    uint8_t*     r    = (uint8_t*) malloc(rcap);
    int          rc   = -1;
    if (!a || !b || !r) goto done;
    uint32_t primary = 0;
    size_t   rn      = 0;
    uint32_t cn      = 0;
    if (crc_raw) *crc_raw = ora_crc32c(in, n, 0);
    if ((naive_bwt ? ora_bwt_encode_naive(in, n, &primary, a) : ora_bwt_encode(in, n, &primary, a)) != 0) goto done;
    ora_mtf_encode(a, n, b);
    if (ora_rle_encode(b, n, r, rcap, &rn) != 0) goto done;
    if (ora_huffman_encode(r, (uint32_t) rn, hdr + 4, payload, cap, &cn) != 0) goto done;
    put_u32(hdr, primary);
    put_u32(hdr + 260, (uint32_t) rn);
    put_u32(hdr + 264, cn);
    rc = 0;
done:
    free(a); free(b); free(r);
    return rc;
}

/* reference src/io/lib_bra_io_file_chunks.c:362-393 (without the BRA_MAX_CHUNK_SIZE header check) */
int ora_decode_block(const uint8_t hdr[ORA_HDR_BYTES], const uint8_t* payload, uint8_t* out, size_t cap, uint32_t* n_out)
{
    const uint32_t primary = get_u32(hdr);
    const uint32_t rn      = get_u32(hdr + 260);
    const uint32_t cn      = get_u32(hdr + 264);
    *n_out = 0;
    if (rn == 0 || cn == 0)
        return -1;
    uint8_t* r  = (uint8_t*) malloc(rn);
    uint8_t* a  = NULL;
    uint8_t* b  = NULL;
    int      rc = -1;
    if (!r) goto done;
    if (ora_huffman_decode(hdr + 4, payload, cn, rn, r) != 0) goto done;
    const size_t n = ora_rle_decode_size(r, rn);
    if (n == 0 || n > cap || primary >= n) goto done;
    a = (uint8_t*) malloc(n);
    b = (uint8_t*) malloc(n);
    if (!a || !b) goto done;
    size_t got = 0;
    if (ora_rle_decode(r, rn, a, n, &got) != 0) goto done;
    ora_mtf_decode(a, n, b);
    if (ora_bwt_decode(b, (uint32_t) n, primary, out) != 0) goto done;
    *n_out = (uint32_t) n;
    rc = 0;
done:
    free(r); free(a); free(b);
    return rc;
}
