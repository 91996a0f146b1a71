// bra_hd.h -- scalar logic shared by host and device code (compiles with g++ and nvcc).
//
// Everything here is small, branchy, per-block work (a 256-symbol Huffman tree, CRC
// polynomial arithmetic) that the kernels run on one thread/warp per block and that the
// host-side C ABI needs as well. Keeping one definition lets the CPU test-suite exercise the
// exact code the GPU executes.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define BRA_HD __host__ __device__ __forceinline__
#else
#define BRA_HD static inline
#endif

#define BRA_ALPHABET 256
#define BRA_CRC_POLY 0x82F63B78u   // reflected Castagnoli polynomial (reference src/utils/lib_bra_crc32c.c:26)
#define BRA_HUF_MAXLEN_DEC 32      // longest code the decoder accepts (Huffman)

// ---------------------------------------------------------------------------------------
// GF(2) polynomial arithmetic modulo the CRC-32C polynomial, reflected bit order
// (bit 31 holds x^0). crc_shift(c, k) advances a CRC register over k zero bytes, i.e. it
// multiplies by x^(8k) mod P -- the "fold" used to combine partial CRCs of adjacent pieces
// (what reference lib_bra_crc32c.c:181-231 does with 32x32 bit-matrix squaring).
// ---------------------------------------------------------------------------------------
BRA_HD uint32_t bra_gf_mul(uint32_t a, uint32_t b)
{
    uint32_t acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
    for (int i = 0; i < 32; ++i)
    {
        acc ^= (0u - ((a >> (31 - i)) & 1u)) & b;
        b = (b >> 1) ^ ((0u - (b & 1u)) & BRA_CRC_POLY);
    }
    return acc;
}

// x^(8*2^i) mod P for i = 0..39, filled by bra_gf_init_pow() (host) and copied to the device.
struct bra_gf_pow_t
{
    uint32_t p2[40];
};

BRA_HD void bra_gf_init_pow(bra_gf_pow_t* t)
{
    uint32_t v = 0x00800000u;  // x^8
    for (int i = 0; i < 40; ++i)
    {
        t->p2[i] = v;
        v        = bra_gf_mul(v, v);
    }
}

// x^(8*nbytes) mod P
BRA_HD uint32_t bra_gf_xpow8(const bra_gf_pow_t* t, uint64_t nbytes)
{
    uint32_t r = 0x80000000u;  // x^0
    for (int i = 0; nbytes != 0 && i < 40; ++i, nbytes >>= 1)
        if (nbytes & 1u) r = bra_gf_mul(r, t->p2[i]);
    return r;
}

// crc(A || B) from crc(A), crc(B), |B|  -- same contract as reference bra_crc32c_combine
BRA_HD uint32_t bra_crc_combine(const bra_gf_pow_t* t, uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    if (len_b == 0) return crc_a;
    return bra_gf_mul(crc_a, bra_gf_xpow8(t, len_b)) ^ crc_b;
}

// ---------------------------------------------------------------------------------------
// Huffman code lengths: exact replay of the reference's sorted-list tree build
// (reference src/encoders/bra_huffman.c:90-118 insert rule, :132-186 build, :188-220 depths).
//
// The reference keeps a singly linked list ascending by frequency. insert(x): x becomes the
// head iff the list is empty or head.freq > x.freq; otherwise it is linked after the last
// node, starting from the head, whose successor has freq < x.freq. On a sorted array that is
//     pos = (a[0].freq > x.freq) ? 0 : max(1, #{j : a[j].freq < x.freq}).
// The array is kept in a window [head, head+m) of a 512-slot buffer; popping the two minima
// frees two slots in front, so the parent is inserted by shifting the (short) front part down.
// ---------------------------------------------------------------------------------------
struct bra_huf_build_ws_t
{
    uint32_t nfreq[512];   // node frequency (uint32 sums wrap like the reference)
    uint16_t parent[512];  // parent node id
    uint16_t list[768];    // sorted window of node ids
    uint16_t leaf_of[256]; // node id of each present symbol
};

BRA_HD uint32_t bra_huf_insert_pos(const bra_huf_build_ws_t* ws, uint32_t head, uint32_t m, uint32_t f)
{
    if (m == 0 || ws->nfreq[ws->list[head]] > f) return 0;
    // lower bound of f in the sorted window
    uint32_t lo = 0, hi = m;
    while (lo < hi)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (ws->nfreq[ws->list[head + mid]] < f)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo < 1 ? 1 : lo;
}

// Returns number of distinct symbols (0 => empty input, caller reports failure like the reference).
BRA_HD uint32_t bra_huf_build_lengths(const uint32_t* freq, uint8_t* lengths, bra_huf_build_ws_t* ws)
{
    uint32_t nn = 0, m = 0;
    uint32_t head = 256;  // window grows to the right during leaf insertion, moves right during merges
    for (int s = 0; s < BRA_ALPHABET; ++s) lengths[s] = 0;
    for (int s = 0; s < BRA_ALPHABET; ++s)
    {
        const uint32_t f = freq[s];
        if (f == 0) continue;
        const uint32_t id  = nn++;
        ws->nfreq[id]      = f;
        ws->leaf_of[id]    = (uint16_t) s;  // leaves get ids 0..k-1 in symbol order
        const uint32_t pos = bra_huf_insert_pos(ws, head, m, f);
        for (uint32_t j = m; j > pos; --j) ws->list[head + j] = ws->list[head + j - 1];
        ws->list[head + pos] = (uint16_t) id;
        ++m;
    }
    const uint32_t nleaves = nn;
    if (nleaves == 0) return 0;
    if (nleaves == 1)
    {
        lengths[ws->leaf_of[0]] = 1;  // lone leaf: length 1, code 0 (bra_huffman.c:201-207)
        return 1;
    }
    while (m > 1)
    {
        const uint32_t l = ws->list[head], r = ws->list[head + 1];
        head += 2;
        m -= 2;
        const uint32_t id = nn++;
        const uint32_t f  = ws->nfreq[l] + ws->nfreq[r];
        ws->nfreq[id]     = f;
        ws->parent[l]     = (uint16_t) id;
        ws->parent[r]     = (uint16_t) id;
        const uint32_t pos = bra_huf_insert_pos(ws, head, m, f);
        // shift the front part [0,pos) one slot down into the space freed by the pops
        for (uint32_t j = 0; j < pos; ++j) ws->list[head - 1 + j] = ws->list[head + j];
        head -= 1;
        ws->list[head + pos] = (uint16_t) id;
        ++m;
    }
    const uint32_t root = nn - 1;
    for (uint32_t i = 0; i < nleaves; ++i)
    {
        uint32_t d = 0, v = i;
        while (v != root)
        {
            v = ws->parent[v];
            ++d;
        }
        lengths[ws->leaf_of[i]] = (uint8_t) d;
    }
    return nleaves;
}

// Canonical code assignment, reference bra_huffman.c:227-261: a uint32 running code shifted once
// per length 1..256 (so it wraps exactly like the reference beyond 32 bits); symbols ascending
// within a length.
BRA_HD void bra_huf_canonical(const uint8_t* lengths, uint32_t* codes)
{
    uint32_t count[257];
    for (int i = 0; i <= 256; ++i) count[i] = 0;
    for (int i = 0; i < 256; ++i)
        if (lengths[i]) ++count[lengths[i]];
    uint32_t code = 0;
    for (int len = 1; len <= 256; ++len)
    {
        code <<= 1;
        const uint32_t c = count[len];
        count[len]       = code;
        code += c;
    }
    for (int i = 0; i < 256; ++i) codes[i] = lengths[i] ? count[lengths[i]]++ : 0u;
}

// ---------------------------------------------------------------------------------------
// Canonical decode tables (valid prefix codes with lengths <= BRA_HUF_MAXLEN_DEC).
// For a left-aligned 32-bit window w, the code length is the smallest L with
// w < limit[L]  (limit is 33-bit, hence uint64), and the symbol is
// sorted_sym[base[L] + (w >> (32-L)) - first[L]].
// Returns false when the lengths are not a usable prefix code (oversubscribed Kraft sum or a
// length above the supported maximum); the reference fails such headers at tree rebuild
// (bra_huffman.c:294-303) or later at the bit walk.
// ---------------------------------------------------------------------------------------
struct bra_huf_dec_t
{
    uint64_t limit[BRA_HUF_MAXLEN_DEC + 2]; // limit[L], L = 1..32; limit[0] = 0
    uint32_t first[BRA_HUF_MAXLEN_DEC + 2]; // first canonical code of length L
    uint16_t base[BRA_HUF_MAXLEN_DEC + 2];  // index of first symbol of length L in sorted_sym
    uint16_t count[BRA_HUF_MAXLEN_DEC + 2];
    uint8_t  sorted_sym[256];
    uint32_t min_len, max_len, nsym;
};

BRA_HD bool bra_huf_make_dec(const uint8_t* lengths, bra_huf_dec_t* d)
{
    for (int i = 0; i < BRA_HUF_MAXLEN_DEC + 2; ++i)
    {
        d->count[i] = 0;
        d->limit[i] = 0;
        d->first[i] = 0;
        d->base[i]  = 0;
    }
    d->min_len = 0;
    d->max_len = 0;
    d->nsym    = 0;
    for (int s = 0; s < 256; ++s)
    {
        const uint32_t l = lengths[s];
        if (l == 0) continue;
        if (l > BRA_HUF_MAXLEN_DEC) return false;
        d->count[l]++;
        d->nsym++;
        if (d->min_len == 0 || l < d->min_len) d->min_len = l;
        if (l > d->max_len) d->max_len = l;
    }
    if (d->nsym == 0) return true;  // empty tree: every bit is an invalid code (max_len == 0)
    uint64_t code = 0;
    uint32_t idx  = 0;
    for (uint32_t l = 1; l <= BRA_HUF_MAXLEN_DEC; ++l)
    {
        code <<= 1;
        d->first[l] = (uint32_t) code;
        d->base[l]  = (uint16_t) idx;
        code += d->count[l];
        idx += d->count[l];
        if (code > (1ull << l)) return false;  // Kraft sum above 1: codes would collide
        d->limit[l] = code << (32 - l);
    }
    uint16_t next[BRA_HUF_MAXLEN_DEC + 2];
    for (int i = 0; i < BRA_HUF_MAXLEN_DEC + 2; ++i) next[i] = d->base[i];
    for (int s = 0; s < 256; ++s)
        if (lengths[s]) d->sorted_sym[next[lengths[s]]++] = (uint8_t) s;
    return true;
}

// Decode one symbol from the left-aligned 32-bit window. Returns the code length, 0 if no
// codeword matches (the reference's "invalid code sequence", bra_huffman.c:466-470).
BRA_HD uint32_t bra_huf_decode_one(const bra_huf_dec_t* d, uint32_t w, uint8_t* sym)
{
    for (uint32_t l = d->min_len; l <= d->max_len; ++l)
    {
        if ((uint64_t) w < d->limit[l])
        {
            // w >= limit[l-1] here, so the prefix is a real code of length l iff count[l] > 0
            const uint32_t c = w >> (32 - l);
            if (c < d->first[l]) return 0;
            *sym = d->sorted_sym[d->base[l] + (c - d->first[l])];
            return l;
        }
    }
    return 0;
}

// ---- move-to-front list of one thread: 256 entries as sixteen 128-bit chunks ------------------------------------------
// Q[0..15]: entry k is byte k of the 256-byte array (little endian inside each 32-bit word). Moving an entry to
// the front shifts everything before it up by one byte: one 128-bit load, four byte-permutes and one 128-bit
// store per sixteen entries (the instruction count is what bounds the replay kernel on high ranks, not HBM).
#if !defined(__CUDACC__)
struct alignas(16) uint4  // host stand-in for the CUDA vector type
{
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#endif

// (cur << 8) | (prev >> 24): the image of one 32-bit word after the shift
BRA_HD uint32_t bra_mtf_shift_word(uint32_t prev, uint32_t cur)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(prev, cur, 0x6543);
#else
    return (cur << 8) | (prev >> 24);
#endif
}
BRA_HD uint32_t bra_ffs32(uint32_t z)  // 1-based index of the lowest set bit, 0 for 0
{
#ifdef __CUDA_ARCH__
    return (uint32_t) __ffs((int) z);
#else
    return (uint32_t) __builtin_ffs((int) z);
#endif
}
BRA_HD uint4 bra_mtf_shift_chunk(uint32_t prev, const uint4 v)
{
    return make_uint4(bra_mtf_shift_word(prev, v.x), bra_mtf_shift_word(v.x, v.y), bra_mtf_shift_word(v.y, v.z), bra_mtf_shift_word(v.z, v.w));
}
// the chunk that holds the moved entry at byte o: bytes <= o take the shifted image, the others stay
BRA_HD uint4 bra_mtf_merge_chunk(const uint4 v, const uint4 n, uint32_t o)
{
    const uint32_t wo = o >> 2;
    const uint32_t pm = 0xFFFFFFFFu >> ((3u - (o & 3u)) * 8u);  // bytes 0..(o&3) of the word that holds the entry
    const uint32_t mx = wo > 0 ? 0xFFFFFFFFu : pm;
    const uint32_t my = wo > 1 ? 0xFFFFFFFFu : (wo == 1 ? pm : 0u);
    const uint32_t mz = wo > 2 ? 0xFFFFFFFFu : (wo == 2 ? pm : 0u);
    const uint32_t mw = wo == 3 ? pm : 0u;
    return make_uint4((n.x & mx) | (v.x & ~mx), (n.y & my) | (v.y & ~my), (n.z & mz) | (v.z & ~mz), (n.w & mw) | (v.w & ~mw));
}
// decode one rank: returns the symbol at position r and moves it to the front
BRA_HD uint32_t bra_mtf_list_decode(uint4* Q, uint32_t r)
{
    const uint32_t sym = reinterpret_cast<const uint8_t*>(Q)[r];
    if (r == 0) return sym;
    const uint32_t nq   = r >> 4;
    uint32_t       prev = sym << 24;  // byte entering the next word from below
    for (uint32_t q = 0; q < nq; ++q)
    {
        const uint4 v = Q[q];
        Q[q]          = bra_mtf_shift_chunk(prev, v);
        prev          = v.w;
    }
    const uint4 v = Q[nq];
    Q[nq]         = bra_mtf_merge_chunk(v, bra_mtf_shift_chunk(prev, v), r & 15u);
    return sym;
}
// first zero byte of t flagged in bit 7 of that byte (higher flags may be spurious, the lowest one never is)
BRA_HD uint32_t bra_mtf_zero_bytes(uint32_t t) { return (t - 0x01010101u) & ~t & 0x80808080u; }
// encode one symbol: returns its position and moves it to the front (single forward pass)
BRA_HD uint32_t bra_mtf_list_encode(uint4* Q, uint32_t x)
{
    const uint32_t x4   = x * 0x01010101u;
    uint32_t       prev = x << 24;
    for (uint32_t q = 0;; ++q)
    {
        const uint4    v  = Q[q];
        const uint32_t z0 = bra_mtf_zero_bytes(v.x ^ x4), z1 = bra_mtf_zero_bytes(v.y ^ x4), z2 = bra_mtf_zero_bytes(v.z ^ x4),
                       z3 = bra_mtf_zero_bytes(v.w ^ x4);
        const uint4    n  = bra_mtf_shift_chunk(prev, v);
        if ((z0 | z1 | z2 | z3) == 0)
        {
            Q[q] = n;
            prev = v.w;
            continue;
        }
        const uint32_t wo = z0 ? 0u : (z1 ? 1u : (z2 ? 2u : 3u));
        const uint32_t z  = z0 ? z0 : (z1 ? z1 : (z2 ? z2 : z3));
        const uint32_t o  = wo * 4 + ((bra_ffs32(z) - 1u) >> 3);
        if (q == 0 && o == 0) return 0;  // already in front
        Q[q] = bra_mtf_merge_chunk(v, n, o);
        return q * 16 + o;
    }
}

// ---- BWT finisher: compare two rotations of T (period p) over the `depth` bytes that start `from` bytes in ----------
// (`from` already reduced mod p; depth a multiple of 4; T 4-byte aligned). Four bytes per step through aligned word
// loads while neither side is about to wrap around the block end, bytes otherwise. EVERY pair is compared to exactly
// `depth` bytes: a deeper look at some pairs only would make "equal" non-transitive, and the finisher's counting sort
// would put two members of a group into the same slot.
BRA_HD uint32_t bra_funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t) ((((uint64_t) hi << 32) | lo) >> (s & 31u));
#endif
}
BRA_HD uint32_t bra_bswap32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
BRA_HD uint32_t bra_load_be32(const uint8_t* T, uint32_t i)  // bytes i..i+3, first byte most significant
{
    const uint32_t* Tw = reinterpret_cast<const uint32_t*>(T);
    const uint32_t  w0 = Tw[i >> 2], w1 = Tw[(i >> 2) + 1];
    return bra_bswap32(bra_funnel_r(w0, w1, (i & 3u) * 8));
}
BRA_HD int bra_rot_cmp_window(const uint8_t* T, uint32_t p, uint32_t a, uint32_t c, uint32_t from, uint32_t depth)
{
    uint32_t ia = a + from, ic = c + from;
    if (ia >= p) ia -= p;
    if (ic >= p) ic -= p;
    if (ia + depth + 4u <= p && ic + depth + 4u <= p)
    {
        // neither window wraps: stream aligned words, one new word per side and step
        const uint32_t* Ta = reinterpret_cast<const uint32_t*>(T) + (ia >> 2);
        const uint32_t* Tc = reinterpret_cast<const uint32_t*>(T) + (ic >> 2);
        const uint32_t  sa = (ia & 3u) * 8u, sc = (ic & 3u) * 8u;
        uint32_t        a0 = Ta[0], c0 = Tc[0];
#ifdef __CUDA_ARCH__
#pragma unroll 4
#endif
        for (uint32_t k = 1; k <= depth / 4u; ++k)
        {
            const uint32_t a1 = Ta[k], c1 = Tc[k];
            const uint32_t x = bra_funnel_r(a0, a1, sa), y = bra_funnel_r(c0, c1, sc);  // bytes in memory order, first byte lowest
            if (x != y) return bra_bswap32(x) < bra_bswap32(y) ? -1 : 1;
            a0 = a1;
            c0 = c1;
        }
        return 0;
    }
    uint32_t k = 0;
    while (k < depth)
    {
        // word steps only while they stay inside the window (after byte steps k is no multiple of four any more), and while
        // the second word of the unaligned read stays inside the block
        if (k + 4 <= depth && ia + 8 <= p && ic + 8 <= p)
        {
            const uint32_t x = bra_load_be32(T, ia), y = bra_load_be32(T, ic);
            if (x != y) return x < y ? -1 : 1;
            ia += 4;
            ic += 4;
            k += 4;
        }
        else
        {
            const uint8_t x = T[ia], y = T[ic];
            if (x != y) return x < y ? -1 : 1;
            if (++ia == p) ia = 0;
            if (++ic == p) ic = 0;
            ++k;
        }
    }
    return 0;
}

// ---- host-path pipeline: blocks per stage ------------------------------------------------------------------------
// Every stage costs a few milliseconds of latency-bound kernels whatever its size; the input copy of a stage hides
// behind the kernels of the stage before it and its output copy behind those of the stage after it, so only the
// first input copy and the last output copy are exposed. Hence a short head (a tenth of the job, long enough for
// its kernels to cover the next input copy), stages as wide as the context allows (`hb` blocks), and a tail that
// shrinks (a quarter, then a sixteenth) so that each output copy fits behind the kernels that follow. The same
// plan serves both directions and any compression ratio. Writes at most nblk / hb + 4 entries; returns the count.
static inline uint32_t bra_stage_plan(uint64_t nblk, uint32_t hb, uint32_t* plan)
{
    uint32_t n = 0;
    if (nblk == 0 || hb == 0) return 0;
    if (nblk < 128 && nblk <= hb)
    {
        if (nblk > 1) plan[n++] = (uint32_t) (nblk / 2);
        plan[n++] = (uint32_t) (nblk - nblk / 2);
        return n;
    }
    const uint64_t tail1 = nblk / 4 < hb ? nblk / 4 : hb, tail2 = nblk / 16 < hb ? nblk / 16 : hb;
    uint64_t       head  = nblk / 10 > 16 ? nblk / 10 : 16;
    if (head > hb) head = hb;
    if (head > nblk - tail1 - tail2) head = nblk - tail1 - tail2;
    if (head) plan[n++] = (uint32_t) head;
    for (uint64_t left = nblk - head - tail1 - tail2; left > 0;)
    {
        const uint64_t take = left < hb ? left : hb;
        plan[n++] = (uint32_t) take;
        left -= take;
    }
    if (tail1) plan[n++] = (uint32_t) tail1;
    if (tail2) plan[n++] = (uint32_t) tail2;
    return n;
}

// Encoding moves a block in and a fifth of it out (or, at worst, all of it): only the first input copy is exposed, the
// output copies are short. So a short head (an eighth of the job: its kernels run as long as the next input copy takes)
// and then stages as wide as the context allows -- every further stage would only add its few milliseconds of
// latency-bound kernels. Writes at most nblk / hb + 2 entries; returns the count.
static inline uint32_t bra_stage_plan_encode(uint64_t nblk, uint32_t hb, uint32_t head_div, uint32_t* plan)
{
    uint32_t n = 0;
    if (nblk == 0 || hb == 0) return 0;
    if (head_div < 2) head_div = 2;
    uint64_t head = nblk / head_div;
    if (head < 16) head = nblk < 32 ? nblk / 2 : 16;
    if (head > hb) head = hb;
    if (head) plan[n++] = (uint32_t) head;
    for (uint64_t left = nblk - head; left > 0;)
    {
        const uint64_t take = left < hb ? left : hb;
        plan[n++] = (uint32_t) take;
        left -= take;
    }
    return n;
}
