// huffman.cu -- canonical Huffman coding for a batch of blocks
// (replaces reference bra_huffman_encode / bra_huffman_decode, src/encoders/bra_huffman.c:352-498).
//
// ENCODE
//   histogram  : 256 bins per block (produced by the RLE emit kernel on the fused path, or by
//                huf_hist_kernel for the stand-alone API), shared-memory privatised per CTA.
//   build      : one thread per block replays the reference's sorted-list tree build exactly
//                (bra_hd.h: bra_huf_build_lengths) -- code lengths depend on its tie rule, so no
//                "optimal lengths" algorithm may be substituted -- then assigns canonical codes.
//   bit lengths: per 4096-symbol tile, sum of code lengths; per-block exclusive scan gives every
//                tile its first output bit (and the block its encoded_size).
//   pack       : each thread owns 16 consecutive symbols, an exclusive prefix sum over code
//                lengths gives its bit offset; codewords are OR-ed MSB-first into a shared-memory
//                image of the tile's output words, which then leaves with plain coalesced word stores: a tile owns
//                every word it starts in (the bits of the preceding codes that share its first word are re-derived
//                by walking back over at most 32 symbols), so there is no payload memset and no global atomic.
// DECODE (the format has no sync markers: bra_huffman_t is lengths + two sizes, lib_bra_types.h:51-56)
//   Self-synchronising decode: the payload is cut into 1024-bit subsequences, one per thread,
//   256 per CTA. Every thread decodes from its current guess of where the first codeword of
//   its subsequence starts and publishes where it crossed into the next subsequence; guesses
//   are refined by fixed-point iteration (in shared memory inside the CTA, across CTAs by
//   re-launching while any CTA's entry changed). Subsequence 0 starts at bit 0, so the fixed
//   point is the true parse. Symbol counts are scanned and a second pass writes the output.
//   Codes that hardly ever re-synchronise (at most 12 bits, nearly uniform length: incompressible data) take the
//   phase mode instead: every subsequence is walked once from each of its max_len possible starts, and the maps
//   start -> (next start, codewords) are chained per sequence and per block (see HD_PHASE_MAX).
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"

namespace bra {

#define HF_TILE 4096
#define HF_THREADS 256

// ------------------------------------------------------------------------------------------------
// histogram (stand-alone path)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HF_THREADS)
    huf_hist_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t h[256];
    const uint32_t      b = blockIdx.y;
    const uint32_t      n = len[b];
    const uint32_t      tile0 = blockIdx.x * HF_TILE;
    if (tile0 >= n) return;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* p    = in + (uint64_t) b * stride;
    const uint32_t tend = min(n, tile0 + HF_TILE);
    for (uint32_t i = tile0 + threadIdx.x; i < tend; i += HF_THREADS) atomicAdd(&h[p[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[(uint64_t) b * 256 + threadIdx.x], h[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// tree build + canonical codes + header fields. One thread per block (tiny, branchy work).
// hdr layout (268 bytes, reference lib_bra_types.h:63-68): u32 primary | lengths[256] | u32 orig | u32 enc
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    huf_build_kernel(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ rlen, uint8_t* __restrict__ hdr, uint32_t* __restrict__ codes,
                     uint32_t* __restrict__ ok)
{
    __shared__ bra_huf_build_ws_t ws;
    __shared__ uint8_t            lengths[256];
    const uint32_t                b = blockIdx.x;
    if (threadIdx.x == 0)
    {
        const uint32_t k = bra_huf_build_lengths(hist + (uint64_t) b * 256, lengths, &ws);
        ok[b]            = k != 0 && rlen[b] != 0;  // empty input: reference returns NULL (bra_huffman.c:155-156)
        bra_huf_canonical(lengths, codes + (uint64_t) b * 256);
    }
    __syncwarp();
    uint8_t* h = hdr + (uint64_t) b * 268;
    for (int i = threadIdx.x; i < 256; i += 32) h[4 + i] = lengths[i];
    if (threadIdx.x == 0)
    {
        const uint32_t r = rlen[b];
        h[260] = (uint8_t) r;
        h[261] = (uint8_t) (r >> 8);
        h[262] = (uint8_t) (r >> 16);
        h[263] = (uint8_t) (r >> 24);
    }
}

// ------------------------------------------------------------------------------------------------
// bits per tile, then per-block scan
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HF_THREADS)
    huf_tile_bits_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ rlen, const uint8_t* __restrict__ hdr,
                         uint32_t tiles, uint32_t* __restrict__ t_bits)
{
    __shared__ uint8_t  slen[256];
    __shared__ uint32_t red[34];
    const uint32_t      b = blockIdx.y, t = blockIdx.x;
    const uint32_t      r = rlen[b];
    const uint32_t      tile0 = t * HF_TILE;
    if (tile0 >= r) return;
    slen[threadIdx.x] = hdr[(uint64_t) b * 268 + 4 + threadIdx.x];
    __syncthreads();
    const uint8_t* p    = in + (uint64_t) b * stride;
    const uint32_t tend = min(r, tile0 + HF_TILE);
    uint32_t       s    = 0;
    for (uint32_t i = tile0 + threadIdx.x; i < tend; i += HF_THREADS) s += slen[p[i]];
    uint32_t total;
    block_excl_add(s, red, &total);
    if (threadIdx.x == 0) t_bits[(uint64_t) b * tiles + t] = total;
}

// exclusive scan of t_bits per block (in place) + encoded_size into the header
__global__ void __launch_bounds__(256)
    huf_scan_bits_kernel(uint32_t* __restrict__ t_bits, const uint32_t* __restrict__ rlen, uint32_t tiles, uint8_t* __restrict__ hdr,
                         uint32_t* __restrict__ clen)
{
    __shared__ uint32_t red[34];
    const uint32_t      b = blockIdx.x;
    const uint32_t      r = rlen[b];
    const uint32_t      ntiles = (r + HF_TILE - 1) / HF_TILE;
    uint32_t*           tb     = t_bits + (uint64_t) b * tiles;
    uint32_t            carry  = 0;
    for (uint32_t base = 0; base < ntiles; base += 256)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? tb[i] : 0u;
        uint32_t       tot;
        const uint32_t ex = block_excl_add(v, red, &tot) + carry;
        if (i < ntiles) tb[i] = ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        const uint32_t c = (carry + 7u) / 8u;  // uint32 arithmetic like bra_huffman.c:390-395
        clen[b]          = c;
        uint8_t* h       = hdr + (uint64_t) b * 268;
        h[264] = (uint8_t) c;
        h[265] = (uint8_t) (c >> 8);
        h[266] = (uint8_t) (c >> 16);
        h[267] = (uint8_t) (c >> 24);
    }
}

// ------------------------------------------------------------------------------------------------
// pack. Output words are big-endian images of the byte stream (bit 31 = first bit), byte-swapped
// on the way out. Max 34 bits per code (uint32 code, zero-extended like the reference's bit array).
// ------------------------------------------------------------------------------------------------
#define HF_WORDS (HF_TILE * 34 / 32 + 8)

__device__ __forceinline__ void huf_put(uint32_t* img, uint32_t bitpos, uint64_t code, uint32_t len)
{
    // code occupies `len` (<= 40) bits, MSB first, starting at stream bit `bitpos`
    while (len)
    {
        const uint32_t w    = bitpos >> 5, o = bitpos & 31u;
        const uint32_t room = 32u - o;
        const uint32_t take = len < room ? len : room;
        const uint32_t bits = (uint32_t) ((code >> (len - take)) & ((take == 32u) ? 0xFFFFFFFFull : ((1ull << take) - 1ull)));
        atomicOr(&img[w], bits << (room - take));
        bitpos += take;
        len -= take;
    }
}

__global__ void __launch_bounds__(HF_THREADS)
    huf_pack_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ rlen, const uint8_t* __restrict__ hdr,
                    const uint32_t* __restrict__ codes, uint32_t tiles, const uint32_t* __restrict__ t_bitoff, uint8_t* __restrict__ out,
                    uint64_t out_stride)
{
    __shared__ uint8_t  slen[256];
    __shared__ uint32_t scode[256];
    __shared__ uint32_t img[HF_WORDS];
    __shared__ uint32_t red[34];
    const uint32_t      b = blockIdx.y, t = blockIdx.x;
    const uint32_t      r = rlen[b];
    const uint32_t      tile0 = t * HF_TILE;
    if (tile0 >= r) return;
    slen[threadIdx.x]  = hdr[(uint64_t) b * 268 + 4 + threadIdx.x];
    scode[threadIdx.x] = codes[(uint64_t) b * 256 + threadIdx.x];
    for (int i = threadIdx.x; i < HF_WORDS; i += HF_THREADS) img[i] = 0;
    __syncthreads();

    const uint8_t* p  = in + (uint64_t) b * stride;
    const uint32_t j0 = tile0 + threadIdx.x * 16;
    const uint32_t m  = j0 < r ? min(16u, r - j0) : 0u;
    uint8_t        x[16];
    uint32_t       mybits = 0;
    if (m == 16)
    {
        const uint4    v    = *reinterpret_cast<const uint4*>(p + j0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (uint32_t) i < m ? p[j0 + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if ((uint32_t) i < m) mybits += slen[x[i]];
    uint32_t       tile_bits;
    uint32_t       off   = block_excl_add(mybits, red, &tile_bits);
    const uint32_t gbit0 = t_bitoff[(uint64_t) b * tiles + t];  // first bit of the tile in the block's stream
    const uint32_t skew  = gbit0 & 31u;                         // image word 0 == global word gbit0/32
    off += skew;
    // The first word of the tile also holds the last `skew` bits of the codes before the tile: thread 0 walks back
    // over those symbols (at most 32 of them) and puts their bits too, so that this CTA owns the whole word and stores
    // it plainly. The word the tile ends in belongs to the next tile in the same way.
    if (threadIdx.x == 0 && skew)
    {
        uint32_t pos = skew;  // bits of the word still to be covered, counted from the word start
        for (uint32_t j = tile0; j > 0 && pos > 0;)
        {
            --j;
            uint32_t l    = slen[p[j]];
            uint64_t code = scode[p[j]];
            if (l == 0) continue;
            if (l > pos)  // the code starts in the previous word: keep its last `pos` bits
            {
                code &= (1ull << pos) - 1ull;
                l = pos;
            }
            pos -= l;
            huf_put(img, pos, code, l);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        if ((uint32_t) i >= m) break;
        const uint32_t l = slen[x[i]];
        huf_put(img, off, (uint64_t) scode[x[i]], l);
        off += l;
    }
    __syncthreads();

    const uint32_t nwords    = (skew + tile_bits + 31u) >> 5;
    const bool     owns_last = ((skew + tile_bits) & 31u) == 0 || tile0 + HF_TILE >= r;  // ends on a word boundary, or nothing follows
    uint32_t*      ow        = reinterpret_cast<uint32_t*>(out + (uint64_t) b * out_stride) + (gbit0 >> 5);
    for (uint32_t i = threadIdx.x; i < nwords; i += HF_THREADS)
        if (i + 1 < nwords || owns_last) ow[i] = __byte_perm(img[i], 0, 0x0123);  // big-endian image -> little-endian store
}

bool huf_encode_batch(const HufEncArgs& a, cudaStream_t st)
{
    if (a.nblk == 0 || a.max_r == 0) return true;
    const uint32_t tiles = bra_div_up(a.max_r, HF_TILE);
    const dim3     grid(tiles, a.nblk);
    if (a.compute_hist)
    {
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_hist, 0, (size_t) a.nblk * 256 * 4, st));
        BRA_LAUNCH(P_HUF_HIST, st, huf_hist_kernel<<<grid, HF_THREADS, 0, st>>>(a.d_in, a.stride, a.d_rlen, a.d_hist));
    }
    BRA_LAUNCH(P_HUF_BUILD, st, huf_build_kernel<<<a.nblk, 32, 0, st>>>(a.d_hist, a.d_rlen, a.d_hdr, a.d_codes, a.d_ok));
    BRA_LAUNCH(P_HUF_BITS, st, huf_tile_bits_kernel<<<grid, HF_THREADS, 0, st>>>(a.d_in, a.stride, a.d_rlen, a.d_hdr, tiles, a.d_t_bits));
    BRA_LAUNCH(P_HUF_BITS, st, huf_scan_bits_kernel<<<a.nblk, 256, 0, st>>>(a.d_t_bits, a.d_rlen, tiles, a.d_hdr, a.d_clen));
    BRA_LAUNCH(P_HUF_PACK, st, huf_pack_kernel<<<grid, HF_THREADS, 0, st>>>(a.d_in, a.stride, a.d_rlen, a.d_hdr, a.d_codes, tiles, a.d_t_bits, a.d_out, a.out_stride));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

// ================================================================================================
// DECODE
// ================================================================================================
#define HD_SUB_BITS 1024u
#define HD_THREADS 256
#define HD_SEQ_BITS (HD_SUB_BITS * HD_THREADS)  // 262144 bits = 32 KiB of payload per CTA
#define HD_SEQ_BYTES (HD_SEQ_BITS / 8)
#define HD_ROW_WORDS (HD_SUB_BITS / 32)         // 32 words per subsequence
#define HD_SMEM_WORDS (HD_SEQ_BYTES / 4 + HD_ROW_WORDS)  // + one row of look-ahead for codes crossing the sequence end
#define HD_LUT_BITS 11
#define HD_WARMUP_BITS 96u

// per-block decode tables, built once per block by one thread
__global__ void __launch_bounds__(32) huf_dec_tables_kernel(const uint8_t* __restrict__ hdr, bra_huf_dec_t* __restrict__ tabs, uint32_t* __restrict__ err)
{
    __shared__ uint8_t lengths[256];
    const uint32_t     b = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += 32) lengths[i] = hdr[(uint64_t) b * 268 + 4 + i];
    __syncwarp();
    if (threadIdx.x == 0)
    {
        if (!bra_huf_make_dec(lengths, &tabs[b])) err[b] = 1;
    }
}

struct HdShared
{
    uint32_t      words[HD_SMEM_WORDS];   // payload slice, big-endian words, swizzled (see hd_word)
    bra_huf_dec_t tab;
    uint16_t      lut[1 << HD_LUT_BITS];  // (sym << 8) | len for codes of at most HD_LUT_BITS bits, 0 = longer / invalid
    uint32_t      start[HD_THREADS + 1];  // bit offset (relative to the sequence start) of the first codeword of each subsequence
    uint32_t      used[HD_THREADS];       // sync kernel: start value each subsequence was last walked from
    uint16_t      count[HD_THREADS];      // sync kernel: codewords that start inside each subsequence
    uint16_t      list[HD_THREADS];       // sync kernel: subsequences to walk this round, compacted
    uint32_t      wcnt[HD_THREADS / 32];
    uint32_t      red[34];
};

// Phase mode of the synchronisation (codes of at most HD_PHASE_MAX bits): a code that crosses a subsequence border ends
// fewer than max_len bits behind it, so a subsequence has at most max_len possible starts ("phases"). Every subsequence
// is walked once from each of them, which gives the map start -> (start of the next subsequence, codewords inside); the
// true starts then follow by chaining the maps from the entry of the sequence -- in a time that does not depend on how
// fast the code re-synchronises (near-uniform code lengths, i.e. incompressible data, hardly ever do).
#define HD_PHASE_MAX 12u
#define HD_PHASE_SEQ_BYTES (HD_PHASE_MAX * HD_THREADS * 3u + 16u)  // per sequence: exits u8 [phase][sub], counts u16 [phase][sub], exit map of the sequence
// Phase mode is for codes that hardly ever re-synchronise: short and of nearly uniform length (incompressible data: 7 to 9
// bits). A code with short and long words (compressible data) re-synchronises within a few words, and walking every
// phase would cost several times the fixed-point iteration.
__host__ __device__ __forceinline__ bool hd_phase_mode(uint32_t min_len, uint32_t max_len)
{
    return max_len >= 1 && max_len <= HD_PHASE_MAX && max_len - min_len <= 3;
}
struct HdSyncShared
{
    HdShared S;
    uint16_t pc[HD_PHASE_MAX][HD_THREADS];
    uint8_t  pe[HD_PHASE_MAX][HD_THREADS];
};

// Thread k streams through words 32k..32k+31: unswizzled, all 256 threads would sit on one bank.
// Word i lives at (i & ~31) | ((i ^ (i >> 5)) & 31), so that threads at the same column hit 32 banks.
__device__ __forceinline__ uint32_t hd_word(const HdShared& S, uint32_t i) { return S.words[(i & ~31u) | ((i ^ (i >> 5)) & 31u)]; }

__device__ __forceinline__ void hd_stage(HdShared& S, const uint8_t* __restrict__ pay, uint32_t seq, uint32_t cbytes)
{
    // stage this CTA's payload slice (+ look-ahead) as big-endian words; bytes past the payload read as 0
    const uint32_t  byte0 = seq * HD_SEQ_BYTES;
    const uint32_t* pw    = reinterpret_cast<const uint32_t*>(pay) + byte0 / 4;
    for (uint32_t i = threadIdx.x; i < HD_SMEM_WORDS; i += HD_THREADS)
    {
        const uint32_t bo = byte0 + i * 4;
        uint32_t       v  = 0;
        if (bo < cbytes)
        {
            v = __byte_perm(pw[i], 0, 0x0123);
            if (bo + 4 > cbytes) v &= 0xFFFFFFFFu << ((bo + 4 - cbytes) * 8);
        }
        S.words[(i & ~31u) | ((i ^ (i >> 5)) & 31u)] = v;
    }
}

// first-level table: one lookup resolves every code of at most HD_LUT_BITS bits
__device__ __forceinline__ void hd_build_lut(HdShared& S)
{
    for (uint32_t i = threadIdx.x; i < (1u << HD_LUT_BITS); i += HD_THREADS)
    {
        uint8_t        sym = 0;
        const uint32_t l   = bra_huf_decode_one(&S.tab, i << (32 - HD_LUT_BITS), &sym);
        S.lut[i]           = (l && l <= HD_LUT_BITS) ? (uint16_t) ((sym << 8) | l) : (uint16_t) 0;
    }
}

// bit reader over the staged words: 64-bit left-aligned buffer, refilled a word at a time
struct HdReader
{
    uint64_t buf;
    uint32_t avail, next_word;
};
__device__ __forceinline__ void hd_open(const HdShared& S, HdReader& R, uint32_t pos)
{
    const uint32_t w = pos >> 5, o = pos & 31u;
    R.buf       = (((uint64_t) hd_word(S, w) << 32) | hd_word(S, w + 1)) << o;
    R.avail     = 64u - o;
    R.next_word = w + 2;
}
__device__ __forceinline__ void hd_skip(const HdShared& S, HdReader& R, uint32_t l)
{
    R.buf <<= l;
    R.avail -= l;
    if (R.avail <= 32u)
    {
        const uint32_t nw = R.next_word < HD_SMEM_WORDS ? hd_word(S, R.next_word) : 0u;
        R.buf |= (uint64_t) nw << (32u - R.avail);
        R.avail += 32u;
        ++R.next_word;
    }
}
__device__ __forceinline__ uint32_t hd_decode(const HdShared& S, const HdReader& R, uint8_t* sym)
{
    const uint32_t w = (uint32_t) (R.buf >> 32);
    const uint32_t e = S.lut[w >> (32 - HD_LUT_BITS)];
    if (e)
    {
        *sym = (uint8_t) (e >> 8);
        return e & 0xFFu;
    }
    return bra_huf_decode_one(&S.tab, w, sym);
}

// Decode subsequence k from relative bit `pos` until crossing its end (or the end of the payload).
// Returns the exit position; *count = number of complete codewords that START inside [pos, end).
// A bit pattern that matches no codeword stops the walk (dead speculative path or corrupt data).
__device__ __forceinline__ uint32_t hd_walk(const HdShared& S, uint32_t pos, uint32_t sub_end, uint32_t data_end, uint32_t* count, bool* dead)
{
    uint32_t c = 0;
    *dead      = false;
    if (pos < sub_end && pos < data_end)
    {
        HdReader R;
        hd_open(S, R, pos);
        while (pos < sub_end && pos < data_end)
        {
            uint8_t        sym;
            const uint32_t l = hd_decode(S, R, &sym);
            if (l == 0)
            {
                *dead = true;
                break;
            }
            if (pos + l > data_end) break;  // incomplete trailing code: padding
            hd_skip(S, R, l);
            pos += l;
            ++c;
        }
    }
    *count = c;
    return pos;
}

// One synchronisation sweep. sub_start/sub_count persist between launches.
//   seq_entry[b][seq] : entry bit offset this CTA last used (0xFFFFFFFF = never ran)
//   seq_exit[b][seq]  : where its last subsequence crossed into the next CTA's sequence (relative to that sequence)
__global__ void __launch_bounds__(HD_THREADS)
    huf_dec_sync_kernel(const uint8_t* __restrict__ pay, uint64_t pay_stride, const uint32_t* __restrict__ clen, const bra_huf_dec_t* __restrict__ tabs,
                        const uint32_t* __restrict__ err, uint32_t seqs, uint8_t* __restrict__ sub_start, uint16_t* __restrict__ sub_count,
                        uint32_t* __restrict__ seq_entry, uint32_t* __restrict__ seq_exit, uint32_t* __restrict__ seq_count,
                        uint32_t* __restrict__ changed_flag, uint8_t* __restrict__ phase_ws)
{
    extern __shared__ __align__(16) uint8_t hd_sync_smem[];
    HdSyncShared&  SS = *reinterpret_cast<HdSyncShared*>(hd_sync_smem);
    HdShared&      S  = SS.S;
    const uint32_t b = blockIdx.y, seq = blockIdx.x;
    const uint32_t      c = clen[b];
    if ((uint64_t) seq * HD_SEQ_BYTES >= c || err[b]) return;
    const uint64_t sidx = (uint64_t) b * seqs + seq;
    // seq_exit of the left neighbour may be rewritten by that CTA during this very launch: read it
    // once and broadcast, so that the whole CTA takes the same decision.
    __shared__ uint32_t s_entry, s_prev;
    if (threadIdx.x == 0)
    {
        s_entry = seq == 0 ? 0u : *reinterpret_cast<volatile const uint32_t*>(&seq_exit[sidx - 1]);
        s_prev  = seq_entry[sidx];
    }
    __syncthreads();
    const uint32_t entry = s_entry;
    if (s_prev == entry) return;  // nothing upstream changed since this CTA last ran

    const bool first_run = s_prev == 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < sizeof(bra_huf_dec_t) / 4; i += HD_THREADS)
        reinterpret_cast<uint32_t*>(&S.tab)[i] = reinterpret_cast<const uint32_t*>(&tabs[b])[i];
    hd_stage(S, pay + (uint64_t) b * pay_stride, seq, c);
    const uint32_t k        = threadIdx.x;
    const uint64_t sub_idx  = sidx * HD_THREADS + k;
    const uint32_t data_end = min((uint32_t) HD_SEQ_BITS + 64u, (c - seq * HD_SEQ_BYTES) * 8u);  // relative bit where the payload ends
    __syncthreads();
    const uint32_t L = S.tab.max_len;
    if (phase_ws && hd_phase_mode(S.tab.min_len, L) && entry < L)
    {
        uint8_t*  g_pe  = phase_ws + sidx * HD_PHASE_SEQ_BYTES;
        uint16_t* g_pc  = reinterpret_cast<uint16_t*>(g_pe + HD_PHASE_MAX * HD_THREADS);
        uint8_t*  g_map = g_pe + HD_PHASE_MAX * HD_THREADS * 3u;
        if (first_run)
        {
            hd_build_lut(S);
            __syncthreads();
            for (uint32_t o = 0; o < L; ++o)
            {
                uint32_t cnt;
                bool     dead;
                uint32_t nx = hd_walk(S, k * HD_SUB_BITS + o, (k + 1) * HD_SUB_BITS, data_end, &cnt, &dead);
                uint32_t ex = (dead || nx < (k + 1) * HD_SUB_BITS) ? 0u : nx - (k + 1) * HD_SUB_BITS;  // dead path / payload ended: neutral
                if (ex >= L) ex = 0;
                SS.pe[o][k] = (uint8_t) ex;
                SS.pc[o][k] = (uint16_t) cnt;
                g_pe[o * HD_THREADS + k] = (uint8_t) ex;
                g_pc[o * HD_THREADS + k] = (uint16_t) cnt;
            }
        }
        else
            for (uint32_t o = 0; o < L; ++o)
            {
                SS.pe[o][k] = g_pe[o * HD_THREADS + k];
                SS.pc[o][k] = g_pc[o * HD_THREADS + k];
            }
        __syncthreads();
        if (k < L)  // thread o chains the maps from entry phase o; the one that starts from the true entry records the path
        {
            const bool mine = k == entry;
            uint32_t   ph   = k;
            for (uint32_t j = 0; j < HD_THREADS; ++j)
            {
                if (mine)
                {
                    S.start[j] = j * HD_SUB_BITS + ph;
                    S.count[j] = SS.pc[ph][j];
                }
                ph = SS.pe[ph][j];
            }
            if (first_run) g_map[k] = (uint8_t) ph;
            if (mine) S.start[HD_THREADS] = HD_SEQ_BITS + ph;
        }
        __syncthreads();
    }
    else
    {
    hd_build_lut(S);
    __syncthreads();
    // current guess for this subsequence's first codeword. First run: decode a short warm-up
    // stretch ahead of the subsequence -- a prefix code re-synchronises within a few codewords on
    // compressible data, so the first boundary crossed is usually already the true one.
    {
        uint32_t guess = k * HD_SUB_BITS;
        if (!first_run)
            guess += sub_start[sub_idx];
        else if (k != 0)
        {
            uint32_t cnt_unused;
            bool     dead;
            const uint32_t g = hd_walk(S, k * HD_SUB_BITS - HD_WARMUP_BITS, k * HD_SUB_BITS, data_end, &cnt_unused, &dead);
            if (!dead && g >= k * HD_SUB_BITS && g < k * HD_SUB_BITS + 32u) guess = g;
        }
        S.start[k] = guess;
    }
    if (k == 0)
    {
        S.start[0]          = entry;
        S.start[HD_THREADS] = first_run ? HD_SEQ_BITS : HD_SEQ_BITS + seq_exit[sidx];
    }
    __syncthreads();

    // Jacobi iteration on the subsequence starts. Only subsequences whose start moved are walked again, and on
    // data that re-synchronises slowly (near-uniform code lengths) those are a few scattered ones per round: they
    // are compacted so that the walks run on dense warps instead of one or two lanes in each of the eight.
    S.used[k]  = 0xFFFFFFFFu;  // start value the current (exit, count) of subsequence k were computed from
    S.count[k] = 0;
    if (!first_run && k != 0)
    {
        // state from the previous launch is still valid unless the start changes
        S.used[k]  = S.start[k];
        S.count[k] = sub_count[sub_idx];
    }
    __syncthreads();
    for (int it = 0; it < HD_THREADS + 2; ++it)
    {
        const bool     need = S.start[k] != S.used[k];
        const uint32_t ball = __ballot_sync(BRA_FULL, need);
        if (lane_id() == 0) S.wcnt[warp_id()] = __popc(ball);
        __syncthreads();
        uint32_t before = 0, nact = 0;
#pragma unroll
        for (uint32_t w = 0; w < HD_THREADS / 32; ++w)
        {
            const uint32_t c = S.wcnt[w];
            before += w < warp_id() ? c : 0u;
            nact += c;
        }
        if (nact == 0) break;
        if (need) S.list[before + __popc(ball & lanemask_lt())] = (uint16_t) k;
        __syncthreads();
        uint32_t j = 0, nx = 0;
        if (k < nact)
        {
            j                 = S.list[k];
            const uint32_t st = S.start[j];
            uint32_t       cnt;
            bool           dead;
            nx = hd_walk(S, st, (j + 1) * HD_SUB_BITS, data_end, &cnt, &dead);
            if (dead || nx < (j + 1) * HD_SUB_BITS) nx = (j + 1) * HD_SUB_BITS;  // dead path / payload ended: neutral guess
            S.used[j]  = st;
            S.count[j] = (uint16_t) cnt;
        }
        __syncthreads();
        if (k < nact) S.start[j + 1] = nx;
        __syncthreads();
    }
    }
    const uint32_t mycount = S.count[k];
    sub_start[sub_idx] = (uint8_t) (S.start[k] - k * HD_SUB_BITS);
    sub_count[sub_idx] = (uint16_t) mycount;
    uint32_t total;
    block_excl_add(mycount, S.red, &total);
    if (k == 0)
    {
        seq_entry[sidx] = entry;
        seq_count[sidx] = total;
        const uint32_t ex = S.start[HD_THREADS] - HD_SEQ_BITS;
        if (first_run || seq_exit[sidx] != ex)
        {
            seq_exit[sidx] = ex;
            atomicAdd(changed_flag, 1u);
        }
    }
}

// Phase mode: the true entry of every sequence of a block, by chaining the sequences' exit maps (one thread per block).
__global__ void huf_dec_phase_chain_kernel(const uint32_t* __restrict__ clen, const bra_huf_dec_t* __restrict__ tabs, const uint32_t* __restrict__ err,
                                           uint32_t seqs, const uint8_t* __restrict__ phase_ws, uint32_t* __restrict__ seq_exit, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk || err[b]) return;
    const uint32_t L = tabs[b].max_len;
    if (!hd_phase_mode(tabs[b].min_len, L)) return;
    const uint32_t nseq = (clen[b] + HD_SEQ_BYTES - 1) / HD_SEQ_BYTES;
    uint32_t       ph   = 0;
    for (uint32_t s = 0; s < nseq; ++s)
    {
        const uint64_t sidx = (uint64_t) b * seqs + s;
        ph                  = phase_ws[sidx * HD_PHASE_SEQ_BYTES + HD_PHASE_MAX * HD_THREADS * 3u + ph];
        seq_exit[sidx]      = ph;
    }
}

// exclusive scan of symbol counts over the sequences of a block; checks total >= orig_size
__global__ void __launch_bounds__(256)
    huf_dec_scan_kernel(uint32_t* __restrict__ seq_count, const uint32_t* __restrict__ clen, const uint8_t* __restrict__ hdr, uint32_t seqs,
                        uint32_t* __restrict__ err)
{
    __shared__ uint32_t red[34];
    const uint32_t      b = blockIdx.x;
    if (err[b]) return;
    const uint32_t c     = clen[b];
    const uint32_t nseq  = (c + HD_SEQ_BYTES - 1) / HD_SEQ_BYTES;
    uint32_t*      sc    = seq_count + (uint64_t) b * seqs;
    uint32_t       carry = 0;
    for (uint32_t base = 0; base < nseq; base += 256)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nseq ? sc[i] : 0u;
        uint32_t       tot;
        const uint32_t ex = block_excl_add(v, red, &tot) + carry;
        if (i < nseq) sc[i] = ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        const uint8_t* h    = hdr + (uint64_t) b * 268;
        const uint32_t orig = (uint32_t) h[260] | ((uint32_t) h[261] << 8) | ((uint32_t) h[262] << 16) | ((uint32_t) h[263] << 24);
        if (carry < orig) err[b] = 1;  // ran out of data (bra_huffman.c:485-489)
    }
}

// Second pass: every subsequence decodes again from its synchronised start and stores its symbols.
// The thread that writes symbol orig_size-1 records the bit position just after it.
__global__ void __launch_bounds__(HD_THREADS)
    huf_dec_write_kernel(const uint8_t* __restrict__ pay, uint64_t pay_stride, const uint32_t* __restrict__ clen, const uint8_t* __restrict__ hdr,
                         const bra_huf_dec_t* __restrict__ tabs, uint32_t seqs, const uint8_t* __restrict__ sub_start,
                         const uint16_t* __restrict__ sub_count, const uint32_t* __restrict__ seq_off, uint8_t* __restrict__ out,
                         uint64_t out_stride, uint32_t* __restrict__ end_bit, uint32_t* __restrict__ err)
{
    __shared__ HdShared S;
    const uint32_t      b = blockIdx.y, seq = blockIdx.x;
    const uint32_t      c = clen[b];
    if ((uint64_t) seq * HD_SEQ_BYTES >= c || err[b]) return;
    const uint8_t* h    = hdr + (uint64_t) b * 268;
    const uint32_t orig = (uint32_t) h[260] | ((uint32_t) h[261] << 8) | ((uint32_t) h[262] << 16) | ((uint32_t) h[263] << 24);
    const uint64_t sidx = (uint64_t) b * seqs + seq;
    const uint32_t o0   = seq_off[sidx];
    if (o0 >= orig) return;  // everything here is padding
    for (uint32_t i = threadIdx.x; i < sizeof(bra_huf_dec_t) / 4; i += HD_THREADS)
        reinterpret_cast<uint32_t*>(&S.tab)[i] = reinterpret_cast<const uint32_t*>(&tabs[b])[i];
    hd_stage(S, pay + (uint64_t) b * pay_stride, seq, c);
    __syncthreads();
    hd_build_lut(S);
    const uint32_t k       = threadIdx.x;
    const uint64_t sub_idx = sidx * HD_THREADS + k;
    const uint32_t cnt     = sub_count[sub_idx];
    uint32_t       dummy;
    uint32_t       o   = o0 + block_excl_add(cnt, S.red, &dummy);  // also orders the staging / table writes
    uint32_t       pos = k * HD_SUB_BITS + sub_start[sub_idx];
    uint8_t*       ob  = out + (uint64_t) b * out_stride;
    const uint32_t sub_end  = (k + 1) * HD_SUB_BITS;
    const uint32_t data_end = min((uint32_t) HD_SEQ_BITS + 64u, (c - seq * HD_SEQ_BYTES) * 8u);
    if (!(pos < sub_end && pos < data_end && o < orig)) return;
    HdReader R;
    hd_open(S, R, pos);
    // same walk as hd_walk (so the counts agree), now storing the symbols: four at a time as one aligned
    // 32-bit store wherever the output position allows it (byte stores cost one L2 transaction each)
    uint32_t acc = 0, nacc = 0;  // symbols not stored yet: they belong to ob[o - nacc .. o)
    while (pos < sub_end && pos < data_end && o < orig)
    {
        uint8_t        sym = 0;
        const uint32_t l   = hd_decode(S, R, &sym);
        if (l == 0)
        {
            err[b] = 1;  // no codeword matches before orig_size symbols: "invalid code sequence" (bra_huffman.c:466-470)
            break;
        }
        if (pos + l > data_end) break;
        acc |= (uint32_t) sym << (8u * nacc);
        ++nacc;
        ++o;
        if (((reinterpret_cast<uintptr_t>(ob) + o) & 3u) == 0)
        {
            if (nacc == 4)
                *reinterpret_cast<uint32_t*>(ob + o - 4) = acc;
            else
                for (uint32_t i = 0; i < nacc; ++i) ob[o - nacc + i] = (uint8_t) (acc >> (8u * i));
            acc  = 0;
            nacc = 0;
        }
        hd_skip(S, R, l);
        pos += l;
        if (o == orig) end_bit[b] = seq * HD_SEQ_BITS + pos;
    }
    for (uint32_t i = 0; i < nacc; ++i) ob[o - nacc + i] = (uint8_t) (acc >> (8u * i));
}

// Bytes after the one holding the last symbol: the reference keeps walking them from the tree root
// (bra_huffman.c:455-481) and fails if that walk completes a codeword or leaves the tree.
__global__ void huf_dec_trailing_kernel(const uint8_t* __restrict__ pay, uint64_t pay_stride, const uint32_t* __restrict__ clen,
                                        const bra_huf_dec_t* __restrict__ tabs, const uint32_t* __restrict__ end_bit, uint32_t* __restrict__ err,
                                        uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk || err[b]) return;
    const uint32_t       c    = clen[b];
    const uint32_t       from = (end_bit[b] + 7u) / 8u;  // first byte the reference's outer loop visits next
    if (from >= c) return;
    const bra_huf_dec_t* d = &tabs[b];
    const uint8_t*       p = pay + (uint64_t) b * pay_stride;
    // bit-by-bit walk with exact tree semantics: prefix v of depth dep is a node iff some code of
    // length L >= dep has it as a prefix; it is a leaf iff L == dep.
    uint32_t v = 0, dep = 0;
    for (uint32_t i = from; i < c; ++i)
        for (int bit = 7; bit >= 0; --bit)
        {
            v = (v << 1) | ((p[i] >> bit) & 1u);
            ++dep;
            bool node = false, leaf = false;
            if (dep <= d->max_len)
                for (uint32_t L = dep; L <= d->max_len; ++L)
                {
                    if (!d->count[L]) continue;
                    const uint32_t lo = d->first[L] >> (L - dep), hi = (d->first[L] + d->count[L] - 1) >> (L - dep);
                    if (v >= lo && v <= hi)
                    {
                        node = true;
                        leaf = (L == dep);
                        break;
                    }
                }
            if (!node || leaf)
            {
                err[b] = 1;  // left the tree, or decoded a symbol beyond orig_size
                return;
            }
        }
}

// sweep status back to the host (through the mail word when there is one, see BwtFwdArgs::h_mail)
static bool read_changed(const HufDecArgs& a, cudaStream_t st, uint32_t* changed)
{
    if (a.h_mail)
    {
        if (!mail_publish(a.h_mail, a.d_changed, 1, st)) return false;
        BRA_CUDA_TRY(cudaStreamSynchronize(st));
        *changed = *reinterpret_cast<volatile uint32_t*>(a.h_mail);
        return true;
    }
    BRA_CUDA_TRY(cudaMemcpyAsync(changed, a.d_changed, 4, cudaMemcpyDeviceToHost, st));
    BRA_CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

bool huf_decode_batch(const HufDecArgs& a, cudaStream_t st)
{
    if (a.nblk == 0 || a.max_c == 0) return true;
    const uint32_t seqs = bra_div_up(a.max_c, HD_SEQ_BYTES);
    const dim3     grid(seqs, a.nblk);
    BRA_LAUNCH(P_HUF_DEC_TABLES, st, huf_dec_tables_kernel<<<a.nblk, 32, 0, st>>>(a.d_hdr, a.d_tabs, a.d_err));
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_seq_entry, 0xFF, (size_t) a.nblk * seqs * 4, st));
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_seq_exit, 0, (size_t) a.nblk * seqs * 4, st));  // first guess: codewords start on sequence boundaries
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_end_bit, 0, (size_t) a.nblk * 4, st));
    uint32_t     sweeps    = 0;
    const size_t sync_smem = sizeof(HdSyncShared);
    BRA_CUDA_TRY(cudaFuncSetAttribute(huf_dec_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sync_smem));
    auto sweep = [&]() -> bool {
        BRA_LAUNCH(P_HUF_DEC_SYNC, st, huf_dec_sync_kernel<<<grid, HD_THREADS, sync_smem, st>>>(a.d_pay, a.pay_stride, a.d_clen, a.d_tabs, a.d_err, seqs, a.d_sub_start,
                                                         a.d_sub_count, a.d_seq_entry, a.d_seq_exit, a.d_seq_count, a.d_changed, a.d_phase));
        ++sweeps;
        return true;
    };
    for (;;)
    {
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_changed, 0, 4, st));
        if (!sweep()) return false;
        if (sweeps == 1 && a.d_phase)
            // blocks in phase mode: every sequence knows its exit for every entry now -- settle the entries in one go
            BRA_LAUNCH(P_HUF_DEC_SCAN, st, huf_dec_phase_chain_kernel<<<bra_div_up(a.nblk, 64), 64, 0, st>>>(a.d_clen, a.d_tabs, a.d_err, seqs, a.d_phase, a.d_seq_exit, a.nblk));
        if (!sweep()) return false;
        uint32_t changed = 0;
        if (!read_changed(a, st, &changed)) return false;
        // the first pair of sweeps always reports changes (every CTA publishes its first exit)
        if (changed == 0 || sweeps > 2 * seqs + 4) break;
        if (sweeps == 2)
        {
            // cheap confirmation sweep: if nothing moves any more we are at the fixed point
            BRA_CUDA_TRY(cudaMemsetAsync(a.d_changed, 0, 4, st));
            if (!sweep()) return false;
            if (!read_changed(a, st, &changed)) return false;
            if (changed == 0) break;
        }
    }
    if (a.h_sweeps) *a.h_sweeps = sweeps;
    BRA_LAUNCH(P_HUF_DEC_SCAN, st, huf_dec_scan_kernel<<<a.nblk, 256, 0, st>>>(a.d_seq_count, a.d_clen, a.d_hdr, seqs, a.d_err));
    BRA_LAUNCH(P_HUF_DEC_WRITE, st, huf_dec_write_kernel<<<grid, HD_THREADS, 0, st>>>(a.d_pay, a.pay_stride, a.d_clen, a.d_hdr, a.d_tabs, seqs, a.d_sub_start, a.d_sub_count,
                                                      a.d_seq_count, a.d_out, a.out_stride, a.d_end_bit, a.d_err));
    BRA_LAUNCH(P_HUF_DEC_TRAILING, st, huf_dec_trailing_kernel<<<bra_div_up(a.nblk, 64), 64, 0, st>>>(a.d_pay, a.pay_stride, a.d_clen, a.d_tabs, a.d_end_bit, a.d_err, a.nblk));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

uint32_t huf_enc_tiles(uint32_t max_r) { return bra_div_up(max_r, HF_TILE); }
uint32_t huf_dec_seqs(uint32_t max_c) { return bra_div_up(max_c, HD_SEQ_BYTES); }
uint32_t huf_dec_subs_per_seq() { return HD_THREADS; }
uint32_t huf_dec_phase_bytes_per_seq() { return HD_PHASE_SEQ_BYTES; }

}  // namespace bra
