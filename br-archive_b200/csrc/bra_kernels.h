// bra_kernels.h -- internal launchers shared between the .cu files (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "bra_hd.h"

namespace bra {

// ---- prof.cu: launch accounting (see include/bra_b200.h, bra_b200_prof_*) ------------------------
enum ProfId
{
    P_CRC, P_RS_HIST, P_RS_SCATTER_IMPL, P_RS_SCATTER, P_RS_SCATTER_U8, P_BWT_PERIOD, P_BWT_KEYS, P_BWT_HEADS, P_BWT_RANKS, P_BWT_PREPARE, P_BWT_GATHER, P_BWT_FINISH, P_BWT_MISC,
    P_MTF_SUMMARY, P_MTF_SCAN, P_MTF_APPLY, P_RLE_ENC_HEADS, P_RLE_ENC_LIT, P_RLE_ENC_SIZE, P_RLE_ENC_EMIT, P_RLE_DEC_EXIT, P_RLE_DEC_CHAIN,
    P_RLE_DEC_MARK, P_RLE_DEC_EXPAND, P_HUF_HIST, P_HUF_BUILD, P_HUF_BITS, P_HUF_PACK, P_HUF_DEC_TABLES, P_HUF_DEC_SYNC, P_HUF_DEC_SCAN,
    P_HUF_DEC_WRITE, P_HUF_DEC_TRAILING, P_IBWT_WALK_LEN, P_IBWT_STITCH, P_IBWT_WALK_EMIT, P_IBWT_COPY, P_GLUE, P_COUNT
};
void prof_pre(int id, cudaStream_t st);
void prof_post(int id, cudaStream_t st);
uint64_t prof_total_launches();
// every kernel launch of the library goes through this macro
#define BRA_LAUNCH(id, st, ...)  \
    do                           \
    {                            \
        bra::prof_pre(id, st);   \
        __VA_ARGS__;             \
        bra::prof_post(id, st);  \
    } while (0)

// ---- crc32c.cu ----------------------------------------------------------------------------
bool                crc_init_tables();
const bra_gf_pow_t* crc_host_pow();
// d_crc[b] = CRC-32C of block b continued from d_prev[b] (or 0). len from d_len[b] or fixed_len.
bool crc_blocks(const uint8_t* d_in, uint64_t stride, const uint32_t* d_len, uint32_t fixed_len, uint32_t max_len, uint32_t nblk,
                const uint32_t* d_prev, uint32_t* d_crc, cudaStream_t st);
bool crc_headers(const uint8_t* d_hdr, uint32_t item_bytes, uint32_t nitems, uint32_t* d_out, cudaStream_t st);

// ---- sort.cu ------------------------------------------------------------------------------
// One stable 8-bit LSD pass = one kernel (decoupled look-back over tiles, values staged by bulk-asynchronous copies).
// `d_hist` is the sort workspace of radix_hist_bytes() bytes. Its first nblk * RS_GHIST_STRIDE words are the digit
// histograms of every block, [b][pass][256], which the producer of the keys must have filled for all passes of the
// sort before the first pass is launched (zeroed by the producer; pass p looks at bits [8p, 8p+8) of the key).
#define RS_GHIST_PASSES 4
#define RS_GHIST_STRIDE (RS_GHIST_PASSES * 256)
size_t radix_hist_bytes(uint32_t max_len, uint32_t nblk);
// small host <-> device transfers done by a copy kernel through pinned, device-visible host memory: a cudaMemcpyAsync of
// a few bytes would wait on its copy engine behind the bulk copies of the neighbouring pipeline stages
bool mail_fetch(uint32_t* d_dst, const uint32_t* h_src, uint32_t words, cudaStream_t st);
bool mail_publish(uint32_t* h_dst, const uint32_t* d_src, uint32_t words, cudaStream_t st);
// vals == nullptr: the values are the element indices (first pass of a sort). The sort key starts at bit `key_shift` of
// the key word; pass p looks at its bits [8p, 8p+8).
bool radix_pass_u32(const uint32_t* keys, const uint32_t* vals, uint32_t* keys_out, uint32_t* vals_out, uint64_t stride, const uint32_t* d_len,
                    const uint8_t* d_skip, uint32_t max_len, uint32_t nblk, uint32_t pass, uint32_t key_shift, uint32_t* d_hist, cudaStream_t st);
// u8 keys, implicit index values, output word (index << 8) | key; computes its own histogram
bool radix_pass_u8_index_packed(const uint8_t* keys, uint32_t* packed_out, uint64_t stride, const uint32_t* d_len, uint32_t max_len,
                                uint32_t nblk, uint32_t* d_hist, cudaStream_t st);

// ---- bwt.cu -------------------------------------------------------------------------------
struct BwtFwdArgs
{
    const uint8_t*  d_in;
    uint8_t*        d_out;      // last column, same stride
    uint64_t        stride;     // elements between consecutive blocks in every per-byte array
    const uint32_t* d_len;      // device copy of block lengths
    const uint32_t* h_len;      // host copy (divisor tables are built on the host)
    uint32_t        max_n, nblk;
    uint32_t*       d_primary;
    // workspace (u32 arrays are nblk*stride elements)
    uint32_t *d_keyA, *d_keyB, *d_valA, *d_valB, *d_rankA, *d_rankB;
    uint8_t * d_flags, *d_flags2;  // nblk*stride bytes each (head flags, double buffered across rounds)
    uint32_t* d_hist;     // radix_hist_bytes(max_n, nblk)
    int*      d_tile_last;  // nblk * ceil(max_n/4096)
    uint32_t *d_period, *d_ngroups, *d_notdone /* 4 words */, *d_maxgroup;
    unsigned long long* d_sumsq;
    uint8_t * d_done, *d_fin, *d_finskip;
    uint32_t *d_div_vals, *d_div_off, *d_div_cnt;
    uint32_t  div_cap;
    uint8_t*  d_bad;
    uint32_t  bad_stride;
    uint32_t* d_tile_heads;  // optional, nblk * ceil(max_n / 4096): with it the first doubling round sorts on dense group numbers (fewer key bits)
    uint8_t*  d_alpha;  // optional, nblk*256: dense symbol codes per block -- the 4-symbol sort keys then need 4*ceil(log2(symbols)) bits only
    uint32_t* h_rounds;  // optional: number of doubling rounds executed
    // optional pinned, device-visible host words: [0,2) loop status, [2,2+div_cap) divisor values, then nblk offsets and
    // nblk counts. With it the loop's small transfers are done by copy kernels and do not queue behind bulk copies
    // on the copy engines; without it they are plain cudaMemcpyAsync calls.
    uint32_t* h_mail;
};
bool bwt_forward_batch(const BwtFwdArgs& a, cudaStream_t st);

struct BwtInvArgs
{
    const uint8_t*  d_in;   // last column
    uint8_t*        d_out;
    uint64_t        stride;
    const uint32_t* d_len;
    const uint32_t* d_primary;
    uint32_t        max_n, nblk;
    uint32_t*       d_W;     // nblk*stride
    uint32_t*       d_hist;
    uint2*          d_walk;  // nblk * ibwt_kmax(max_n)
    uint32_t*       d_woff;  // same
    uint32_t*       d_orbit; // nblk
    uint32_t*       d_ovf;   // optional: nblk * ibwt_ovf_words() -- overflow walker pool (needs d_tmp)
    uint8_t*        d_tmp;   // optional: nblk * ibwt_kmax(max_n) * ibwt_tmp_cap(max_n) bytes -- the first walk keeps the bytes it passes
};
uint32_t ibwt_tmp_cap(uint32_t max_n);
uint32_t ibwt_row_stride(uint32_t max_n);
uint32_t ibwt_kmax(uint32_t max_n);
uint32_t ibwt_ovf_words();
bool     bwt_inverse_batch(const BwtInvArgs& a, cudaStream_t st);

// ---- mtf.cu -------------------------------------------------------------------------------
uint32_t mtf_segments(uint32_t max_n);
bool     mtf_encode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                          uint8_t* d_summ, uint16_t* d_scnt, uint8_t* d_state, cudaStream_t st);
bool     mtf_decode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                          uint8_t* d_summ, uint8_t* d_state, cudaStream_t st);

// ---- rle.cu -------------------------------------------------------------------------------
struct RleEncArgs
{
    const uint8_t*  d_in;
    uint64_t        stride;
    const uint32_t* d_len;
    uint32_t        max_n, nblk;
    uint8_t*        d_out;
    uint64_t        out_stride;
    uint32_t*       d_rlen;  // encoded size per block
    uint32_t*       d_hist;  // 256 bins per block of the encoded bytes (input of the Huffman build)
    int *d_t_first_head, *d_t_last_head, *d_t_first_nl, *d_t_last_nl;  // nblk * rle_enc_tiles(max_n) each
    uint32_t* d_t_cnt;
};
bool     rle_encode_batch(const RleEncArgs& a, cudaStream_t st);
uint32_t rle_enc_tiles(uint32_t max_n);

struct RleDecArgs
{
    const uint8_t*  d_in;
    uint64_t        stride;
    const uint32_t* d_rlen;
    uint32_t        max_r, nblk;
    uint8_t*        d_out;
    uint64_t        out_stride;
    uint32_t        out_cap;   // blocks decoding to more than this many bytes fail
    uint32_t*       d_nlen;    // decoded size per block, 0 on error
    uint32_t*       d_err;     // per block, must be zeroed by the caller
    uint8_t *       d_t_exit, *d_t_entry;  // nblk*tiles*rle_dec_entries(), nblk*tiles
    uint32_t *      d_t_tok, *d_t_ocnt;    // nblk*tiles*32, nblk*tiles
    bool            size_only;
};
bool     rle_decode_batch(const RleDecArgs& a, cudaStream_t st);
uint32_t rle_dec_tiles(uint32_t max_r);
uint32_t rle_dec_entries();

// ---- huffman.cu ---------------------------------------------------------------------------
struct HufEncArgs
{
    const uint8_t*  d_in;
    uint64_t        stride;
    const uint32_t* d_rlen;
    uint32_t        max_r, nblk;
    bool            compute_hist;  // false when the RLE emit kernel already filled d_hist
    uint32_t*       d_hist;        // nblk*256
    uint8_t*        d_hdr;         // nblk*268: lengths, orig_size, encoded_size are written here
    uint32_t*       d_codes;       // nblk*256
    uint32_t*       d_ok;          // nblk
    uint32_t*       d_t_bits;      // nblk*huf_enc_tiles(max_r)
    uint32_t*       d_clen;        // nblk
    uint8_t*        d_out;
    uint64_t        out_stride;
};
bool     huf_encode_batch(const HufEncArgs& a, cudaStream_t st);
uint32_t huf_enc_tiles(uint32_t max_r);

struct HufDecArgs
{
    const uint8_t*  d_pay;
    uint64_t        pay_stride;
    const uint32_t* d_clen;
    const uint8_t*  d_hdr;  // nblk*268
    uint32_t        max_c, nblk;
    uint8_t*        d_out;
    uint64_t        out_stride;  // must hold orig_size bytes per block
    bra_huf_dec_t*  d_tabs;
    uint32_t*       d_err;  // per block, zeroed by the caller
    uint8_t*        d_sub_start;   // nblk*seqs*256
    uint16_t*       d_sub_count;   // nblk*seqs*256
    uint32_t *      d_seq_entry, *d_seq_exit, *d_seq_count;  // nblk*seqs
    uint32_t *      d_end_bit, *d_changed;
    uint8_t*        d_phase;       // optional, nblk*seqs*huf_dec_phase_bytes_per_seq(): enables the phase mode of the synchronisation
    uint32_t*       h_sweeps;
    uint32_t*       h_mail;  // optional pinned, device-visible host word for the sweep status (see BwtFwdArgs::h_mail)
};
bool     huf_decode_batch(const HufDecArgs& a, cudaStream_t st);
uint32_t huf_dec_seqs(uint32_t max_c);
uint32_t huf_dec_subs_per_seq();
uint32_t huf_dec_phase_bytes_per_seq();

}  // namespace bra
