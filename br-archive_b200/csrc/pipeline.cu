// pipeline.cu -- context, device workspace and the batched encode / decode chains
// (the seam under reference src/io/lib_bra_io_file_chunks.c:199-266 and :338-414).
//
// Device memory layout. One arena per context, carved per call:
//   per-byte arrays (input copy, last column, MTF ranks, decode scratch) : nblk x stride bytes
//   per-element u32 arrays (sort keys/values, ranks, inverse-BWT words)  : nblk x stride x 4
//   RLE output   : nblk x rle_stride   (rle_stride   = stride*129/128 + slack: literals-only worst case)
//   payload      : nblk x pay_stride   (pay_stride   = rle_stride*9/8 + slack: H+1 <= 9 bits/symbol)
//   tile summaries, histograms, per-block scalars: small.
// Block b of a batch always lives at offset b*stride (or b*rle_stride, ...) of an array, so a
// kernel addresses it from blockIdx.y alone and lengths stay in device memory between stages.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"
#include "pipeline.h"

#include <algorithm>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

// ---- logging: use the host program's bra_log_error when it is linked in (reference src/log/bra_log.h) ----
extern "C" void bra_log_error(const char* fmt, ...) __attribute__((weak));

void bra_b200_log_error(const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (bra_log_error)
        bra_log_error("%s", buf);
    else
        fprintf(stderr, "ERROR: %s\n", buf);
}

namespace bra {

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

uint64_t rle_stride_for(uint32_t block) { return align_up((uint64_t) block + block / 128 + 64, 256); }
uint64_t pay_stride_for(uint32_t block) { return align_up(rle_stride_for(block) * 9 / 8 + 64, 256); }

// ---- arena ------------------------------------------------------------------------------------
struct Arena
{
    uint8_t* base = nullptr;
    uint64_t cap = 0, used = 0;
    template <typename T>
    T* take(uint64_t count)
    {
        used = align_up(used, 256);
        T* p = reinterpret_cast<T*>(base + used);
        used += count * sizeof(T);
        return p;
    }
};

// workspace requirement of one batch, by dry-running the carve
struct EncWs
{
    uint8_t * L, *M, *R, *flags, *flags2, *bad, *done, *fin, *finskip, *summ, *state, *alpha;
    uint32_t* maxgroup;
    unsigned long long* sumsq;
    uint32_t *keyA, *keyB, *valA, *valB, *rankA, *rankB, *hist, *len, *primary, *period, *ngroups, *notdone, *div_vals, *div_off, *div_cnt;
    uint32_t *rlen, *clen, *rhist, *codes, *ok, *t_bits, *t_cnt;
    int *     tile_last, *t_first_head, *t_last_head, *t_first_nl, *t_last_nl;
    uint16_t* scnt;
};
struct DecWs
{
    uint8_t * R, *M, *L, *summ, *state, *sub_start, *t_exit, *t_entry, *wtmp, *phase;
    uint16_t* sub_count;
    uint32_t *W, *hist, *rlen, *clen, *nlen, *primary, *err, *seq_entry, *seq_exit, *seq_count, *end_bit, *changed, *t_tok, *t_ocnt, *woff,
        *orbit, *ovf;
    uint2*         walk;
    bra_huf_dec_t* tabs;
};

#define BRA_DIV_CAP 4096u
#define BRA_BAD_STRIDE 1024u

static void carve_enc(Arena& A, uint32_t S, uint32_t nb, EncWs& w)
{
    const uint64_t N = (uint64_t) nb * S, RS = rle_stride_for(S);
    w.keyA = A.take<uint32_t>(N); w.keyB = A.take<uint32_t>(N); w.valA = A.take<uint32_t>(N); w.valB = A.take<uint32_t>(N);
    w.rankA = A.take<uint32_t>(N); w.rankB = A.take<uint32_t>(N);
    w.L = A.take<uint8_t>(N); w.M = A.take<uint8_t>(N); w.flags = A.take<uint8_t>(N); w.flags2 = A.take<uint8_t>(N);
    w.R = A.take<uint8_t>((uint64_t) nb * RS);
    w.hist = A.take<uint32_t>(radix_hist_bytes(S, nb) / 4);
    const uint64_t tiles = bra_div_up(S, 4096);
    w.tile_last = A.take<int>(nb * tiles);
    w.t_first_head = A.take<int>(nb * tiles); w.t_last_head = A.take<int>(nb * tiles);
    w.t_first_nl = A.take<int>(nb * tiles); w.t_last_nl = A.take<int>(nb * tiles);
    w.t_cnt = A.take<uint32_t>(nb * tiles + 64);  // (+ the RLE emit kernel's ticket word)
    w.t_bits = A.take<uint32_t>((uint64_t) nb * huf_enc_tiles((uint32_t) RS));
    const uint64_t segs = mtf_segments(S);
    w.summ = A.take<uint8_t>(nb * segs * 256); w.state = A.take<uint8_t>(nb * segs * 256); w.scnt = A.take<uint16_t>(nb * segs);
    w.len = A.take<uint32_t>(nb); w.primary = A.take<uint32_t>(nb); w.period = A.take<uint32_t>(nb); w.ngroups = A.take<uint32_t>(nb);
    w.notdone = A.take<uint32_t>(4); w.done = A.take<uint8_t>(nb); w.fin = A.take<uint8_t>(nb); w.finskip = A.take<uint8_t>(nb);
    w.maxgroup = A.take<uint32_t>(nb); w.sumsq = A.take<unsigned long long>(nb);
    w.div_vals = A.take<uint32_t>(BRA_DIV_CAP); w.div_off = A.take<uint32_t>(nb); w.div_cnt = A.take<uint32_t>(nb);
    w.alpha = A.take<uint8_t>((uint64_t) nb * 256);
    w.bad = A.take<uint8_t>((uint64_t) nb * BRA_BAD_STRIDE);
    w.rlen = A.take<uint32_t>(nb); w.clen = A.take<uint32_t>(nb); w.rhist = A.take<uint32_t>((uint64_t) nb * 256);
    w.codes = A.take<uint32_t>((uint64_t) nb * 256); w.ok = A.take<uint32_t>(nb);
}

static void carve_dec(Arena& A, uint32_t S, uint32_t nb, DecWs& w)
{
    const uint64_t N = (uint64_t) nb * S, RS = rle_stride_for(S), PS = pay_stride_for(S);
    w.W = A.take<uint32_t>(N);
    w.R = A.take<uint8_t>((uint64_t) nb * RS); w.M = A.take<uint8_t>(N); w.L = A.take<uint8_t>(N);
    w.hist = A.take<uint32_t>(radix_hist_bytes(S, nb) / 4);
    const uint64_t segs = mtf_segments(S);
    w.summ = A.take<uint8_t>(nb * segs * 256); w.state = A.take<uint8_t>(nb * segs * 256);
    const uint64_t seqs = huf_dec_seqs((uint32_t) PS);
    w.sub_start = A.take<uint8_t>(nb * seqs * huf_dec_subs_per_seq()); w.sub_count = A.take<uint16_t>(nb * seqs * huf_dec_subs_per_seq());
    w.phase = A.take<uint8_t>(nb * seqs * huf_dec_phase_bytes_per_seq());
    w.seq_entry = A.take<uint32_t>(nb * seqs); w.seq_exit = A.take<uint32_t>(nb * seqs); w.seq_count = A.take<uint32_t>(nb * seqs);
    const uint64_t rt = rle_dec_tiles((uint32_t) RS);
    w.t_exit = A.take<uint8_t>(nb * rt * rle_dec_entries()); w.t_entry = A.take<uint8_t>(nb * rt);
    w.t_tok = A.take<uint32_t>(nb * rt * 32); w.t_ocnt = A.take<uint32_t>(nb * rt);
    const uint64_t km = ibwt_kmax(S);
    w.walk = A.take<uint2>(nb * km); w.woff = A.take<uint32_t>(nb * km); w.orbit = A.take<uint32_t>(nb); w.ovf = A.take<uint32_t>((uint64_t) nb * ibwt_ovf_words());
    w.wtmp = A.take<uint8_t>(nb * km * ibwt_tmp_cap(S));  // scratch rows of the inverse-BWT walks (about 8 bytes per input byte)
    w.rlen = A.take<uint32_t>(nb); w.clen = A.take<uint32_t>(nb); w.nlen = A.take<uint32_t>(nb); w.primary = A.take<uint32_t>(nb);
    w.err = A.take<uint32_t>(nb); w.end_bit = A.take<uint32_t>(nb); w.changed = A.take<uint32_t>(4);
    w.tabs = A.take<bra_huf_dec_t>(nb);
}

}  // namespace bra

using namespace bra;

struct bra_b200_ctx
{
    int          device = 0;
    uint32_t     block = 0, max_batch = 0;
    uint64_t     rle_stride = 0, pay_stride = 0;
    Arena        arena;
    cudaStream_t own_stream = nullptr;
    // pinned staging for the host path
    uint8_t* h_stage = nullptr;
    uint64_t h_stage_bytes = 0;
    uint8_t* d_io = nullptr;  // device staging for the host path: input blocks / headers / payloads / output
    uint64_t d_io_bytes = 0;
    uint32_t last_rounds = 0, last_sweeps = 0;
    uint64_t last_launches = 0;
    // pinned, device-visible host words for the small transfers of the control loops (mail_fetch / mail_publish):
    // [bwt: 2 + BRA_DIV_CAP + 2*max_batch][huffman: 16][block lengths: max_batch][host path: 8*max_batch + 16]
    uint32_t* h_mail = nullptr;
    uint64_t  crc_pending = 0;  // bytes of the CRC submission in flight (bra_b200_crc32c_submit), 0 = none
};

static inline uint32_t* mail_bwt(bra_b200_ctx* c) { return c->h_mail; }
static inline uint32_t* mail_huffman(bra_b200_ctx* c) { return c->h_mail + 2 + BRA_DIV_CAP + 2 * (size_t) c->max_batch; }
static inline uint32_t* mail_lengths(bra_b200_ctx* c) { return mail_huffman(c) + 16; }

// ---- small glue kernels -------------------------------------------------------------------------
__global__ void mail_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t words)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) dst[i] = src[i];
}
namespace bra {
bool mail_fetch(uint32_t* d_dst, const uint32_t* h_src, uint32_t words, cudaStream_t st)
{
    if (words == 0) return true;
    BRA_LAUNCH(P_GLUE, st, mail_copy_kernel<<<bra_div_up(words, 256), 256, 0, st>>>(d_dst, h_src, words));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}
bool mail_publish(uint32_t* h_dst, const uint32_t* d_src, uint32_t words, cudaStream_t st)
{
    if (words == 0) return true;
    BRA_LAUNCH(P_GLUE, st, mail_copy_kernel<<<bra_div_up(words, 256), 256, 0, st>>>(h_dst, d_src, words));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}
}  // namespace bra
__global__ void set_primary_kernel(uint8_t* __restrict__ hdr, const uint32_t* __restrict__ primary, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint32_t p = primary[b];
    uint8_t*       h = hdr + (uint64_t) b * 268;
    h[0] = (uint8_t) p; h[1] = (uint8_t) (p >> 8); h[2] = (uint8_t) (p >> 16); h[3] = (uint8_t) (p >> 24);
}

// parse + validate headers on the device (reference chunks.c:31-48, with the run-time block size)
__global__ void parse_hdr_kernel(const uint8_t* __restrict__ hdr, uint32_t nblk, uint32_t max_r, uint32_t max_c, uint32_t* __restrict__ rlen,
                                 uint32_t* __restrict__ clen, uint32_t* __restrict__ primary, uint32_t* __restrict__ err)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t* h = hdr + (uint64_t) b * 268;
    auto           u32 = [&](int o) { return (uint32_t) h[o] | ((uint32_t) h[o + 1] << 8) | ((uint32_t) h[o + 2] << 16) | ((uint32_t) h[o + 3] << 24); };
    const uint32_t pi = u32(0), r = u32(260), c = u32(264);
    const bool     bad = r == 0 || c == 0 || r > max_r || c > max_c;
    rlen[b]    = bad ? 0u : r;
    clen[b]    = bad ? 0u : c;
    primary[b] = pi;
    err[b]     = bad ? 1u : 0u;
}

// after RLE decode: errors zero the block length so that later stages skip it; primary must be < n
__global__ void post_rle_kernel(uint32_t* __restrict__ nlen, const uint32_t* __restrict__ primary, uint32_t* __restrict__ err, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    if (err[b] || nlen[b] == 0 || primary[b] >= nlen[b])  // reference chunks.c:385-389
    {
        err[b]  = 1;
        nlen[b] = 0;
    }
}
__global__ void zero_len_on_err_kernel(uint32_t* __restrict__ len, const uint32_t* __restrict__ err, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk && err[b]) len[b] = 0;
}

// ---- context --------------------------------------------------------------------------------------
extern "C" int bra_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
    return n;
}

extern "C" bra_b200_ctx_t* bra_b200_ctx_create(int device, uint32_t block_size, uint32_t max_batch)
{
    if (block_size == 0 || block_size > BRA_B200_MAX_BLOCK || (block_size % 16) != 0 || max_batch == 0 || max_batch > 32768)
    {
        bra_b200_log_error("bra_b200_ctx_create: invalid block_size %u / max_batch %u", block_size, max_batch);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    {
        bra_b200_log_error("bra_b200_ctx_create: no usable CUDA device %d (found %d); there is no CPU fallback", device, ndev);
        return nullptr;
    }
    bra_b200_ctx* c = new bra_b200_ctx();
    c->device     = device;
    c->block      = block_size;
    c->max_batch  = max_batch;
    c->rle_stride = rle_stride_for(block_size);
    c->pay_stride = pay_stride_for(block_size);
    BraDeviceGuard dg(device);
    if (!dg.ok) { delete c; return nullptr; }
    if (!crc_init_tables()) { delete c; return nullptr; }  // per device, thread-safe
    // size the arena by dry-running both carves
    Arena dry;
    EncWs ew;
    DecWs dw;
    carve_enc(dry, block_size, max_batch, ew);
    const uint64_t need_enc = dry.used;
    dry.used = 0;
    carve_dec(dry, block_size, max_batch, dw);
    const uint64_t need = std::max(need_enc, dry.used) + 4096;
    if (cudaMalloc(&c->arena.base, need) != cudaSuccess)
    {
        bra_b200_log_error("bra_b200_ctx_create: cudaMalloc of %llu workspace bytes failed", (unsigned long long) need);
        delete c;
        return nullptr;
    }
    c->arena.cap = need;
    const size_t mail_words = (2 + BRA_DIV_CAP + 2 * (size_t) max_batch) + 16 + max_batch + (8 * (size_t) max_batch + 16);
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&c->h_mail), mail_words * 4, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
    {
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        cudaFree(c->arena.base);
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void bra_b200_ctx_destroy(bra_b200_ctx_t* c)
{
    if (!c) return;
    BraDeviceGuard dg(c->device);
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->d_io) cudaFree(c->d_io);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_mail) cudaFreeHost(c->h_mail);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" uint32_t bra_b200_block_size(const bra_b200_ctx_t* c) { return c ? c->block : 0; }
extern "C" uint32_t bra_b200_max_batch(const bra_b200_ctx_t* c) { return c ? c->max_batch : 0; }
extern "C" uint64_t bra_b200_payload_stride(const bra_b200_ctx_t* c) { return c ? c->pay_stride : 0; }
extern "C" uint64_t bra_b200_workspace_bytes(const bra_b200_ctx_t* c) { return c ? c->arena.cap + c->d_io_bytes : 0; }
extern "C" void     bra_b200_last_stats(const bra_b200_ctx_t* c, uint32_t* r, uint32_t* s, uint64_t* l)
{
    if (!c) return;
    if (r) *r = c->last_rounds;
    if (s) *s = c->last_sweeps;
    if (l) *l = c->last_launches;
}

// ---- encode one batch (nb <= max_batch) ---------------------------------------------------------------
namespace bra {

bool encode_batch(bra_b200_ctx* c, const uint8_t* d_in, uint32_t nb, uint32_t last_len, uint8_t* d_hdr, uint8_t* d_payload, uint32_t* d_crc,
                  cudaStream_t st)
{
    const uint32_t S = c->block;
    Arena          A = c->arena;
    A.used           = 0;
    EncWs w;
    carve_enc(A, S, nb, w);

    std::vector<uint32_t> h_len(nb, S);
    h_len[nb - 1] = last_len;
    uint32_t* mail_len = mail_lengths(c);
    memcpy(mail_len, h_len.data(), (size_t) nb * 4);
    if (!mail_fetch(w.len, mail_len, nb, st)) return false;  // rewritten by the next batch only, after many stream syncs

    if (!crc_blocks(d_in, S, w.len, 0, S, nb, nullptr, d_crc, st)) return false;

    BwtFwdArgs ba{};
    ba.d_in = d_in; ba.d_out = w.L; ba.stride = S; ba.d_len = w.len; ba.h_len = h_len.data(); ba.max_n = S; ba.nblk = nb;
    ba.d_primary = w.primary;
    ba.d_keyA = w.keyA; ba.d_keyB = w.keyB; ba.d_valA = w.valA; ba.d_valB = w.valB; ba.d_rankA = w.rankA; ba.d_rankB = w.rankB;
    ba.d_flags = w.flags; ba.d_flags2 = w.flags2; ba.d_hist = w.hist; ba.d_tile_last = w.tile_last;
    ba.d_period = w.period; ba.d_ngroups = w.ngroups; ba.d_notdone = w.notdone; ba.d_done = w.done; ba.d_fin = w.fin; ba.d_finskip = w.finskip; ba.d_maxgroup = w.maxgroup; ba.d_sumsq = w.sumsq;
    ba.d_div_vals = w.div_vals; ba.d_div_off = w.div_off; ba.d_div_cnt = w.div_cnt; ba.div_cap = BRA_DIV_CAP;
    ba.d_bad = w.bad; ba.bad_stride = BRA_BAD_STRIDE; ba.d_alpha = w.alpha; ba.d_tile_heads = w.t_cnt;  // (the RLE stage's per-tile counters: same tiling, not in use yet)
    uint32_t rounds = 0;
    ba.h_rounds = &rounds;
    ba.h_mail   = mail_bwt(c);
    if (!bwt_forward_batch(ba, st)) return false;
    c->last_rounds = std::max(c->last_rounds, rounds);

    if (!mtf_encode_batch(w.L, w.M, S, w.len, S, nb, w.summ, w.scnt, w.state, st)) return false;

    RleEncArgs ra{};
    ra.d_in = w.M; ra.stride = S; ra.d_len = w.len; ra.max_n = S; ra.nblk = nb;
    ra.d_out = w.R; ra.out_stride = c->rle_stride; ra.d_rlen = w.rlen; ra.d_hist = w.rhist;
    ra.d_t_first_head = w.t_first_head; ra.d_t_last_head = w.t_last_head; ra.d_t_first_nl = w.t_first_nl; ra.d_t_last_nl = w.t_last_nl;
    ra.d_t_cnt = w.t_cnt;
    if (!rle_encode_batch(ra, st)) return false;

    HufEncArgs ha{};
    ha.d_in = w.R; ha.stride = c->rle_stride; ha.d_rlen = w.rlen; ha.max_r = (uint32_t) std::min<uint64_t>(c->rle_stride, (uint64_t) S + S / 128 + 2);
    ha.nblk = nb; ha.compute_hist = false; ha.d_hist = w.rhist; ha.d_hdr = d_hdr; ha.d_codes = w.codes; ha.d_ok = w.ok;
    ha.d_t_bits = w.t_bits; ha.d_clen = w.clen; ha.d_out = d_payload; ha.out_stride = c->pay_stride;
    if (!huf_encode_batch(ha, st)) return false;

    BRA_LAUNCH(P_GLUE, st, set_primary_kernel<<<bra_div_up(nb, 128), 128, 0, st>>>(d_hdr, w.primary, nb));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

bool decode_batch(bra_b200_ctx* c, const uint8_t* d_hdr, const uint8_t* d_payload, uint32_t nb, uint32_t hint_r, uint32_t hint_c, uint8_t* d_out,
                  uint32_t* d_out_len, uint32_t* d_crc, uint32_t* d_status, cudaStream_t st, bool sizes_only)
{
    const uint32_t S = c->block;
    Arena          A = c->arena;
    A.used           = 0;
    DecWs w;
    carve_dec(A, S, nb, w);
    const uint32_t cap_r = (uint32_t) std::min<uint64_t>(c->rle_stride - 32, (uint64_t) S + S / 128 + 2);
    const uint32_t cap_c = (uint32_t) (c->pay_stride - 32);
    const uint32_t max_r = hint_r ? std::min(hint_r, cap_r) : cap_r;
    const uint32_t max_c = hint_c ? std::min(hint_c, cap_c) : cap_c;

    BRA_LAUNCH(P_GLUE, st, parse_hdr_kernel<<<bra_div_up(nb, 128), 128, 0, st>>>(d_hdr, nb, max_r, max_c, w.rlen, w.clen, w.primary, w.err));

    HufDecArgs ha{};
    ha.d_pay = d_payload; ha.pay_stride = c->pay_stride; ha.d_clen = w.clen; ha.d_hdr = d_hdr; ha.max_c = max_c; ha.nblk = nb;
    ha.d_out = w.R; ha.out_stride = c->rle_stride; ha.d_tabs = w.tabs; ha.d_err = w.err;
    ha.d_sub_start = w.sub_start; ha.d_sub_count = w.sub_count; ha.d_seq_entry = w.seq_entry; ha.d_seq_exit = w.seq_exit;
    ha.d_seq_count = w.seq_count; ha.d_end_bit = w.end_bit; ha.d_changed = w.changed; ha.d_phase = w.phase;
    uint32_t sweeps = 0;
    ha.h_sweeps = &sweeps;
    ha.h_mail   = mail_huffman(c);
    if (!huf_decode_batch(ha, st)) return false;
    c->last_sweeps = std::max(c->last_sweeps, sweeps);
    BRA_LAUNCH(P_GLUE, st, zero_len_on_err_kernel<<<bra_div_up(nb, 128), 128, 0, st>>>(w.rlen, w.err, nb));

    RleDecArgs ra{};
    ra.d_in = w.R; ra.stride = c->rle_stride; ra.d_rlen = w.rlen; ra.max_r = max_r; ra.nblk = nb;
    ra.d_out = w.M; ra.out_stride = S; ra.out_cap = S; ra.d_nlen = w.nlen; ra.d_err = w.err;
    ra.d_t_exit = w.t_exit; ra.d_t_entry = w.t_entry; ra.d_t_tok = w.t_tok; ra.d_t_ocnt = w.t_ocnt; ra.size_only = sizes_only;
    BRA_CUDA_TRY(cudaMemsetAsync(w.nlen, 0, nb * 4, st));
    if (sizes_only)
    {
        // list mode (reference chunks.c:369-373): Huffman decode, then bra_rle_decode_compute_size only. A chunk whose
        // Huffman stream is broken fails; one whose RLE tokens are truncated counts as 0 bytes, as in the reference.
        BRA_CUDA_TRY(cudaMemcpyAsync(d_status, w.err, nb * 4, cudaMemcpyDeviceToDevice, st));
        if (!rle_decode_batch(ra, st)) return false;
        BRA_CUDA_TRY(cudaMemcpyAsync(d_out_len, w.nlen, nb * 4, cudaMemcpyDeviceToDevice, st));
        BRA_CUDA_TRY(cudaMemsetAsync(d_crc, 0, nb * 4, st));
        BRA_CUDA_TRY(cudaGetLastError());
        return true;
    }
    if (!rle_decode_batch(ra, st)) return false;
    BRA_LAUNCH(P_GLUE, st, post_rle_kernel<<<bra_div_up(nb, 128), 128, 0, st>>>(w.nlen, w.primary, w.err, nb));

    if (!mtf_decode_batch(w.M, w.L, S, w.nlen, S, nb, w.summ, w.state, st)) return false;

    BwtInvArgs ia{};
    ia.d_in = w.L; ia.d_out = d_out; ia.stride = S; ia.d_len = w.nlen; ia.d_primary = w.primary; ia.max_n = S; ia.nblk = nb;
    ia.d_W = w.W; ia.d_hist = w.hist; ia.d_walk = w.walk; ia.d_woff = w.woff; ia.d_orbit = w.orbit; ia.d_tmp = w.wtmp; ia.d_ovf = w.ovf;
    if (!bwt_inverse_batch(ia, st)) return false;

    if (!crc_blocks(d_out, S, w.nlen, 0, S, nb, nullptr, d_crc, st)) return false;
    BRA_CUDA_TRY(cudaMemcpyAsync(d_out_len, w.nlen, nb * 4, cudaMemcpyDeviceToDevice, st));
    BRA_CUDA_TRY(cudaMemcpyAsync(d_status, w.err, nb * 4, cudaMemcpyDeviceToDevice, st));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

}  // namespace bra

// ---- public device-resident entry points -----------------------------------------------------------
extern "C" int bra_b200_encode_device(bra_b200_ctx_t* c, const uint8_t* d_in, uint32_t nblk, uint32_t last_len, uint8_t* d_hdr,
                                      uint8_t* d_payload, uint32_t* d_crc_raw, void* stream)
{
    if (!c || !d_in || !d_hdr || !d_payload || !d_crc_raw || nblk == 0 || last_len == 0 || last_len > c->block ||
        (reinterpret_cast<uintptr_t>(d_in) & 15u))
    {
        bra_b200_log_error("bra_b200_encode_device: invalid arguments");
        return 1;
    }
    BraDeviceGuard dg(c->device);
    if (!dg.ok) return 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    c->last_rounds  = 0;
    const uint64_t launches0 = prof_total_launches();
    for (uint32_t b0 = 0; b0 < nblk; b0 += c->max_batch)
    {
        const uint32_t nb = std::min(c->max_batch, nblk - b0);
        const uint32_t ll = (b0 + nb == nblk) ? last_len : c->block;
        if (!encode_batch(c, d_in + (uint64_t) b0 * c->block, nb, ll, d_hdr + (uint64_t) b0 * 268, d_payload + (uint64_t) b0 * c->pay_stride,
                          d_crc_raw + b0, st))
            return 3;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return 4;
    c->last_launches = prof_total_launches() - launches0;
    return 0;
}

extern "C" int bra_b200_decode_device(bra_b200_ctx_t* c, const uint8_t* d_hdr, const uint8_t* d_payload, uint32_t nblk, uint32_t hint_max_r,
                                      uint32_t hint_max_c, uint8_t* d_out, uint32_t* d_out_len, uint32_t* d_crc_raw, uint32_t* d_status,
                                      void* stream)
{
    if (!c || !d_hdr || !d_payload || !d_out || !d_out_len || !d_crc_raw || !d_status || nblk == 0)
    {
        bra_b200_log_error("bra_b200_decode_device: invalid arguments");
        return 1;
    }
    BraDeviceGuard dg(c->device);
    if (!dg.ok) return 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    c->last_sweeps  = 0;
    const uint64_t launches0 = prof_total_launches();
    for (uint32_t b0 = 0; b0 < nblk; b0 += c->max_batch)
    {
        const uint32_t nb = std::min(c->max_batch, nblk - b0);
        if (!decode_batch(c, d_hdr + (uint64_t) b0 * 268, d_payload + (uint64_t) b0 * c->pay_stride, nb, hint_max_r, hint_max_c,
                          d_out + (uint64_t) b0 * c->block, d_out_len + b0, d_crc_raw + b0, d_status + b0, st, false))
            return 3;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return 4;
    c->last_launches = prof_total_launches() - launches0;
    return 0;
}

// ---- streaming CRC-32C of host memory (the STORED path, reference chunks.c:114-167) -------------------------
extern "C" int bra_b200_crc32c_submit(bra_b200_ctx_t* c, const void* data, uint64_t len)
{
    if (!c || !data || len == 0 || len > (1ull << 30) || c->crc_pending)
    {
        bra_b200_log_error("bra_b200_crc32c_submit: invalid arguments or a submission is already in flight");
        return 1;
    }
    BraDeviceGuard dg(c->device);
    if (!dg.ok) return 2;
    const uint64_t padded = (len + 15) / 16 * 16;
    uint8_t*       io     = ctx_io_buffer(c, padded + 256);
    if (!io) return 3;
    cudaStream_t st    = c->own_stream;
    uint32_t*    d_crc = reinterpret_cast<uint32_t*>(io + padded);
    if (cudaMemcpyAsync(io, data, len, cudaMemcpyHostToDevice, st) != cudaSuccess) return 4;
    if (!crc_blocks(io, padded, nullptr, (uint32_t) len, (uint32_t) len, 1, nullptr, d_crc, st)) return 5;
    if (!mail_publish(ctx_mail_host(c), d_crc, 1, st)) return 5;
    c->crc_pending = len;
    return 0;
}

extern "C" int bra_b200_crc32c_finish(bra_b200_ctx_t* c, uint32_t* crc)
{
    if (!c || c->crc_pending == 0) return 1;
    BraDeviceGuard dg(c->device);
    if (!dg.ok) return 2;
    const uint64_t len = c->crc_pending;
    c->crc_pending     = 0;
    if (cudaStreamSynchronize(c->own_stream) != cudaSuccess) return 4;
    if (crc)
    {
        const uint32_t piece = *reinterpret_cast<volatile uint32_t*>(ctx_mail_host(c));  // crc32c(data, len, 0)
        *crc                 = bra_crc_combine(crc_host_pow(), *crc, piece, len);        // == crc32c(data, len, *crc)
    }
    return 0;
}

namespace bra {
uint8_t* ctx_io_buffer(bra_b200_ctx* c, uint64_t bytes)
{
    if (bytes <= c->d_io_bytes) return c->d_io;
    if (c->d_io) cudaFree(c->d_io);
    c->d_io       = nullptr;
    c->d_io_bytes = 0;
    if (cudaMalloc(&c->d_io, bytes) != cudaSuccess)
    {
        bra_b200_log_error("bra_b200: cudaMalloc of %llu staging bytes failed", (unsigned long long) bytes);
        return nullptr;
    }
    c->d_io_bytes = bytes;
    return c->d_io;
}
cudaStream_t ctx_stream(bra_b200_ctx* c) { return c->own_stream; }
uint32_t*    ctx_mail_host(bra_b200_ctx* c) { return mail_lengths(c) + c->max_batch; }  // 8*max_batch + 16 words for the host path
int          ctx_device(const bra_b200_ctx* c) { return c->device; }
}  // namespace bra
