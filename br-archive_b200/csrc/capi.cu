// capi.cu -- the reference's per-stage C API (include/encoders/*.h, include/utils/lib_bra_crc32c.h,
// include/lib_bra.h) implemented on the GPU: host pointer in, H2D, the same batched kernels with a
// batch of one block, D2H into plain malloc memory the caller frees -- the ownership rules of
// reference src/encoders/*.c. Failures return false / NULL / 0 and log; nothing aborts.
//
// These calls move one block over PCIe each and are latency bound by construction; they exist for
// drop-in correctness (the reference's unit tests run against them unchanged). Throughput comes
// from the batched entry points in pipeline.cu / hostpath.cu.
#include "bra_common.cuh"
#include "bra_kernels.h"
#include "pipeline.h"

#include <encoders/bra_bwt.h>
#include <encoders/bra_huffman.h>
#include <encoders/bra_mtf.h>
#include <encoders/bra_rle.h>
#include <lib_bra.h>
#include <utils/lib_bra_crc32c.h>

#include <mutex>
#include <stdlib.h>
#include <string.h>

using namespace bra;

namespace {

std::mutex g_mu;
bool       g_ready = false;
uint8_t*   g_buf   = nullptr;  // device scratch, grown on demand
uint64_t   g_cap   = 0;
cudaStream_t g_st  = nullptr;
int        g_dev   = 0;  // the device that was current at the first call: the per-stage API stays on it (scratch, stream, CRC tables)

bool stage_init()
{
    if (g_ready) return true;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        bra_b200_log_error("bra_b200: no CUDA device available and there is no CPU fallback");
        return false;
    }
    BRA_CUDA_TRY(cudaGetDevice(&g_dev));
    if (!crc_init_tables()) return false;
    BRA_CUDA_TRY(cudaStreamCreateWithFlags(&g_st, cudaStreamNonBlocking));
    g_ready = true;
    return true;
}

// every per-stage entry: library initialised, and the stage device current until the entry returns
#define STAGE_ENTER(fail)            \
    if (!stage_init()) return fail;  \
    BraDeviceGuard _dg(g_dev);       \
    if (!_dg.ok) return fail

struct Bump
{
    uint64_t used = 0;
    template <typename T>
    uint64_t take(uint64_t count)
    {
        used = (used + 255) / 256 * 256;
        const uint64_t o = used;
        used += count * sizeof(T);
        return o;
    }
};

bool scratch(uint64_t bytes)
{
    if (bytes <= g_cap) return true;
    if (g_buf) cudaFree(g_buf);
    g_buf = nullptr;
    g_cap = 0;
    const uint64_t want = bytes + bytes / 4 + (1 << 20);
    if (cudaMalloc(&g_buf, want) != cudaSuccess)
    {
        bra_b200_log_error("bra_b200: cudaMalloc of %llu scratch bytes failed", (unsigned long long) want);
        return false;
    }
    g_cap = want;
    return true;
}

template <typename T>
T* at(uint64_t off) { return reinterpret_cast<T*>(g_buf + off); }

inline uint64_t pad16(uint64_t n) { return (n + 15) / 16 * 16; }

bool sync_ok()
{
    BRA_CUDA_TRY(cudaStreamSynchronize(g_st));
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// lib_bra.h
// ------------------------------------------------------------------------------------------------
extern "C" __attribute__((weak)) bool bra_init(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return stage_init();
}
extern "C" __attribute__((weak)) bool bra_quit(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_buf)
    {
        BraDeviceGuard dg(g_dev);
        cudaFree(g_buf);
    }
    g_buf = nullptr;
    g_cap = 0;
    return true;
}
extern "C" __attribute__((weak)) bool bra_has_sse42(void) { return false; }  // no CPU CRC path exists in this library

// ------------------------------------------------------------------------------------------------
// CRC-32C
// ------------------------------------------------------------------------------------------------
static uint32_t crc_device(const void* data, uint64_t length, uint32_t prev)
{
    if (length == 0 || data == nullptr) return prev;  // reference: empty/NULL input leaves the CRC unchanged (test_bra_crc32c.cpp:38-43)
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(0);
    // inputs above 2^32-1 bytes are folded piecewise (kernel lengths are 32-bit)
    const uint8_t* p   = static_cast<const uint8_t*>(data);
    uint32_t       crc = prev;
    while (length)
    {
        const uint32_t n = (uint32_t) std::min<uint64_t>(length, 1u << 30);
        Bump           B;
        const uint64_t o_in = B.take<uint8_t>(pad16(n)), o_prev = B.take<uint32_t>(1), o_crc = B.take<uint32_t>(1);
        if (!scratch(B.used)) return 0;
        if (cudaMemcpyAsync(at<uint8_t>(o_in), p, n, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return 0;
        if (cudaMemcpyAsync(at<uint32_t>(o_prev), &crc, 4, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return 0;
        if (!crc_blocks(at<uint8_t>(o_in), pad16(n), nullptr, n, n, 1, at<uint32_t>(o_prev), at<uint32_t>(o_crc), g_st)) return 0;
        if (cudaMemcpyAsync(&crc, at<uint32_t>(o_crc), 4, cudaMemcpyDeviceToHost, g_st) != cudaSuccess) return 0;
        if (!sync_ok()) return 0;
        p += n;
        length -= n;
    }
    return crc;
}

extern "C" uint32_t bra_crc32c(const void* data, const uint64_t length, const uint32_t previous_crc) { return crc_device(data, length, previous_crc); }
extern "C" uint32_t bra_crc32c_table(const void* data, const uint64_t length, const uint32_t previous_crc) { return crc_device(data, length, previous_crc); }
extern "C" uint32_t bra_crc32c_sse42(const void* data, const uint64_t length, const uint32_t previous_crc) { return crc_device(data, length, previous_crc); }
extern "C" void     bra_crc32c_use_sse42(const bool) {}
extern "C" uint32_t bra_crc32c_combine(uint32_t crc32a, uint32_t crc32b, uint32_t len_b)
{
    return bra_crc_combine(crc_host_pow(), crc32a, crc32b, len_b);
}

// ------------------------------------------------------------------------------------------------
// BWT
// ------------------------------------------------------------------------------------------------
extern "C" bool bra_bwt_encode2(const uint8_t* buf, const bra_bwt_index_t n, bra_bwt_index_t* primary_index, uint8_t* out_buf)
{
    if (!buf || !primary_index || !out_buf || n == 0 || n > BRA_B200_MAX_BLOCK)
    {
        bra_b200_log_error("bra_bwt_encode2: invalid arguments (size %u)", n);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(false);
    const uint64_t S     = pad16(n);
    const uint64_t tiles = bra_div_up(S, 4096);
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(S), o_out = B.take<uint8_t>(S), o_flags = B.take<uint8_t>(S), o_flags2 = B.take<uint8_t>(S);
    uint64_t       o_u32[6];
    for (auto& o : o_u32) o = B.take<uint32_t>(S);
    const uint64_t o_hist = B.take<uint8_t>(radix_hist_bytes((uint32_t) S, 1)), o_tl = B.take<int>(tiles);
    const uint64_t o_small = B.take<uint32_t>(64), o_div = B.take<uint32_t>(4096), o_bad = B.take<uint8_t>(1024), o_done = B.take<uint8_t>(16);
    if (!scratch(B.used)) return false;
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st));
    uint32_t* sm = at<uint32_t>(o_small);  // [0]=len [1]=primary [2]=period [3]=ngroups [4]=notdone [5]=div_off [6]=div_cnt
    BRA_CUDA_TRY(cudaMemcpyAsync(sm, &n, 4, cudaMemcpyHostToDevice, g_st));
    BwtFwdArgs a{};
    a.d_in = at<uint8_t>(o_in); a.d_out = at<uint8_t>(o_out); a.stride = S; a.d_len = sm; a.h_len = &n; a.max_n = n; a.nblk = 1;
    a.d_primary = sm + 1;
    a.d_keyA = at<uint32_t>(o_u32[0]); a.d_keyB = at<uint32_t>(o_u32[1]); a.d_valA = at<uint32_t>(o_u32[2]); a.d_valB = at<uint32_t>(o_u32[3]);
    a.d_rankA = at<uint32_t>(o_u32[4]); a.d_rankB = at<uint32_t>(o_u32[5]);
    a.d_flags = at<uint8_t>(o_flags); a.d_flags2 = at<uint8_t>(o_flags2); a.d_hist = at<uint32_t>(o_hist); a.d_tile_last = at<int>(o_tl);
    a.d_period = sm + 2; a.d_ngroups = sm + 3; a.d_notdone = sm + 8; a.d_done = at<uint8_t>(o_done); a.d_fin = at<uint8_t>(o_done) + 4;
    a.d_finskip = at<uint8_t>(o_done) + 8; a.d_maxgroup = sm + 4; a.d_sumsq = reinterpret_cast<unsigned long long*>(sm + 12);
    // sm: [0]len [1]primary [2]period [3]ngroups [4]maxgroup [5]div_off [6]div_cnt [8,9]stat [12,13]sumsq
    a.d_div_vals = at<uint32_t>(o_div); a.d_div_off = sm + 5; a.d_div_cnt = sm + 6; a.div_cap = 4096; a.d_bad = at<uint8_t>(o_bad);
    a.bad_stride = 1024;
    if (!bwt_forward_batch(a, g_st)) return false;
    uint32_t pi = 0;
    BRA_CUDA_TRY(cudaMemcpyAsync(out_buf, a.d_out, n, cudaMemcpyDeviceToHost, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(&pi, a.d_primary, 4, cudaMemcpyDeviceToHost, g_st));
    if (!sync_ok()) return false;
    *primary_index = pi;
    return true;
}

extern "C" uint8_t* bra_bwt_encode(const uint8_t* buf, const bra_bwt_index_t n, bra_bwt_index_t* primary_index)
{
    uint8_t* out = static_cast<uint8_t*>(malloc(n ? n : 1));
    if (!out) return nullptr;
    if (!bra_bwt_encode2(buf, n, primary_index, out))
    {
        free(out);
        return nullptr;
    }
    return out;
}

static bool bwt_decode_impl(const uint8_t* buf, uint32_t n, uint32_t pi, uint32_t* transform, uint8_t* out_buf)
{
    if (!buf || !out_buf || n == 0 || n > BRA_B200_MAX_BLOCK || pi >= n)
    {
        bra_b200_log_error("bra_bwt_decode: invalid arguments (size %u, primary index %u)", n, pi);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(false);
    const uint64_t S = pad16(n), km = ibwt_kmax(n);
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(S), o_out = B.take<uint8_t>(S), o_W = B.take<uint32_t>(S);
    const uint64_t o_hist = B.take<uint8_t>(radix_hist_bytes((uint32_t) S, 1)), o_walk = B.take<uint2>(km), o_woff = B.take<uint32_t>(km);
    const uint64_t o_small = B.take<uint32_t>(16);
    if (!scratch(B.used)) return false;
    uint32_t* sm = at<uint32_t>(o_small);  // [0]=len [1]=primary [2]=orbit
    const uint32_t hv[2] = {n, pi};
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(sm, hv, 8, cudaMemcpyHostToDevice, g_st));
    BwtInvArgs a{};
    a.d_in = at<uint8_t>(o_in); a.d_out = at<uint8_t>(o_out); a.stride = S; a.d_len = sm; a.d_primary = sm + 1; a.max_n = n; a.nblk = 1;
    a.d_W = at<uint32_t>(o_W); a.d_hist = at<uint32_t>(o_hist); a.d_walk = at<uint2>(o_walk); a.d_woff = at<uint32_t>(o_woff); a.d_orbit = sm + 2;
    if (!bwt_inverse_batch(a, g_st)) return false;
    BRA_CUDA_TRY(cudaMemcpyAsync(out_buf, a.d_out, n, cudaMemcpyDeviceToHost, g_st));
    if (transform)
    {
        // the reference leaves its LF map in the caller's scratch (bra_bwt.c:155-159); W >> 8 is that map
        BRA_CUDA_TRY(cudaMemcpyAsync(transform, a.d_W, (size_t) n * 4, cudaMemcpyDeviceToHost, g_st));
    }
    if (!sync_ok()) return false;
    if (transform)
        for (uint32_t i = 0; i < n; ++i) transform[i] >>= 8;
    return true;
}

extern "C" void bra_bwt_decode2(const uint8_t* buf, const bra_bwt_index_t n, const bra_bwt_index_t primary_index, bra_bwt_index_t* transform,
                                uint8_t* out_buf)
{
    (void) bwt_decode_impl(buf, n, primary_index, transform, out_buf);
}

extern "C" uint8_t* bra_bwt_decode(const uint8_t* buf, const bra_bwt_index_t n, const bra_bwt_index_t primary_index)
{
    uint8_t* out = static_cast<uint8_t*>(malloc(n ? n : 1));
    if (!out) return nullptr;
    if (!bwt_decode_impl(buf, n, primary_index, nullptr, out))
    {
        free(out);
        return nullptr;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// MTF
// ------------------------------------------------------------------------------------------------
static bool mtf_impl(const uint8_t* buf, size_t n, uint8_t* out_buf, bool encode)
{
    if (!buf || !out_buf || n == 0 || n > 0x40000000u)
    {
        bra_b200_log_error("bra_mtf: invalid arguments (size %zu)", n);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(false);
    const uint32_t n32 = (uint32_t) n;
    const uint64_t S = pad16(n), segs = mtf_segments(n32);
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(S), o_out = B.take<uint8_t>(S), o_summ = B.take<uint8_t>(segs * 256), o_state = B.take<uint8_t>(segs * 256);
    const uint64_t o_scnt = B.take<uint16_t>(segs), o_len = B.take<uint32_t>(4);
    if (!scratch(B.used)) return false;
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint32_t>(o_len), &n32, 4, cudaMemcpyHostToDevice, g_st));
    const bool ok = encode ? mtf_encode_batch(at<uint8_t>(o_in), at<uint8_t>(o_out), S, at<uint32_t>(o_len), n32, 1, at<uint8_t>(o_summ),
                                              at<uint16_t>(o_scnt), at<uint8_t>(o_state), g_st)
                           : mtf_decode_batch(at<uint8_t>(o_in), at<uint8_t>(o_out), S, at<uint32_t>(o_len), n32, 1, at<uint8_t>(o_summ),
                                              at<uint8_t>(o_state), g_st);
    if (!ok) return false;
    BRA_CUDA_TRY(cudaMemcpyAsync(out_buf, at<uint8_t>(o_out), n, cudaMemcpyDeviceToHost, g_st));
    return sync_ok();
}

extern "C" bool bra_mtf_encode2(const uint8_t* buf, const size_t n, uint8_t* out_buf) { return mtf_impl(buf, n, out_buf, true); }
extern "C" void bra_mtf_decode2(const uint8_t* buf, const size_t n, uint8_t* out_buf) { (void) mtf_impl(buf, n, out_buf, false); }
extern "C" uint8_t* bra_mtf_encode(const uint8_t* buf, const size_t n)
{
    uint8_t* out = static_cast<uint8_t*>(malloc(n ? n : 1));
    if (!out) return nullptr;
    if (!mtf_impl(buf, n, out, true))
    {
        free(out);
        return nullptr;
    }
    return out;
}
extern "C" uint8_t* bra_mtf_decode(const uint8_t* buf, const size_t n)
{
    uint8_t* out = static_cast<uint8_t*>(malloc(n ? n : 1));
    if (!out) return nullptr;
    if (!mtf_impl(buf, n, out, false))
    {
        free(out);
        return nullptr;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// RLE
// ------------------------------------------------------------------------------------------------
extern "C" bool bra_rle_encode(const uint8_t* buf, const size_t n, uint8_t** out_buf, size_t* out_buf_size)
{
    if (out_buf) *out_buf = nullptr;
    if (out_buf_size) *out_buf_size = 0;
    if (!buf || !out_buf || !out_buf_size || n == 0 || n > 0x40000000u)
    {
        bra_b200_log_error("bra_rle_encode: invalid arguments (size %zu)", n);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(false);
    const uint32_t n32 = (uint32_t) n;
    const uint64_t S = pad16(n), RS = pad16(n + n / 128 + 64), tiles = rle_enc_tiles(n32);
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(S), o_out = B.take<uint8_t>(RS), o_hist = B.take<uint32_t>(256);
    uint64_t       o_t[5];
    for (auto& o : o_t) o = B.take<int>(tiles);
    const uint64_t o_small = B.take<uint32_t>(8);
    if (!scratch(B.used)) return false;
    uint32_t* sm = at<uint32_t>(o_small);  // [0]=len [1]=rlen
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(sm, &n32, 4, cudaMemcpyHostToDevice, g_st));
    RleEncArgs a{};
    a.d_in = at<uint8_t>(o_in); a.stride = S; a.d_len = sm; a.max_n = n32; a.nblk = 1; a.d_out = at<uint8_t>(o_out); a.out_stride = RS;
    a.d_rlen = sm + 1; a.d_hist = at<uint32_t>(o_hist);
    a.d_t_first_head = at<int>(o_t[0]); a.d_t_last_head = at<int>(o_t[1]); a.d_t_first_nl = at<int>(o_t[2]); a.d_t_last_nl = at<int>(o_t[3]);
    a.d_t_cnt = at<uint32_t>(o_t[4]);
    if (!rle_encode_batch(a, g_st)) return false;
    uint32_t r = 0;
    BRA_CUDA_TRY(cudaMemcpyAsync(&r, sm + 1, 4, cudaMemcpyDeviceToHost, g_st));
    if (!sync_ok()) return false;
    if (r == 0) return false;
    uint8_t* o = static_cast<uint8_t*>(malloc(r));
    if (!o) return false;
    if (cudaMemcpy(o, a.d_out, r, cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        free(o);
        return false;
    }
    *out_buf      = o;
    *out_buf_size = r;
    return true;
}

static bool rle_decode_impl(const uint8_t* buf, size_t n, bool size_only, uint8_t** out_buf, size_t* out_size)
{
    *out_size = 0;
    if (out_buf) *out_buf = nullptr;
    if (!buf || n == 0) return false;
    // tile output counts are summed in 32 bits and a run token expands 64x: 64 MiB of input can never wrap them
    // (the format's chunks are at most 16 MiB + 1/128)
    if (n > 0x03FFFFFFu)
    {
        bra_b200_log_error("bra_rle_decode: input of %zu bytes is above this implementation's limit", n);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(false);
    const uint32_t r32 = (uint32_t) n;
    const uint64_t RS = pad16(n + 16), tiles = rle_dec_tiles(r32);
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(RS), o_exit = B.take<uint8_t>(tiles * rle_dec_entries()), o_entry = B.take<uint8_t>(tiles);
    const uint64_t o_tok = B.take<uint32_t>(tiles * 32), o_ocnt = B.take<uint32_t>(tiles), o_small = B.take<uint32_t>(8);
    if (!scratch(B.used)) return false;
    uint32_t*      sm    = at<uint32_t>(o_small);  // [0]=rlen [1]=nlen [2]=err
    const uint32_t hv[3] = {r32, 0, 0};
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(sm, hv, 12, cudaMemcpyHostToDevice, g_st));
    RleDecArgs a{};
    a.d_in = at<uint8_t>(o_in); a.stride = RS; a.d_rlen = sm; a.max_r = r32; a.nblk = 1; a.d_out = nullptr; a.out_stride = 0; a.out_cap = 0xFFFFFFFFu;
    a.d_nlen = sm + 1; a.d_err = sm + 2; a.d_t_exit = at<uint8_t>(o_exit); a.d_t_entry = at<uint8_t>(o_entry); a.d_t_tok = at<uint32_t>(o_tok);
    a.d_t_ocnt = at<uint32_t>(o_ocnt); a.size_only = true;
    if (!rle_decode_batch(a, g_st)) return false;
    uint32_t hn[2] = {0, 0};
    BRA_CUDA_TRY(cudaMemcpyAsync(hn, sm + 1, 8, cudaMemcpyDeviceToHost, g_st));
    if (!sync_ok()) return false;
    if (hn[1] != 0 || hn[0] == 0) return false;  // truncated token or empty output (bra_rle.c:136,148,176-178)
    *out_size = hn[0];
    if (size_only) return true;

    // second phase: expand into a device buffer sized from the first phase (the scratch may move, so redo the carve)
    const uint64_t NS = pad16(hn[0]);
    Bump           B2;
    const uint64_t p_in = B2.take<uint8_t>(RS), p_exit = B2.take<uint8_t>(tiles * rle_dec_entries()), p_entry = B2.take<uint8_t>(tiles);
    const uint64_t p_tok = B2.take<uint32_t>(tiles * 32), p_ocnt = B2.take<uint32_t>(tiles), p_small = B2.take<uint32_t>(8), p_out = B2.take<uint8_t>(NS);
    if (B2.used > g_cap)
    {
        // growing reallocates: stage everything again
        if (!scratch(B2.used)) return false;
    }
    sm = at<uint32_t>(p_small);
    BRA_CUDA_TRY(cudaMemcpyAsync(at<uint8_t>(p_in), buf, n, cudaMemcpyHostToDevice, g_st));
    BRA_CUDA_TRY(cudaMemcpyAsync(sm, hv, 12, cudaMemcpyHostToDevice, g_st));
    a.d_in = at<uint8_t>(p_in); a.d_rlen = sm; a.d_nlen = sm + 1; a.d_err = sm + 2; a.d_t_exit = at<uint8_t>(p_exit); a.d_t_entry = at<uint8_t>(p_entry);
    a.d_t_tok = at<uint32_t>(p_tok); a.d_t_ocnt = at<uint32_t>(p_ocnt); a.d_out = at<uint8_t>(p_out); a.out_stride = NS; a.out_cap = hn[0];
    a.size_only = false;
    if (!rle_decode_batch(a, g_st)) return false;
    uint8_t* o = static_cast<uint8_t*>(malloc(hn[0]));
    if (!o) return false;
    if (cudaMemcpyAsync(o, a.d_out, hn[0], cudaMemcpyDeviceToHost, g_st) != cudaSuccess || !sync_ok())
    {
        free(o);
        return false;
    }
    *out_buf = o;
    return true;
}

extern "C" size_t bra_rle_decode_compute_size(const uint8_t* buf, const size_t n)
{
    size_t s = 0;
    return rle_decode_impl(buf, n, true, nullptr, &s) ? s : 0;
}

extern "C" bool bra_rle_decode(const uint8_t* buf, const size_t n, uint8_t** out_buf, size_t* out_buf_size)
{
    if (!out_buf || !out_buf_size) return false;
    return rle_decode_impl(buf, n, false, out_buf, out_buf_size);
}

// ------------------------------------------------------------------------------------------------
// Huffman
// ------------------------------------------------------------------------------------------------
extern "C" bra_huffman_chunk_t* bra_huffman_encode(const uint8_t* buf, const uint32_t n)
{
    if (!buf) return nullptr;
    if (n == 0)
    {
        bra_b200_log_error("unable to build huffman tree");  // reference bra_huffman.c:155-156,180 on empty input
        return nullptr;
    }
    if (n > 0x20000000u)
    {
        bra_b200_log_error("bra_huffman_encode: input of %u bytes is above this implementation's limit", n);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(nullptr);
    const uint64_t RS = pad16(n), tiles = huf_enc_tiles(n);
    const uint64_t PS = pad16((uint64_t) n * 34 / 8 + 64);  // payload bound for ANY length assignment the tree can produce
    Bump           B;
    const uint64_t o_in = B.take<uint8_t>(RS), o_hist = B.take<uint32_t>(256), o_hdr = B.take<uint8_t>(272), o_codes = B.take<uint32_t>(256);
    const uint64_t o_bits = B.take<uint32_t>(tiles), o_small = B.take<uint32_t>(8), o_out = B.take<uint8_t>(PS);
    if (!scratch(B.used)) return nullptr;
    uint32_t* sm = at<uint32_t>(o_small);  // [0]=rlen [1]=ok [2]=clen
    if (cudaMemcpyAsync(at<uint8_t>(o_in), buf, n, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return nullptr;
    if (cudaMemcpyAsync(sm, &n, 4, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return nullptr;
    HufEncArgs a{};
    a.d_in = at<uint8_t>(o_in); a.stride = RS; a.d_rlen = sm; a.max_r = n; a.nblk = 1; a.compute_hist = true; a.d_hist = at<uint32_t>(o_hist);
    a.d_hdr = at<uint8_t>(o_hdr); a.d_codes = at<uint32_t>(o_codes); a.d_ok = sm + 1; a.d_t_bits = at<uint32_t>(o_bits); a.d_clen = sm + 2;
    a.d_out = at<uint8_t>(o_out); a.out_stride = PS;
    if (!huf_encode_batch(a, g_st)) return nullptr;
    uint8_t hdr[268];
    if (cudaMemcpyAsync(hdr, a.d_hdr, 268, cudaMemcpyDeviceToHost, g_st) != cudaSuccess || !sync_ok()) return nullptr;
    bra_huffman_chunk_t* ch = static_cast<bra_huffman_chunk_t*>(malloc(sizeof(bra_huffman_chunk_t)));
    if (!ch) return nullptr;
    memcpy(&ch->meta, hdr + 4, 264);
    ch->data = static_cast<uint8_t*>(malloc(ch->meta.encoded_size ? ch->meta.encoded_size : 1));
    if (!ch->data || cudaMemcpy(ch->data, a.d_out, ch->meta.encoded_size, cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        free(ch->data);
        free(ch);
        return nullptr;
    }
    return ch;
}

extern "C" uint8_t* bra_huffman_decode(const bra_huffman_t* meta, const uint8_t* data, uint32_t* out_size)
{
    if (out_size) *out_size = 0;
    if (!meta || !data || !out_size) return nullptr;
    const uint32_t r = meta->orig_size, c = meta->encoded_size;
    if (c == 0)
    {
        // nothing to walk: the reference succeeds only when nothing was expected either (bra_huffman.c:455,485)
        if (r != 0)
        {
            bra_b200_log_error("huffman decode error: decoded data:0 - original_data:%u", r);
            return nullptr;
        }
        return static_cast<uint8_t*>(malloc(1));
    }
    if (r > 0x20000000u || c >= 0x20000000u)  // bit offsets (c * 8) are 32-bit
    {
        bra_b200_log_error("bra_huffman_decode: sizes above this implementation's limit (%u, %u)", r, c);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    STAGE_ENTER(nullptr);
    const uint64_t PS = pad16((uint64_t) c + 64), RS = pad16((uint64_t) r + 16), seqs = huf_dec_seqs(c);
    Bump           B;
    const uint64_t o_pay = B.take<uint8_t>(PS), o_hdr = B.take<uint8_t>(272), o_out = B.take<uint8_t>(RS), o_tab = B.take<bra_huf_dec_t>(1);
    const uint64_t o_ss = B.take<uint8_t>(seqs * huf_dec_subs_per_seq()), o_sc = B.take<uint16_t>(seqs * huf_dec_subs_per_seq());
    const uint64_t o_se = B.take<uint32_t>(seqs), o_sx = B.take<uint32_t>(seqs), o_sn = B.take<uint32_t>(seqs), o_small = B.take<uint32_t>(8);
    const uint64_t o_ph = B.take<uint8_t>((uint64_t) seqs * huf_dec_phase_bytes_per_seq());
    if (!scratch(B.used)) return nullptr;
    uint8_t hdr[268];
    memset(hdr, 0, 4);
    memcpy(hdr + 4, meta, 264);
    uint32_t*      sm    = at<uint32_t>(o_small);  // [0]=clen [1]=err [2]=end_bit [3]=changed
    const uint32_t hv[4] = {c, 0, 0, 0};
    if (cudaMemcpyAsync(at<uint8_t>(o_pay), data, c, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return nullptr;
    if (cudaMemcpyAsync(at<uint8_t>(o_hdr), hdr, 268, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return nullptr;
    if (cudaMemcpyAsync(sm, hv, 16, cudaMemcpyHostToDevice, g_st) != cudaSuccess) return nullptr;
    if (cudaStreamSynchronize(g_st) != cudaSuccess) return nullptr;  // hdr/hv live on this stack frame
    HufDecArgs a{};
    a.d_pay = at<uint8_t>(o_pay); a.pay_stride = PS; a.d_clen = sm; a.d_hdr = at<uint8_t>(o_hdr); a.max_c = c; a.nblk = 1; a.d_out = at<uint8_t>(o_out);
    a.out_stride = RS; a.d_tabs = at<bra_huf_dec_t>(o_tab); a.d_err = sm + 1; a.d_sub_start = at<uint8_t>(o_ss); a.d_sub_count = at<uint16_t>(o_sc);
    a.d_seq_entry = at<uint32_t>(o_se); a.d_seq_exit = at<uint32_t>(o_sx); a.d_seq_count = at<uint32_t>(o_sn); a.d_end_bit = sm + 2;
    a.d_changed = sm + 3; a.d_phase = at<uint8_t>(o_ph);
    if (!huf_decode_batch(a, g_st)) return nullptr;
    uint32_t err = 1;
    if (cudaMemcpyAsync(&err, sm + 1, 4, cudaMemcpyDeviceToHost, g_st) != cudaSuccess || !sync_ok()) return nullptr;
    if (err)
    {
        bra_b200_log_error("huffman decode error: invalid code sequence or size mismatch");
        return nullptr;
    }
    uint8_t* o = static_cast<uint8_t*>(malloc(r ? r : 1));
    if (!o) return nullptr;
    if (r && cudaMemcpy(o, a.d_out, r, cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        free(o);
        return nullptr;
    }
    *out_size = r;
    return o;
}

extern "C" void bra_huffman_chunk_free(bra_huffman_chunk_t* chunk)
{
    if (!chunk) return;
    free(chunk->data);
    chunk->data = nullptr;
    free(chunk);
}
