// hostpath.cu -- host-buffer entry points: what bra_io_file_chunks_compress_file /
// decompress_file need (reference src/io/lib_bra_io_file_chunks.c:169-441), minus the FILE I/O.
//
// encode: host bytes -> H2D -> batched chain -> device-side gather into the on-disk chunk stream
//         (3-byte index + 264-byte Huffman header + payload per chunk, chunks.c:81-92,252-256)
//         -> one D2H copy per batch. The per-entry CRC chain of chunks.c:248-249 is folded on the
//         host from the per-chunk CRCs the GPU produced (two GF(2) multiplications per chunk).
// decode: the stream is walked on the host only to find chunk boundaries (each header names its
//         payload size), uploaded as is, scattered on the device into the batched layout, decoded,
//         and the plain bytes come back with one D2H copy per batch.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"
#include "pipeline.h"

#include <algorithm>
#include <string.h>
#include <vector>

using namespace bra;

// BRA_B200_TRACE=1: per-stage host timestamps of the host-buffer path on stderr (diagnostics only)
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
static bool trace_on()
{
    static const bool on = getenv("BRA_B200_TRACE") != nullptr;
    return on;
}
static double trace_ms()
{
    using namespace std::chrono;
    static const steady_clock::time_point t0 = steady_clock::now();
    return duration<double, std::milli>(steady_clock::now() - t0).count();
}
#define BRA_TRACE(...)                          \
    do {                                        \
        if (trace_on())                         \
        {                                       \
            fprintf(stderr, "[bra trace %9.3f] ", trace_ms()); \
            fprintf(stderr, __VA_ARGS__);       \
            fputc('\n', stderr);                \
        }                                       \
    } while (0)

namespace {

// out offsets (exclusive scan of 267 + clen) for up to 32768 blocks: one CTA
// (64-bit throughout: 1024 chunks of 4 MiB or more of incompressible data exceed 2^32 stream bytes)
__global__ void __launch_bounds__(1024) stream_offsets_kernel(const uint8_t* __restrict__ hdr, uint32_t nblk, uint64_t* __restrict__ off)
{
    __shared__ unsigned long long red[33];
    unsigned long long            carry = 0;
    const uint32_t                w = warp_id(), l = lane_id();
    for (uint32_t base = 0; base < nblk; base += 1024)
    {
        const uint32_t     b = base + threadIdx.x;
        unsigned long long v = 0;
        if (b < nblk)
        {
            const uint8_t* h = hdr + (uint64_t) b * 268 + 264;
            v = 267ull + ((uint32_t) h[0] | ((uint32_t) h[1] << 8) | ((uint32_t) h[2] << 16) | ((uint32_t) h[3] << 24));
        }
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned long long t = __shfl_up_sync(BRA_FULL, inc, d);
            if (l >= (uint32_t) d) inc += t;
        }
        if (l == 31) red[w] = inc;
        __syncthreads();
        if (w == 0)
        {
            const unsigned long long x = red[l];
            unsigned long long       s = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const unsigned long long t = __shfl_up_sync(BRA_FULL, s, d);
                if (l >= (uint32_t) d) s += t;
            }
            red[l] = s - x;
            if (l == 31) red[32] = s;
        }
        __syncthreads();
        if (b < nblk) off[b] = carry + red[w] + inc - v;
        carry += red[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nblk] = carry;
}

// gather: out[off[b] ..) = hdr268[b][0..3) | hdr268[b][4..268) | payload[b][0..clen)
__global__ void __launch_bounds__(256)
    stream_gather_kernel(const uint8_t* __restrict__ hdr, const uint8_t* __restrict__ pay, uint64_t pay_stride, const uint64_t* __restrict__ off,
                         uint8_t* __restrict__ out)
{
    const uint32_t b    = blockIdx.y;
    const uint64_t o    = off[b];
    const uint32_t size = (uint32_t) (off[b + 1] - o);  // 267 + clen
    const uint8_t* h    = hdr + (uint64_t) b * 268;
    const uint8_t* p    = pay + (uint64_t) b * pay_stride;
    for (uint32_t i = blockIdx.x * 256 * 16 + threadIdx.x; i < min(size, (blockIdx.x + 1) * 256u * 16u); i += 256)
        out[o + i] = i < 3 ? h[i] : (i < 267 ? h[i + 1] : p[i - 267]);
}

// scatter: inverse of the gather (bytes past a payload are never read as data: the Huffman staging masks them)
__global__ void __launch_bounds__(256)
    stream_scatter_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ off, uint8_t* __restrict__ hdr, uint8_t* __restrict__ pay,
                          uint64_t pay_stride)
{
    const uint32_t b    = blockIdx.y;
    const uint64_t o    = off[b];
    const uint32_t size = (uint32_t) (off[b + 1] - o);
    uint8_t*       h    = hdr + (uint64_t) b * 268;
    uint8_t*       p    = pay + (uint64_t) b * pay_stride;
    for (uint32_t i = blockIdx.x * 256 * 16 + threadIdx.x; i < min(size, (blockIdx.x + 1) * 256u * 16u); i += 256)
    {
        const uint8_t v = in[o + i];
        if (i < 3)
            h[i] = v;
        else if (i < 267)
            h[i + 1] = v;
        else
            p[i - 267] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) h[3] = 0;  // index is 3 bytes on disk, 4 in memory (chunks.c:52-64)
}

__global__ void compact_out_kernel(const uint8_t* __restrict__ src, uint32_t block, const uint32_t* __restrict__ len, const uint64_t* __restrict__ off,
                                   uint8_t* __restrict__ dst)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b];
    const uint64_t o = off[b];
    for (uint32_t i = blockIdx.x * 256 * 16 + threadIdx.x; i < min(n, (blockIdx.x + 1) * 256u * 16u); i += 256)
        dst[o + i] = src[(uint64_t) b * block + i];
}

}  // namespace

// page-locked host memory for callers that have no CUDA headers (the seam's window buffers)
extern "C" void* bra_b200_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess)
    {
        cudaGetLastError();  // a failed allocation must not poison the next call
        return nullptr;
    }
    return p;
}
extern "C" void bra_b200_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

extern "C" uint64_t bra_b200_encode_bound(const bra_b200_ctx_t* c, uint64_t total)
{
    if (!c) return 0;
    const uint64_t nblk = (total + bra_b200_block_size(c) - 1) / bra_b200_block_size(c);
    return nblk * (267 + bra_b200_payload_stride(c));
}

// RAII bundle of the two copy streams and the events that order them against the compute stream
namespace {
struct Pipe
{
    cudaStream_t s_in = nullptr, s_out = nullptr, s_comp = nullptr;  // s_comp: the context's compute stream (not owned)
    cudaEvent_t  ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool         ok = false;
    explicit Pipe(cudaStream_t comp) : s_comp(comp)
    {
        ok = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2 && ok; ++i)
            ok = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ~Pipe()
    {
        // Every return path, error paths included, leaves with no copy or kernel in flight: the transfers touch the
        // caller's buffers and the context's staging memory, which the caller may free or reuse right after the call.
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_comp) cudaStreamSynchronize(s_comp);
        if (s_out) cudaStreamSynchronize(s_out);
        for (int i = 0; i < 2; ++i)
        {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    }
};
// Largest pipeline stage: the context's whole batch.
inline uint32_t stage_blocks(uint32_t max_batch) { return max_batch; }
// blocks per stage: bra_stage_plan (bra_hd.h)
inline std::vector<uint32_t> stage_plan(uint64_t nblk, uint32_t hb)
{
    std::vector<uint32_t> plan(nblk / std::max(1u, hb) + 4);
    plan.resize(bra_stage_plan(nblk, hb, plan.data()));
    return plan;
}
inline std::vector<uint32_t> stage_plan_encode(uint64_t nblk, uint32_t hb)
{
    static const uint32_t head_div = getenv("BRA_B200_ENC_HEAD_DIV") ? (uint32_t) atoi(getenv("BRA_B200_ENC_HEAD_DIV")) : 8u;  // tuning switch, read once
    std::vector<uint32_t> plan(nblk / std::max(1u, hb) + 4);
    plan.resize(bra_stage_plan_encode(nblk, hb, head_div, plan.data()));
    return plan;
}
}  // namespace

// Three streams: input copies run one stage ahead of the kernels, output copies one stage behind
// (double-buffered device staging). The kernels' own host syncs (BWT round control) only block the
// host thread; the copy streams keep moving underneath.
namespace bra {
// `place(user, bytes)` names the host destination of the next `bytes` stream bytes (nullptr: no room). The plain entry
// point hands out consecutive pieces of the caller's buffer; the multi-GPU pool (pool.cu) first waits until the sizes of
// all earlier block ranges are known, so that every range lands at its final offset of the ordered stream.
int encode_host_impl(bra_b200_ctx_t* c, const uint8_t* in, uint64_t total, uint8_t* (*place)(void*, uint64_t), void* user, uint64_t* out_size,
                     uint32_t* crc_chain)
{
    if (!c || !in || !place || !out_size || total == 0)
    {
        bra_b200_log_error("bra_b200_encode_host: invalid arguments");
        return 1;
    }
    BraDeviceGuard dg(ctx_device(c));
    if (!dg.ok) return 2;
    const uint32_t S  = bra_b200_block_size(c);
    const uint32_t HB = stage_blocks(bra_b200_max_batch(c));
    const uint64_t PS = bra_b200_payload_stride(c);
    const uint64_t nblk_total = (total + S - 1) / S;
    const std::vector<uint32_t> plan = stage_plan_encode(nblk_total, HB);
    const uint64_t        nstage = plan.size();
    std::vector<uint64_t> first(nstage + 1, 0);  // first block of every stage
    for (uint64_t i = 0; i < nstage; ++i) first[i + 1] = first[i] + plan[i];
    cudaStream_t   st = ctx_stream(c);
    *out_size         = 0;
    Pipe P(st);
    if (!P.ok) return 2;

    // device staging: input x2 | hdr | payload | stream x2 | offsets | crcs
    auto           up    = [](uint64_t v) { return (v + 255) / 256 * 256; };
    const uint64_t in_b  = up((uint64_t) HB * S), hdr_b = up((uint64_t) HB * 268), pay_b = up((uint64_t) HB * PS),
                   str_b = up((uint64_t) HB * (267 + PS)), off_b = up(((uint64_t) HB + 1) * 8), crc_b = up((uint64_t) HB * 8);
    uint8_t* io = ctx_io_buffer(c, 2 * in_b + hdr_b + pay_b + 2 * str_b + off_b + crc_b + 4096);
    if (!io) return 3;
    uint8_t*  d_in[2]  = {io, io + in_b};
    uint8_t*  d_hdr    = io + 2 * in_b;
    uint8_t*  d_pay    = d_hdr + hdr_b;
    uint8_t*  d_str[2] = {d_pay + pay_b, d_pay + pay_b + str_b};
    uint64_t* d_off    = reinterpret_cast<uint64_t*>(d_str[1] + str_b);
    uint32_t* d_crc    = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(d_off) + off_b);
    uint32_t* d_hcrc   = d_crc + HB;

    const bra_gf_pow_t*   pw = crc_host_pow();
    uint32_t*             mail = ctx_mail_host(c);
    std::vector<uint32_t> h_crc(2 * (size_t) HB);
    uint32_t              crc = crc_chain ? *crc_chain : 0;
    uint64_t              produced = 0;
    auto stage_bytes = [&](uint64_t i) { return std::min<uint64_t>((uint64_t) plan[i] * S, total - first[i] * S); };

    if (cudaMemcpyAsync(d_in[0], in, stage_bytes(0), cudaMemcpyHostToDevice, P.s_in) != cudaSuccess) return 4;
    cudaEventRecord(P.ev_in[0], P.s_in);
    for (uint64_t i = 0; i < nstage; ++i)
    {
        const int      slot  = (int) (i & 1);
        const uint64_t bytes = stage_bytes(i);
        const uint32_t nb    = (uint32_t) ((bytes + S - 1) / S);
        const uint32_t last  = (uint32_t) (bytes - (uint64_t) (nb - 1) * S);
        if (i + 1 < nstage)
        {
            // prefetch the next stage's input; its slot was last read by the kernels of stage i-1
            if (i >= 1) cudaStreamWaitEvent(P.s_in, P.ev_comp[slot ^ 1], 0);
            if (cudaMemcpyAsync(d_in[slot ^ 1], in + first[i + 1] * S, stage_bytes(i + 1), cudaMemcpyHostToDevice, P.s_in) != cudaSuccess) return 4;
            cudaEventRecord(P.ev_in[slot ^ 1], P.s_in);
        }
        cudaStreamWaitEvent(st, P.ev_in[slot], 0);
        if (i >= 2) cudaStreamWaitEvent(st, P.ev_out[slot], 0);  // the stream slot is free once stage i-2 has left the device
        BRA_TRACE("encode stage %llu: %u blocks, kernels enqueued from here", (unsigned long long) i, nb);
        if (!encode_batch(c, d_in[slot], nb, last, d_hdr, d_pay, d_crc, st)) return 5;
        if (!crc_headers(d_hdr, 268, nb, d_hcrc, st)) return 5;
        BRA_LAUNCH(P_GLUE, st, stream_offsets_kernel<<<1, 1024, 0, st>>>(d_hdr, nb, d_off));
        const uint32_t gx = bra_div_up(267 + PS, 4096);
        BRA_LAUNCH(P_GLUE, st, stream_gather_kernel<<<dim3(gx, nb), 256, 0, st>>>(d_hdr, d_pay, PS, d_off, d_str[slot]));
        // stream size and the per-chunk CRCs come back through the mail words: a small cudaMemcpy would wait on the
        // device-to-host copy engine behind the previous stage's stream
        if (!mail_publish(mail, reinterpret_cast<const uint32_t*>(d_off + nb), 2, st) || !mail_publish(mail + 16, d_crc, 2 * HB, st)) return 5;
        cudaEventRecord(P.ev_comp[slot], st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return 4;
        uint64_t str_size = 0;
        memcpy(&str_size, mail, 8);
        BRA_TRACE("encode stage %llu: kernels done, %llu stream bytes", (unsigned long long) i, (unsigned long long) str_size);
        memcpy(h_crc.data(), mail + 16, (size_t) 2 * HB * 4);
        uint8_t* dst = place(user, str_size);
        if (!dst)
        {
            bra_b200_log_error("bra_b200_encode_host: output buffer too small");
            return 6;
        }
        cudaStreamWaitEvent(P.s_out, P.ev_comp[slot], 0);
        if (cudaMemcpyAsync(dst, d_str[slot], str_size, cudaMemcpyDeviceToHost, P.s_out) != cudaSuccess) return 4;
        cudaEventRecord(P.ev_out[slot], P.s_out);
        // CRC chain of reference chunks.c:248-249 while the copy runs
        for (uint32_t b = 0; b < nb; ++b)
        {
            const uint32_t n = (b + 1 == nb) ? last : S;
            crc = bra_crc_combine(pw, crc, h_crc[HB + b], 268);
            crc = bra_crc_combine(pw, crc, h_crc[b], n);
        }
        produced += str_size;
    }
    if (cudaStreamSynchronize(P.s_out) != cudaSuccess || cudaStreamSynchronize(P.s_in) != cudaSuccess) return 4;
    BRA_TRACE("encode: last output copy done");
    *out_size = produced;
    if (crc_chain) *crc_chain = crc;
    return 0;
}
}  // namespace bra

namespace {
struct LinearSink
{
    uint8_t* out;
    uint64_t cap, used;
};
uint8_t* linear_place(void* user, uint64_t bytes)
{
    LinearSink* s = static_cast<LinearSink*>(user);
    if (s->used + bytes > s->cap) return nullptr;
    uint8_t* p = s->out + s->used;
    s->used += bytes;
    return p;
}
}  // namespace

extern "C" int bra_b200_encode_host(bra_b200_ctx_t* c, const uint8_t* in, uint64_t total, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                                    uint32_t* crc_chain)
{
    if (!out)
    {
        bra_b200_log_error("bra_b200_encode_host: invalid arguments");
        return 1;
    }
    LinearSink s{out, out_cap, 0};
    return encode_host_impl(c, in, total, linear_place, &s, out_size, crc_chain);
}

// sizes_only: list mode of the reference (chunks.c:369-373) -- Huffman decode + RLE size pass, nothing is copied back
static int decode_host_impl(bra_b200_ctx_t* c, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                            uint32_t* crc_chain, bool sizes_only)
{
    if (!c || !in || (!out && !sizes_only) || !out_size)
    {
        bra_b200_log_error("bra_b200_decode_host: invalid arguments");
        return 1;
    }
    BraDeviceGuard dg(ctx_device(c));
    if (!dg.ok) return 2;
    const uint32_t S  = bra_b200_block_size(c);
    const uint32_t HB = stage_blocks(bra_b200_max_batch(c));
    const uint64_t PS = bra_b200_payload_stride(c);
    cudaStream_t   st = ctx_stream(c);
    *out_size         = 0;
    Pipe P(st);
    if (!P.ok) return 2;

    auto           up    = [](uint64_t v) { return (v + 255) / 256 * 256; };
    const uint64_t str_b = up((uint64_t) HB * (267 + PS)), hdr_b = up((uint64_t) HB * 268), pay_b = up((uint64_t) HB * PS), out_b = up((uint64_t) HB * S),
                   off_b = up(((uint64_t) HB + 1) * 8), misc_b = up((uint64_t) HB * 16);
    uint8_t* io = ctx_io_buffer(c, 2 * str_b + hdr_b + pay_b + 3 * out_b + 3 * off_b + misc_b + 8192);
    if (!io) return 3;
    uint8_t*  d_str[2] = {io, io + str_b};
    uint8_t*  d_hdr    = io + 2 * str_b;
    uint8_t*  d_pay    = d_hdr + hdr_b;
    uint8_t*  d_out[2] = {d_pay + pay_b, d_pay + pay_b + out_b};
    uint8_t*  d_cmp    = d_out[1] + out_b;
    uint64_t* d_off[2] = {reinterpret_cast<uint64_t*>(d_cmp + out_b), reinterpret_cast<uint64_t*>(d_cmp + out_b + off_b)};
    uint64_t* d_ooff   = reinterpret_cast<uint64_t*>(d_cmp + out_b + 2 * off_b);
    uint32_t* d_len    = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(d_ooff) + off_b);
    uint32_t* d_crc    = d_len + HB;
    uint32_t* d_stat   = d_crc + HB;
    uint32_t* d_hcrc   = d_stat + HB;

    // walk the stream once on the host: chunk boundaries of every stage (reference chunks.c:338-357, :414)
    struct Stage
    {
        uint64_t pos, end;
        uint32_t nb, max_r, max_c;
    };
    std::vector<Stage>    stages;
    std::vector<uint64_t> offs;  // per stage: nb+1 offsets relative to the stage start, concatenated
    {
        struct Chunk
        {
            uint64_t pos;
            uint32_t r, c;
        };
        std::vector<Chunk> chunks;
        uint64_t           p = 0;
        while (p < in_size)
        {
            if (in_size - p < 267)
            {
                bra_b200_log_error("bra_b200_decode_host: truncated chunk header at offset %llu", (unsigned long long) p);
                return 7;
            }
            const uint8_t* h = in + p + 3 + 256;
            const uint32_t r = (uint32_t) h[0] | ((uint32_t) h[1] << 8) | ((uint32_t) h[2] << 16) | ((uint32_t) h[3] << 24);
            const uint32_t cc = (uint32_t) h[4] | ((uint32_t) h[5] << 8) | ((uint32_t) h[6] << 16) | ((uint32_t) h[7] << 24);
            if (cc == 0 || r == 0 || cc > PS - 32 || in_size - p - 267 < cc)
            {
                bra_b200_log_error("bra_b200_decode_host: chunk header not valid at offset %llu", (unsigned long long) p);
                return 7;
            }
            chunks.push_back({p, r, cc});
            p += 267 + cc;
        }
        size_t k = 0;
        for (const uint32_t nb : stage_plan(chunks.size(), HB))
        {
            Stage sg{chunks[k].pos, 0, nb, 0, 0};
            for (uint32_t b = 0; b < nb; ++b, ++k)
            {
                offs.push_back(chunks[k].pos - sg.pos);
                sg.max_r = std::max(sg.max_r, chunks[k].r);
                sg.max_c = std::max(sg.max_c, chunks[k].c);
            }
            sg.end = k < chunks.size() ? chunks[k].pos : in_size;
            offs.push_back(sg.end - sg.pos);
            stages.push_back(sg);
        }
    }
    const bra_gf_pow_t*   pw = crc_host_pow();
    uint32_t*             mail = ctx_mail_host(c);
    std::vector<uint64_t> h_ooff(HB + 1);
    std::vector<uint32_t> h_misc(4 * (size_t) HB);
    uint32_t              crc = crc_chain ? *crc_chain : 0;
    uint64_t              produced = 0;
    std::vector<size_t>   off_base(stages.size() + 1, 0);
    for (size_t i = 0; i < stages.size(); ++i) off_base[i + 1] = off_base[i] + stages[i].nb + 1;

    auto upload = [&](size_t i, int slot) -> bool {
        const Stage& sg = stages[i];
        if (cudaMemcpyAsync(d_str[slot], in + sg.pos, sg.end - sg.pos, cudaMemcpyHostToDevice, P.s_in) != cudaSuccess) return false;
        if (cudaMemcpyAsync(d_off[slot], offs.data() + off_base[i], ((size_t) sg.nb + 1) * 8, cudaMemcpyHostToDevice, P.s_in) != cudaSuccess) return false;
        cudaEventRecord(P.ev_in[slot], P.s_in);
        return true;
    };
    if (!stages.empty() && !upload(0, 0)) return 4;
    for (size_t i = 0; i < stages.size(); ++i)
    {
        const int    slot = (int) (i & 1);
        const Stage& sg   = stages[i];
        if (i + 1 < stages.size())
        {
            if (i >= 1) cudaStreamWaitEvent(P.s_in, P.ev_comp[slot ^ 1], 0);
            if (!upload(i + 1, slot ^ 1)) return 4;
        }
        cudaStreamWaitEvent(st, P.ev_in[slot], 0);
        if (i >= 2) cudaStreamWaitEvent(st, P.ev_out[slot], 0);  // output slot free once stage i-2 has been copied out
        const uint32_t gx = bra_div_up(267 + (uint64_t) sg.max_c, 4096);
        BRA_TRACE("decode stage %zu: %u chunks, kernels enqueued from here", i, sg.nb);
        BRA_LAUNCH(P_GLUE, st, stream_scatter_kernel<<<dim3(gx, sg.nb), 256, 0, st>>>(d_str[slot], d_off[slot], d_hdr, d_pay, PS));
        if (!decode_batch(c, d_hdr, d_pay, sg.nb, sg.max_r, sg.max_c, d_out[slot], d_len, d_crc, d_stat, st, sizes_only)) return 5;
        if (!crc_headers(d_hdr, 268, sg.nb, d_hcrc, st)) return 5;
        // lengths, CRCs and status come back through the mail words (see encode)
        if (!mail_publish(mail, d_len, 4 * HB, st)) return 5;
        if (cudaStreamSynchronize(st) != cudaSuccess) return 4;
        memcpy(h_misc.data(), mail, (size_t) 4 * HB * 4);
        BRA_TRACE("decode stage %zu: kernels done", i);
        if (sizes_only)
        {
            for (uint32_t b = 0; b < sg.nb; ++b)
            {
                if (h_misc[2 * (size_t) HB + b] != 0)
                {
                    bra_b200_log_error("bra_b200_list_host: chunk %u of the batch at offset %llu is corrupt", b, (unsigned long long) sg.pos);
                    return 8;
                }
                produced += h_misc[b];
            }
            cudaEventRecord(P.ev_comp[slot], st);
            cudaEventRecord(P.ev_out[slot], st);
            continue;
        }
        uint64_t o = 0;
        bool     contiguous = true;
        for (uint32_t b = 0; b < sg.nb; ++b)
        {
            if (h_misc[2 * (size_t) HB + b] != 0)
            {
                bra_b200_log_error("bra_b200_decode_host: chunk %u of the batch at offset %llu is corrupt", b, (unsigned long long) sg.pos);
                return 8;
            }
            h_ooff[b] = o;
            o += h_misc[b];
            if (b + 1 < sg.nb) contiguous &= h_misc[b] == S;
            crc = bra_crc_combine(pw, crc, h_misc[3 * (size_t) HB + b], 268);   // chunks.c:396
            crc = bra_crc_combine(pw, crc, h_misc[(size_t) HB + b], h_misc[b]); // chunks.c:397
        }
        h_ooff[sg.nb] = o;
        if (produced + o > out_cap)
        {
            bra_b200_log_error("bra_b200_decode_host: output buffer too small");
            return 6;
        }
        // full blocks are already contiguous; compact only when some block is short
        const uint8_t* src = d_out[slot];
        if (!contiguous)
        {
            if (cudaMemcpyAsync(d_ooff, h_ooff.data(), ((size_t) sg.nb + 1) * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return 4;
            BRA_LAUNCH(P_GLUE, st, compact_out_kernel<<<dim3(bra_div_up(S, 4096), sg.nb), 256, 0, st>>>(d_out[slot], S, d_len, d_ooff, d_cmp));
            if (cudaStreamSynchronize(st) != cudaSuccess) return 4;  // d_cmp is single-buffered: drain it before the next stage
            if (cudaMemcpyAsync(out + produced, d_cmp, o, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 4;
            if (cudaStreamSynchronize(st) != cudaSuccess) return 4;
            cudaEventRecord(P.ev_comp[slot], st);
            cudaEventRecord(P.ev_out[slot], st);
        }
        else
        {
            cudaEventRecord(P.ev_comp[slot], st);
            cudaStreamWaitEvent(P.s_out, P.ev_comp[slot], 0);
            if (cudaMemcpyAsync(out + produced, src, o, cudaMemcpyDeviceToHost, P.s_out) != cudaSuccess) return 4;
            cudaEventRecord(P.ev_out[slot], P.s_out);
        }
        produced += o;
    }
    if (cudaStreamSynchronize(P.s_out) != cudaSuccess || cudaStreamSynchronize(P.s_in) != cudaSuccess) return 4;
    BRA_TRACE("decode: last output copy done");
    *out_size = produced;
    if (crc_chain) *crc_chain = crc;
    return 0;
}

extern "C" int bra_b200_decode_host(bra_b200_ctx_t* c, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                                    uint32_t* crc_chain)
{
    return decode_host_impl(c, in, in_size, out, out_cap, out_size, crc_chain, false);
}

extern "C" int bra_b200_list_host(bra_b200_ctx_t* c, const uint8_t* in, uint64_t in_size, uint64_t* plain_size)
{
    return decode_host_impl(c, in, in_size, nullptr, 0, plain_size, nullptr, true);
}
