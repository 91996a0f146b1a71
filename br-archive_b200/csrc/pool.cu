// pool.cu -- ONE job over several GPUs of one process (BASELINE config 5).
//
// The reference walks a file chunk by chunk and carries nothing from one chunk to the next but the running
// CRC (reference src/io/lib_bra_io_file_chunks.c:199-266), which is associative through bra_crc32c_combine
// (src/utils/lib_bra_crc32c.c:181-231). So the block list of one input is cut into ranges of `range_blocks`
// blocks and the ranges are handed out dynamically (an atomic ticket: blocks of long-repeat data cost ten times
// what text costs, a static split would leave GPUs idle) to worker threads, two per GPU, each with its own
// context, stream and workspace -- the second worker's copies overlap the first one's kernels. No collective:
// NVLink is not on the data path; the only ordering is the host-side one of the output stream.
//
//   encode: a range's stream bytes go straight from the GPU to their final offset of the ordered chunk stream.
//           The offset is the sum of the sizes of all earlier ranges, known once their kernels have finished:
//           a worker waits for that prefix (ranges are started in order, so the wait is short) before it issues
//           its device-to-host copy. CRC chains are computed per range from 0 and folded in order.
//   decode: chunk boundaries are found by one host pass over the headers (each names its payload size); ranges
//           decode to `range_blocks * block` bytes each at their final offset; a stream whose inner chunks are
//           short is compacted afterwards.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"
#include "pipeline.h"

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>

using namespace bra;

struct bra_b200_pool
{
    uint32_t                     block = 0, range_blocks = 0;
    std::vector<int>             devices;   // one entry per worker
    std::vector<bra_b200_ctx_t*> ctx;       // one context per worker
    std::vector<uint32_t>        last_ranges;  // ranges each worker processed in the last call
};

extern "C" bra_b200_pool_t* bra_b200_pool_create(const int* devices, int ndev, uint32_t block_size, uint32_t range_blocks, int workers_per_device)
{
    if (!devices || ndev <= 0 || ndev > 64 || range_blocks == 0 || workers_per_device <= 0 || workers_per_device > 4)
    {
        bra_b200_log_error("bra_b200_pool_create: invalid arguments");
        return nullptr;
    }
    bra_b200_pool* p = new bra_b200_pool();
    p->block         = block_size;
    p->range_blocks  = range_blocks;
    for (int w = 0; w < workers_per_device; ++w)
        for (int d = 0; d < ndev; ++d)
        {
            bra_b200_ctx_t* c = bra_b200_ctx_create(devices[d], block_size, range_blocks);
            if (!c)
            {
                bra_b200_pool_destroy(p);
                return nullptr;
            }
            p->devices.push_back(devices[d]);
            p->ctx.push_back(c);
        }
    p->last_ranges.assign(p->ctx.size(), 0);
    return p;
}

extern "C" void bra_b200_pool_destroy(bra_b200_pool_t* p)
{
    if (!p) return;
    for (bra_b200_ctx_t* c : p->ctx) bra_b200_ctx_destroy(c);
    delete p;
}

extern "C" int bra_b200_pool_workers(const bra_b200_pool_t* p) { return p ? (int) p->ctx.size() : 0; }

extern "C" int bra_b200_pool_worker_ranges(const bra_b200_pool_t* p, int worker, int* device, uint32_t* ranges)
{
    if (!p || worker < 0 || worker >= (int) p->ctx.size()) return 1;
    if (device) *device = p->devices[worker];
    if (ranges) *ranges = p->last_ranges[worker];
    return 0;
}

extern "C" uint64_t bra_b200_pool_encode_bound(const bra_b200_pool_t* p, uint64_t total)
{
    return (p && !p->ctx.empty()) ? bra_b200_encode_bound(p->ctx[0], total) : 0;
}

namespace {

// ordered hand-over of output offsets between ranges
struct Order
{
    std::mutex              mu;
    std::condition_variable cv;
    std::vector<uint64_t>   start;  // start[r] = stream offset of range r, valid once known[r]
    std::vector<uint8_t>    known;
    bool                    failed = false;
    explicit Order(size_t nranges) : start(nranges + 1, 0), known(nranges + 1, 0) { known[0] = 1; }
    bool wait(size_t r, uint64_t* off)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return known[r] || failed; });
        *off = start[r];
        return !failed;
    }
    void publish(size_t r, uint64_t off)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            start[r] = off;
            known[r] = 1;
        }
        cv.notify_all();
    }
    void fail()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            failed = true;
        }
        cv.notify_all();
    }
};

struct RangeSink
{
    Order*   order;
    size_t   r;
    uint8_t* out;
    uint64_t cap, base, used;
    bool     have_base;
};

uint8_t* range_place(void* user, uint64_t bytes)
{
    RangeSink* s = static_cast<RangeSink*>(user);
    if (!s->have_base)
    {
        if (!s->order->wait(s->r, &s->base)) return nullptr;
        s->have_base = true;
    }
    if (s->base + s->used + bytes > s->cap) return nullptr;
    uint8_t* p = s->out + s->base + s->used;
    s->used += bytes;
    return p;
}

}  // namespace

extern "C" int bra_b200_pool_encode_host(bra_b200_pool_t* p, const uint8_t* in, uint64_t total, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                                         uint32_t* crc_chain)
{
    if (!p || !in || !out || !out_size || total == 0)
    {
        bra_b200_log_error("bra_b200_pool_encode_host: invalid arguments");
        return 1;
    }
    const uint64_t S = p->block, RB = (uint64_t) p->range_blocks * S;
    const size_t   nranges = (size_t) ((total + RB - 1) / RB);
    Order          order(nranges);
    std::vector<uint32_t> crc(nranges, 0);
    std::vector<uint64_t> covered(nranges, 0);  // bytes each range's CRC chain covers: 268 per chunk + the chunk
    std::atomic<size_t>   next{0};
    std::atomic<int>      rc{0};
    auto work = [&](size_t w) {
        uint32_t done = 0;
        for (;;)
        {
            const size_t r = next.fetch_add(1);
            if (r >= nranges) break;
            const uint64_t off = r * RB, bytes = std::min<uint64_t>(RB, total - off);
            RangeSink      sink{&order, r, out, out_cap, 0, 0, false};
            uint64_t       sz = 0;
            int            e  = rc.load() ? 9 : encode_host_impl(p->ctx[w], in + off, bytes, range_place, &sink, &sz, &crc[r]);
            if (e == 0 && !sink.have_base) e = order.wait(r, &sink.base) ? 0 : 9;
            if (e != 0)
            {
                int zero = 0;
                rc.compare_exchange_strong(zero, e);
                order.fail();
                break;
            }
            covered[r] = bytes + 268ull * ((bytes + S - 1) / S);
            order.publish(r + 1, sink.base + sz);
            ++done;
        }
        p->last_ranges[w] = done;
    };
    std::vector<std::thread> th;
    for (size_t w = 1; w < p->ctx.size(); ++w) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    if (rc.load()) return rc.load();
    const bra_gf_pow_t* pw = crc_host_pow();
    uint32_t            c  = crc_chain ? *crc_chain : 0;
    for (size_t r = 0; r < nranges; ++r) c = bra_crc_combine(pw, c, crc[r], covered[r]);  // chunks.c:248-249 over the whole entry
    if (crc_chain) *crc_chain = c;
    *out_size = order.start[nranges];
    return 0;
}

extern "C" int bra_b200_pool_decode_host(bra_b200_pool_t* p, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                                         uint32_t* crc_chain)
{
    if (!p || !in || !out || !out_size)
    {
        bra_b200_log_error("bra_b200_pool_decode_host: invalid arguments");
        return 1;
    }
    *out_size = 0;
    // chunk boundaries: one pass over the headers (reference chunks.c:338-357: each header names its payload size)
    const uint64_t        S = p->block, PS = bra_b200_payload_stride(p->ctx[0]);
    std::vector<uint64_t> range_pos;
    std::vector<uint32_t> range_chunks;
    {
        uint64_t pos = 0;
        uint32_t n   = 0;
        while (pos < in_size)
        {
            if (in_size - pos < 267)
            {
                bra_b200_log_error("bra_b200_pool_decode_host: truncated chunk header at offset %llu", (unsigned long long) pos);
                return 7;
            }
            const uint8_t* h  = in + pos + 3 + 256 + 4;
            const uint32_t cc = (uint32_t) h[0] | ((uint32_t) h[1] << 8) | ((uint32_t) h[2] << 16) | ((uint32_t) h[3] << 24);
            if (cc == 0 || cc > PS - 32 || in_size - pos - 267 < cc)
            {
                bra_b200_log_error("bra_b200_pool_decode_host: chunk header not valid at offset %llu", (unsigned long long) pos);
                return 7;
            }
            if (n % p->range_blocks == 0)
            {
                range_pos.push_back(pos);
                range_chunks.push_back(0);
            }
            ++range_chunks.back();
            ++n;
            pos += 267 + cc;
        }
        range_pos.push_back(in_size);
    }
    const size_t nranges = range_chunks.size();
    if (nranges == 0) return 0;
    const uint64_t RB = (uint64_t) p->range_blocks * S;
    if ((nranges - 1) * RB > out_cap)  // (the last range is checked by its own decode call, once its sizes are known)
    {
        bra_b200_log_error("bra_b200_pool_decode_host: output buffer too small");
        return 6;
    }
    std::vector<uint32_t> crc(nranges, 0);
    std::vector<uint64_t> produced(nranges, 0);
    std::atomic<size_t>   next{0};
    std::atomic<int>      rc{0};
    auto work = [&](size_t w) {
        uint32_t done = 0;
        for (;;)
        {
            const size_t r = next.fetch_add(1);
            if (r >= nranges || rc.load()) break;
            const uint64_t o   = r * RB;
            const uint64_t cap = o < out_cap ? std::min<uint64_t>((uint64_t) range_chunks[r] * S, out_cap - o) : 0;
            const int      e   = bra_b200_decode_host(p->ctx[w], in + range_pos[r], range_pos[r + 1] - range_pos[r], out + o, cap, &produced[r], &crc[r]);
            if (e != 0)
            {
                int zero = 0;
                rc.compare_exchange_strong(zero, e);
                break;
            }
            ++done;
        }
        p->last_ranges[w] = done;
    };
    std::vector<std::thread> th;
    for (size_t w = 1; w < p->ctx.size(); ++w) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    if (rc.load()) return rc.load();
    // ordered assembly: every range but the last normally decodes to exactly range_blocks * block bytes
    const bra_gf_pow_t* pw = crc_host_pow();
    uint32_t            c  = crc_chain ? *crc_chain : 0;
    uint64_t            total = 0;
    for (size_t r = 0; r < nranges; ++r)
    {
        if (total != r * RB && produced[r]) memmove(out + total, out + r * RB, produced[r]);  // an earlier range was short (never for streams this library wrote)
        c = bra_crc_combine(pw, c, crc[r], produced[r] + 268ull * range_chunks[r]);            // chunks.c:396-397 over the whole entry
        total += produced[r];
    }
    if (crc_chain) *crc_chain = c;
    *out_size = total;
    return 0;
}
