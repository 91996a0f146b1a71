// prof.cu -- launch accounting for the measurement contract (bench.py):
//   * every kernel launch of the library is counted per kernel family (always on);
//   * when enabled, each launch is bracketed by CUDA events on the stream it is launched on, so
//     bench.py can report the dominant kernel's average launch duration from a live run
//     (events are read after the stream has been synchronised; nothing here blocks the stream).
#include "bra_common.cuh"
#include "bra_kernels.h"

#include <atomic>
#include <mutex>
#include <vector>

namespace bra {

static const char* const kProfNames[P_COUNT] = {
    "crc32c", "radix_hist", "radix_scatter_implicit", "radix_scatter", "radix_scatter_u8", "bwt_period", "bwt_keys", "bwt_heads", "bwt_ranks", "bwt_prepare", "bwt_gather",
    "bwt_finish", "bwt_misc", "mtf_summary", "mtf_scan", "mtf_apply", "rle_enc_heads", "rle_enc_lit", "rle_enc_size", "rle_enc_emit", "rle_dec_exit",
    "rle_dec_chain", "rle_dec_mark", "rle_dec_expand", "huf_hist", "huf_build", "huf_bits", "huf_pack", "huf_dec_tables", "huf_dec_sync",
    "huf_dec_scan", "huf_dec_write", "huf_dec_trailing", "ibwt_walk_len", "ibwt_stitch", "ibwt_walk_emit", "ibwt_copy", "glue"};

struct ProfSlot
{
    std::atomic<uint64_t> launches{0};
    double                ms = 0;
    struct Bracket
    {
        cudaEvent_t a, b;
        int         dev;
    };
    std::vector<Bracket> pending;
};
static ProfSlot   g_slots[P_COUNT];
static bool       g_timing = false;
static std::mutex g_prof_mu;
// CUDA events belong to the device that was current when they were created: one pool per device, so that several
// contexts on different GPUs of one process (bra_b200_pool) can be timed at once.
#define PROF_MAX_DEV 64
static std::vector<cudaEvent_t> g_pool[PROF_MAX_DEV];

static int prof_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < PROF_MAX_DEV) ? d : 0;
}

static cudaEvent_t take_event(int dev)
{
    if (!g_pool[dev].empty())
    {
        cudaEvent_t e = g_pool[dev].back();
        g_pool[dev].pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void prof_pre(int id, cudaStream_t st)
{
    g_slots[id].launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    const int   dev = prof_device();
    cudaEvent_t a = take_event(dev), b = take_event(dev);
    cudaEventRecord(a, st);
    g_slots[id].pending.push_back({a, b, dev});
}

void prof_post(int id, cudaStream_t st)
{
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_slots[id].pending.empty()) cudaEventRecord(g_slots[id].pending.back().b, st);
}

uint64_t prof_total_launches()
{
    uint64_t t = 0;
    for (int i = 0; i < P_COUNT; ++i) t += g_slots[i].launches.load(std::memory_order_relaxed);
    return t;
}

static void prof_drain()
{
    for (int i = 0; i < P_COUNT; ++i)
    {
        for (auto& pr : g_slots[i].pending)
        {
            float ms = 0;
            if (cudaEventSynchronize(pr.b) == cudaSuccess && cudaEventElapsedTime(&ms, pr.a, pr.b) == cudaSuccess) g_slots[i].ms += ms;
            g_pool[pr.dev].push_back(pr.a);
            g_pool[pr.dev].push_back(pr.b);
        }
        g_slots[i].pending.clear();
    }
}

}  // namespace bra

using namespace bra;

extern "C" void bra_b200_prof_enable(int timing_on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain();
    g_timing = timing_on != 0;
}

extern "C" void bra_b200_prof_reset(void)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain();
    for (int i = 0; i < P_COUNT; ++i)
    {
        g_slots[i].launches.store(0);
        g_slots[i].ms = 0;
    }
}

extern "C" int bra_b200_prof_count(void) { return P_COUNT; }

extern "C" int bra_b200_prof_read(int id, const char** name, uint64_t* launches, double* ms)
{
    if (id < 0 || id >= P_COUNT) return 1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain();
    if (name) *name = kProfNames[id];
    if (launches) *launches = g_slots[id].launches.load();
    if (ms) *ms = g_slots[id].ms;
    return 0;
}
