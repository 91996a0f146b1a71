// sort.cu -- batched, segmented, stable LSD radix-sort pass, one sweep per digit (8-bit digits).
//
// Used three ways on the hot path:
//   * forward BWT: initial sort of rotation indices by their first 4 symbols (2-4 passes) and the
//     per-doubling-round stable re-bucketing by rank (1-3 passes) -- replaces the qsort_r call of
//     reference src/encoders/bra_bwt.c:91. (A 10-bit digit, 2 passes per round, was measured in round 1: the wider
//     shared-memory counters and 4-element output runs make each pass ~1.7x slower, a net loss.)
//   * inverse BWT: one 8-bit pass with the BWT bytes as keys builds `transform[]`, i.e. the stable
//     counting sort of reference bra_bwt.c:142-159.
//
// One pass = ONE kernel over every block of the batch (blocks never mix): read key+value once, write key+value once.
//   * The digit histogram of each block (256 counters) is produced beforehand by whoever writes the keys (all digits of
//     a sort at once), so no pass re-reads its keys to count them.
//   * Each CTA takes a tile of 4096 elements by ticket (an atomic counter: a tile only ever waits for tiles with smaller
//     tickets, which are running or finished, so the waits cannot deadlock whatever the launch order; the tickets of up to
//     64 blocks are interleaved, so that a tile's predecessors in its own block are many tickets old and have usually
//     published their inclusive counts), ranks its elements stably per warp -- match.any, or lane bits OR-ed into
//     shared-memory words per digit when the warp's digits are diverse -- and obtains the number of equal digits in the
//     earlier tiles of its block by decoupled look-back over per-tile status words (count | flag in one 32-bit word).
//   * The tile is reordered in shared memory so that each digit's run leaves the SM as one contiguous store.
//   * The values of a tile are not needed before the reorder: they are fetched by one bulk-asynchronous copy (TMA,
//     cp.async.bulk + mbarrier) issued when the CTA starts and land in shared memory while the ranking runs.
// Algorithmic traffic per pass and element: 16 bytes (u32 key + value, read + write). HBM/L2 bound in bytes, issue
// and shared-memory bound in practice (ranking, reorder); no tensor-core work.
#include "bra_common.cuh"
#include "bra_kernels.h"

#include <algorithm>
#include <stdlib.h>

namespace bra {

#define RS_TILE 4096
#define RS_THREADS 256
#define RS_ITEMS 16  // per thread
#define RS_RADIX 256
#define RS_LOOKBACK 8  // predecessor tiles whose status words are fetched together

// status word of the look-back: [31:30] 0 = empty, 1 = tile count, 2 = inclusive prefix; [29:0] value
#define RS_FLAG_LOCAL 1u
#define RS_FLAG_INCL 2u
#define RS_VAL_MASK 0x3FFFFFFFu

enum
{
    RS_PAIRS,     // keys and values from memory (values staged by a bulk-asynchronous copy)
    RS_IMPLICIT,  // values are the element indices 0, 1, 2, ...
    RS_U8_PACK    // inverse BWT: u8 keys, output word (index << 8) | key
};

struct RsSmem
{
    unsigned long long mbar;               // completion barrier of the bulk copy
    uint32_t           ticket;
    uint32_t           red[34];
    unsigned short     wcnt[8][RS_RADIX];  // per-warp digit counters, then position of the warp's first element of the digit in the sorted tile
    uint32_t           gdelta[RS_RADIX];   // (global destination of the digit's run for this tile) - (its start in the sorted tile), mod 2^32
    __align__(128) uint2    pair[RS_TILE]; // sorted tile: (key, value)   [RS_U8_PACK: the first half, as packed words]
    __align__(128) uint32_t vin[RS_TILE];  // values in input order (landing buffer of the bulk copy)
};

struct RsParams
{
    const void*     keys;      // u32 keys (u8 for RS_U8_PACK)
    const uint32_t* vals;
    uint32_t*       keys_out;
    uint32_t*       vals_out;
    uint64_t        stride;
    const uint32_t* len;
    const uint8_t*  skip;
    uint32_t        shift, tiles, group, nblk;
    uint32_t        match_max_distinct;  // warps whose first round has more distinct digits rank through shared-memory masks
    const uint32_t* ghist;     // digit histogram of block b for this pass at ghist[b * RS_GHIST_STRIDE + d]
    uint32_t*       status;    // [nblk][tiles][256], zero before the launch
    uint32_t*       ticket;    // one counter, zero before the launch
};

// ---- bulk-asynchronous copy (TMA, non-tensor form) and its completion barrier --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int MODE>
__device__ __forceinline__ uint32_t rs_digit(uint32_t k, uint32_t shift)
{
    return MODE == RS_U8_PACK ? k : ((k >> shift) & 0xFFu);
}

// The body of the kernel for one tile; FULL = the tile has all RS_TILE elements (no bounds tests anywhere).
template <int MODE, bool FULL>
__device__ __forceinline__ void rs_tile(const RsParams& P, RsSmem& S, const uint32_t b, const uint32_t t, const uint32_t n, const uint32_t tn)
{
    const uint32_t w = warp_id(), l = lane_id();
    const uint32_t tile0 = t * RS_TILE;
    const uint64_t base  = (uint64_t) b * P.stride;
#define RS_IN(e) (FULL || (e) < tn)

    // ---- values of a full, 16-byte aligned tile: one bulk-asynchronous copy that lands while the ranking runs
    const bool bulk = MODE == RS_PAIRS && FULL && (((base + tile0) & 3u) == 0);
    if (MODE == RS_PAIRS)
    {
        if (bulk)
        {
            if (threadIdx.x == 0)
            {
                mbar_init(&S.mbar, 1);
                fence_mbar_init();
                mbar_expect_tx(&S.mbar, RS_TILE * 4);
                bulk_g2s(S.vin, P.vals + base + tile0, RS_TILE * 4, &S.mbar);
            }
        }
        else
            for (uint32_t e = threadIdx.x; e < tn; e += RS_THREADS) S.vin[e] = P.vals[base + tile0 + e];
    }

    // ---- keys. Warp w owns elements [w*512, w*512+512) of the tile, visited in 16 rounds of 32 (memory order)
    uint32_t k[RS_ITEMS];
    {
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
        {
            const uint32_t e = w * 512 + r * 32 + l;
            if (MODE == RS_U8_PACK)
                k[r] = RS_IN(e) ? (uint32_t) static_cast<const uint8_t*>(P.keys)[base + tile0 + e] : 0u;
            else
                k[r] = RS_IN(e) ? static_cast<const uint32_t*>(P.keys)[base + tile0 + e] : 0u;
        }
    }
    // start of every digit's bucket inside the block (the loads above are in flight meanwhile)
    const uint32_t gbase = block_excl_add(P.ghist[(uint64_t) b * RS_GHIST_STRIDE + threadIdx.x], S.red, nullptr);

    // ---- ranking: position of every element among the elements of its warp with the same digit, in memory order.
    // The peers of a lane (the lanes of its round with the same digit) come from match.any, whose cost grows with the
    // number of distinct digits in the warp, or -- when the digits of the warp are diverse -- from an OR of lane bits
    // into a shared-memory word per digit, whose cost grows with the number of EQUAL digits instead. Each warp picks by
    // looking at its first round. (Digit RS_RADIX = padding of the ragged last tile, matches only padding.)
    unsigned short rk[RS_ITEMS];  // rank among equal digits inside the warp (< 512)
    const uint32_t peers0   = __match_any_sync(BRA_FULL, RS_IN(w * 512 + l) ? rs_digit<MODE>(k[0], P.shift) : (uint32_t) RS_RADIX);
    const uint32_t distinct = __popc(__ballot_sync(BRA_FULL, (uint32_t) (__ffs(peers0) - 1) == l));
    if (distinct <= P.match_max_distinct)
    {
        // Phase A: the peer masks of all 16 rounds are independent of each other -- issue every match first so that
        // their latency overlaps.
        uint32_t peers[RS_ITEMS];
        peers[0] = peers0;
#pragma unroll
        for (int r = 1; r < RS_ITEMS; ++r)
        {
            const uint32_t e = w * 512 + r * 32 + l;
            peers[r]         = __match_any_sync(BRA_FULL, RS_IN(e) ? rs_digit<MODE>(k[r], P.shift) : (uint32_t) RS_RADIX);
        }
        // Phase B: running per-warp digit counters, one round after the other (memory order = stability)
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
        {
            const uint32_t e      = w * 512 + r * 32 + l;
            const uint32_t d      = rs_digit<MODE>(k[r], P.shift);
            const uint32_t before = __popc(peers[r] & lanemask_lt());
            const int      leader = __ffs(peers[r]) - 1;
            uint32_t       old    = 0;
            if (RS_IN(e) && (int) l == leader)
            {
                old          = S.wcnt[w][d];
                S.wcnt[w][d] = (unsigned short) (old + __popc(peers[r]));
            }
            old   = __shfl_sync(BRA_FULL, old, leader);
            rk[r] = (unsigned short) (old + before);
            __syncwarp();
        }
    }
    else
    {
        // Three mask buffers in rotation ([3][256] words per warp, in the not yet used sorted-tile buffer, zeroed at kernel
        // start): the words of round r are cleared by the round's leaders after the warp barrier of round r+1 (every lane
        // has read them by then) and used again in round r+3, i.e. after the barrier of round r+2.
        uint32_t* const mk        = reinterpret_cast<uint32_t*>(S.pair) + w * (3 * RS_RADIX);
        uint32_t*       prev_slot = nullptr;
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
        {
            const uint32_t e    = w * 512 + r * 32 + l;
            const uint32_t d    = rs_digit<MODE>(k[r], P.shift);
            uint32_t* const slot = mk + (r % 3) * RS_RADIX + d;
            if (RS_IN(e)) atomicOr(slot, 1u << l);
            __syncwarp();
            const uint32_t pm = RS_IN(e) ? *reinterpret_cast<volatile uint32_t*>(slot) : 0u;
            if (prev_slot) *prev_slot = 0;
            const uint32_t before = __popc(pm & lanemask_lt());
            const int      leader = __ffs(pm) - 1;
            uint32_t       old    = 0;
            prev_slot             = nullptr;
            if (RS_IN(e) && (int) l == leader)
            {
                old          = S.wcnt[w][d];
                S.wcnt[w][d] = (unsigned short) (old + __popc(pm));
                prev_slot    = slot;
            }
            old   = __shfl_sync(BRA_FULL, old, leader);
            rk[r] = (unsigned short) (old + before);
        }
    }
    __syncthreads();

    // ---- thread d owns digit d: tile count, position of every warp's share, look-back over the earlier tiles of the block
    {
        const uint32_t d = threadIdx.x;
        uint32_t       cw[8], cnt = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww)
        {
            cw[ww] = S.wcnt[ww][d];
            cnt += cw[ww];
        }
        uint32_t* const tile_status = P.status + ((uint64_t) b * P.tiles) * RS_RADIX + d;  // + tile * RS_RADIX
        st_relaxed_u32(tile_status + (uint64_t) t * RS_RADIX, ((t == 0 ? RS_FLAG_INCL : RS_FLAG_LOCAL) << 30) | cnt);
        // the nearest predecessor's word is fetched now and looked at after the scan below: with the blocks interleaved it
        // nearly always carries the inclusive prefix already, and its L2 latency hides behind the scan
        const uint32_t early = t != 0 ? ld_relaxed_u32(tile_status + (uint64_t) (t - 1) * RS_RADIX) : 0u;
        const uint32_t dstart = block_excl_add(cnt, S.red, nullptr);  // start of the digit's run in the sorted tile
        uint32_t       run    = dstart;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww)
        {
            S.wcnt[ww][d] = (unsigned short) run;
            run += cw[ww];
        }
        uint32_t excl = 0;
        if (t != 0)
        {
            // Decoupled look-back, RS_LOOKBACK predecessors per trip: their status words are independent loads. A tile
            // only waits for tiles with smaller tickets, which are running or finished.
            int      tt    = (int) t - 1;
            bool     open  = true;
            uint32_t trips = 0;
            if ((early >> 30) != 0u)
            {
                excl = early & RS_VAL_MASK;
                --tt;
                open = (early >> 30) != RS_FLAG_INCL;
            }
            while (open)
            {
                if (++trips > (1u << 24)) __trap();  // seconds of waiting: a predecessor never published -- fail loudly instead of hanging
                uint32_t v[RS_LOOKBACK];
#pragma unroll
                for (int i = 0; i < RS_LOOKBACK; ++i)
                    v[i] = tt - i >= 0 ? ld_relaxed_u32(tile_status + (uint64_t) (tt - i) * RS_RADIX) : (RS_FLAG_INCL << 30);
                const int tt0 = tt;
#pragma unroll
                for (int i = 0; i < RS_LOOKBACK; ++i)
                {
                    if (open && (v[i] >> 30) != 0u)
                    {
                        excl += v[i] & RS_VAL_MASK;
                        --tt;
                        if ((v[i] >> 30) == RS_FLAG_INCL) open = false;
                    }
                    else if (open)
                        break;  // not published yet: fetch again from this tile on
                }
                if (open && tt == tt0) __nanosleep(100);
            }
            st_relaxed_u32(tile_status + (uint64_t) t * RS_RADIX, (RS_FLAG_INCL << 30) | (excl + cnt));
        }
        S.gdelta[d] = gbase + excl - dstart;
    }
    if (bulk) mbar_wait(&S.mbar, 0);
    __syncthreads();

    // ---- reorder into the sorted tile
    uint32_t* const pw = reinterpret_cast<uint32_t*>(S.pair);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t e = w * 512 + r * 32 + l;
        if (RS_IN(e))
        {
            const uint32_t pos = (uint32_t) S.wcnt[w][rs_digit<MODE>(k[r], P.shift)] + rk[r];
            if (MODE == RS_U8_PACK)
                pw[pos] = ((tile0 + e) << 8) | k[r];
            else
                S.pair[pos] = make_uint2(k[r], MODE == RS_IMPLICIT ? (tile0 + e) : S.vin[e]);
        }
    }
    __syncthreads();

    // ---- every digit's run leaves as one contiguous store
    if (MODE == RS_U8_PACK)
    {
        uint32_t* const out = P.vals_out + base;
#pragma unroll 4
        for (uint32_t e = threadIdx.x; e < tn; e += RS_THREADS)
        {
            const uint32_t pv                = pw[e];
            out[S.gdelta[pv & 0xFFu] + e] = pv;
        }
    }
    else
    {
        uint32_t* const ko = P.keys_out + base;
        uint32_t* const vo = P.vals_out + base;
#pragma unroll 4
        for (uint32_t e = threadIdx.x; e < tn; e += RS_THREADS)
        {
            const uint2    kv  = S.pair[e];
            const uint32_t dst = S.gdelta[(kv.x >> P.shift) & 0xFFu] + e;
            ko[dst]            = kv.x;
            vo[dst]            = kv.y;
        }
    }
#undef RS_IN
}

template <int MODE>
__global__ void __launch_bounds__(RS_THREADS, 4) rs_onesweep_kernel(const RsParams P)
{
    extern __shared__ __align__(128) uint8_t rs_smem_raw[];
    RsSmem& S = *reinterpret_cast<RsSmem*>(rs_smem_raw);

    if (threadIdx.x == 0) S.ticket = atomicAdd(P.ticket, 1u);
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&S.wcnt[0][0]);
        for (uint32_t i = threadIdx.x; i < 8 * RS_RADIX / 2; i += RS_THREADS) z[i] = 0;
        uint4* m = reinterpret_cast<uint4*>(S.pair);  // lane masks of the ranking, 8 warps x 3 x 256 words
        for (uint32_t i = threadIdx.x; i < 8 * 3 * RS_RADIX / 4; i += RS_THREADS) m[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    // Tickets run over groups of P.group blocks, tile-major inside a group: the tiles of one block arrive P.group
    // tickets apart, so a tile's predecessors have usually finished their own look-back when it starts its own, while
    // the blocks of a group (ranks, keys, partially written output lines) still share the L2.
    const uint32_t per_group = P.group * P.tiles;
    const uint32_t g = S.ticket / per_group, r = S.ticket % per_group;
    const uint32_t b = g * P.group + r % P.group, t = r / P.group;
    if (b >= P.nblk) return;
    if (P.skip && P.skip[b]) return;
    const uint32_t n = P.len[b];
    if (t * RS_TILE >= n) return;
    const uint32_t tn = min((uint32_t) RS_TILE, n - t * RS_TILE);
    if (tn == RS_TILE)
        rs_tile<MODE, true>(P, S, b, t, n, tn);
    else
        rs_tile<MODE, false>(P, S, b, t, n, tn);
}

// ---- per-block byte histogram (the inverse BWT's only pass sorts the BWT bytes themselves) ---------------------------
__global__ void __launch_bounds__(256) rs_ghist_u8_kernel(const uint8_t* __restrict__ keys, uint64_t stride, const uint32_t* __restrict__ len,
                                                          uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[RS_RADIX];
    const uint32_t      b     = blockIdx.y;
    const uint32_t      n     = len[b];
    const uint32_t      tile0 = blockIdx.x * 16384;
    if (tile0 >= n) return;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* p   = keys + (uint64_t) b * stride;
    const uint32_t end = min(n, tile0 + 16384u);
    const uint32_t mis = (uint32_t) ((16u - (reinterpret_cast<uintptr_t>(p + tile0) & 15u)) & 15u);  // bytes before the first aligned 16
    for (uint32_t j = tile0 + threadIdx.x; j < min(end, tile0 + mis); j += 256) atomicAdd(&h[p[j]], 1u);
    for (uint32_t i = tile0 + mis + threadIdx.x * 16; i < end; i += 256 * 16)
    {
        if (i + 16 <= end)
        {
            const uint4    v    = *reinterpret_cast<const uint4*>(p + i);
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(&h[(x[j >> 2] >> ((j & 3) * 8)) & 0xFFu], 1u);
        }
        else
            for (uint32_t j = i; j < end; ++j) atomicAdd(&h[p[j]], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&ghist[(uint64_t) b * RS_GHIST_STRIDE + threadIdx.x], h[threadIdx.x]);
}

// Workspace of the sort passes (u32 words): [nblk][RS_GHIST_PASSES][256] digit histograms, filled by whoever produces the
// keys; then 64 words holding the ticket counter; then the look-back status [nblk][tiles][256].
size_t radix_hist_bytes(uint32_t max_len, uint32_t nblk)
{
    return ((size_t) nblk * RS_GHIST_STRIDE + 64 + (size_t) nblk * RS_RADIX * bra_div_up(max_len, RS_TILE)) * sizeof(uint32_t);
}

template <int MODE>
static bool radix_launch(RsParams& P, uint32_t* d_hist, uint32_t pass, uint32_t max_len, uint32_t nblk, int prof_id, cudaStream_t st)
{
    if (nblk == 0 || max_len == 0) return true;
    P.tiles  = bra_div_up(max_len, RS_TILE);
    P.nblk   = nblk;
    static const uint32_t thresh = getenv("BRA_B200_MATCH_MAX") ? (uint32_t) atoi(getenv("BRA_B200_MATCH_MAX")) : 16u;  // tuning switch, read once
    P.match_max_distinct = thresh;
    // blocks whose tiles are interleaved in ticket order (each block streams through its own 256 write fronts: the partially
    // written lines of sixteen blocks are a negligible share of L2)
    static const uint32_t group_max = getenv("BRA_B200_RS_GROUP") ? (uint32_t) std::max(1, atoi(getenv("BRA_B200_RS_GROUP"))) : 64u;  // tuning switch, read once
    P.group  = std::min<uint32_t>(group_max, nblk);
    P.ghist  = d_hist + (size_t) pass * RS_RADIX;
    P.ticket = d_hist + (size_t) nblk * RS_GHIST_STRIDE;
    P.status = P.ticket + 64;
    BRA_CUDA_TRY(cudaMemsetAsync(P.ticket, 0, (64 + (size_t) nblk * P.tiles * RS_RADIX) * sizeof(uint32_t), st));
    const size_t smem = MODE == RS_PAIRS ? sizeof(RsSmem) : offsetof(RsSmem, vin);
    // per-device attribute: set it on every call (cheap) so that multi-GPU processes are covered
    BRA_CUDA_TRY(cudaFuncSetAttribute(rs_onesweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    BRA_LAUNCH(prof_id, st, rs_onesweep_kernel<MODE><<<bra_div_up(nblk, P.group) * P.group * P.tiles, RS_THREADS, smem, st>>>(P));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

bool radix_pass_u32(const uint32_t* keys, const uint32_t* vals, uint32_t* keys_out, uint32_t* vals_out, uint64_t stride, const uint32_t* d_len,
                    const uint8_t* d_skip, uint32_t max_len, uint32_t nblk, uint32_t pass, uint32_t key_shift, uint32_t* d_hist, cudaStream_t st)
{
    RsParams P{};
    P.keys = keys; P.vals = vals; P.keys_out = keys_out; P.vals_out = vals_out; P.stride = stride; P.len = d_len; P.skip = d_skip; P.shift = key_shift + pass * 8;
    if (vals == nullptr)  // values are the element indices 0, 1, 2, ... (first pass of a sort): nothing to read
        return radix_launch<RS_IMPLICIT>(P, d_hist, pass, max_len, nblk, P_RS_SCATTER_IMPL, st);
    return radix_launch<RS_PAIRS>(P, d_hist, pass, max_len, nblk, P_RS_SCATTER, st);
}

bool radix_pass_u8_index_packed(const uint8_t* keys, uint32_t* packed_out, uint64_t stride, const uint32_t* d_len, uint32_t max_len, uint32_t nblk,
                                uint32_t* d_hist, cudaStream_t st)
{
    if (nblk == 0 || max_len == 0) return true;
    BRA_CUDA_TRY(cudaMemsetAsync(d_hist, 0, (size_t) nblk * RS_GHIST_STRIDE * sizeof(uint32_t), st));
    BRA_LAUNCH(P_RS_HIST, st, rs_ghist_u8_kernel<<<dim3(bra_div_up(max_len, 16384), nblk), 256, 0, st>>>(keys, stride, d_len, d_hist));
    RsParams P{};
    P.keys = keys; P.vals_out = packed_out; P.stride = stride; P.len = d_len; P.shift = 0;
    return radix_launch<RS_U8_PACK>(P, d_hist, 0, max_len, nblk, P_RS_SCATTER_U8, st);
}

}  // namespace bra
