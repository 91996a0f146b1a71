// sort.cu -- batched, segmented, stable LSD radix-sort pass (8-bit digit; the kernels are templated on the digit width).
//
// Used three ways on the hot path:
//   * forward BWT: initial sort of rotation indices by their first 4 bytes (4 passes) and the
//     per-doubling-round stable re-bucketing by rank (2-3 passes) -- replaces the qsort_r call of
//     reference src/encoders/bra_bwt.c:91. (A 10-bit digit, 2 passes per round, was measured: the wider
//     shared-memory counters and 4-element output runs make each pass ~1.7x slower, a net loss.)
//   * inverse BWT: one 8-bit pass with the BWT bytes as keys builds `transform[]`, i.e. the stable
//     counting sort of reference bra_bwt.c:142-159.
//
// One pass = three kernels over every block of the batch (blocks never mix):
//   hist    : per 4096-element tile, shared-memory histogram            -> hist[b][digit][tile]
//   scan    : per block, exclusive scan in (digit, tile) order          -> global offsets
//   scatter : per tile, stable in-tile ranking with warp match/ballot, shared-memory reorder so
//             that each digit's run leaves the SM as one contiguous store, then the scatter.
// Algorithmic traffic per pass and element: read key+value, write key+value (+ key re-read by
// hist). HBM/L2 bound; no tensor-core work. The CTAs of one block are adjacent in launch order, so
// the partially written output lines of a block are completed in L2 before they are evicted.
#include "bra_common.cuh"
#include "bra_kernels.h"

namespace bra {

#define RS_TILE 4096
#define RS_THREADS 256
#define RS_ITEMS 16  // per thread

template <int BITS, typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, uint32_t shift)
{
    return ((uint32_t) k >> shift) & ((1u << BITS) - 1u);
}

// ------------------------------------------------------------------------------------ hist
template <int BITS, typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KeyT* __restrict__ keys, uint64_t stride, const uint32_t* __restrict__ len,
                                                             const uint8_t* __restrict__ skip, uint32_t shift, uint32_t tiles,
                                                             uint32_t* __restrict__ hist)
{
    constexpr uint32_t  RADIX = 1u << BITS;
    __shared__ uint32_t h[RADIX];
    const uint32_t      b = blockIdx.y, t = blockIdx.x;
    if (skip && skip[b]) return;
    const uint32_t n     = len[b];
    const uint32_t tile0 = t * RS_TILE;
    uint32_t*      out   = hist + ((uint64_t) b * RADIX) * tiles + t;  // [b][digit][tile]
    if (tile0 >= n)
    {
        for (uint32_t d = threadIdx.x; d < RADIX; d += RS_THREADS) out[(uint64_t) d * tiles] = 0;
        return;
    }
    for (uint32_t d = threadIdx.x; d < RADIX; d += RS_THREADS) h[d] = 0;
    __syncthreads();
    const KeyT*    k  = keys + (uint64_t) b * stride + tile0;
    const uint32_t tn = min((uint32_t) RS_TILE, n - tile0);
    KeyT           kv[RS_ITEMS];  // all loads in flight before the first shared-memory atomic
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t i = threadIdx.x + r * RS_THREADS;
        kv[r]            = i < tn ? k[i] : (KeyT) 0;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
        if (threadIdx.x + r * RS_THREADS < tn) atomicAdd(&h[rs_digit<BITS>(kv[r], shift)], 1u);
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < RADIX; d += RS_THREADS) out[(uint64_t) d * tiles] = h[d];
}

// ------------------------------------------------------------------------------------ scan
// One CTA per block: exclusive scan of radix*tiles counters in place.
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ hist, const uint8_t* __restrict__ skip, uint32_t tiles, uint32_t radix)
{
    __shared__ uint32_t red[33];
    const uint32_t      b = blockIdx.x;
    if (skip && skip[b]) return;
    uint32_t*      h     = hist + ((uint64_t) b * radix) * tiles;
    const uint32_t total = radix * tiles;
    uint32_t       carry = 0;
    for (uint32_t base = 0; base < total; base += 1024 * 4)
    {
        // 4 consecutive counters per thread
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t       v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < total) ? h[i0 + j] : 0u;
        const uint32_t s = v[0] + v[1] + v[2] + v[3];
        uint32_t       tot;
        uint32_t       ex = block_excl_add(s, red, &tot) + carry;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (i0 + j < total) h[i0 + j] = ex;
            ex += v[j];
        }
        carry += tot;
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------- scatter
// OUT_MODE 0: write keys_out and vals_out; 1: write vals_out only;
//          2: write packed (val << 8) | key8 into vals_out (inverse-BWT "next row | byte" word)
template <int BITS, typename KeyT>
struct RsSmem
{
    static constexpr uint32_t RADIX = 1u << BITS;
    unsigned short            wcnt[8][RADIX];  // per-warp digit counters -> per-warp bases (< 4096)
    unsigned short            dstart[RADIX];   // start of each digit's run inside the sorted tile
    uint32_t                  gdelta[RADIX];   // (global destination of the digit's run for this tile) - dstart, mod 2^32
    KeyT                      skey[RS_TILE];
    uint32_t                  sval[RS_TILE];
    uint32_t                  red[34];
};

template <int BITS, typename KeyT, bool IMPLICIT_VALS, int OUT_MODE>
__global__ void __launch_bounds__(RS_THREADS, 4)
    rs_scatter_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, KeyT* __restrict__ keys_out,
                      uint32_t* __restrict__ vals_out, uint64_t stride, const uint32_t* __restrict__ len,
                      const uint8_t* __restrict__ skip, uint32_t shift, uint32_t tiles, const uint32_t* __restrict__ hist)
{
    constexpr uint32_t RADIX = 1u << BITS;
    constexpr uint32_t DPT   = RADIX / RS_THREADS;  // digits per thread in the digit scan
    extern __shared__ __align__(16) uint8_t rs_smem_raw[];
    RsSmem<BITS, KeyT>& S = *reinterpret_cast<RsSmem<BITS, KeyT>*>(rs_smem_raw);

    const uint32_t b = blockIdx.y, t = blockIdx.x;
    if (skip && skip[b]) return;
    const uint32_t n     = len[b];
    const uint32_t tile0 = t * RS_TILE;
    if (tile0 >= n) return;
    const uint32_t tn   = min((uint32_t) RS_TILE, n - tile0);
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t w = warp_id(), l = lane_id();

    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&S.wcnt[0][0]);
        for (uint32_t i = threadIdx.x; i < 8 * RADIX / 2; i += RS_THREADS) z[i] = 0;
    }
    __syncthreads();

    // warp w owns elements [w*512, w*512+512) of the tile, visited in 16 rounds of 32 (memory order)
    KeyT           k[RS_ITEMS];
    unsigned short rk[RS_ITEMS];  // rank among equal digits inside the warp (< 512)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t e = w * 512 + r * 32 + l;
        k[r]             = e < tn ? keys[base + tile0 + e] : (KeyT) 0;
    }
    // Phase A: the peer masks of all 16 rounds are independent of each other -- issue every match first so
    // that their latency overlaps (the profile showed the warp waiting on one match at a time otherwise).
    uint32_t peers[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t e = w * 512 + r * 32 + l;
        const uint32_t d = e < tn ? rs_digit<BITS>(k[r], shift) : RADIX;  // RADIX = padding, matches only padding
        peers[r]         = __match_any_sync(BRA_FULL, d);
    }
    // Phase B: running per-warp digit counters, one round after the other (memory order = stability)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t e  = w * 512 + r * 32 + l;
        const bool     ok = e < tn;
        const uint32_t d  = ok ? rs_digit<BITS>(k[r], shift) : RADIX;
        const uint32_t before = __popc(peers[r] & lanemask_lt());
        const int      leader = __ffs(peers[r]) - 1;
        uint32_t       old    = 0;
        if (ok && (int) l == leader)
        {
            old          = S.wcnt[w][d];
            S.wcnt[w][d] = (unsigned short) (old + __popc(peers[r]));
        }
        old   = __shfl_sync(BRA_FULL, old, leader);
        rk[r] = (unsigned short) (old + before);
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive scan over warps; digit totals -> exclusive scan over digits (DPT consecutive digits per thread)
    {
        uint32_t tot[DPT], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < DPT; ++i)
        {
            const uint32_t d   = threadIdx.x * DPT + i;
            uint32_t       run = 0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww)
            {
                const uint32_t c = S.wcnt[ww][d];
                S.wcnt[ww][d]    = (unsigned short) run;
                run += c;
            }
            tot[i] = run;
            sum += run;
        }
        uint32_t ex = block_excl_add(sum, S.red, nullptr);
#pragma unroll
        for (uint32_t i = 0; i < DPT; ++i)
        {
            const uint32_t d = threadIdx.x * DPT + i;
            S.dstart[d]      = (unsigned short) ex;
            S.gdelta[d]      = hist[((uint64_t) b * RADIX + d) * tiles + t] - ex;
            ex += tot[i];
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const uint32_t e = w * 512 + r * 32 + l;
        if (e < tn)
        {
            const uint32_t d   = rs_digit<BITS>(k[r], shift);
            const uint32_t pos = (uint32_t) S.dstart[d] + S.wcnt[w][d] + rk[r];
            S.skey[pos]        = k[r];
            S.sval[pos]        = IMPLICIT_VALS ? (tile0 + e) : vals[base + tile0 + e];  // values are only touched here (keeps registers low)
        }
    }
    __syncthreads();

    for (uint32_t e = threadIdx.x; e < tn; e += RS_THREADS)
    {
        const KeyT     kk  = S.skey[e];
        const uint32_t d   = rs_digit<BITS>(kk, shift);
        const uint64_t dst = base + (uint32_t) (S.gdelta[d] + e);
        if (OUT_MODE == 0)
        {
            keys_out[dst] = kk;
            vals_out[dst] = S.sval[e];
        }
        else if (OUT_MODE == 1)
            vals_out[dst] = S.sval[e];
        else
            vals_out[dst] = (S.sval[e] << 8) | (uint32_t) kk;
    }
}

template <int BITS, typename KeyT, bool IMPLICIT, int OUT_MODE>
static bool radix_pass_t(const KeyT* keys, const uint32_t* vals, KeyT* keys_out, uint32_t* vals_out, uint64_t stride, const uint32_t* d_len,
                         const uint8_t* d_skip, uint32_t max_len, uint32_t nblk, uint32_t shift, bool hist_ready, uint32_t* d_hist, cudaStream_t st)
{
    if (nblk == 0 || max_len == 0) return true;
    const uint32_t tiles = bra_div_up(max_len, RS_TILE);
    const size_t   smem  = sizeof(RsSmem<BITS, KeyT>);
    // per-device attribute: set it on every call (cheap) so that multi-GPU processes are covered
    BRA_CUDA_TRY(cudaFuncSetAttribute(rs_scatter_kernel<BITS, KeyT, IMPLICIT, OUT_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const int scatter_id = sizeof(KeyT) == 1 ? P_RS_SCATTER_U8 : P_RS_SCATTER;
    if (!hist_ready)  // the producer of the keys may already have filled d_hist for this digit
        BRA_LAUNCH(P_RS_HIST, st, rs_hist_kernel<BITS, KeyT><<<dim3(tiles, nblk), RS_THREADS, 0, st>>>(keys, stride, d_len, d_skip, shift, tiles, d_hist));
    BRA_LAUNCH(P_RS_SCAN, st, rs_scan_kernel<<<nblk, 1024, 0, st>>>(d_hist, d_skip, tiles, 1u << BITS));
    BRA_LAUNCH(scatter_id, st, rs_scatter_kernel<BITS, KeyT, IMPLICIT, OUT_MODE>
        <<<dim3(tiles, nblk), RS_THREADS, smem, st>>>(keys, vals, keys_out, vals_out, stride, d_len, d_skip, shift, tiles, d_hist));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

size_t radix_hist_bytes(uint32_t max_len, uint32_t nblk) { return (size_t) nblk * 256 * bra_div_up(max_len, RS_TILE) * sizeof(uint32_t); }

bool radix_pass_u32(const uint32_t* keys, const uint32_t* vals, uint32_t* keys_out, uint32_t* vals_out, uint64_t stride,
                    const uint32_t* d_len, const uint8_t* d_skip, uint32_t max_len, uint32_t nblk, uint32_t shift, uint32_t bits,
                    bool hist_ready, uint32_t* d_hist, cudaStream_t st)
{
    (void) bits;  // only the 8-bit digit is instantiated
    if (vals == nullptr)  // values are the element indices 0, 1, 2, ... (first pass of a sort): nothing to read
        return radix_pass_t<8, uint32_t, true, 0>(keys, nullptr, keys_out, vals_out, stride, d_len, d_skip, max_len, nblk, shift, hist_ready, d_hist, st);
    return radix_pass_t<8, uint32_t, false, 0>(keys, vals, keys_out, vals_out, stride, d_len, d_skip, max_len, nblk, shift, hist_ready, d_hist, st);
}

bool radix_pass_u8_index_packed(const uint8_t* keys, uint32_t* packed_out, uint64_t stride, const uint32_t* d_len, uint32_t max_len,
                                uint32_t nblk, uint32_t* d_hist, cudaStream_t st)
{
    return radix_pass_t<8, uint8_t, true, 2>(keys, nullptr, nullptr, packed_out, stride, d_len, nullptr, max_len, nblk, 0, false, d_hist, st);
}

}  // namespace bra
