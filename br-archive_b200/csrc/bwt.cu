// bwt.cu -- forward and inverse Burrows-Wheeler transform for a batch of blocks.
//
// FORWARD (replaces reference bra_bwt_encode2, src/encoders/bra_bwt.c:73-108)
//   The reference sorts the n cyclic rotations with qsort_r and a byte-wise comparator and
//   breaks ties between identical rotations by ascending start index (glibc's stable merge
//   sort). Here the same order is built as a suffix array of the cyclic string:
//     1. period detection: the smallest p | n with T[i] == T[(i+p) mod n]. Rotations i and i+p
//        are identical, so it suffices to sort the p rotations of the primitive root (all
//        distinct) and replicate each n/p times in ascending index -- exactly the tie rule. A 32-byte pretest refutes
//        the divisors; only a block with a surviving candidate compares itself with its shift, one candidate per round.
//     2. keys of the first 4 symbols (dense per-block symbol codes) + 2-4 LSD radix passes (sort.cu).
//     3. prefix doubling, Manber-Myers style: walking the current order j = 0..p-1, rotation
//        SA[j]-h is appended to its own h-group, which a stable radix sort on rank[SA[j]-h]
//        does in 1-3 passes; new group heads give the 2h-ranks. Where the fields fit 64 bits the sort moves packed
//        records (key, group of the traversal slot, rotation): the new heads are then a neighbour compare, no gather.
//        Repeats until every group is a singleton (all rotations of a primitive string differ, so this terminates with
//        h < 2p); blocks whose remaining groups are small are finished by direct comparison of the rotations.
//     4. gather L[j] = T[SA[j/k] - 1], primary = k * (row of rotation 0), k = n/p.
//
// INVERSE (replaces reference bra_bwt_decode2, bra_bwt.c:133-168)
//   transform[] is the stable sort of positions by byte value = one 8-bit radix pass, emitted
//   as W[j] = transform[j] << 8 | F[j] so that the chase needs one load per output byte. The
//   n-step dependent chase is cut into ~n/R independent walks (16384 per block: twice the walkers halve the blocks whose W
//   arrays are in flight at once, and with them the DRAM sectors fetched per step) that start at every R-th row
//   (and at the primary row) and stop at the next start row, keeping their bytes in scratch rows (a walk that fills its
//   row continues as an overflow walker); a per-block stitch orders the walks from the primary row by pointer doubling
//   and a copy kernel assembles the output. Periodic blocks write their orbit once and replicate it.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace bra {

#define EW_TILE 4096  // elementwise tile: 256 threads x 16 consecutive elements
#define EW_THREADS 256

// ------------------------------------------------------------------------------------------------
// 1. period detection
// ------------------------------------------------------------------------------------------------
// Divisors are tried in ascending order, so the first one that survives is the smallest period. A cheap pretest on
// the first 32 bytes refutes almost every divisor of almost every block (all of them on text and random data: the full
// comparison is then never launched); a block with a surviving candidate compares itself with its shift by that
// candidate, tile by tile, and moves on to the next unrefuted divisor if it fails. While a block is unresolved period[b]
// holds 0x80000000 | index of its current candidate.
#define BWT_PERIOD_PENDING 0x80000000u
__global__ void __launch_bounds__(32) bwt_period_pretest_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len,
                                                                const uint32_t* __restrict__ div_vals, const uint32_t* __restrict__ div_off,
                                                                const uint32_t* __restrict__ div_cnt, uint8_t* __restrict__ bad, uint32_t bad_stride,
                                                                uint32_t* __restrict__ period, uint32_t* __restrict__ pending)
{
    const uint32_t b = blockIdx.x;
    const uint32_t n = len[b];
    const uint8_t* T   = in + (uint64_t) b * stride;
    const uint32_t cnt = div_cnt[b];
    const uint32_t* dv = div_vals + div_off[b];
    const uint32_t q   = min(n, 32u);
    const uint32_t l   = lane_id();
    const uint8_t  mine = l < q ? T[l] : 0;
    uint32_t       first = cnt;
    for (uint32_t di = 0; di < cnt; ++di)
    {
        const uint32_t d  = dv[di];  // proper divisor of n
        bool           ne = false;
        if (l < q)
        {
            uint32_t j = l + d;
            if (j >= n) j -= n;
            ne = mine != T[j];
        }
        if (__any_sync(BRA_FULL, ne))
        {
            if (l == 0) bad[(uint64_t) b * bad_stride + di] = 1;
        }
        else if (first == cnt)
            first = di;
    }
    if (l == 0)
    {
        if (first == cnt)
            period[b] = n;  // no divisor left: the block is primitive
        else
        {
            period[b] = BWT_PERIOD_PENDING | first;
            atomicAdd(pending, 1u);
        }
    }
}

__global__ void __launch_bounds__(256) bwt_period_full_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len,
                                                              const uint32_t* __restrict__ div_vals, const uint32_t* __restrict__ div_off,
                                                              const uint32_t* __restrict__ period, uint8_t* __restrict__ bad, uint32_t bad_stride)
{
    const uint32_t b = blockIdx.y;
    const uint32_t v = period[b];
    if (!(v & BWT_PERIOD_PENDING)) return;
    const uint32_t n = len[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= n) return;
    const uint32_t di   = v & ~BWT_PERIOD_PENDING;
    const uint32_t d    = div_vals[div_off[b] + di];
    const uint8_t* T    = in + (uint64_t) b * stride;
    const uint32_t tend = min(n, tile0 + EW_TILE);
    bool           mism = false;
    for (uint32_t i = tile0 + threadIdx.x; i < tend; i += 256)
    {
        uint32_t j = i + d;
        if (j >= n) j -= n;
        mism |= T[i] != T[j];
    }
    if (__syncthreads_or(mism) && threadIdx.x == 0) bad[(uint64_t) b * bad_stride + di] = 1;
}

__global__ void bwt_period_advance_kernel(const uint32_t* __restrict__ len, const uint32_t* __restrict__ div_vals, const uint32_t* __restrict__ div_off,
                                          const uint32_t* __restrict__ div_cnt, const uint8_t* __restrict__ bad, uint32_t bad_stride,
                                          uint32_t* __restrict__ period, uint32_t* __restrict__ pending, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint32_t v = period[b];
    if (!(v & BWT_PERIOD_PENDING)) return;
    uint32_t di = v & ~BWT_PERIOD_PENDING;
    if (!bad[(uint64_t) b * bad_stride + di])
    {
        period[b] = div_vals[div_off[b] + di];  // the candidate held everywhere: it is the smallest period
        return;
    }
    const uint32_t cnt = div_cnt[b];
    for (++di; di < cnt && bad[(uint64_t) b * bad_stride + di]; ++di) {}
    if (di == cnt)
        period[b] = len[b];
    else
    {
        period[b] = BWT_PERIOD_PENDING | di;
        atomicAdd(pending, 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// 2a. dense symbol codes: code[s] = number of smaller symbols present in the block. Packing the codes
//     instead of the bytes keeps the order of the 4-symbol keys and shortens them to 4*bits(alphabet)
//     bits: text (<= 64 symbols) sorts in 3 radix passes instead of 4, a 16-symbol alphabet in 2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    bwt_alpha_present_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint8_t* __restrict__ alpha)
{
    __shared__ uint8_t seen[256];
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= n) return;
    seen[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* T  = in + (uint64_t) b * stride;
    const uint32_t j0 = tile0 + threadIdx.x * 16;
    if (j0 + 16 <= n)
    {
        const uint4    v    = *reinterpret_cast<const uint4*>(T + j0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) seen[(w[i >> 2] >> ((i & 3) * 8)) & 0xFFu] = 1;  // racing stores of the same value
    }
    else
        for (uint32_t j = j0; j < n; ++j) seen[T[j]] = 1;
    __syncthreads();
    if (seen[threadIdx.x]) alpha[(uint64_t) b * 256 + threadIdx.x] = 1;
}

__global__ void __launch_bounds__(256) bwt_alpha_codes_kernel(uint8_t* __restrict__ alpha, uint32_t* __restrict__ max_bits)
{
    __shared__ uint32_t red[34];
    const uint32_t b = blockIdx.x;
    const uint32_t p = alpha[(uint64_t) b * 256 + threadIdx.x] ? 1u : 0u;
    uint32_t       sigma;
    const uint32_t code = block_excl_add(p, red, &sigma);
    alpha[(uint64_t) b * 256 + threadIdx.x] = (uint8_t) (p ? code : 0u);
    if (threadIdx.x == 0)
    {
        uint32_t bits = 1;
        while ((1u << bits) < sigma) ++bits;
        atomicMax(max_bits, bits);
    }
}

// ------------------------------------------------------------------------------------------------
// 2b. initial keys: first 4 symbols of every rotation of the primitive root, most significant first
//     (raw bytes, or dense codes of `cbits` bits each)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) bwt_init_keys_kernel(const uint8_t* __restrict__ in, uint64_t stride,
                                                                   const uint32_t* __restrict__ period, const uint8_t* __restrict__ alpha,
                                                                   uint32_t cbits, uint32_t* __restrict__ keys, uint32_t npass,
                                                                   uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t              sh[RS_GHIST_STRIDE];  // digit histograms of every radix pass of the sort that follows
    __shared__ uint8_t               code[256];
    __shared__ __align__(16) uint8_t s_code[EW_TILE + 16];  // symbol codes of the tile and of the three symbols behind it
    const uint32_t b = blockIdx.y;
    const uint32_t p = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS) sh[i] = 0;
    code[threadIdx.x] = alpha ? alpha[(uint64_t) b * 256 + threadIdx.x] : (uint8_t) threadIdx.x;
    __syncthreads();
    const uint8_t* T    = in + (uint64_t) b * stride;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t tend = min(p, tile0 + EW_TILE);
    const uint32_t tn   = tend - tile0;
    // stage the codes: sixteen symbols per thread (one 128-bit load where the address allows), then the three that follow
    // the tile -- cyclically, the last tile's wrap to the start of the period
    {
        const uint32_t i0 = threadIdx.x * 16;
        uint32_t       w4[4] = {0, 0, 0, 0};
        if (i0 + 16 <= tn && ((reinterpret_cast<uintptr_t>(T) + tile0 + i0) & 15u) == 0)
        {
            const uint4 v = *reinterpret_cast<const uint4*>(T + tile0 + i0);
            w4[0] = v.x; w4[1] = v.y; w4[2] = v.z; w4[3] = v.w;
        }
        else
            for (uint32_t k = 0; k < 16; ++k)
            {
                const uint32_t g = i0 + k;  // slot of the staging array: the tile, then three followers, then nothing
                if (g >= tn + 3) break;
                uint32_t j = tile0 + g;
                if (j >= p) j %= p;
                w4[k >> 2] |= (uint32_t) T[j] << ((k & 3) * 8);
            }
        uint32_t c4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            c4[q] = (uint32_t) code[w4[q] & 0xFFu] | ((uint32_t) code[(w4[q] >> 8) & 0xFFu] << 8) | ((uint32_t) code[(w4[q] >> 16) & 0xFFu] << 16) |
                    ((uint32_t) code[w4[q] >> 24] << 24);
        *reinterpret_cast<uint4*>(s_code + i0) = make_uint4(c4[0], c4[1], c4[2], c4[3]);
        if (threadIdx.x < 3 && tn + threadIdx.x >= EW_TILE)  // (followers behind slot 4095: nobody's sixteen slots reach there)
        {
            uint32_t j = tend + threadIdx.x;
            if (j >= p) j %= p;
            s_code[tn + threadIdx.x] = code[T[j]];
        }
    }
    __syncthreads();
    // keys of sixteen consecutive rotations per thread, rolled: key(j+1) = key(j) << cbits | code(j+4), cut to 4 symbols
    const uint32_t i0 = threadIdx.x * 16;
    if (i0 < tn)
    {
        const uint32_t m    = min(16u, tn - i0);
        const uint32_t mask = cbits == 8 ? 0xFFFFFFFFu : ((1u << (4 * cbits)) - 1u);
        uint32_t       key  = ((uint32_t) s_code[i0] << (2 * cbits)) | ((uint32_t) s_code[i0 + 1] << cbits) | s_code[i0 + 2];
        uint32_t       kk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            key   = ((key << cbits) | s_code[min(i0 + k + 3, tn + 2)]) & mask;
            kk[k] = key;
            if ((uint32_t) k < m)
                for (uint32_t ps = 0; ps < npass; ++ps) atomicAdd(&sh[ps * 256 + ((key >> (8 * ps)) & 0xFFu)], 1u);
        }
        uint32_t* ko = keys + base + tile0 + i0;
        if (m == 16 && (reinterpret_cast<uintptr_t>(ko) & 15u) == 0)
        {
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(ko)[q] = make_uint4(kk[4 * q], kk[4 * q + 1], kk[4 * q + 2], kk[4 * q + 3]);
        }
        else
            for (uint32_t k = 0; k < m; ++k) ko[k] = kk[k];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS)
        if (sh[i]) atomicAdd(&ghist[(uint64_t) b * RS_GHIST_STRIDE + i], sh[i]);
}

// Digit histograms of the ranks a doubling round sorts on. The keys of the round are rank[SA[j] - h] over all j, i.e.
// every rank exactly once, so the kernel that assigns the ranks counts their digits: runs of equal ranks (the members of
// one group are neighbours) are counted with one shared-memory atomic per pass.
struct RankHist
{
    uint32_t* sh;
    uint32_t  npass, val, cnt;
    __device__ __forceinline__ RankHist(uint32_t* s, uint32_t np) : sh(s), npass(np), val(0), cnt(0) {}
    __device__ __forceinline__ void flush()
    {
        if (cnt)
            for (uint32_t ps = 0; ps < npass; ++ps) atomicAdd(&sh[ps * 256 + ((val >> (8 * ps)) & 0xFFu)], cnt);
        cnt = 0;
    }
    __device__ __forceinline__ void add(uint32_t v)
    {
        if (v != val) flush();
        val = v;
        ++cnt;
    }
};

// ------------------------------------------------------------------------------------------------
// 3. group heads and ranks
// ------------------------------------------------------------------------------------------------
// Pass A: head flag per sorted slot j (1 byte), last head of each tile, group count per block.
//   MODE 0: head <=> sorted key differs from its left neighbour          (after the 4-byte sort)
//   MODE 1: doubling round. The stable re-bucketing keeps every old group in its slot range, so old
//           heads stay heads and a slot inside an old group becomes a head iff rank[SA[j]+h] differs
//           from its left neighbour's. Slots that already were singleton groups are settled: they
//           need no gather at all (that is most of the block in the late rounds).
template <int MODE>
__global__ void __launch_bounds__(EW_THREADS)
    bwt_heads_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint32_t h,
                     uint64_t stride, const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip, const uint8_t* __restrict__ flags_old,
                     uint8_t* __restrict__ flags, int* __restrict__ tile_last, uint32_t tiles, uint32_t* __restrict__ ngroups,
                     uint32_t* __restrict__ tile_heads /* optional: heads per tile */)
{
    __shared__ int      s_last[8];
    __shared__ uint32_t s_cnt[8];
    const uint32_t      b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t j0   = tile0 + threadIdx.x * 16;
    const uint32_t hm   = h % p;

    int      last = -1;
    uint32_t cnt  = 0;
    if (j0 < p)
    {
        const uint32_t m = min(16u, p - j0);
        uint32_t       fl[4] = {0, 0, 0, 0};
        if (MODE == 0)
        {
            uint32_t prevk = j0 > 0 ? skeys[base + j0 - 1] : 0u;
            for (uint32_t i = 0; i < m; ++i)
            {
                const uint32_t j = j0 + i;
                const uint32_t k = skeys[base + j];
                if (j == 0 || k != prevk)
                {
                    fl[i >> 2] |= 1u << ((i & 3) * 8);
                    last = (int) j;
                    ++cnt;
                }
                prevk = k;
            }
        }
        else
        {
            // old flags of my 16 slots + the one after (end of block counts as a head)
            uint32_t oldmask = 0;
            if (m == 16)
            {
                const uint4    v    = *reinterpret_cast<const uint4*>(flags_old + base + j0);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((w[i >> 2] >> ((i & 3) * 8)) & 0xFFu) oldmask |= 1u << i;
            }
            else
                for (uint32_t i = 0; i < m; ++i)
                    if (flags_old[base + j0 + i]) oldmask |= 1u << i;
            if (j0 + m >= p || flags_old[base + j0 + m]) oldmask |= 1u << m;
            uint32_t prevg = 0;
            if (!(oldmask & 1u))  // slot j0 continues the old group of slot j0-1: need that neighbour's second key
            {
                uint32_t x = sa[base + j0 - 1] + hm;
                if (x >= p) x -= p;
                prevg = rank[base + x];
            }
            for (uint32_t i = 0; i < m; ++i)
            {
                const uint32_t j       = j0 + i;
                const bool     oldhead = (oldmask >> i) & 1u, nexthead = (oldmask >> (i + 1)) & 1u;
                bool           head    = true;
                if (!(oldhead && nexthead))
                {
                    uint32_t x = sa[base + j] + hm;
                    if (x >= p) x -= p;
                    const uint32_t g = rank[base + x];
                    head             = oldhead || g != prevg;
                    prevg            = g;
                }
                if (head)
                {
                    fl[i >> 2] |= 1u << ((i & 3) * 8);
                    last = (int) j;
                    ++cnt;
                }
            }
        }
        if (m == 16)
            *reinterpret_cast<uint4*>(flags + base + j0) = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        else
            for (uint32_t i = 0; i < m; ++i) flags[base + j0 + i] = (fl[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        last = max(last, __shfl_xor_sync(BRA_FULL, last, d));
        cnt += __shfl_xor_sync(BRA_FULL, cnt, d);
    }
    if (lane_id() == 0)
    {
        s_last[warp_id()] = last;
        s_cnt[warp_id()]  = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int      L = -1;
        uint32_t c = 0;
        for (int i = 0; i < 8; ++i)
        {
            L = max(L, s_last[i]);
            c += s_cnt[i];
        }
        tile_last[(uint64_t) b * tiles + blockIdx.x] = L;
        if (tile_heads) tile_heads[(uint64_t) b * tiles + blockIdx.x] = c;
        atomicAdd(&ngroups[b], c);
    }
}

// Pass B': rank[SA[j]] = NUMBER of j's group (0, 1, 2, ... in sorted order) for every slot. Used for the first
// doubling round only: the groups of the 4-symbol sort are few, so their numbers need fewer key bits -- and
// radix passes -- than their positions. The order of the keys, which is all the round looks at, is the same.
__global__ void __launch_bounds__(EW_THREADS)
    bwt_dense_ranks_kernel(const uint32_t* __restrict__ sa, const uint8_t* __restrict__ flags, uint64_t stride, const uint32_t* __restrict__ period,
                           const uint8_t* __restrict__ skip, const uint32_t* __restrict__ tile_heads, uint32_t tiles, uint32_t* __restrict__ rank_out,
                           const uint32_t* __restrict__ ngroups, uint32_t npass, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t red[34];
    __shared__ uint32_t sh[RS_GHIST_STRIDE];
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p = period[b];
    if (ngroups[b] >= p) return;  // finished: nobody will read its ranks
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS) sh[i] = 0;
    const uint64_t base = (uint64_t) b * stride;
    uint32_t       before = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += EW_THREADS) before += tile_heads[(uint64_t) b * tiles + t];
    uint32_t carry;
    block_excl_add(before, red, &carry);
    const uint32_t j0 = tile0 + threadIdx.x * 16;
    const uint32_t m  = j0 < p ? min(16u, p - j0) : 0u;
    uint32_t       mask = 0;
    for (uint32_t i = 0; i < m; ++i)
        if (flags[base + j0 + i]) mask |= 1u << i;
    uint32_t run = carry + block_excl_add((uint32_t) __popc(mask), red, nullptr);  // heads before my first slot
    RankHist rh(sh, npass);
    for (uint32_t i = 0; i < m; ++i)
    {
        run += (mask >> i) & 1u;
        rank_out[base + sa[base + j0 + i]] = run - 1u;  // slot 0 is a head, so run >= 1 here
        rh.add(run - 1u);
    }
    rh.flush();
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS)
        if (sh[i]) atomicAdd(&ghist[(uint64_t) b * RS_GHIST_STRIDE + i], sh[i]);
}

// Pass B: rank[SA[j]] = index of the head of j's group, in place. Slots that were singleton groups
// before this round (flags_old) keep their rank and are not touched.
// WHAT 1: group statistics only (the ranks are written later, and only if the block goes on doubling -- a block the
// finisher completes never reads them); WHAT 2: ranks only.
template <int WHAT>
__global__ void __launch_bounds__(EW_THREADS)
    bwt_ranks_kernel(const uint32_t* __restrict__ sa, const uint8_t* __restrict__ flags, const uint8_t* __restrict__ flags_old, uint64_t stride,
                     const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip, const int* __restrict__ tile_last, uint32_t tiles,
                     uint32_t* __restrict__ rank_out, uint32_t* __restrict__ maxgroup, unsigned long long* __restrict__ sumsq,
                     const uint32_t* __restrict__ ngroups, uint32_t npass, uint32_t* __restrict__ ghist /* WHAT 2: digit histograms of the ranks */)
{
    __shared__ int      red[33];
    __shared__ uint32_t sh[WHAT == 2 ? RS_GHIST_STRIDE : 1];
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    if (ngroups[b] >= p) return;  // every group is a singleton: the block is finished, nobody will read its ranks
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    if (WHAT == 2)
        for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS) sh[i] = 0;

    // carry-in: last head before this tile
    int carry = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += EW_THREADS) carry = max(carry, tile_last[(uint64_t) b * tiles + t]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) carry = max(carry, __shfl_xor_sync(BRA_FULL, carry, d));
    if (lane_id() == 0) red[warp_id()] = carry;
    __syncthreads();
    carry = red[0];
    for (int i = 1; i < 8; ++i) carry = max(carry, red[i]);
    __syncthreads();

    const uint32_t j0 = tile0 + threadIdx.x * 16;
    const uint32_t m  = j0 < p ? min(16u, p - j0) : 0u;
    uint32_t       newmask = 0, oldmask = 0;
    int            mylast = -1;
    for (uint32_t i = 0; i < m; ++i)
    {
        if (flags[base + j0 + i])
        {
            newmask |= 1u << i;
            mylast = (int) (j0 + i);
        }
        if (flags_old && flags_old[base + j0 + i]) oldmask |= 1u << i;
    }
    if (flags_old && m && (j0 + m >= p || flags_old[base + j0 + m])) oldmask |= 1u << m;
    int run = max(block_excl_max(mylast, -1, red), carry);
    // group statistics for the finisher decision: every head closes the group before it
    uint32_t           mg = 0;
    unsigned long long sq = 0;
    RankHist           rh(sh, WHAT == 2 ? npass : 0u);
    for (uint32_t i = 0; i < m; ++i)
    {
        if (newmask & (1u << i))
        {
            const uint32_t j = j0 + i;
            if (j > 0)
            {
                const uint32_t g = j - (uint32_t) run;
                mg               = max(mg, g);
                if (g > 1) sq += (unsigned long long) g * g;
            }
            run = (int) j;
        }
        const bool settled = ((oldmask >> i) & 3u) == 3u;  // was a singleton group already: rank unchanged
        if (WHAT != 1 && !settled) rank_out[base + sa[base + j0 + i]] = (uint32_t) run;
        if (WHAT == 2) rh.add((uint32_t) run);
    }
    if (WHAT == 2)
    {
        rh.flush();
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < npass * 256; i += EW_THREADS)
            if (sh[i]) atomicAdd(&ghist[(uint64_t) b * RS_GHIST_STRIDE + i], sh[i]);
        return;
    }
    if (m && j0 + m == p)  // the last group of the block ends at p
    {
        const uint32_t g = p - (uint32_t) run;
        mg               = max(mg, g);
        if (g > 1) sq += (unsigned long long) g * g;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        mg = max(mg, __shfl_xor_sync(BRA_FULL, mg, d));
        sq += __shfl_xor_sync(BRA_FULL, sq, d);
    }
    if (lane_id() == 0 && mg)
    {
        atomicMax(&maxgroup[b], mg);
        if (sq) atomicAdd(&sumsq[b], sq);
    }
}

// Block state machine kept in done[b]: 0 = still sorting, 1 = order complete this round (its last
// column is gathered right away from the buffer that currently holds its suffix array, because
// finished blocks are skipped by later rounds and the ping-pong buffers move on), 2 = finished.
// fin[b]: 0 = finisher not tried, 1 = selected for the finisher now, 2 = tried.
// stat[0] counts the blocks still sorting, stat[1] those selected for the finisher.
#define FIN_MAX_GROUP 512u   // longest group a finisher thread will scan
#define FIN_WORK_PER_ELEM 16 // finisher is used when sum(g^2) <= this * p (bounded extra work)
#define FIN_DEPTH 64u        // bytes compared beyond the h already known equal

__global__ void bwt_check_done_kernel(const uint32_t* __restrict__ period, const uint32_t* __restrict__ ngroups, uint8_t* __restrict__ done,
                                      uint8_t* __restrict__ fin, uint8_t* __restrict__ finskip, const uint32_t* __restrict__ maxgroup,
                                      const unsigned long long* __restrict__ sumsq, uint32_t* __restrict__ stat, uint32_t nblk, bool allow_finisher)
{
    // stat[0]: blocks still sorting, stat[1]: of those, selected for the finisher, stat[2]: most groups in a block still sorting
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t d = done[b];
    uint8_t       f = fin[b];
    if (f == 1) f = 2;
    if (d == 1)
        done[b] = 2;
    else if (d == 0)
    {
        if (ngroups[b] >= period[b])
            done[b] = 1;
        else
        {
            atomicAdd(&stat[0], 1u);
            atomicMax(&stat[2], ngroups[b]);
            if (allow_finisher && f == 0 && maxgroup[b] <= FIN_MAX_GROUP && sumsq[b] <= (unsigned long long) FIN_WORK_PER_ELEM * period[b])
            {
                f = 1;
                atomicAdd(&stat[1], 1u);
            }
        }
    }
    fin[b]     = f;
    finskip[b] = !(done[b] == 0 && f == 1);
}

__global__ void bwt_reset_stats_kernel(const uint8_t* __restrict__ skip, uint32_t* __restrict__ ngroups, uint32_t* __restrict__ maxgroup,
                                       unsigned long long* __restrict__ sumsq, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk || skip[b]) return;
    ngroups[b]  = 0;
    maxgroup[b] = 0;
    sumsq[b]    = 0;
}

__global__ void bwt_reset_tile_last_kernel(const uint8_t* __restrict__ skip, int* __restrict__ tile_last, uint32_t tiles)
{
    const uint32_t b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (!skip[b] && t < tiles) tile_last[(uint64_t) b * tiles + t] = -1;
}

// Finisher: when the remaining groups are small, order each of them by comparing the rotations
// directly (bytes h .. h+FIN_DEPTH-1; everything before is known equal). Every member counts the
// members that sort before it, so the kernel is fully parallel. Members still equal after FIN_DEPTH
// bytes stay one group (in their current order) and go on with prefix doubling.
// The comparison itself is bra_rot_cmp_window (bra_hd.h): shared with the host so that the CPU suite can check it,
// wrap-around cases included, against a plain byte comparison.
__device__ __forceinline__ int rot_cmp_window(const uint8_t* __restrict__ T, uint32_t p, uint32_t a, uint32_t c, uint32_t from)
{
    return bra_rot_cmp_window(T, p, a, c, from, FIN_DEPTH);
}

// Slots that are singleton groups already (most of them) are copied through; the members of the remaining
// groups are first compacted into a shared-memory list, so that the comparison loops run on full warps
// instead of on the few lanes of each warp that happen to sit in a group. (Spreading the comparisons
// evenly as (member, other member) work items was measured too: the bookkeeping costs what the balance gains.)
__global__ void __launch_bounds__(EW_THREADS)
    bwt_finish_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip,
                      uint32_t h, const uint32_t* __restrict__ sa, const uint8_t* __restrict__ flags_old, uint32_t* __restrict__ sa_out,
                      uint8_t* __restrict__ flags_new, int* __restrict__ tile_last, uint32_t tiles, uint32_t* __restrict__ ngroups)
{
    __shared__ uint16_t s_list[EW_TILE];
    __shared__ uint32_t s_n;
    __shared__ int      s_best[8];
    __shared__ uint32_t s_heads[8];
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint8_t* T    = in + base;
    const uint8_t* fo   = flags_old + base;
    const uint32_t tend = min(p, tile0 + EW_TILE);
    const uint32_t hm   = h % p;
    uint32_t       heads = 0;
    int            best  = -1;  // last head that lands in this tile
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (uint32_t j = tile0 + threadIdx.x; j < tend; j += EW_THREADS)
    {
        if (fo[j] && (j + 1 == p || fo[j + 1]))
        {
            sa_out[base + j]    = sa[base + j];
            flags_new[base + j] = 1;
            ++heads;
            best = (int) j;  // j grows along the loop
        }
        else
            s_list[atomicAdd(&s_n, 1u)] = (uint16_t) (j - tile0);
    }
    __syncthreads();
    const uint32_t nlist = s_n;
    for (uint32_t x = threadIdx.x; x < nlist; x += EW_THREADS)
    {
        const uint32_t j  = tile0 + s_list[x];
        const uint32_t me = sa[base + j];
        uint32_t       s = j, e = j + 1;
        while (!fo[s]) --s;  // slot 0 is always a head
        while (e < p && !fo[e]) ++e;
        uint32_t less = 0, tie_before = 0;
        for (uint32_t m = s; m < e; ++m)
        {
            if (m == j) continue;
            const int c = rot_cmp_window(T, p, me, sa[base + m], hm);
            if (c > 0)
                ++less;
            else if (c == 0 && m < j)
            {
                ++less;
                ++tie_before;
            }
        }
        const uint32_t pos    = s + less;
        const bool     head   = tie_before == 0;
        sa_out[base + pos]    = me;
        flags_new[base + pos] = head ? 1 : 0;
        if (head)
        {
            ++heads;
            if (pos >= tile0 && pos < tend)
                best = max(best, (int) pos);
            else
                atomicMax(&tile_last[(uint64_t) b * tiles + pos / EW_TILE], (int) pos);  // member moved across a tile border (rare)
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        heads += __shfl_xor_sync(BRA_FULL, heads, d);
        best = max(best, __shfl_xor_sync(BRA_FULL, best, d));
    }
    if (lane_id() == 0)
    {
        s_best[warp_id()]  = best;
        s_heads[warp_id()] = heads;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; ++i)
        {
            best = max(best, s_best[i]);
            heads += s_heads[i];
        }
        if (best >= 0) atomicMax(&tile_last[(uint64_t) b * tiles + blockIdx.x], best);
        if (heads) atomicAdd(&ngroups[b], heads);
    }
}

// after the finisher: bring the selected blocks' order and flags back into the current buffers
__global__ void __launch_bounds__(EW_THREADS)
    bwt_copyback_kernel(uint64_t stride, const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip, const uint32_t* __restrict__ sa_src,
                        uint32_t* __restrict__ sa_dst, const uint8_t* __restrict__ fl_src, uint8_t* __restrict__ fl_dst)
{
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t tend = min(p, tile0 + EW_TILE);
    for (uint32_t j = tile0 + threadIdx.x; j < tend; j += EW_THREADS)
    {
        sa_dst[base + j] = sa_src[base + j];
        fl_dst[base + j] = fl_src[base + j];
    }
}

// doubling round, step 1: (key, value) = (rank[SA[j] - h], SA[j] - h) in current-order traversal. (Fusing this gather into
// the first sort pass was measured: the sort kernel runs at half occupancy and exposes the gather latency -- 3.9 ms
// against 1.35 + 1.1 ms per 2^28 elements for the two kernels.)
__global__ void __launch_bounds__(EW_THREADS)
    bwt_dbl_prepare_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint32_t h, uint64_t stride,
                           const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip, uint32_t* __restrict__ keys,
                           uint32_t* __restrict__ vals)
{
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t tend = min(p, tile0 + EW_TILE);
    const uint32_t hm   = h % p;
    for (uint32_t j = tile0 + threadIdx.x; j < tend; j += EW_THREADS)
    {
        const uint32_t s = sa[base + j];
        const uint32_t v = s >= hm ? s - hm : s + p - hm;
        vals[base + j]   = v;
        keys[base + j]   = rank[base + v];
    }
}

// ---- packed rounds (blocks of at most 2^21 bytes; up to 2^24 bytes while the group numbers fit 20 bits) -----------------------------------------------------------------------
// A doubling round orders element v = SA[j] - h by (group of v, group of v + h). The second component is the group of
// traversal slot j itself, i.e. the position of the last head at or before j -- known while the round's records are
// written, no gather needed. With 21 bits per field the triple (key, second key, v) fits the 8 bytes per element the sort
// moves anyway: hi = key << 11 | second >> 10, lo = (second & 1023) << 22 | v. After the sort, new heads are where
// (key, second) differs from the left neighbour: a streaming pass instead of one more random gather per element.
#define BWT_PACK_MAX_N (1u << 21)

// Both kernels are warp-striped: warp w owns slots [512 w, 512 w + 512) of the tile and visits them in 16 rounds of 32
// consecutive slots, so every load and store is one contiguous 128-byte (or 32-byte) piece per warp, and "last head at
// or before slot j" is a ballot and a count-leading-zeros away.
// Two layouts. WIDE == false (blocks up to 2 MiB): rotation in 22 bits, second key = position of the slot's group head
// (21 bits), key up to 21 bits. WIDE == true (blocks up to 16 MiB, rounds on dense group numbers of at most 20 bits):
// rotation in 24 bits, second key = NUMBER of the slot's group (20 bits), key up to 20 bits.
template <bool WIDE>
struct BwtPack
{
    static constexpr uint32_t VBITS = WIDE ? 24u : 22u;          // rotation index
    static constexpr uint32_t RLO   = 32u - VBITS;               // low bits of the second key that share the word with it
    static constexpr uint32_t KS    = (WIDE ? 20u : 21u) - RLO;  // the key starts at this bit of the high word
};
template <bool WIDE>
__global__ void __launch_bounds__(EW_THREADS, 4)
    bwt_dbl_prepare_packed_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, const uint8_t* __restrict__ flags, uint32_t h,
                                  uint64_t stride, const uint32_t* __restrict__ period, const uint8_t* __restrict__ skip,
                                  const int* __restrict__ tile_last, const uint32_t* __restrict__ tile_heads, uint32_t tiles,
                                  uint32_t* __restrict__ hi_out, uint32_t* __restrict__ lo_out)
{
    using L = BwtPack<WIDE>;
    __shared__ int red[8], s_wl[8];
    const uint32_t b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t hm   = h % p;
    const uint32_t w = warp_id(), l = lane_id();
    const uint32_t seg0 = tile0 + w * 512;
    // WIDE: heads in the warp's slots (their count gives group numbers); else: the last head of the warp's slots
    int wl = WIDE ? 0 : -1;
#pragma unroll 4
    for (int r = 0; r < 16; ++r)
    {
        const uint32_t j  = seg0 + r * 32 + l;
        const uint32_t hb = __ballot_sync(BRA_FULL, j < p && flags[base + j] != 0);
        if (WIDE)
            wl += __popc(hb);
        else if (hb)
            wl = (int) (seg0 + r * 32 + (31 - __clz(hb)));
    }
    // carry-in from the tiles before this one (head count / last head), then from the warps before me
    int carry = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += EW_THREADS)
    {
        if (WIDE)
            carry += (int) tile_heads[(uint64_t) b * tiles + t];
        else
            carry = max(carry, tile_last[(uint64_t) b * tiles + t]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const int o = __shfl_xor_sync(BRA_FULL, carry, d);
        carry       = WIDE ? carry + o : max(carry, o);
    }
    if (l == 0)
    {
        red[w]  = carry;
        s_wl[w] = wl;
    }
    __syncthreads();
    int run = red[0];  // WIDE: heads before the current round; else: last head at or before it
    for (int i = 1; i < 8; ++i) run = WIDE ? run + red[i] : max(run, red[i]);
    for (uint32_t i = 0; i < w; ++i) run = WIDE ? run + s_wl[i] : max(run, s_wl[i]);
    // two halves of eight rounds: eight independent gathers per lane in flight, registers for four CTAs per SM
#pragma unroll 1
    for (int half = 0; half < 2; ++half)
    {
        uint32_t v[8], k[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const uint32_t j = seg0 + (half * 8 + r) * 32 + l;
            const uint32_t s = j < p ? sa[base + j] : hm;
            v[r]             = s >= hm ? s - hm : s + p - hm;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) k[r] = seg0 + (half * 8 + r) * 32 + l < p ? rank[base + v[r]] : 0u;
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const uint32_t j0r  = seg0 + (half * 8 + r) * 32;
            const uint32_t j    = j0r + l;
            const uint32_t hb   = __ballot_sync(BRA_FULL, j < p && flags[base + j] != 0);
            const uint32_t mine = hb & (lanemask_lt() | (1u << l));
            uint32_t       r2;
            if (WIDE)
            {
                r2 = (uint32_t) run + __popc(mine) - 1u;  // slot 0 is a head: at least one head at or before j
                run += __popc(hb);
            }
            else
            {
                r2 = mine ? j0r + (31 - __clz(mine)) : (uint32_t) run;
                if (hb) run = (int) (j0r + (31 - __clz(hb)));
            }
            if (j < p)
            {
                hi_out[base + j] = (k[r] << L::KS) | (r2 >> L::RLO);
                lo_out[base + j] = ((r2 & ((1u << L::RLO) - 1u)) << L::VBITS) | v[r];
            }
        }
    }
}

// Heads of a freshly sorted order plus the group statistics the finisher decision looks at, warp-striped.
//   KIND 0: initial sort -- head <=> the sorted key differs from its left neighbour
//   KIND 1: packed round -- head <=> (key, second key) differs; also unpacks the rotation indices into sa_out
// Statistics: every group piece between two consecutive heads counts with its length; pieces are cut at tile borders
// only (a group spanning a border counts as two pieces -- the decision it feeds is a cost heuristic, not a correctness
// condition, and a tile without any head still reports a piece of 4096).
template <int KIND>
__global__ void __launch_bounds__(EW_THREADS, 4)
    bwt_heads_stats_kernel(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, uint64_t stride, const uint32_t* __restrict__ period,
                           const uint8_t* __restrict__ skip, uint8_t* __restrict__ flags, uint32_t* __restrict__ sa_out, int* __restrict__ tile_last,
                           uint32_t tiles, uint32_t* __restrict__ ngroups, uint32_t* __restrict__ tile_heads, uint32_t* __restrict__ maxgroup,
                           unsigned long long* __restrict__ sumsq, uint32_t vbits /* KIND 1: width of the rotation field */)
{
    __shared__ int                s_first[8], s_last[8];
    __shared__ uint32_t           s_cnt[8], s_mg[8];
    __shared__ unsigned long long s_sq[8];
    const uint32_t                b = blockIdx.y;
    if (skip[b]) return;
    const uint32_t p     = period[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= p) return;
    const uint64_t base = (uint64_t) b * stride;
    const uint32_t w = warp_id(), l = lane_id();
    const uint32_t seg0 = tile0 + w * 512;
    int                first = -1, last = -1;
    uint32_t           cnt = 0, mg = 0;
    unsigned long long sq = 0;
    uint32_t           xh[16], xl[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)
    {
        const uint32_t j = seg0 + r * 32 + l;
        xh[r]            = j < p ? hi[base + j] : 0u;
        xl[r]            = (KIND == 1 && j < p) ? lo[base + j] : 0u;
    }
    // the element left of the warp's first slot
    uint32_t ph = 0, pl = 0;
    if (seg0 > 0 && seg0 < p)
    {
        ph = hi[base + seg0 - 1];
        if (KIND == 1) pl = lo[base + seg0 - 1] >> vbits;
    }
#pragma unroll
    for (int r = 0; r < 16; ++r)
    {
        const uint32_t j0r = seg0 + r * 32;
        const uint32_t j   = j0r + l;
        const uint32_t s2  = xl[r] >> vbits;
        uint32_t       lh = __shfl_up_sync(BRA_FULL, xh[r], 1), ll = __shfl_up_sync(BRA_FULL, s2, 1);
        if (l == 0)
        {
            lh = ph;
            ll = pl;
        }
        const bool     head = j < p && (j == 0 || xh[r] != lh || (KIND == 1 && s2 != ll));
        const uint32_t hb   = __ballot_sync(BRA_FULL, head);
        if (head)
        {
            // the piece this head closes starts at the previous head of the warp's slots (pieces that start before them
            // are closed below, once the warps know of each other)
            const uint32_t below = hb & lanemask_lt();
            const int      prev  = below ? (int) (j0r + (31 - __clz(below))) : last;
            if (prev >= 0)
            {
                const uint32_t g = j - (uint32_t) prev;
                mg               = max(mg, g);
                if (g > 1) sq += (unsigned long long) g * g;
            }
        }
        if (hb)
        {
            if (first < 0) first = (int) (j0r + (__ffs(hb) - 1));
            last = (int) (j0r + (31 - __clz(hb)));
        }
        cnt += __popc(hb);
        if (j < p)
        {
            flags[base + j] = head ? 1 : 0;
            if (KIND == 1) sa_out[base + j] = xl[r] & ((1u << vbits) - 1u);
        }
        ph = __shfl_sync(BRA_FULL, xh[r], 31);
        pl = __shfl_sync(BRA_FULL, s2, 31);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        mg = max(mg, __shfl_xor_sync(BRA_FULL, mg, d));
        sq += __shfl_xor_sync(BRA_FULL, sq, d);
    }
    if (l == 0)
    {
        s_first[w] = first;
        s_last[w]  = last;
        s_cnt[w]   = cnt;
        s_mg[w]    = mg;
        s_sq[w]    = sq;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int                L = -1;
        uint32_t           c = 0, m = 0;
        unsigned long long q = 0;
        int                open = (int) tile0;  // start of the piece still open: the tile border, then the last head seen
        for (int i = 0; i < 8; ++i)
        {
            c += s_cnt[i];
            m = max(m, s_mg[i]);
            q += s_sq[i];
            if (s_first[i] >= 0)
            {
                const uint32_t g = (uint32_t) (s_first[i] - open);  // 0 when the piece starts right at a head
                m                = max(m, g);
                if (g > 1) q += (unsigned long long) g * g;
                open = s_last[i];
                L    = s_last[i];
            }
        }
        const uint32_t g = min(p, tile0 + EW_TILE) - (uint32_t) open;  // the piece that runs into the next tile, or ends the block
        m                = max(m, g);
        if (g > 1) q += (unsigned long long) g * g;
        tile_last[(uint64_t) b * tiles + blockIdx.x] = L;
        if (tile_heads) tile_heads[(uint64_t) b * tiles + blockIdx.x] = c;
        atomicAdd(&ngroups[b], c);
        atomicMax(&maxgroup[b], m);
        if (q) atomicAdd(&sumsq[b], q);
    }
}

// 4. last column + primary index
__global__ void __launch_bounds__(EW_THREADS)
    bwt_gather_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ sa, uint64_t stride, const uint32_t* __restrict__ len,
                      const uint32_t* __restrict__ period, const uint8_t* __restrict__ done, uint8_t* __restrict__ out,
                      uint32_t* __restrict__ primary)
{
    const uint32_t b = blockIdx.y;
    if (done[b] != 1) return;
    const uint32_t n = len[b];
    const uint32_t tile0 = blockIdx.x * EW_TILE;
    if (tile0 >= n) return;
    const uint32_t p    = period[b];
    const uint32_t k    = n / p;
    const uint64_t base = (uint64_t) b * stride;
    const uint8_t* T    = in + base;
    const uint32_t tend = min(n, tile0 + EW_TILE);
    for (uint32_t j = tile0 + threadIdx.x; j < tend; j += EW_THREADS)
    {
        const uint32_t q = (k == 1) ? j : j / k;
        const uint32_t s = sa[base + q];
        out[base + j]    = T[s == 0 ? p - 1 : s - 1];
        if (s == 0 && q * k == j) primary[b] = j;  // identical rotations are in ascending index: rotation 0 is the first
    }
}

static void host_divisors(uint32_t n, std::vector<uint32_t>& out)
{
    std::vector<uint32_t> hi;
    for (uint32_t d = 1; (uint64_t) d * d <= n; ++d)
        if (n % d == 0)
        {
            if (d < n) out.push_back(d);
            const uint32_t e = n / d;
            if (e != d && e < n) hi.push_back(e);
        }
    for (size_t i = hi.size(); i-- > 0;) out.push_back(hi[i]);
}

bool bwt_forward_batch(const BwtFwdArgs& a, cudaStream_t st)
{
    const uint32_t nblk = a.nblk, max_n = a.max_n;
    if (nblk == 0) return true;
    const uint32_t tiles = bra_div_up(max_n, EW_TILE);
    const dim3     grid(tiles, nblk);

    // ---- period detection (divisor tables built on the host from the known block lengths)
    {
        std::vector<uint32_t> vals, off(nblk), cnt(nblk);
        std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> seen;  // n -> (off, cnt)
        uint32_t maxcnt = 1;
        for (uint32_t b = 0; b < nblk; ++b)
        {
            const uint32_t n = a.h_len[b];
            bool           found = false;
            for (auto& s : seen)
                if (s.first == n)
                {
                    off[b] = s.second.first;
                    cnt[b] = s.second.second;
                    found  = true;
                    break;
                }
            if (!found)
            {
                const uint32_t o = (uint32_t) vals.size();
                host_divisors(n, vals);
                off[b] = o;
                cnt[b] = (uint32_t) vals.size() - o;
                seen.push_back({n, {o, cnt[b]}});
            }
            maxcnt = std::max(maxcnt, cnt[b]);
        }
        if (vals.empty()) vals.push_back(1);
        if (vals.size() > a.div_cap || maxcnt > a.bad_stride)
        {
            bra_b200_log_error("bwt: divisor table overflow (%zu values, %u per block)", vals.size(), maxcnt);
            return false;
        }
        if (a.h_mail)
        {
            uint32_t* m = a.h_mail + 2;
            memcpy(m, vals.data(), vals.size() * 4);
            memcpy(m + a.div_cap, off.data(), (size_t) nblk * 4);
            memcpy(m + a.div_cap + nblk, cnt.data(), (size_t) nblk * 4);
            if (!mail_fetch(a.d_div_vals, m, (uint32_t) vals.size(), st) || !mail_fetch(a.d_div_off, m + a.div_cap, nblk, st) ||
                !mail_fetch(a.d_div_cnt, m + a.div_cap + nblk, nblk, st))
                return false;
            BRA_CUDA_TRY(cudaMemsetAsync(a.d_bad, 0, (size_t) nblk * a.bad_stride, st));
            // the mail words are rewritten by the next batch only, and the loop below synchronises long before that
        }
        else
        {
            BRA_CUDA_TRY(cudaMemcpyAsync(a.d_div_vals, vals.data(), vals.size() * 4, cudaMemcpyHostToDevice, st));
            BRA_CUDA_TRY(cudaMemcpyAsync(a.d_div_off, off.data(), nblk * 4, cudaMemcpyHostToDevice, st));
            BRA_CUDA_TRY(cudaMemcpyAsync(a.d_div_cnt, cnt.data(), nblk * 4, cudaMemcpyHostToDevice, st));
            BRA_CUDA_TRY(cudaMemsetAsync(a.d_bad, 0, (size_t) nblk * a.bad_stride, st));
            // the pageable-host staging vectors die at scope exit: the copies above must have landed
            BRA_CUDA_TRY(cudaStreamSynchronize(st));
        }
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_notdone, 0, 16, st));  // [0] alphabet bits (below), [1] blocks with a period candidate left
        BRA_LAUNCH(P_BWT_PERIOD, st, bwt_period_pretest_kernel<<<nblk, 32, 0, st>>>(a.d_in, a.stride, a.d_len, a.d_div_vals, a.d_div_off, a.d_div_cnt, a.d_bad, a.bad_stride,
                                                                                 a.d_period, a.d_notdone + 1));
    }

    BRA_CUDA_TRY(cudaMemsetAsync(a.d_done, 0, nblk, st));
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_ngroups, 0, nblk * 4, st));

    // ---- radix sort on the first 4 symbols
    uint32_t *kA = a.d_keyA, *kB = a.d_keyB, *vA = a.d_valA, *vB = a.d_valB;
    // parity triage switches (tools/fuzz_probe.py): each turns one optional optimisation off. Read once per process.
    static const uint32_t diag = (getenv("BRA_B200_NO_ALPHA") ? 1u : 0u) | (getenv("BRA_B200_NO_DENSE") ? 2u : 0u) | (getenv("BRA_B200_NO_FINISH") ? 4u : 0u);
    const bool            no_alpha = diag & 1u, no_dense = diag & 2u, no_finish = diag & 4u;
    uint32_t   cbits = 8;
    if (a.d_alpha && !no_alpha)
    {
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_alpha, 0, (size_t) nblk * 256, st));
        BRA_LAUNCH(P_BWT_KEYS, st, bwt_alpha_present_kernel<<<grid, 256, 0, st>>>(a.d_in, a.stride, a.d_len, a.d_alpha));
        BRA_LAUNCH(P_BWT_KEYS, st, bwt_alpha_codes_kernel<<<nblk, 256, 0, st>>>(a.d_alpha, a.d_notdone));
    }
    // one round trip for the alphabet width and the number of blocks whose period is still open; blocks with a surviving
    // candidate divisor (periodic inputs only) then run full comparisons, one candidate per round
    for (uint32_t round = 0;; ++round)
    {
        uint32_t st2[2] = {0, 0};
        if (a.h_mail)
        {
            if (!mail_publish(a.h_mail, a.d_notdone, 2, st)) return false;
            BRA_CUDA_TRY(cudaStreamSynchronize(st));
            st2[0] = reinterpret_cast<volatile uint32_t*>(a.h_mail)[0];
            st2[1] = reinterpret_cast<volatile uint32_t*>(a.h_mail)[1];
        }
        else
        {
            BRA_CUDA_TRY(cudaMemcpyAsync(st2, a.d_notdone, 8, cudaMemcpyDeviceToHost, st));
            BRA_CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (round == 0 && a.d_alpha && !no_alpha) cbits = st2[0];
        if (st2[1] == 0) break;
        if (round > a.bad_stride)
        {
            bra_b200_log_error("bwt: period detection did not settle");
            return false;
        }
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_notdone + 1, 0, 4, st));
        BRA_LAUNCH(P_BWT_PERIOD, st, bwt_period_full_kernel<<<grid, 256, 0, st>>>(a.d_in, a.stride, a.d_len, a.d_div_vals, a.d_div_off, a.d_period, a.d_bad, a.bad_stride));
        BRA_LAUNCH(P_BWT_PERIOD, st, bwt_period_advance_kernel<<<bra_div_up(nblk, 128), 128, 0, st>>>(a.d_len, a.d_div_vals, a.d_div_off, a.d_div_cnt, a.d_bad, a.bad_stride,
                                                                                               a.d_period, a.d_notdone + 1, nblk));
    }
    if (cbits < 1 || cbits > 8) cbits = 8;
    const size_t   ghist_bytes = (size_t) nblk * RS_GHIST_STRIDE * sizeof(uint32_t);
    const uint32_t init_passes = (4 * cbits + 7) / 8;
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_hist, 0, ghist_bytes, st));
    BRA_LAUNCH(P_BWT_KEYS, st, bwt_init_keys_kernel<<<grid, EW_THREADS, 0, st>>>(a.d_in, a.stride, a.d_period, no_alpha ? nullptr : a.d_alpha, cbits, kA, init_passes, a.d_hist));
    for (uint32_t pass = 0; pass < init_passes; ++pass)
    {
        // the first pass takes the rotation indices as implicit values
        if (!radix_pass_u32(kA, pass == 0 ? nullptr : vA, kB, vB, a.stride, a.d_period, nullptr, max_n, nblk, pass, 0, a.d_hist, st)) return false;
        std::swap(kA, kB);
        std::swap(vA, vB);
    }
    uint32_t* rk = a.d_rankA;  // ranks are updated in place
    uint8_t * fcur = a.d_flags, *fnext = a.d_flags2;
    const dim3 g1(bra_div_up(nblk, 128));
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_fin, 0, nblk, st));
    BRA_LAUNCH(P_BWT_MISC, st, bwt_reset_stats_kernel<<<g1, 128, 0, st>>>(a.d_done, a.d_ngroups, a.d_maxgroup, a.d_sumsq, nblk));
    BRA_LAUNCH(P_BWT_HEADS, st, bwt_heads_stats_kernel<0><<<grid, EW_THREADS, 0, st>>>(kA, nullptr, a.stride, a.d_period, a.d_done, fcur, nullptr, a.d_tile_last, tiles, a.d_ngroups,
                                                                                   a.d_tile_heads, a.d_maxgroup, a.d_sumsq, 0));
    // the ranks themselves are written when (and for the blocks that) a doubling round follows
    const uint8_t* ranks_old     = nullptr;

    uint32_t h = 4, rounds = 0, finishes = 0;
    uint32_t key_bits = 1;
    while ((1ull << key_bits) < max_n) ++key_bits;
    for (;;)
    {
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_notdone, 0, 16, st));
        BRA_LAUNCH(P_BWT_MISC, st, bwt_check_done_kernel<<<g1, 128, 0, st>>>(a.d_period, a.d_ngroups, a.d_done, a.d_fin, a.d_finskip, a.d_maxgroup, a.d_sumsq,
                                                                          a.d_notdone, nblk, !no_finish));
        BRA_LAUNCH(P_BWT_GATHER, st, bwt_gather_kernel<<<grid, EW_THREADS, 0, st>>>(a.d_in, vA, a.stride, a.d_len, a.d_period, a.d_done, a.d_out, a.d_primary));
        uint32_t stat[3] = {0, 0, 0};
        if (a.h_mail)
        {
            // (the mail words [0,2) are followed by the divisor table, long consumed by now: three words are free to use)
            if (!mail_publish(a.h_mail, a.d_notdone, 3, st)) return false;
            BRA_CUDA_TRY(cudaStreamSynchronize(st));
            for (int i = 0; i < 3; ++i) stat[i] = reinterpret_cast<volatile uint32_t*>(a.h_mail)[i];
        }
        else
        {
            BRA_CUDA_TRY(cudaMemcpyAsync(stat, a.d_notdone, 12, cudaMemcpyDeviceToHost, st));
            BRA_CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (stat[0] == 0) break;
        if (stat[1] != 0)
        {
            // finisher pass over the selected blocks (everything else keeps its state)
            BRA_LAUNCH(P_BWT_MISC, st, bwt_reset_stats_kernel<<<g1, 128, 0, st>>>(a.d_finskip, a.d_ngroups, a.d_maxgroup, a.d_sumsq, nblk));
            // (only the selected blocks: the others still need their tile carries for the pending rank pass)
            BRA_LAUNCH(P_BWT_MISC, st, bwt_reset_tile_last_kernel<<<dim3(bra_div_up(tiles, 256), nblk), 256, 0, st>>>(a.d_finskip, a.d_tile_last, tiles));
            BRA_LAUNCH(P_BWT_FINISH, st, bwt_finish_kernel<<<grid, EW_THREADS, 0, st>>>(a.d_in, a.stride, a.d_period, a.d_finskip, h, vA, fcur, vB, fnext,
                                                                                     a.d_tile_last, tiles, a.d_ngroups));
            // (no group statistics afterwards: they only feed the decision to try the finisher, which a block gets once)
            if (stat[1] == stat[0])
            {
                // every block still sorting went through the finisher: its outputs simply become the current buffers
                std::swap(vA, vB);
                std::swap(fcur, fnext);
            }
            else
                BRA_LAUNCH(P_BWT_FINISH, st, bwt_copyback_kernel<<<grid, EW_THREADS, 0, st>>>(a.stride, a.d_period, a.d_finskip, vB, vA, fnext, fcur));
            // A block the finisher could not complete (rotations equal beyond its depth) goes on doubling: it needs every
            // rank written, its slots were reordered. The pending rank pass therefore stops trusting the older flags.
            ranks_old     = nullptr;
            ++finishes;
            continue;
        }
        if (h >= 2u * max_n + 8u)
        {
            bra_b200_log_error("bwt: prefix doubling did not converge (h=%u, max_n=%u, %u blocks left)", h, max_n, stat[0]);
            return false;
        }
        // While the groups are few (stat[2] = the largest group count of a block still sorting), their dense numbers need
        // fewer key bits -- and radix passes -- than their positions: the first round on text-like data, and most rounds
        // of long-repeat data. (Not after a finisher pass: it moves heads without keeping the per-tile counts.)
        uint32_t dense_bits = 1;
        while ((1ull << dense_bits) < stat[2]) ++dense_bits;
        const bool dense = !no_dense && finishes == 0 && a.d_tile_heads != nullptr && (dense_bits + 7) / 8 < (key_bits + 7) / 8;
        const uint32_t round_bits = dense ? dense_bits : key_bits, round_passes = (round_bits + 7) / 8;
        // the kernel that assigns the ranks also counts their digits for every pass of the round's sort
        BRA_CUDA_TRY(cudaMemsetAsync(a.d_hist, 0, ghist_bytes, st));
        // (a warp-striped version of the two rank kernels -- coalesced loads, ballot scans -- was measured slower: 17.7 against
        // 15.5 ms per GiB of text; the scatter, not the bookkeeping, is what they wait for)
        if (dense)
            BRA_LAUNCH(P_BWT_RANKS, st, bwt_dense_ranks_kernel<<<grid, EW_THREADS, 0, st>>>(vA, fcur, a.stride, a.d_period, a.d_done, a.d_tile_heads, tiles, rk, a.d_ngroups,
                                                                                        round_passes, a.d_hist));
        else
            // (every slot is visited for the histogram; ranks that are known to be in place already are not rewritten)
            BRA_LAUNCH(P_BWT_RANKS, st, bwt_ranks_kernel<2><<<grid, EW_THREADS, 0, st>>>(vA, fcur, ranks_old, a.stride, a.d_period, a.d_done,
                                                                                     a.d_tile_last, tiles, rk, a.d_maxgroup, a.d_sumsq, a.d_ngroups, round_passes, a.d_hist));
        BRA_LAUNCH(P_BWT_MISC, st, bwt_reset_stats_kernel<<<g1, 128, 0, st>>>(a.d_done, a.d_ngroups, a.d_maxgroup, a.d_sumsq, nblk));
        const bool pack_narrow = max_n <= BWT_PACK_MAX_N;
        const bool pack_wide   = !pack_narrow && dense && dense_bits <= 20 && max_n <= (1u << 24);
        if (pack_narrow || pack_wide)
        {
            // packed records: the second sort key travels with the element, the new heads need no gather
            const uint32_t ks = pack_wide ? BwtPack<true>::KS : BwtPack<false>::KS, vbits = pack_wide ? BwtPack<true>::VBITS : BwtPack<false>::VBITS;
            if (pack_wide)
                BRA_LAUNCH(P_BWT_PREPARE, st, bwt_dbl_prepare_packed_kernel<true><<<grid, EW_THREADS, 0, st>>>(vA, rk, fcur, h, a.stride, a.d_period, a.d_done, a.d_tile_last,
                                                                                                       a.d_tile_heads, tiles, kB, vB));
            else
                BRA_LAUNCH(P_BWT_PREPARE, st, bwt_dbl_prepare_packed_kernel<false><<<grid, EW_THREADS, 0, st>>>(vA, rk, fcur, h, a.stride, a.d_period, a.d_done, a.d_tile_last,
                                                                                                        a.d_tile_heads, tiles, kB, vB));
            std::swap(kA, kB);
            std::swap(vA, vB);
            for (uint32_t pass = 0; pass < round_passes; ++pass)
            {
                if (!radix_pass_u32(kA, vA, kB, vB, a.stride, a.d_period, a.d_done, max_n, nblk, pass, ks, a.d_hist, st)) return false;
                std::swap(kA, kB);
                std::swap(vA, vB);
            }
            BRA_LAUNCH(P_BWT_HEADS, st, bwt_heads_stats_kernel<1><<<grid, EW_THREADS, 0, st>>>(kA, vA, a.stride, a.d_period, a.d_done, fnext, vB, a.d_tile_last, tiles, a.d_ngroups,
                                                                                           a.d_tile_heads, a.d_maxgroup, a.d_sumsq, vbits));
            std::swap(vA, vB);  // the plain rotation indices
        }
        else
        {
            BRA_LAUNCH(P_BWT_PREPARE, st, bwt_dbl_prepare_kernel<<<grid, EW_THREADS, 0, st>>>(vA, rk, h, a.stride, a.d_period, a.d_done, kB, vB));
            std::swap(kA, kB);
            std::swap(vA, vB);
            for (uint32_t pass = 0; pass < round_passes; ++pass)
            {
                if (!radix_pass_u32(kA, vA, kB, vB, a.stride, a.d_period, a.d_done, max_n, nblk, pass, 0, a.d_hist, st)) return false;
                std::swap(kA, kB);
                std::swap(vA, vB);
            }
            BRA_LAUNCH(P_BWT_HEADS, st, bwt_heads_kernel<1><<<grid, EW_THREADS, 0, st>>>(nullptr, vA, rk, h, a.stride, a.d_period, a.d_done, fcur, fnext, a.d_tile_last, tiles,
                                                              a.d_ngroups, a.d_tile_heads));
        }
        if (!(pack_narrow || pack_wide))  // (the packed path's head kernel has produced the group statistics already)
            BRA_LAUNCH(P_BWT_RANKS, st, bwt_ranks_kernel<1><<<grid, EW_THREADS, 0, st>>>(vA, fnext, fcur, a.stride, a.d_period, a.d_done, a.d_tile_last, tiles, rk,
                                                                                     a.d_maxgroup, a.d_sumsq, a.d_ngroups, 0, nullptr));
        std::swap(fcur, fnext);
        ranks_old     = dense ? nullptr : fnext;  // the flags of before this round; after a round on group numbers every rank is rewritten
        h *= 2;
        ++rounds;
    }
    (void) finishes;
    if (a.h_rounds) *a.h_rounds = rounds;
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

// ================================================================================================
// INVERSE
// ================================================================================================
#define IB_THREADS 128
#define IB_INVALID 0xFFFFFFFFu

__device__ __forceinline__ uint32_t ib_rows(uint32_t n, uint32_t R) { return (n + R - 1) / R; }

// Walk 1: from every start row (multiples of R, plus the primary row as walker K) follow
// W[idx] >> 8 until the next start row; record the walk length and which walker it runs into.
// With `tmp`, the walk also keeps the first `cap` bytes it passes (F[idx] = W[idx] & 0xFF) in its own scratch
// row: walks are about R steps long, so a row of 8R bytes holds nearly every walk completely, and the output
// can then be assembled by a sequential copy instead of a second pass of dependent random loads.
// A walk that fills its row (cap steps; about 3 walks in 10 000 at cap = 8R) takes a fresh walker slot from the block's
// small overflow pool and goes on in that slot's row: it becomes two (or more) chained walkers, all complete in their
// rows, and no second, emitting walk of a thousand dependent loads is needed for it. ovf[b*(IB_OVF+1)] counts the slots
// taken, the words behind it keep the start rows of the overflow walkers.
#define IB_OVF 64u
__global__ void __launch_bounds__(IB_THREADS)
    ibwt_walk_len_kernel(const uint32_t* __restrict__ W, uint64_t stride, const uint32_t* __restrict__ len, const uint32_t* __restrict__ primary,
                         uint32_t R, uint32_t kmax, uint2* __restrict__ walk /* (len, succ) */, uint32_t* __restrict__ ovf,
                         uint8_t* __restrict__ tmp, uint32_t cap)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b];
    if (n == 0) return;
    const uint32_t K  = ib_rows(n, R);
    uint32_t       w  = blockIdx.x * IB_THREADS + threadIdx.x;
    if (w > K) return;
    const uint32_t pi = primary[b];
    const uint32_t* Wb = W + (uint64_t) b * stride;
    uint32_t idx = (w == K) ? pi : w * R;
    uint32_t steps = 0;
    if (tmp == nullptr)
    {
        do
        {
            idx = Wb[idx] >> 8;
            ++steps;
        } while (idx != pi && (idx & (R - 1u)) != 0);  // R is a power of two
    }
    else
    {
        uint4*   row = reinterpret_cast<uint4*>(tmp + ((uint64_t) b * kmax + w) * cap);  // cap is a multiple of 16
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // sixteen bytes gathered per store
        do
        {
            const uint32_t e  = Wb[idx];
            idx               = e >> 8;
            const uint32_t by = (e & 0xFFu) << ((steps & 3u) * 8u);
            const uint32_t q  = (steps >> 2) & 3u;
            a0 |= q == 0 ? by : 0u;
            a1 |= q == 1 ? by : 0u;
            a2 |= q == 2 ? by : 0u;
            a3 |= q == 3 ? by : 0u;
            ++steps;
            if ((steps & 15u) == 0)
            {
                if (steps <= cap) row[(steps >> 4) - 1] = make_uint4(a0, a1, a2, a3);
                a0 = a1 = a2 = a3 = 0;
                if (steps == cap && ovf != nullptr && idx != pi && (idx & (R - 1u)) != 0)
                {
                    // row full and the walk goes on: continue as an overflow walker, if the pool has a slot left
                    uint32_t* pool = ovf + (uint64_t) b * (IB_OVF + 1u);
                    const uint32_t o = atomicAdd(pool, 1u);
                    if (o < IB_OVF && K + 1u + o < kmax)
                    {
                        const uint32_t w2 = K + 1u + o;
                        walk[(uint64_t) b * kmax + w] = make_uint2(steps, w2);
                        pool[1u + o] = idx;
                        w     = w2;
                        row   = reinterpret_cast<uint4*>(tmp + ((uint64_t) b * kmax + w) * cap);
                        steps = 0;
                    }
                }
            }
        } while (idx != pi && (idx & (R - 1u)) != 0);  // R is a power of two
        if ((steps & 15u) != 0 && steps <= cap) row[steps >> 4] = make_uint4(a0, a1, a2, a3);
    }
    const uint32_t succ = (idx == pi) ? K : idx / R;
    walk[(uint64_t) b * kmax + w] = make_uint2(steps, succ);
}

// number of walkers of block b: start rows, the primary row, the overflow slots taken
__device__ __forceinline__ uint32_t ib_walkers(uint32_t K, const uint32_t* __restrict__ ovf, uint32_t b, uint32_t kmax)
{
    const uint32_t o = ovf ? min(ovf[(uint64_t) b * (IB_OVF + 1u)], IB_OVF) : 0u;
    return min(K + 1u + o, kmax);
}

// Assemble the output from the scratch rows. A warp takes 32 consecutive walkers: every lane fetches the record of one
// (offset, length: two coalesced loads instead of a chain of dependent ones per walker), then the warp copies the rows
// one after the other, a row being a few aligned 32-bit words per lane.
// `l` of `nl` lanes copy one row
__device__ __forceinline__ void ibwt_copy_row(const uint8_t* __restrict__ row, uint8_t* __restrict__ dst, uint32_t m, uint32_t l, uint32_t nl)
{
    // aligned 32-bit stores: the row is 16-byte aligned, the destination is not -- every lane funnels the two row words
    // that straddle its output word together; the few bytes before the first and after the last whole word go singly
    const uint32_t mis = min(m, (uint32_t) ((4u - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
    if (l < mis) dst[l] = row[l];
    const uint32_t  nw    = (m - mis) >> 2;
    const uint32_t* row32 = reinterpret_cast<const uint32_t*>(row);
    uint32_t*       dst32 = reinterpret_cast<uint32_t*>(dst + mis);
    for (uint32_t k = l; k < nw; k += nl)
    {
        const uint32_t sb = mis + 4u * k;
        const uint32_t w0 = row32[sb >> 2];
        const uint32_t w1 = (sb & 3u) ? row32[(sb >> 2) + 1u] : 0u;
        dst32[k]          = __funnelshift_r(w0, w1, (sb & 3u) * 8u);
    }
    const uint32_t tail0 = mis + 4u * nw;
    if (l < m - tail0) dst[tail0 + l] = row[tail0 + l];
}

__global__ void __launch_bounds__(256)
    ibwt_copy_kernel(const uint32_t* __restrict__ len, uint64_t stride, uint32_t R, uint32_t kmax, const uint2* __restrict__ walk,
                     const uint32_t* __restrict__ woff, const uint32_t* __restrict__ orbit, const uint8_t* __restrict__ tmp, uint32_t cap,
                     uint8_t* __restrict__ out, const uint32_t* __restrict__ ovf)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b];
    if (n == 0 || orbit[b] < n) return;  // periodic blocks are replicated by the emit walk
    const uint32_t K  = ib_rows(n, R);
    const uint32_t NW = ib_walkers(K, ovf, b, kmax);
    const uint32_t l  = lane_id();
    const uint32_t w0 = (blockIdx.x * 8 + warp_id()) * 32;
    if (w0 >= NW) return;
    uint32_t my_off = IB_INVALID, my_steps = 0;
    if (w0 + l < NW)
    {
        my_off   = woff[(uint64_t) b * kmax + w0 + l];
        my_steps = walk[(uint64_t) b * kmax + w0 + l].x;
        if (my_steps > cap) my_off = IB_INVALID;  // did not fit: the emit walk does it
    }
    uint8_t* ob = out + (uint64_t) b * stride;
    // walks are about R bytes long: with short rows (R = 64: sixteen words) the two halves of the warp copy one row each
    const uint32_t nl = R <= 64 ? 16u : 32u, per = 32u / nl, sub = l % nl, grp = l / nl;
    for (uint32_t i = 0; i < 32; i += per)
    {
        const uint32_t o0 = __shfl_sync(BRA_FULL, my_off, i + grp), steps = __shfl_sync(BRA_FULL, my_steps, i + grp);
        if (o0 == IB_INVALID) continue;
        ibwt_copy_row(tmp + ((uint64_t) b * kmax + w0 + i + grp) * cap, ob + o0, min(steps, n - min(n, o0)), sub, nl);
    }
}

// Stitch: order the walks from the primary row. If the chain closes before n bytes are covered the
// text is a repetition of that orbit (periodic input): orbit[b] < n and every walk is replicated.
// One CTA per block. The successor links form a cycle through walker K (the primary row); the offset of a
// walk is its distance from K along that cycle. It is found by pointer doubling in shared memory (about
// log2(K) rounds over all walkers) instead of following the chain link by link.
#define IB_STITCH_THREADS 256
#define IB_TERM 0xFFFFFFFFu
#define IB_TERM16 0xFFFFu  // (walker indices stay below 65535: ibwt_kmax)
__global__ void __launch_bounds__(IB_STITCH_THREADS)
    ibwt_stitch_kernel(const uint32_t* __restrict__ len, const uint32_t* __restrict__ primary, uint32_t R, uint32_t kmax,
                       const uint2* __restrict__ walk, uint32_t* __restrict__ woff, uint32_t* __restrict__ orbit, const uint32_t* __restrict__ ovf)
{
    extern __shared__ uint32_t s_dyn[];  // two (distance to the end of the cycle: u32, link: u16) pairs of kmax entries each, ping-pong
    const uint32_t b = blockIdx.x;
    const uint32_t n = len[b];
    if (n == 0)
    {
        if (threadIdx.x == 0) orbit[b] = 0;
        return;
    }
    const uint32_t K  = ib_rows(n, R);
    const uint32_t NW = ib_walkers(K, ovf, b, kmax);  // walkers 0..NW-1 (K is the primary row's, the ones behind it are overflow walkers)
    const uint32_t kp = (kmax + 1u) & ~1u;            // (keeps the u16 arrays behind the u32 ones 4-byte aligned)
    uint32_t *     d0 = s_dyn, *d1 = s_dyn + kp;
    uint16_t *     n0 = reinterpret_cast<uint16_t*>(s_dyn + 2 * kp), *n1 = n0 + kp;
    for (uint32_t w = threadIdx.x; w < NW; w += IB_STITCH_THREADS)
    {
        const uint2 e = walk[(uint64_t) b * kmax + w];
        d0[w]         = e.x;
        n0[w]         = e.y == K ? (uint16_t) IB_TERM16 : (uint16_t) e.y;  // the cycle is cut where it returns to K
    }
    __syncthreads();
    for (uint32_t span = 1;;)  // links every pointer has jumped so far
    {
        bool open = false;
        for (uint32_t w = threadIdx.x; w < NW; w += IB_STITCH_THREADS)
        {
            const uint32_t nx = n0[w];
            uint32_t       dd = d0[w], nn = IB_TERM16;
            if (nx != IB_TERM16)
            {
                dd += d0[nx];  // wraps for walkers that are not on K's cycle; they are discarded below
                nn   = n0[nx];
                open = true;
            }
            d1[w] = dd;
            n1[w] = (uint16_t) nn;
        }
        uint32_t* t = d0; d0 = d1; d1 = t;
        uint16_t* u = n0; n0 = n1; n1 = u;
        const bool any = __syncthreads_or(open);
        span <<= 1;
        // a walker on K's cycle is at most NW links from the cut; walkers on other cycles never reach it
        if (!any || span > NW) break;
    }
    const uint32_t q = d0[K];  // length of the cycle through the primary row
    const uint32_t pi = primary[b];
    for (uint32_t w = threadIdx.x; w < NW; w += IB_STITCH_THREADS)
    {
        uint32_t off = IB_INVALID;
        if (n0[w] == IB_TERM16 && !(w != K && pi % R == 0 && w == pi / R)) off = q - d0[w];  // a start row equal to the primary row is walker K's double
        woff[(uint64_t) b * kmax + w] = off;
    }
    if (threadIdx.x == 0) orbit[b] = q;
}

// Walk 2: emit F[idx] = W[idx] & 0xFF at the stitched offsets (replicated every `orbit` bytes).
__global__ void __launch_bounds__(IB_THREADS)
    ibwt_walk_emit_kernel(const uint32_t* __restrict__ W, uint64_t stride, const uint32_t* __restrict__ len, const uint32_t* __restrict__ primary,
                          uint32_t R, uint32_t kmax, const uint2* __restrict__ walk, const uint32_t* __restrict__ woff,
                          const uint32_t* __restrict__ orbit, uint8_t* __restrict__ out, uint32_t cap /* 0: no scratch rows */,
                          const uint32_t* __restrict__ ovf)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b];
    if (n == 0) return;
    const uint32_t K = ib_rows(n, R);
    const uint32_t w = blockIdx.x * IB_THREADS + threadIdx.x;
    if (w >= ib_walkers(K, ovf, b, kmax)) return;
    const uint32_t o0 = woff[(uint64_t) b * kmax + w];
    if (o0 == IB_INVALID) return;  // not on the primary row's orbit
    const uint32_t steps = walk[(uint64_t) b * kmax + w].x;
    const uint32_t q     = orbit[b];
    if (cap && q >= n && steps <= cap) return;  // assembled from the scratch rows by ibwt_copy_kernel
    const uint32_t* Wb   = W + (uint64_t) b * stride;
    uint8_t*        ob   = out + (uint64_t) b * stride;
    uint32_t idx = (w == K) ? primary[b] : (w < K ? w * R : ovf[(uint64_t) b * (IB_OVF + 1u) + 1u + (w - K - 1u)]);
    if (q >= n)
    {
        // a walk's bytes are consecutive: gather four into one aligned 32-bit store (a byte store costs a whole L2 transaction)
        const uint32_t m   = min(steps, n - min(n, o0));
        uint32_t       acc = 0, nacc = 0;
        for (uint32_t t = 0; t < m; ++t)
        {
            const uint32_t e = Wb[idx];
            idx              = e >> 8;
            acc |= (e & 0xFFu) << (8u * nacc);
            ++nacc;
            const uint32_t o = o0 + t + 1;  // bytes [o - nacc, o) are pending
            if (((reinterpret_cast<uintptr_t>(ob) + o) & 3u) == 0)
            {
                if (nacc == 4)
                    *reinterpret_cast<uint32_t*>(ob + o - 4) = acc;
                else
                    for (uint32_t i = 0; i < nacc; ++i) ob[o - nacc + i] = (uint8_t) (acc >> (8u * i));
                acc  = 0;
                nacc = 0;
            }
        }
        for (uint32_t i = 0; i < nacc; ++i) ob[o0 + m - nacc + i] = (uint8_t) (acc >> (8u * i));
    }
    else
    {
        // periodic block: the walkers of the primary row's cycle write its q bytes once; ibwt_replicate_kernel repeats them
        for (uint32_t t = 0; t < steps; ++t)
        {
            const uint32_t e = Wb[idx];
            if (o0 + t < n) ob[o0 + t] = (uint8_t) e;
            idx = e >> 8;
        }
    }
}

// Periodic blocks (orbit q shorter than n): out[i] = out[i mod q] for i >= q, sixteen bytes per thread.
__global__ void __launch_bounds__(256)
    ibwt_replicate_kernel(const uint32_t* __restrict__ len, uint64_t stride, const uint32_t* __restrict__ orbit, uint8_t* __restrict__ out)
{
    const uint32_t b = blockIdx.y;
    const uint32_t n = len[b], q = orbit[b];
    if (q == 0 || q >= n) return;
    uint8_t*       ob = out + (uint64_t) b * stride;
    const uint64_t i0 = (uint64_t) q + ((uint64_t) blockIdx.x * 256 + threadIdx.x) * 16;
    if (i0 >= n) return;
    const uint32_t m = (uint32_t) min((uint64_t) 16, (uint64_t) n - i0);
    uint32_t       x = (uint32_t) (i0 % q);
    uint32_t       w[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < m; ++k)
    {
        w[k >> 2] |= (uint32_t) ob[x] << ((k & 3u) * 8u);
        if (++x == q) x = 0;
    }
    uint8_t* dst = ob + i0;
    if (m == 16 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0)
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    else
        for (uint32_t k = 0; k < m; ++k) dst[k] = (uint8_t) (w[k >> 2] >> ((k & 3u) * 8u));
}

uint32_t ibwt_row_stride(uint32_t max_n)
{
    static const uint32_t max_walkers =  // tuning switch, read once; the stitch kernel holds 12 bytes per walker in shared memory and links them in 16 bits
        getenv("BRA_B200_IBWT_WALKERS") ? (uint32_t) std::min(16384, std::max(1024, atoi(getenv("BRA_B200_IBWT_WALKERS")))) : 16384u;
    uint32_t R = 64;
    while ((uint64_t) R * max_walkers < max_n) R *= 2;
    return R;
}
uint32_t ibwt_tmp_cap(uint32_t max_n) { return 8u * ibwt_row_stride(max_n); }
uint32_t ibwt_kmax(uint32_t max_n) { return (max_n + ibwt_row_stride(max_n) - 1) / ibwt_row_stride(max_n) + 1 + IB_OVF; }
uint32_t ibwt_ovf_words() { return IB_OVF + 1u; }

static bool bwt_inverse_chunk(const BwtInvArgs& a, cudaStream_t st)
{
    const uint32_t R = ibwt_row_stride(a.max_n), kmax = ibwt_kmax(a.max_n);
    if (!radix_pass_u8_index_packed(a.d_in, a.d_W, a.stride, a.d_len, a.max_n, a.nblk, a.d_hist, st)) return false;
    const dim3 grid(bra_div_up(kmax, IB_THREADS), a.nblk);
    const uint32_t cap = a.d_tmp ? ibwt_tmp_cap(a.max_n) : 0u;
    uint32_t* ovf = (a.d_tmp && a.d_ovf) ? a.d_ovf : nullptr;
    if (ovf) BRA_CUDA_TRY(cudaMemsetAsync(ovf, 0, (size_t) a.nblk * (IB_OVF + 1u) * sizeof(uint32_t), st));
    BRA_LAUNCH(P_IBWT_WALK_LEN, st, ibwt_walk_len_kernel<<<grid, IB_THREADS, 0, st>>>(a.d_W, a.stride, a.d_len, a.d_primary, R, kmax, a.d_walk, ovf, a.d_tmp, cap));
    const size_t stitch_smem = (size_t) ((kmax + 1u) & ~1u) * 2 * (sizeof(uint32_t) + sizeof(uint16_t));
    BRA_CUDA_TRY(cudaFuncSetAttribute(ibwt_stitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) stitch_smem));
    BRA_LAUNCH(P_IBWT_STITCH, st, ibwt_stitch_kernel<<<a.nblk, IB_STITCH_THREADS, stitch_smem, st>>>(a.d_len, a.d_primary, R, kmax, a.d_walk, a.d_woff, a.d_orbit, ovf));
    if (cap)
        BRA_LAUNCH(P_IBWT_COPY, st, ibwt_copy_kernel<<<dim3(bra_div_up(kmax, 256), a.nblk), 256, 0, st>>>(a.d_len, a.stride, R, kmax, a.d_walk, a.d_woff, a.d_orbit, a.d_tmp, cap, a.d_out, ovf));
    BRA_LAUNCH(P_IBWT_WALK_EMIT, st, ibwt_walk_emit_kernel<<<grid, IB_THREADS, 0, st>>>(a.d_W, a.stride, a.d_len, a.d_primary, R, kmax, a.d_walk, a.d_woff, a.d_orbit, a.d_out, cap, ovf));
    BRA_LAUNCH(P_IBWT_WALK_EMIT, st, ibwt_replicate_kernel<<<dim3(bra_div_up(a.max_n, 4096), a.nblk), 256, 0, st>>>(a.d_len, a.stride, a.d_orbit, a.d_out));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

// One call over the whole batch. (Building and walking W in L2-sized chunks of blocks was measured and is
// slower: with ~24 blocks per chunk the walk kernels are bound by the longest single walk, not by bandwidth.)
bool bwt_inverse_batch(const BwtInvArgs& a, cudaStream_t st)
{
    if (a.nblk == 0 || a.max_n == 0) return true;
    return bwt_inverse_chunk(a, st);
}

}  // namespace bra
