// mtf.cu -- move-to-front coding and decoding for a batch of blocks
// (replaces reference bra_mtf_encode2 / bra_mtf_decode2, src/encoders/bra_mtf.c:67-82, :98-115).
//
// The reference walks the block with one 256-entry list (identity at block start,
// bra_mtf.c:9-13). Here every block is cut into 4 KiB segments that run in parallel:
//   summary : what a segment does to ANY incoming list, without knowing it
//               encode -> the segment's distinct symbols in most-recent-first order
//               decode -> the permutation of list positions the segment's ranks perform
//   scan    : per block, compose the summaries left to right -> the list entering each segment
//   apply   : replay each segment from its entry list.
// The list lives in registers, 8 entries per lane of a warp; one symbol costs a byte-compare,
// a ballot and a funnel shift across lanes, independent of the rank (random data has mean rank
// ~128, which is what makes the reference's linear search slow).
// Traffic: read n + write n (+ 256 B of state per 4 KiB segment, twice). Latency/issue bound.
#include "bra_common.cuh"
#include "bra_kernels.h"

namespace bra {

#define MTF_SEG 4096
#define MTF_WARPS 4  // warps (= segments) per CTA

// ---- warp-resident list ---------------------------------------------------------------------------
// Entries 0..7 (where almost every hit lands on BWT output) are replicated in every lane as the
// uniform pair (f0, f1): hits there cost no cross-lane traffic. Entries 8..255 are distributed,
// lane l (1..31) holding entries 8l..8l+7 in (lo, hi), entry 8l in the low byte of lo. Lane 0's
// (lo, hi) are unused.
struct WarpList
{
    uint32_t f0, f1;  // uniform: entries 0-3, 4-7
    uint32_t lo, hi;  // per lane: entries 8l..8l+3, 8l+4..8l+7
};

__device__ __forceinline__ WarpList wl_from_lane_words(uint32_t lo, uint32_t hi)
{
    WarpList r;
    r.lo = lo;
    r.hi = hi;
    r.f0 = __shfl_sync(BRA_FULL, lo, 0);
    r.f1 = __shfl_sync(BRA_FULL, hi, 0);
    return r;
}
__device__ __forceinline__ WarpList wl_identity()
{
    const uint32_t b = lane_id() * 8;
    return wl_from_lane_words((b) | ((b + 1) << 8) | ((b + 2) << 16) | ((b + 3) << 24), (b + 4) | ((b + 5) << 8) | ((b + 6) << 16) | ((b + 7) << 24));
}
__device__ __forceinline__ WarpList wl_load(const uint8_t* p)  // 256 bytes, 8-byte aligned
{
    const uint2 v = reinterpret_cast<const uint2*>(p)[lane_id()];
    return wl_from_lane_words(v.x, v.y);
}
__device__ __forceinline__ void wl_store(uint8_t* p, const WarpList& l)
{
    reinterpret_cast<uint2*>(p)[lane_id()] = lane_id() == 0 ? make_uint2(l.f0, l.f1) : make_uint2(l.lo, l.hi);
}

// 0x80 in the lowest byte of `w` equal to the replicated byte s4 (higher marks may be spurious: use the lowest)
__device__ __forceinline__ uint32_t byte_match(uint32_t w, uint32_t s4)
{
    const uint32_t t = w ^ s4;
    return (t - 0x01010101u) & ~t & 0x80808080u;
}

// front part: move entry p (1..7, uniform) to position 0; x is its value
__device__ __forceinline__ void wl_rotate_front(WarpList& L, uint32_t p, uint32_t x)
{
    const uint64_t v    = ((uint64_t) L.f1 << 32) | L.f0;
    const uint64_t low  = (p == 7) ? ~0ull : ((1ull << ((p + 1) * 8)) - 1ull);
    const uint64_t nv   = (v & ~low) | (((v << 8) | x) & low);
    L.f0 = (uint32_t) nv;
    L.f1 = (uint32_t) (nv >> 32);
}

// distributed part: entry at position pos >= 8 (uniform) moves to the front; x is its value
__device__ __forceinline__ void wl_move_far(WarpList& L, uint32_t pos, uint32_t x)
{
    const uint32_t lane = lane_id();
    const uint32_t hl = pos >> 3, hb = pos & 7u;
    // byte entering each lane from its left neighbour; lane 1 receives entry 7 of the front part
    uint32_t incoming = __shfl_up_sync(BRA_FULL, (lane == 0 ? L.f1 : L.hi) >> 24, 1);
    const uint64_t v       = ((uint64_t) L.hi << 32) | L.lo;
    const uint64_t shifted = (v << 8) | incoming;
    uint64_t       nv      = v;
    if (lane < hl)
        nv = shifted;
    else if (lane == hl)
    {
        const uint64_t keep = (hb == 7) ? 0ull : (~0ull << ((hb + 1) * 8));  // entries above the hit stay
        nv                  = (v & keep) | (shifted & ~keep);
    }
    L.lo = (uint32_t) nv;
    L.hi = (uint32_t) (nv >> 32);
    L.f1 = (L.f1 << 8) | (L.f0 >> 24);
    L.f0 = (L.f0 << 8) | x;
}

// encode one symbol (uniform): returns its rank and updates the list
__device__ __forceinline__ uint32_t wl_encode(WarpList& L, uint32_t x)
{
    const uint32_t s4 = x * 0x01010101u;
    const uint32_t z0 = byte_match(L.f0, s4);
    if (z0)
    {
        const uint32_t p = (__ffs(z0) - 1) >> 3;
        if (p) wl_rotate_front(L, p, x);
        return p;
    }
    const uint32_t z1 = byte_match(L.f1, s4);
    if (z1)
    {
        const uint32_t p = 4 + ((__ffs(z1) - 1) >> 3);
        wl_rotate_front(L, p, x);
        return p;
    }
    const uint32_t m0 = byte_match(L.lo, s4), m1 = byte_match(L.hi, s4);
    const uint32_t ball = __ballot_sync(BRA_FULL, lane_id() != 0 && (m0 | m1) != 0u);
    const uint32_t hl   = __ffs(ball) - 1;
    const uint32_t in_lane = m0 ? ((__ffs(m0) - 1) >> 3) : (4 + ((__ffs(m1) - 1) >> 3));
    const uint32_t pos = hl * 8 + __shfl_sync(BRA_FULL, in_lane, hl);
    wl_move_far(L, pos, x);
    return pos;
}

// decode one rank (uniform): returns the symbol and updates the list
__device__ __forceinline__ uint32_t wl_decode(WarpList& L, uint32_t r)
{
    if (r < 8)
    {
        const uint32_t x = ((r & 4u ? L.f1 : L.f0) >> ((r & 3u) * 8)) & 0xFFu;
        if (r) wl_rotate_front(L, r, x);
        return x;
    }
    const uint32_t w = (r & 4u) ? L.hi : L.lo;
    const uint32_t x = __shfl_sync(BRA_FULL, (w >> ((r & 3u) * 8)) & 0xFFu, r >> 3);
    wl_move_far(L, r, x);
    return x;
}

// ---- segment replay (shared by summary/apply) ---------------------------------------------------
// ENCODE: in = symbols, out = ranks. DECODE: in = ranks, out = symbols. out may be null (summary).
// Four symbols travel per shuffle; a word that repeats the front symbol (encode) or is all-zero
// ranks (decode) -- the common case on BWT output -- is retired without touching the list.
template <bool ENCODE, bool WRITE>
__device__ __forceinline__ void mtf_replay(WarpList& L, const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t m)
{
    const uint32_t lane = lane_id();
    for (uint32_t base = 0; base < m; base += 128)
    {
        const uint32_t cnt = min(128u, m - base);
        uint32_t       wrd = 0;
        if (base + lane * 4 + 4 <= m)
            wrd = *reinterpret_cast<const uint32_t*>(in + base + lane * 4);
        else
            for (uint32_t k = 0; k < 4; ++k)
                if (base + lane * 4 + k < m) wrd |= (uint32_t) in[base + lane * 4 + k] << (8 * k);
        uint32_t       myow   = 0;
        const uint32_t nwords = (cnt + 3) >> 2;
        for (uint32_t wi = 0; wi < nwords; ++wi)
        {
            const uint32_t w    = __shfl_sync(BRA_FULL, wrd, wi);
            const uint32_t nsym = min(4u, cnt - wi * 4);
            const uint32_t fr   = L.f0 & 0xFFu;
            uint32_t       ow;
            if (nsym == 4 && (ENCODE ? (w == fr * 0x01010101u) : (w == 0u)))
                ow = ENCODE ? 0u : fr * 0x01010101u;
            else
            {
                ow = 0;
                for (uint32_t k = 0; k < nsym; ++k)
                {
                    const uint32_t x   = (w >> (k * 8)) & 0xFFu;
                    const uint32_t res = ENCODE ? wl_encode(L, x) : wl_decode(L, x);
                    ow |= res << (k * 8);
                }
            }
            if (WRITE && lane == wi) myow = ow;
        }
        if (WRITE)
        {
            if (base + lane * 4 + 4 <= m)
                *reinterpret_cast<uint32_t*>(out + base + lane * 4) = myow;
            else
                for (uint32_t k = 0; k < 4; ++k)
                    if (base + lane * 4 + k < m) out[base + lane * 4 + k] = (myow >> (8 * k)) & 0xFFu;
        }
    }
}

// ---- summaries ------------------------------------------------------------------------------
// encode: recency list of the segment's distinct symbols. Computed from last-occurrence positions
// (no sequential dependence): order[r] = symbol with the r-th largest last occurrence; cnt = #distinct.
__global__ void __launch_bounds__(MTF_WARPS * 32)
    mtf_enc_summary_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                           uint8_t* __restrict__ summ /* [b][seg][256] */, uint16_t* __restrict__ scnt /* [b][seg] */)
{
    __shared__ int s_last[MTF_WARPS][256];
    const uint32_t b = blockIdx.y, w = warp_id(), lane = lane_id();
    const uint32_t seg = blockIdx.x * MTF_WARPS + w;
    const uint32_t n   = len[b];
    if ((uint64_t) seg * MTF_SEG >= n) return;  // whole warps leave; no CTA barrier below
    const uint32_t m = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    const uint8_t* p = in + (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    int*           last = s_last[w];
    for (int i = lane; i < 256; i += 32) last[i] = -1;
    __syncwarp();
    for (uint32_t i = lane * 4; i < m; i += 128)
    {
        if (i + 4 <= m)
        {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(p + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicMax(&last[(v >> (8 * k)) & 0xFFu], (int) (i + k));
        }
        else
            for (uint32_t k = 0; i + k < m; ++k) atomicMax(&last[p[i + k]], (int) (i + k));
    }
    __syncwarp();
    // rank each present symbol by counting symbols with a later last occurrence
    uint8_t* o = summ + ((uint64_t) b * segs + seg) * 256;
    uint32_t present = 0;
    int      mine[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        mine[k] = last[lane * 8 + k];
        present += mine[k] >= 0;
    }
    uint32_t rk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < 256; ++s)
    {
        const int ls = last[s];  // broadcast read
#pragma unroll
        for (int k = 0; k < 8; ++k) rk[k] += ls > mine[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (mine[k] >= 0) o[rk[k]] = (uint8_t) (lane * 8 + k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) present += __shfl_xor_sync(BRA_FULL, present, d);
    if (lane == 0) scnt[(uint64_t) b * segs + seg] = (uint16_t) present;
}

// decode: permutation of positions performed by the segment = replay from the identity list
__global__ void __launch_bounds__(MTF_WARPS * 32)
    mtf_dec_summary_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                           uint8_t* __restrict__ summ)
{
    const uint32_t b = blockIdx.y, w = warp_id();
    const uint32_t seg = blockIdx.x * MTF_WARPS + w;
    const uint32_t n   = len[b];
    if ((uint64_t) seg * MTF_SEG >= n) return;
    const uint32_t m = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    WarpList       L = wl_identity();
    mtf_replay<false, false>(L, in + (uint64_t) b * stride + (uint64_t) seg * MTF_SEG, nullptr, m);
    wl_store(summ + ((uint64_t) b * segs + seg) * 256, L);
}

// ---- scan: one warp per block composes the summaries; state[b][seg] = list entering the segment ----
template <bool ENCODE>
__global__ void __launch_bounds__(32)
    mtf_scan_kernel(const uint8_t* __restrict__ summ, const uint16_t* __restrict__ scnt, const uint32_t* __restrict__ len, uint32_t segs,
                    uint8_t* __restrict__ state)
{
    __shared__ __align__(16) uint8_t cur[256];
    __shared__ __align__(16) uint8_t nxt[256];
    __shared__ __align__(16) uint8_t member[256];
    const uint32_t b = blockIdx.x, lane = lane_id();
    const uint32_t n = len[b];
    const uint32_t nseg = (n + MTF_SEG - 1) / MTF_SEG;
    for (int i = lane; i < 256; i += 32) cur[i] = (uint8_t) i;
    __syncwarp();
    for (uint32_t s = 0; s < nseg; ++s)
    {
        uint8_t*       st = state + ((uint64_t) b * segs + s) * 256;
        const uint8_t* sm = summ + ((uint64_t) b * segs + s) * 256;
        reinterpret_cast<uint2*>(st)[lane] = reinterpret_cast<const uint2*>(cur)[lane];
        if (s + 1 == nseg) break;
        if (ENCODE)
        {
            // new list = segment's recency list, then the old list minus those symbols (order kept)
            const uint32_t k = scnt[(uint64_t) b * segs + s];
            for (int i = lane; i < 256; i += 32) member[i] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < k; i += 32)
            {
                const uint8_t sy = sm[i];
                nxt[i]           = sy;
                member[sy]       = 1;
            }
            __syncwarp();
            uint32_t outp = k;
            for (int r = 0; r < 8; ++r)
            {
                const uint8_t  sy   = cur[r * 32 + lane];
                const bool     keep = !member[sy];
                const uint32_t ball = __ballot_sync(BRA_FULL, keep);
                if (keep) nxt[outp + __popc(ball & lanemask_lt())] = sy;
                outp += __popc(ball);
            }
        }
        else
        {
            // positions are permuted: new[j] = old[perm[j]]
            for (int i = lane; i < 256; i += 32) nxt[i] = cur[sm[i]];
        }
        __syncwarp();
        for (int i = lane; i < 256; i += 32) cur[i] = nxt[i];
        __syncwarp();
    }
}

// ---- apply ---------------------------------------------------------------------------------------
template <bool ENCODE>
__global__ void __launch_bounds__(MTF_WARPS * 32)
    mtf_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                     const uint8_t* __restrict__ state)
{
    const uint32_t b = blockIdx.y, w = warp_id();
    const uint32_t seg = blockIdx.x * MTF_WARPS + w;
    const uint32_t n   = len[b];
    if ((uint64_t) seg * MTF_SEG >= n) return;
    const uint32_t m   = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    const uint64_t off = (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    WarpList       L   = wl_load(state + ((uint64_t) b * segs + seg) * 256);
    mtf_replay<ENCODE, true>(L, in + off, out + off, m);
}

// ---- encode apply with the INVERSE list ------------------------------------------------------------
// For encoding only the rank of the incoming symbol matters, so the warp keeps pos[symbol] instead of
// the list: lane l holds the positions of symbols 8l..8l+7 as bytes of (plo, phi). One symbol costs a
// shuffle (read pos[x]), a branch-free SIMD-within-register "+1 to every position below pos[x]" on
// the lane's eight bytes, and a byte clear -- independent of the rank, which is what uniform random
// input (mean rank ~128) needs.
__device__ __forceinline__ uint32_t swar_inc_below(uint32_t a, uint32_t pl, bool p_high)
{
    // per byte: a += (a < p), p replicated; pl = low 7 bits of p replicated, p_high = bit 7 of p
    const uint32_t t  = ((a & 0x7F7F7F7Fu) | 0x80808080u) - pl;  // bit 7 of each byte: low7(a) >= low7(p)
    const uint32_t lt = p_high ? ~(a & t) : ~(a | t);            // bit 7: a < p
    return a + ((lt & 0x80808080u) >> 7);
}

__global__ void __launch_bounds__(MTF_WARPS * 32)
    mtf_enc_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                         const uint8_t* __restrict__ state)
{
    __shared__ __align__(8) uint8_t s_pos[MTF_WARPS][256];
    const uint32_t b = blockIdx.y, w = warp_id(), lane = lane_id();
    const uint32_t seg = blockIdx.x * MTF_WARPS + w;
    const uint32_t n   = len[b];
    if ((uint64_t) seg * MTF_SEG >= n) return;  // whole warps leave; only warp-level syncs below
    const uint32_t m   = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    const uint64_t off = (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    const uint8_t* ip  = in + off;
    uint8_t*       op  = out + off;

    // invert the entry list: pos[list[j]] = j
    const uint2 ent = reinterpret_cast<const uint2*>(state + ((uint64_t) b * segs + seg) * 256)[lane];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        s_pos[w][(ent.x >> (8 * k)) & 0xFFu] = (uint8_t) (lane * 8 + k);
        s_pos[w][(ent.y >> (8 * k)) & 0xFFu] = (uint8_t) (lane * 8 + 4 + k);
    }
    __syncwarp();
    const uint2 pv  = reinterpret_cast<const uint2*>(s_pos[w])[lane];
    uint32_t    plo = pv.x, phi = pv.y;
    uint32_t    front = __shfl_sync(BRA_FULL, ent.x, 0) & 0xFFu;  // symbol at position 0

    for (uint32_t base = 0; base < m; base += 128)
    {
        const uint32_t cnt = min(128u, m - base);
        uint32_t       wrd = 0;
        if (base + lane * 4 + 4 <= m)
            wrd = *reinterpret_cast<const uint32_t*>(ip + base + lane * 4);
        else
            for (uint32_t k = 0; k < 4; ++k)
                if (base + lane * 4 + k < m) wrd |= (uint32_t) ip[base + lane * 4 + k] << (8 * k);
        uint32_t       myow   = 0;
        const uint32_t nwords = (cnt + 3) >> 2;
        for (uint32_t wi = 0; wi < nwords; ++wi)
        {
            const uint32_t wv   = __shfl_sync(BRA_FULL, wrd, wi);
            const uint32_t nsym = min(4u, cnt - wi * 4);
            uint32_t       ow   = 0;
            if (!(nsym == 4 && wv == front * 0x01010101u))
            {
                for (uint32_t k = 0; k < nsym; ++k)
                {
                    const uint32_t x = (wv >> (k * 8)) & 0xFFu;
                    if (x == front) continue;  // rank 0, list unchanged
                    const uint32_t sh = (x & 3u) * 8;
                    const uint32_t p  = (__shfl_sync(BRA_FULL, (x & 4u) ? phi : plo, x >> 3) >> sh) & 0xFFu;
                    const uint32_t pl = (p & 0x7Fu) * 0x01010101u;
                    const bool     ph = (p & 0x80u) != 0;
                    plo = swar_inc_below(plo, pl, ph);
                    phi = swar_inc_below(phi, pl, ph);
                    if (lane == (x >> 3))
                    {
                        if (x & 4u)
                            phi &= ~(0xFFu << sh);
                        else
                            plo &= ~(0xFFu << sh);
                    }
                    front = x;
                    ow |= p << (k * 8);
                }
            }
            if (lane == wi) myow = ow;
        }
        if (base + lane * 4 + 4 <= m)
            *reinterpret_cast<uint32_t*>(op + base + lane * 4) = myow;
        else
            for (uint32_t k = 0; k < 4; ++k)
                if (base + lane * 4 + k < m) op[base + lane * 4 + k] = (myow >> (8 * k)) & 0xFFu;
    }
}

uint32_t mtf_segments(uint32_t max_n) { return bra_div_up(max_n, MTF_SEG); }

bool mtf_encode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                      uint8_t* d_summ, uint16_t* d_scnt, uint8_t* d_state, cudaStream_t st)
{
    if (nblk == 0 || max_n == 0) return true;
    const uint32_t segs = mtf_segments(max_n);
    const dim3     grid(bra_div_up(segs, MTF_WARPS), nblk);
    BRA_LAUNCH(P_MTF_SUMMARY, st, mtf_enc_summary_kernel<<<grid, MTF_WARPS * 32, 0, st>>>(d_in, stride, d_len, segs, d_summ, d_scnt));
    BRA_LAUNCH(P_MTF_SCAN, st, mtf_scan_kernel<true><<<nblk, 32, 0, st>>>(d_summ, d_scnt, d_len, segs, d_state));
    BRA_LAUNCH(P_MTF_APPLY, st, mtf_enc_apply_kernel<<<grid, MTF_WARPS * 32, 0, st>>>(d_in, d_out, stride, d_len, segs, d_state));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

bool mtf_decode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                      uint8_t* d_summ, uint8_t* d_state, cudaStream_t st)
{
    if (nblk == 0 || max_n == 0) return true;
    const uint32_t segs = mtf_segments(max_n);
    const dim3     grid(bra_div_up(segs, MTF_WARPS), nblk);
    BRA_LAUNCH(P_MTF_SUMMARY, st, mtf_dec_summary_kernel<<<grid, MTF_WARPS * 32, 0, st>>>(d_in, stride, d_len, segs, d_summ));
    BRA_LAUNCH(P_MTF_SCAN, st, mtf_scan_kernel<false><<<nblk, 32, 0, st>>>(d_summ, nullptr, d_len, segs, d_state));
    BRA_LAUNCH(P_MTF_APPLY, st, mtf_apply_kernel<false><<<grid, MTF_WARPS * 32, 0, st>>>(d_in, d_out, stride, d_len, segs, d_state));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

}  // namespace bra
