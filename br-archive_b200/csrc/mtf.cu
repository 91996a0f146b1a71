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
// Replay is one segment per THREAD: each lane owns a 256-byte list in shared memory (stride 65
// eight-byte words, so lanes working at the same depth hit different banks) and does what the reference
// does -- find / shift -- but eight list entries per shared-memory word. A warp therefore advances
// 32 segments at once, and the cost per symbol is (rank/8 + 1) word steps of the slowest lane:
// about one warp instruction per symbol on BWT output (ranks are small), ~20 on uniform random
// ranks. Traffic: read n + write n (+ 256 B of state per 4 KiB segment, twice). Issue/latency bound.
#include "bra_common.cuh"
#include "bra_kernels.h"

namespace bra {

#define MTF_SEG 4096
#define MTF_WARPS 4          // warps per CTA in the warp-per-segment summary kernel
#define MTF_LANE_WARPS 4     // warps per CTA in the lane-per-segment replay kernels
#define MTF_LIST_QW 33       // 32 eight-byte words of list + 1 pad: lane stride 33 -> lanes at the same depth use distinct banks

// ---- per-lane list in shared memory -----------------------------------------------------------------
// Q[0..31]: entry k lives in byte (k & 7) of the 64-bit word (k >> 3).

// decode one rank: returns the symbol at position r and moves it to the front
__device__ __forceinline__ uint32_t lane_mtf_decode(uint64_t* Q, uint32_t r)
{
    const uint32_t wr = r >> 3, br = r & 7u;
    uint64_t       cur = Q[wr];
    const uint32_t sym = (uint32_t) (cur >> (br * 8)) & 0xFFu;
    if (r == 0) return sym;
    uint64_t       below    = wr ? Q[wr - 1] : 0ull;
    uint64_t       incoming = wr ? (below >> 56) : (uint64_t) sym;
    const uint64_t mask     = br == 7 ? ~0ull : ((1ull << ((br + 1) * 8)) - 1ull);  // bytes <= br take the shifted image
    Q[wr] = (((cur << 8) | incoming) & mask) | (cur & ~mask);
    for (int w = (int) wr - 1; w >= 0; --w)
    {
        cur      = below;
        below    = w ? Q[w - 1] : 0ull;
        incoming = w ? (below >> 56) : (uint64_t) sym;
        Q[w]     = (cur << 8) | incoming;
    }
    return sym;
}

// encode one symbol: returns its position and moves it to the front (single forward pass)
__device__ __forceinline__ uint32_t lane_mtf_encode(uint64_t* Q, uint32_t x)
{
    const uint64_t x8    = (uint64_t) x * 0x0101010101010101ull;
    uint64_t       carry = x;  // byte entering the next word from below
    for (uint32_t w = 0;; ++w)
    {
        const uint64_t cur = Q[w];
        const uint64_t t   = cur ^ x8;
        const uint64_t z   = (t - 0x0101010101010101ull) & ~t & 0x8080808080808080ull;  // lowest marker = first byte equal to x
        if (z == 0)
        {
            Q[w]  = (cur << 8) | carry;
            carry = cur >> 56;
            continue;
        }
        const uint32_t b = (uint32_t) (__ffsll((long long) z) - 1) >> 3;
        if (w == 0 && b == 0) return 0;  // already in front
        const uint64_t mask = b == 7 ? ~0ull : ((1ull << ((b + 1) * 8)) - 1ull);
        Q[w] = (((cur << 8) | carry) & mask) | (cur & ~mask);
        return w * 8 + b;
    }
}

// One segment per lane. MODE 0: encode (symbols -> ranks) from the segment's entry list.
// MODE 2: decode from the IDENTITY list: the output is, per rank, the entry-list POSITION of the decoded
// symbol ("relative symbol"), and the final list is the position permutation of the segment (its summary).
// After the per-block scan has produced the entry lists, mtf_map_kernel turns relative symbols into symbols
// with a 256-entry table lookup -- so decoding replays each segment once, not twice.
// (MODE 1, decode from a known entry list, is kept for completeness; the batch path does not use it.)
template <int MODE>
__global__ void __launch_bounds__(MTF_LANE_WARPS * 32)
    mtf_lane_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                    const uint8_t* __restrict__ state_in, uint8_t* __restrict__ summ_out)
{
    __shared__ uint64_t s_list[MTF_LANE_WARPS * 32 * MTF_LIST_QW];
    const uint32_t b   = blockIdx.y;
    const uint32_t n   = len[b];
    const uint32_t seg = blockIdx.x * (MTF_LANE_WARPS * 32) + threadIdx.x;
    if ((uint64_t) seg * MTF_SEG >= n) return;  // no barriers below: lanes are independent
    const uint32_t m   = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    const uint64_t off = (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    uint64_t*      W   = s_list + threadIdx.x * MTF_LIST_QW;

    if (MODE == 2)
    {
#pragma unroll 8
        for (uint32_t w = 0; w < 32; ++w) W[w] = 0x0706050403020100ull + (uint64_t) (w * 8) * 0x0101010101010101ull;
    }
    else
    {
        const uint4* st = reinterpret_cast<const uint4*>(state_in + ((uint64_t) b * segs + seg) * 256);
#pragma unroll 4
        for (uint32_t q = 0; q < 16; ++q)
        {
            const uint4 v = st[q];
            W[q * 2 + 0]  = ((uint64_t) v.y << 32) | v.x;
            W[q * 2 + 1]  = ((uint64_t) v.w << 32) | v.z;
        }
    }

    const uint8_t* ip = in + off;
    uint8_t*       op = out + off;
    uint32_t       i  = 0;
    for (; i + 16 <= m; i += 16)
    {
        const uint4    v     = *reinterpret_cast<const uint4*>(ip + i);
        const uint32_t iw[4] = {v.x, v.y, v.z, v.w};
        uint32_t       ow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const uint32_t x = (iw[q] >> (k * 8)) & 0xFFu;
                const uint32_t r = (MODE == 0) ? lane_mtf_encode(W, x) : lane_mtf_decode(W, x);
                o |= r << (k * 8);
            }
            ow[q] = o;
        }
        *reinterpret_cast<uint4*>(op + i) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    for (; i < m; ++i)
    {
        const uint32_t x = ip[i];
        const uint32_t r = (MODE == 0) ? lane_mtf_encode(W, x) : lane_mtf_decode(W, x);
        op[i] = (uint8_t) r;
    }
    if (MODE == 2)
    {
        uint4* so = reinterpret_cast<uint4*>(summ_out + ((uint64_t) b * segs + seg) * 256);
#pragma unroll 4
        for (uint32_t q = 0; q < 16; ++q)
            so[q] = make_uint4((uint32_t) W[q * 2], (uint32_t) (W[q * 2] >> 32), (uint32_t) W[q * 2 + 1], (uint32_t) (W[q * 2 + 1] >> 32));
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

// relative symbols -> symbols: out[i] = entry_list_of_segment[out[i]], in place
__global__ void __launch_bounds__(256)
    mtf_map_kernel(uint8_t* __restrict__ data, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs, const uint8_t* __restrict__ state)
{
    __shared__ uint8_t lut[256];
    const uint32_t b = blockIdx.y, seg = blockIdx.x;
    const uint32_t n = len[b];
    if ((uint64_t) seg * MTF_SEG >= n) return;
    const uint32_t m = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    lut[threadIdx.x] = state[((uint64_t) b * segs + seg) * 256 + threadIdx.x];
    __syncthreads();
    uint8_t*       p = data + (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    const uint32_t i = threadIdx.x * 16;
    if (i + 16 <= m)
    {
        uint4    v    = *reinterpret_cast<uint4*>(p + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[q] = (uint32_t) lut[w[q] & 0xFFu] | ((uint32_t) lut[(w[q] >> 8) & 0xFFu] << 8) | ((uint32_t) lut[(w[q] >> 16) & 0xFFu] << 16) |
                   ((uint32_t) lut[w[q] >> 24] << 24);
        *reinterpret_cast<uint4*>(p + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    else
        for (uint32_t k = i; k < m; ++k) p[k] = lut[p[k]];
}

uint32_t mtf_segments(uint32_t max_n) { return bra_div_up(max_n, MTF_SEG); }

bool mtf_encode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                      uint8_t* d_summ, uint16_t* d_scnt, uint8_t* d_state, cudaStream_t st)
{
    if (nblk == 0 || max_n == 0) return true;
    const uint32_t segs = mtf_segments(max_n);
    const dim3     grid(bra_div_up(segs, MTF_WARPS), nblk);
    const dim3     lgrid(bra_div_up(segs, MTF_LANE_WARPS * 32), nblk);
    BRA_LAUNCH(P_MTF_SUMMARY, st, mtf_enc_summary_kernel<<<grid, MTF_WARPS * 32, 0, st>>>(d_in, stride, d_len, segs, d_summ, d_scnt));
    BRA_LAUNCH(P_MTF_SCAN, st, mtf_scan_kernel<true><<<nblk, 32, 0, st>>>(d_summ, d_scnt, d_len, segs, d_state));
    BRA_LAUNCH(P_MTF_APPLY, st, mtf_lane_kernel<0><<<lgrid, MTF_LANE_WARPS * 32, 0, st>>>(d_in, d_out, stride, d_len, segs, d_state, nullptr));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

bool mtf_decode_batch(const uint8_t* d_in, uint8_t* d_out, uint64_t stride, const uint32_t* d_len, uint32_t max_n, uint32_t nblk,
                      uint8_t* d_summ, uint8_t* d_state, cudaStream_t st)
{
    if (nblk == 0 || max_n == 0) return true;
    const uint32_t segs = mtf_segments(max_n);
    const dim3     lgrid(bra_div_up(segs, MTF_LANE_WARPS * 32), nblk);
    BRA_LAUNCH(P_MTF_APPLY, st, mtf_lane_kernel<2><<<lgrid, MTF_LANE_WARPS * 32, 0, st>>>(d_in, d_out, stride, d_len, segs, nullptr, d_summ));
    BRA_LAUNCH(P_MTF_SCAN, st, mtf_scan_kernel<false><<<nblk, 32, 0, st>>>(d_summ, nullptr, d_len, segs, d_state));
    BRA_LAUNCH(P_MTF_SUMMARY, st, mtf_map_kernel<<<dim3(segs, nblk), 256, 0, st>>>(d_out, stride, d_len, segs, d_state));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

}  // namespace bra
