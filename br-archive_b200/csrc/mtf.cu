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
// Replay is one segment per THREAD: each lane owns a 256-byte list in shared memory (stride 17
// sixteen-byte words, so lanes working at the same depth hit different bank groups) and does what the
// reference does -- find / shift -- but sixteen list entries per shared-memory access. A warp therefore
// advances 32 segments at once, and the cost per symbol is (rank/16 + 1) chunk steps of the slowest lane:
// a few warp instructions per symbol on BWT output (ranks are small), the instruction count of ~16 chunk
// steps on uniform random ranks. Traffic: read n + write n (+ 256 B of state per 4 KiB segment, twice).
// Issue bound.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"

namespace bra {

#define MTF_SEG 4096
#define MTF_WARPS 4          // warps per CTA in the warp-per-segment summary kernel
#define MTF_LANE_WARPS 4     // warps per CTA in the lane-per-segment replay kernels
#define MTF_LIST_V4 17       // 16 sixteen-byte words of list + 1 pad: odd lane stride -> lanes at the same depth use distinct bank groups

// ---- per-lane list in shared memory -----------------------------------------------------------------
// The list operations (bra_mtf_list_decode / bra_mtf_list_encode: find, shift by one byte, insert at the front,
// sixteen entries per 128-bit access) live in bra_hd.h so that the CPU suite runs the very same code against a
// naive move-to-front.
__device__ __forceinline__ uint32_t lane_mtf_decode(uint4* Q, uint32_t r) { return bra_mtf_list_decode(Q, r); }
__device__ __forceinline__ uint32_t lane_mtf_encode(uint4* Q, uint32_t x) { return bra_mtf_list_encode(Q, x); }

// One segment per lane. MODE 0: encode (symbols -> ranks) from the segment's entry list.
// MODE 2: decode from the IDENTITY list: the output is, per rank, the entry-list POSITION of the decoded
// symbol ("relative symbol"), and the final list is the position permutation of the segment (its summary).
// After the per-block scan has produced the entry lists, mtf_map_kernel turns relative symbols into symbols
// with a 256-entry table lookup -- so decoding replays each segment once, not twice.
template <int MODE>
__global__ void __launch_bounds__(MTF_LANE_WARPS * 32)
    mtf_lane_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t stride, const uint32_t* __restrict__ len, uint32_t segs,
                    const uint8_t* __restrict__ state_in, uint8_t* __restrict__ summ_out)
{
    __shared__ uint4 s_list[MTF_LANE_WARPS * 32 * MTF_LIST_V4];
    const uint32_t b   = blockIdx.y;
    const uint32_t n   = len[b];
    const uint32_t seg = blockIdx.x * (MTF_LANE_WARPS * 32) + threadIdx.x;
    if ((uint64_t) seg * MTF_SEG >= n) return;  // no barriers below: lanes are independent
    const uint32_t m   = min((uint32_t) MTF_SEG, n - seg * MTF_SEG);
    const uint64_t off = (uint64_t) b * stride + (uint64_t) seg * MTF_SEG;
    uint4*         W   = s_list + threadIdx.x * MTF_LIST_V4;

    if (MODE == 2)
    {
#pragma unroll 4
        for (uint32_t q = 0; q < 16; ++q)
        {
            const uint32_t w0 = 0x03020100u + (q * 16) * 0x01010101u;
            W[q]              = make_uint4(w0, w0 + 0x04040404u, w0 + 0x08080808u, w0 + 0x0C0C0C0Cu);
        }
    }
    else
    {
        const uint4* st = reinterpret_cast<const uint4*>(state_in + ((uint64_t) b * segs + seg) * 256);
#pragma unroll 4
        for (uint32_t q = 0; q < 16; ++q) W[q] = st[q];
    }

    const uint8_t* ip = in + off;
    uint8_t*       op = out + off;
    uint32_t       i  = 0;
    for (; i + 16 <= m; i += 16)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(ip + i);
        uint4       ov;
        // four symbols per trip, the trip not unrolled: the replay code is long and sixteen copies of it overflow the instruction cache
#pragma unroll 1
        for (int q = 0; q < 4; ++q)
        {
            const uint32_t iw = q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w));
            uint32_t       o  = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const uint32_t x = (iw >> (k * 8)) & 0xFFu;
                const uint32_t r = (MODE == 0) ? lane_mtf_encode(W, x) : lane_mtf_decode(W, x);
                o |= r << (k * 8);
            }
            if (q == 0) ov.x = o;
            else if (q == 1) ov.y = o;
            else if (q == 2) ov.z = o;
            else ov.w = o;
        }
        *reinterpret_cast<uint4*>(op + i) = ov;
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
        for (uint32_t q = 0; q < 16; ++q) so[q] = W[q];
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
    __shared__ int s_lp[MTF_WARPS][256];
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
            // a byte whose symbol occurs again later in the same word cannot be the last occurrence: BWT output is
            // mostly runs, so this removes most of the atomics (and of their same-address serialisation)
            const uint32_t v  = *reinterpret_cast<const uint32_t*>(p + i);
            const uint32_t b0 = v & 0xFFu, b1 = (v >> 8) & 0xFFu, b2 = (v >> 16) & 0xFFu, b3 = v >> 24;
            atomicMax(&last[b3], (int) (i + 3));
            if (b2 != b3) atomicMax(&last[b2], (int) (i + 2));
            if (b1 != b2 && b1 != b3) atomicMax(&last[b1], (int) (i + 1));
            if (b0 != b1 && b0 != b2 && b0 != b3) atomicMax(&last[b0], (int) i);
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
    // the last occurrences of the PRESENT symbols, compacted (BWT output of text uses a few dozen of the 256): the rank of
    // a symbol is the number of those that are later than its own
    int*     lp  = s_lp[w];
    uint32_t cnt = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        const int      v    = last[r * 32 + lane];
        const uint32_t ball = __ballot_sync(BRA_FULL, v >= 0);
        if (v >= 0) lp[cnt + __popc(ball & lanemask_lt())] = v;
        cnt += __popc(ball);
    }
    __syncwarp();
    uint32_t rk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t s = 0; s < cnt; ++s)
    {
        const int ls = lp[s];  // broadcast read
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
    // the walk is a chain of dependent steps: the next segment's summary is fetched while the current one is applied
    const uint2* sm2   = reinterpret_cast<const uint2*>(summ + (uint64_t) b * segs * 256);
    uint2        nextv = nseg > 1 ? sm2[lane] : make_uint2(0, 0);
    uint32_t     nextk = (ENCODE && nseg > 1) ? scnt[(uint64_t) b * segs] : 0u;
    for (uint32_t s = 0; s < nseg; ++s)
    {
        uint8_t* st = state + ((uint64_t) b * segs + s) * 256;
        reinterpret_cast<uint2*>(st)[lane] = reinterpret_cast<const uint2*>(cur)[lane];
        if (s + 1 == nseg) break;
        const uint2    v = nextv;  // summary of segment s: bytes 8*lane .. 8*lane+7
        const uint32_t k = nextk;
        if (s + 2 < nseg)
        {
            nextv = sm2[(uint64_t) (s + 1) * 32 + lane];
            if (ENCODE) nextk = scnt[(uint64_t) b * segs + s + 1];
        }
        const uint32_t vw[2] = {v.x, v.y};
        if (ENCODE)
        {
            // new list = segment's recency list, then the old list minus those symbols (order kept)
            for (int i = lane; i < 256; i += 32) member[i] = 0;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const uint32_t i = lane * 8 + j;
                if (i < k)
                {
                    const uint8_t sy = (uint8_t) (vw[j >> 2] >> ((j & 3) * 8));
                    nxt[i]           = sy;
                    member[sy]       = 1;
                }
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
#pragma unroll
            for (int j = 0; j < 8; ++j) nxt[lane * 8 + j] = cur[(vw[j >> 2] >> ((j & 3) * 8)) & 0xFFu];
        }
        __syncwarp();
        reinterpret_cast<uint2*>(cur)[lane] = reinterpret_cast<const uint2*>(nxt)[lane];
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
