// rle.cu -- PackBits-style run-length coding for a batch of blocks
// (replaces reference bra_rle_encode / bra_rle_decode / bra_rle_decode_compute_size,
//  src/encoders/bra_rle.c:60-120, :162-224, :122-160; constants src/lib_bra_defs.h:95-98).
//
// ENCODE. The reference parses greedily, byte by byte. Its output has a closed form over the
// maximal runs of the input (validated against the reference in tests/test_host_logic.py):
//   * a run of length L >= 3 starting at s emits, for every k = 0,128,256,.. below L - rem,
//     the token {1 - min(128, L - rem - k), value}; rem = L mod 128 if that is 1 or 2, else 0;
//   * every other byte (runs of 1-2 bytes and the 1-2 leftover bytes of a long run) is a
//     literal; maximal stretches of consecutive literals are cut from their start into groups
//     of 128, each prefixed by {group_len - 1}.
// So every byte knows what it emits once it knows (run start, run end, stretch start, stretch
// end): four max/min scans of head positions plus a sum scan for the output offset. Three
// streaming kernels per block, each recomputing the cheap per-byte state from the input and
// taking its cross-tile carries from the tile summaries the previous kernel wrote:
//   heads -> literal flags -> sizes + emit (+ the 256-bin histogram Huffman needs; output offsets by look-back).
//
// DECODE. Token boundaries depend on all previous tokens. Per 1 KiB tile the map "entry offset
// -> exit offset" is built for every possible entry (a token overhangs by at most 128 bytes):
// one warp walks the token path from entry 0, the paths from the other entries are followed until
// they merge with it; one warp per block chains the tiles; tiles then mark their true token
// starts by the same warp walk, count, and expand with a binary search per output byte.
#include "bra_common.cuh"
#include "bra_kernels.h"

#include <limits.h>

namespace bra {

#define RL_TILE 4096
#define RL_THREADS 256
#define RL_INF 0x7FFFFFFF

// ---- shared per-tile machinery ------------------------------------------------------------------
struct RleTile
{
    uint8_t  x[16];
    int      s[16];  // start of the maximal run containing the byte
    int      e[16];  // end (exclusive) of that run
    uint32_t m;      // valid bytes of this thread
    int      j0;     // block-relative index of x[0]
};

// reduce tile summaries of other tiles: max over tiles < t of a[], min over tiles > t of c[]
__device__ __forceinline__ void rle_carries(const int* __restrict__ last_arr, const int* __restrict__ first_arr, uint32_t t, uint32_t ntiles,
                                            int none_before, int none_after, int* red, int& before, int& after)
{
    int mx = none_before, mn = none_after;
    for (uint32_t i = threadIdx.x; i < ntiles; i += RL_THREADS)
    {
        if (i < t) mx = max(mx, last_arr[i]);
        if (i > t) mn = min(mn, first_arr[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(BRA_FULL, mx, d));
        mn = min(mn, __shfl_xor_sync(BRA_FULL, mn, d));
    }
    __syncthreads();
    if (lane_id() == 0)
    {
        red[warp_id()]     = mx;
        red[8 + warp_id()] = mn;
    }
    __syncthreads();
    before = red[0];
    after  = red[8];
    for (int i = 1; i < 8; ++i)
    {
        before = max(before, red[i]);
        after  = min(after, red[8 + i]);
    }
    __syncthreads();
}

__device__ __forceinline__ void rle_load(RleTile& T, const uint8_t* __restrict__ xb, uint32_t n, uint32_t tile0, uint32_t& headmask)
{
    T.j0 = (int) (tile0 + threadIdx.x * 16);
    T.m  = (uint32_t) T.j0 < n ? min(16u, n - T.j0) : 0u;
    uint8_t prev = 0;
    if (T.m == 16)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(xb + T.j0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) T.x[i] = (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 16; ++i) T.x[i] = (uint32_t) i < T.m ? xb[T.j0 + i] : 0;
    }
    if (T.m && T.j0 > 0) prev = xb[T.j0 - 1];
    headmask = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        if ((uint32_t) i < T.m && ((T.j0 + i) == 0 || T.x[i] != prev)) headmask |= 1u << i;
        prev = T.x[i];
    }
}

// fill s[], e[] given the carries (last head before the tile, first head after the tile or n)
__device__ __forceinline__ void rle_runs(RleTile& T, uint32_t headmask, int head_before, int head_after, int* red)
{
    int mylast = -1, myfirst = RL_INF;
    if (headmask)
    {
        mylast  = T.j0 + (31 - __clz(headmask));
        myfirst = T.j0 + (__ffs(headmask) - 1);
    }
    int run_s = max(block_excl_max(mylast, -1, red), head_before);
    int nxt   = min(block_excl_min_rev(myfirst, RL_INF, red), head_after);
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        if (headmask & (1u << i)) run_s = T.j0 + i;
        T.s[i] = run_s;
    }
#pragma unroll
    for (int i = 15; i >= 0; --i)
    {
        T.e[i] = nxt;
        if (headmask & (1u << i)) nxt = T.j0 + i;
    }
}

__device__ __forceinline__ uint32_t rle_litmask(const RleTile& T)
{
    uint32_t lit = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        if ((uint32_t) i >= T.m) break;
        const int L = T.e[i] - T.s[i], k = T.j0 + i - T.s[i], rho = L & 127;
        if (L < 3 || ((rho == 1 || rho == 2) && k >= L - rho)) lit |= 1u << i;
    }
    return lit;
}

// ---- pass 1: run heads per tile ------------------------------------------------------------------
__global__ void __launch_bounds__(RL_THREADS)
    rle_enc_heads_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t tiles,
                         int* __restrict__ t_first_head, int* __restrict__ t_last_head)
{
    __shared__ int red[16];
    const uint32_t b = blockIdx.y, t = blockIdx.x;
    const uint32_t n = len[b];
    const uint32_t tile0 = t * RL_TILE;
    if (tile0 >= n) return;
    RleTile  T;
    uint32_t hm;
    rle_load(T, in + (uint64_t) b * stride, n, tile0, hm);
    int mx = hm ? T.j0 + (31 - __clz(hm)) : -1;
    int mn = hm ? T.j0 + (__ffs(hm) - 1) : RL_INF;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(BRA_FULL, mx, d));
        mn = min(mn, __shfl_xor_sync(BRA_FULL, mn, d));
    }
    if (lane_id() == 0)
    {
        red[warp_id()]     = mx;
        red[8 + warp_id()] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; ++i)
        {
            mx = max(mx, red[i]);
            mn = min(mn, red[8 + i]);
        }
        t_last_head[(uint64_t) b * tiles + t]  = mx;
        t_first_head[(uint64_t) b * tiles + t] = mn;
    }
}

// ---- pass 2: literal flags -> first/last non-literal byte per tile ---------------------------------
__global__ void __launch_bounds__(RL_THREADS)
    rle_enc_lit_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t tiles,
                       const int* __restrict__ t_first_head, const int* __restrict__ t_last_head, int* __restrict__ t_first_nl,
                       int* __restrict__ t_last_nl)
{
    __shared__ int red[34];
    const uint32_t b = blockIdx.y, t = blockIdx.x;
    const uint32_t n = len[b];
    const uint32_t tile0 = t * RL_TILE;
    if (tile0 >= n) return;
    const uint32_t ntiles = (n + RL_TILE - 1) / RL_TILE;
    int            hb, ha;
    rle_carries(t_last_head + (uint64_t) b * tiles, t_first_head + (uint64_t) b * tiles, t, ntiles, -1, (int) n, red, hb, ha);
    ha = min(ha, (int) n);
    RleTile  T;
    uint32_t hm;
    rle_load(T, in + (uint64_t) b * stride, n, tile0, hm);
    rle_runs(T, hm, hb, ha, red);
    const uint32_t lit = rle_litmask(T);
    const uint32_t nl  = ~lit & (T.m == 16 ? 0xFFFFu : ((1u << T.m) - 1u));
    int mx = nl ? T.j0 + (31 - __clz(nl)) : -1;
    int mn = nl ? T.j0 + (__ffs(nl) - 1) : RL_INF;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(BRA_FULL, mx, d));
        mn = min(mn, __shfl_xor_sync(BRA_FULL, mn, d));
    }
    __syncthreads();
    if (lane_id() == 0)
    {
        red[warp_id()]     = mx;
        red[8 + warp_id()] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; ++i)
        {
            mx = max(mx, red[i]);
            mn = min(mn, red[8 + i]);
        }
        t_last_nl[(uint64_t) b * tiles + t]  = mx;
        t_first_nl[(uint64_t) b * tiles + t] = mn;
    }
}

// ---- pass 3: sizes, output offsets and emit in one kernel --------------------------------------------
// The output offset of a tile is the sum of the sizes of the tiles before it: every tile publishes its size in a status
// word and adds up its predecessors' by decoupled look-back (one warp, 32 predecessors per trip). Tiles are handed out
// by ticket, a group of blocks interleaved, so a tile only ever waits for tiles that are running or done, and its
// predecessors have usually published their inclusive sums already. Also writes r_len and the histogram Huffman needs.
// status word: [31:30] 0 = empty, 1 = tile size, 2 = inclusive sum; [29:0] value
#define RL_GROUP 16u
#ifndef RL_EMIT_CTAS
#define RL_EMIT_CTAS 4
#endif
__device__ __forceinline__ uint32_t rl_ld_status(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rl_st_status(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(RL_THREADS, RL_EMIT_CTAS)
    rle_enc_out_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len, uint32_t tiles, uint32_t nblk,
                       const int* __restrict__ t_first_head, const int* __restrict__ t_last_head, const int* __restrict__ t_first_nl,
                       const int* __restrict__ t_last_nl, uint32_t* __restrict__ status /* [nblk][tiles], then the ticket counter; zeroed */,
                       uint8_t* __restrict__ out, uint64_t out_stride, uint32_t* __restrict__ r_len, uint32_t* __restrict__ hist)
{
    __shared__ int      red[34];
    __shared__ uint32_t ured[34];
    __shared__ uint8_t  stage[RL_TILE + 64];
    __shared__ uint32_t shist[256];
    __shared__ uint32_t s_ticket, s_out0;
    if (threadIdx.x == 0) s_ticket = atomicAdd(status + (uint64_t) nblk * tiles, 1u);
    __syncthreads();
    const uint32_t per_group = RL_GROUP * tiles;
    const uint32_t b = (s_ticket / per_group) * RL_GROUP + (s_ticket % per_group) % RL_GROUP, t = (s_ticket % per_group) / RL_GROUP;
    if (b >= nblk) return;
    const uint32_t n = len[b];
    const uint32_t tile0 = t * RL_TILE;
    if (tile0 >= n) return;
    const uint32_t ntiles = (n + RL_TILE - 1) / RL_TILE;
    int            hb, ha, nlb, nla;
    rle_carries(t_last_head + (uint64_t) b * tiles, t_first_head + (uint64_t) b * tiles, t, ntiles, -1, (int) n, red, hb, ha);
    rle_carries(t_last_nl + (uint64_t) b * tiles, t_first_nl + (uint64_t) b * tiles, t, ntiles, -1, (int) n, red, nlb, nla);
    ha  = min(ha, (int) n);
    nla = min(nla, (int) n);
    RleTile  T;
    uint32_t hm;
    rle_load(T, in + (uint64_t) b * stride, n, tile0, hm);
    rle_runs(T, hm, hb, ha, red);
    const uint32_t lit   = rle_litmask(T);
    const uint32_t valid = T.m == 16 ? 0xFFFFu : ((1u << T.m) - 1u);
    const uint32_t nl    = ~lit & valid;

    // stretch start (last non-literal before the byte, +1) and stretch end (first non-literal after it)
    int mylast = nl ? T.j0 + (31 - __clz(nl)) : -1;
    int myfirst = nl ? T.j0 + (__ffs(nl) - 1) : RL_INF;
    int last_nl = max(block_excl_max(mylast, -1, red), nlb);
    int next_nl = min(block_excl_min_rev(myfirst, RL_INF, red), nla);
    int ss[16], se[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        if (nl & (1u << i)) last_nl = T.j0 + i;
        ss[i] = last_nl + 1;
    }
#pragma unroll
    for (int i = 15; i >= 0; --i)
    {
        se[i] = next_nl;
        if (nl & (1u << i)) next_nl = T.j0 + i;
    }

    uint32_t contrib[16], mysum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        uint32_t c = 0;
        if ((uint32_t) i < T.m)
        {
            if (lit & (1u << i))
                c = 1u + (((T.j0 + i - ss[i]) & 127) == 0);
            else
                c = (((T.j0 + i - T.s[i]) & 127) == 0) ? 2u : 0u;
        }
        contrib[i] = c;
        mysum += c;
    }
    uint32_t tile_total;
    uint32_t off = block_excl_add(mysum, ured, &tile_total);
    {
        // output offset of the tile = sum of the sizes of the tiles before it (look-back by warp 0)
        if (warp_id() == 0)
        {
            uint32_t* const st = status + (uint64_t) b * tiles;
            const uint32_t  l  = lane_id();
            if (l == 0) rl_st_status(st + t, ((t == 0 ? 2u : 1u) << 30) | tile_total);
            uint32_t excl = 0;
            if (t != 0)
            {
                int tt = (int) t - 1;
                for (uint32_t trips = 0;; ++trips)
                {
                    if (trips > (1u << 24)) __trap();  // seconds of waiting: a predecessor never published -- fail loudly instead of hanging
                    const uint32_t v     = tt - (int) l >= 0 ? rl_ld_status(st + (tt - (int) l)) : (2u << 30);  // before tile 0: inclusive sum 0
                    const uint32_t incl  = __ballot_sync(BRA_FULL, (v >> 30) == 2u);
                    const uint32_t empty = __ballot_sync(BRA_FULL, (v >> 30) == 0u);
                    const uint32_t upto  = incl ? (uint32_t) (__ffs(incl) - 1) : 31u;  // lanes 0..upto are summed
                    const uint32_t need  = upto == 31u ? 0xFFFFFFFFu : ((2u << upto) - 1u);
                    if (empty & need)
                    {
                        __nanosleep(100);
                        continue;  // a tile in the window has not published yet
                    }
                    uint32_t x = l <= upto ? (v & 0x3FFFFFFFu) : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(BRA_FULL, x, d);
                    excl += x;
                    if (incl) break;
                    tt -= 32;
                }
                if (l == 0) rl_st_status(st + t, (2u << 30) | (excl + tile_total));
            }
            if (l == 0) s_out0 = excl;
        }
        shist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t tile_out0 = s_out0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            if ((uint32_t) i >= T.m) break;
            if (lit & (1u << i))
            {
                const int ks = T.j0 + i - ss[i];
                if ((ks & 127) == 0)
                {
                    stage[off]     = (uint8_t) (min(128, se[i] - ss[i] - ks) - 1);
                    stage[off + 1] = T.x[i];
                }
                else
                    stage[off] = T.x[i];
            }
            else
            {
                const int k = T.j0 + i - T.s[i];
                if ((k & 127) == 0)
                {
                    const int L = T.e[i] - T.s[i], rho = L & 127;
                    const int tok = min(128, L - (rho < 3 ? rho : 0) - k);
                    stage[off]     = (uint8_t) (1 - tok);
                    stage[off + 1] = T.x[i];
                }
            }
            off += contrib[i];
        }
        __syncthreads();
        uint8_t* o = out + (uint64_t) b * out_stride + tile_out0;
        for (uint32_t i = threadIdx.x; i < tile_total; i += RL_THREADS)
        {
            const uint8_t v = stage[i];
            o[i]            = v;
            atomicAdd(&shist[v], 1u);
        }
        __syncthreads();
        if (shist[threadIdx.x]) atomicAdd(&hist[(uint64_t) b * 256 + threadIdx.x], shist[threadIdx.x]);
        if (t + 1 == ntiles && threadIdx.x == 0) r_len[b] = tile_out0 + tile_total;
    }
}

bool rle_encode_batch(const RleEncArgs& a, cudaStream_t st)
{
    if (a.nblk == 0 || a.max_n == 0) return true;
    const uint32_t tiles = bra_div_up(a.max_n, RL_TILE);
    const dim3     grid(tiles, a.nblk);
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_hist, 0, (size_t) a.nblk * 256 * 4, st));
    BRA_LAUNCH(P_RLE_ENC_HEADS, st, rle_enc_heads_kernel<<<grid, RL_THREADS, 0, st>>>(a.d_in, a.stride, a.d_len, tiles, a.d_t_first_head, a.d_t_last_head));
    BRA_LAUNCH(P_RLE_ENC_LIT, st, rle_enc_lit_kernel<<<grid, RL_THREADS, 0, st>>>(a.d_in, a.stride, a.d_len, tiles, a.d_t_first_head, a.d_t_last_head, a.d_t_first_nl,
                                                    a.d_t_last_nl));
    BRA_CUDA_TRY(cudaMemsetAsync(a.d_t_cnt, 0, ((size_t) a.nblk * tiles + 1) * sizeof(uint32_t), st));
    const uint32_t ctas = bra_div_up(a.nblk, RL_GROUP) * RL_GROUP * tiles;
    BRA_LAUNCH(P_RLE_ENC_EMIT, st, rle_enc_out_kernel<<<ctas, RL_THREADS, 0, st>>>(a.d_in, a.stride, a.d_len, tiles, a.nblk, a.d_t_first_head, a.d_t_last_head,
                                                    a.d_t_first_nl, a.d_t_last_nl, a.d_t_cnt, a.d_out, a.out_stride, a.d_rlen, a.d_hist));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

// ================================================================================================
// DECODE
// ================================================================================================
#define RD_TILE 1024
#define RD_THREADS 256
#define RD_ENTRIES 130  // a token starting in the previous tile can overhang by 0..129 bytes... exit offsets 0..128

__device__ __forceinline__ uint32_t rle_tok_len(uint8_t c)  // bytes of input the token occupies
{
    const int8_t s = (int8_t) c;
    return s >= 0 ? (uint32_t) s + 2u : (s == -128 ? 1u : 2u);
}
__device__ __forceinline__ uint32_t rle_tok_out(uint8_t c)  // bytes of output it produces
{
    const int8_t s = (int8_t) c;
    return s >= 0 ? (uint32_t) s + 1u : (s == -128 ? 0u : (uint32_t) (1 - s));
}

// The token path through a tile, one WARP per tile. The lanes look at 32 consecutive positions at a time (their control
// bytes and token lengths); the path inside the window is then followed by shuffles -- position -> position + length of
// the token there -- which costs a couple of instructions per token for the whole warp. `pos` is uniform over the warp.
#define RD_WARPS 8  // tiles per CTA
__device__ __forceinline__ void rle_dec_stage_tile(const uint8_t* __restrict__ xb, uint32_t tile0, uint32_t tile_n, uint8_t* sx)
{
    const uint32_t l = lane_id();
    if (tile_n == RD_TILE && ((reinterpret_cast<uintptr_t>(xb) + tile0) & 15u) == 0)
    {
        const uint4* q = reinterpret_cast<const uint4*>(xb + tile0);
        reinterpret_cast<uint4*>(sx)[l]      = q[l];
        reinterpret_cast<uint4*>(sx)[l + 32] = q[l + 32];
    }
    else
        for (uint32_t i = l; i < RD_TILE; i += 32) sx[i] = i < tile_n ? xb[tile0 + i] : 0;
    __syncwarp();
}

// Exit offsets of a tile for every possible entry (a token starting in the previous tile overhangs by 0..129 bytes): the
// path from entry 0 is walked and marked; the path from any other entry is followed only until it meets a marked
// position -- from there on the two paths are one -- or leaves the tile. RLE streams re-synchronise within a few tokens,
// so the other 129 walks are short (and never longer than the tile has tokens).
__global__ void __launch_bounds__(RD_WARPS * 32)
    rle_dec_exit_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ rlen, uint32_t tiles,
                        uint8_t* __restrict__ t_exit /* [b][tile][RD_ENTRIES] */)
{
    __shared__ __align__(16) uint8_t s_x[RD_WARPS][RD_TILE];
    __shared__ uint32_t              s_mark[RD_WARPS][RD_TILE / 32];
    const uint32_t b = blockIdx.y, w = warp_id(), l = lane_id();
    const uint32_t t = blockIdx.x * RD_WARPS + w;
    const uint32_t r = rlen[b];
    const uint32_t tile0 = t * RD_TILE;
    if (tile0 >= r) return;  // whole warps leave; no CTA barrier below
    const uint32_t tile_n = min((uint32_t) RD_TILE, r - tile0);
    uint8_t*       sx     = s_x[w];
    rle_dec_stage_tile(in + (uint64_t) b * stride, tile0, tile_n, sx);
    uint32_t pos = 0;
    for (uint32_t win = 0; win < RD_TILE / 32; ++win)
    {
        const uint32_t wb  = win * 32;
        const uint32_t len = rle_tok_len(sx[wb + l]);
        uint32_t       m   = 0;
        while (pos < wb + 32 && pos < tile_n)
        {
            m |= 1u << (pos - wb);
            pos += __shfl_sync(BRA_FULL, len, pos - wb);
        }
        if (l == 0) s_mark[w][win] = m;
    }
    __syncwarp();
    const uint32_t exit0 = pos - tile_n;  // pos >= tile_n here
    uint8_t*       ex    = t_exit + ((uint64_t) b * tiles + t) * RD_ENTRIES;
    for (uint32_t e = l; e < RD_ENTRIES; e += 32)
    {
        uint32_t x = 0;  // entries beyond the tile (only possible in a short last tile) are never followed
        if (e < tile_n)
        {
            uint32_t p = e;
            for (;;)
            {
                if (p >= tile_n)
                {
                    x = p - tile_n;
                    break;
                }
                if ((s_mark[w][p >> 5] >> (p & 31u)) & 1u)
                {
                    x = exit0;
                    break;
                }
                p += rle_tok_len(sx[p]);
            }
        }
        ex[e] = (uint8_t) min(x, 255u);
    }
}

// Entry offset of every tile: e(t+1) = exit_t[e(t)], a chain over up to a few thousand tiles per block. One warp per block
// cuts it into 32 pieces: every lane first composes the exit maps of its piece for all RD_ENTRIES possible entries (the
// chains of one lane are independent loads), lane 0 then links the 32 composed maps, and every lane replays its piece
// from its now known entry -- three short dependent stretches instead of one long one.
__global__ void __launch_bounds__(32) rle_dec_chain_kernel(const uint32_t* __restrict__ rlen, uint32_t tiles, const uint8_t* __restrict__ t_exit,
                                                           uint8_t* __restrict__ t_entry, uint32_t nblk)
{
    __shared__ uint8_t s_map[32][RD_ENTRIES + 2];
    __shared__ uint8_t s_entry[32];
    const uint32_t     b = blockIdx.x, l = threadIdx.x;
    if (b >= nblk) return;
    const uint32_t r      = rlen[b];
    const uint32_t ntiles = (r + RD_TILE - 1) / RD_TILE;
    const uint32_t per    = (ntiles + 31) / 32;
    const uint32_t t0 = min(ntiles, l * per), t1 = min(ntiles, t0 + per);
    const uint8_t* ex = t_exit + (uint64_t) b * tiles * RD_ENTRIES;
    for (uint32_t e0 = 0; e0 < RD_ENTRIES; e0 += 10)
    {
        uint32_t e[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) e[i] = min(e0 + i, (uint32_t) RD_ENTRIES - 1);
        for (uint32_t t = t0; t < t1; ++t)
#pragma unroll
            for (int i = 0; i < 10; ++i) e[i] = ex[(uint64_t) t * RD_ENTRIES + e[i]];
#pragma unroll
        for (int i = 0; i < 10; ++i)
            if (e0 + i < RD_ENTRIES) s_map[l][e0 + i] = (uint8_t) e[i];
    }
    __syncwarp();
    if (l == 0)
    {
        uint32_t e = 0;
        for (int c = 0; c < 32; ++c)
        {
            s_entry[c] = (uint8_t) e;
            e          = s_map[c][e];
        }
    }
    __syncwarp();
    uint32_t e = s_entry[l];
    for (uint32_t t = t0; t < t1; ++t)
    {
        t_entry[(uint64_t) b * tiles + t] = (uint8_t) e;
        e = ex[(uint64_t) t * RD_ENTRIES + e];
    }
}

// True token starts of a tile (bitmap) and its output size, from the tile's entry offset: the same warp walk.
__global__ void __launch_bounds__(RD_WARPS * 32)
    rle_dec_mark_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ rlen, uint32_t tiles,
                        const uint8_t* __restrict__ t_entry, uint32_t* __restrict__ t_tok /* [b][tile][32] */, uint32_t* __restrict__ t_ocnt,
                        uint32_t* __restrict__ err)
{
    __shared__ __align__(16) uint8_t s_x[RD_WARPS][RD_TILE];
    const uint32_t b = blockIdx.y, w = warp_id(), l = lane_id();
    const uint32_t t = blockIdx.x * RD_WARPS + w;
    const uint32_t r = rlen[b];
    const uint32_t tile0 = t * RD_TILE;
    if (tile0 >= r) return;  // whole warps leave; no CTA barrier below
    const uint32_t tile_n = min((uint32_t) RD_TILE, r - tile0);
    uint8_t*       sx     = s_x[w];
    rle_dec_stage_tile(in + (uint64_t) b * stride, tile0, tile_n, sx);
    uint32_t pos = t_entry[(uint64_t) b * tiles + t], total = 0;
    bool     bad = false;
    uint32_t* tok = t_tok + ((uint64_t) b * tiles + t) * 32;
    for (uint32_t win = 0; win < RD_TILE / 32; ++win)
    {
        const uint32_t wb  = win * 32;
        const uint8_t  c   = sx[wb + l];
        const uint32_t len = rle_tok_len(c), out = rle_tok_out(c);
        uint32_t       m   = 0;
        while (pos < wb + 32 && pos < tile_n)
        {
            const uint32_t k = pos - wb;
            m |= 1u << k;
            total += __shfl_sync(BRA_FULL, out, k);
            pos += __shfl_sync(BRA_FULL, len, k);
            bad |= (uint64_t) tile0 + pos > r;  // truncated token: the reference's error exits (bra_rle.c:136,148)
        }
        if (l == 0) tok[win] = m;
    }
    if (l == 0)
    {
        t_ocnt[(uint64_t) b * tiles + t] = total;
        if (bad) err[b] = 1;
    }
}

__global__ void __launch_bounds__(RD_THREADS)
    rle_dec_expand_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ rlen, uint32_t tiles,
                          const uint32_t* __restrict__ t_tok, const uint32_t* __restrict__ t_ocnt, const uint32_t* __restrict__ err,
                          uint8_t* __restrict__ out, uint64_t out_stride, uint32_t out_cap, uint32_t* __restrict__ n_len)
{
    __shared__ uint32_t tstart[RD_TILE / 1 + 1];  // output start of each token of the tile (compacted)
    __shared__ uint16_t tsrc[RD_TILE];            // input position (tile-relative) of each token
    __shared__ uint32_t ured[34];
    const uint32_t      b = blockIdx.y, t = blockIdx.x;
    const uint32_t      r = rlen[b];
    const uint32_t      tile0 = t * RD_TILE;
    if (tile0 >= r) return;
    const uint32_t ntiles = (r + RD_TILE - 1) / RD_TILE;
    const uint8_t* xb     = in + (uint64_t) b * stride;

    // offset of this tile's output; the last tile also publishes the decoded size (0 on error)
    uint32_t before = 0;
    for (uint32_t i = threadIdx.x; i < t; i += RD_THREADS) before += t_ocnt[(uint64_t) b * tiles + i];
    uint32_t out0;
    block_excl_add(before, ured, &out0);
    const uint32_t mine  = t_ocnt[(uint64_t) b * tiles + t];
    const bool     isbad = err[b] != 0 || (uint64_t) out0 + mine > out_cap;
    if (t + 1 == ntiles && threadIdx.x == 0) n_len[b] = isbad ? 0u : out0 + mine;
    if (isbad || out == nullptr) return;  // out == nullptr: size query only (bra_rle_decode_compute_size)

    // compact the tokens: 4 consecutive positions per thread
    const uint32_t i0   = threadIdx.x * 4;
    const uint32_t bits = (t_tok[((uint64_t) b * tiles + t) * 32 + (i0 >> 5)] >> (i0 & 31)) & 0xFu;
    uint32_t       olen[4], nt = 0, osum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        olen[k] = 0;
        if (bits & (1u << k))
        {
            olen[k] = rle_tok_out(xb[tile0 + i0 + k]);
            ++nt;
            osum += olen[k];
        }
    }
    uint32_t ntok, dummy;
    uint32_t ti = block_excl_add(nt, ured, &ntok);
    uint32_t oo = block_excl_add(osum, ured, &dummy);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (bits & (1u << k))
        {
            tstart[ti] = oo;
            tsrc[ti]   = (uint16_t) (i0 + k);
            ++ti;
            oo += olen[k];
        }
    if (threadIdx.x == 0) tstart[ntok] = mine;
    __syncthreads();

    // Four consecutive output bytes per thread: one binary search for the first, then the token list is walked
    // forward, and the four bytes leave as one aligned 32-bit store (groups are aligned to the output address).
    uint8_t*       ob  = out + (uint64_t) b * out_stride + out0;
    const uint32_t mis = (uint32_t) (reinterpret_cast<uintptr_t>(ob) & 3u);
    const uint32_t ngroups = (mine + mis + 3) / 4;
    for (uint32_t g = threadIdx.x; g < ngroups; g += RD_THREADS)
    {
        const uint32_t o_begin = g * 4 >= mis ? g * 4 - mis : 0u;
        const uint32_t o_end   = min(mine, g * 4 + 4 - mis);
        uint32_t lo = 0, hi = ntok;  // last token with tstart <= o_begin
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (tstart[mid] <= o_begin)
                lo = mid;
            else
                hi = mid;
        }
        uint32_t acc = 0;
        for (uint32_t o = o_begin; o < o_end; ++o)
        {
            while (tstart[lo + 1] <= o) ++lo;  // tstart[ntok] == mine > o ends the walk
            const uint32_t src = tile0 + tsrc[lo];
            const int8_t   c   = (int8_t) xb[src];
            const uint32_t v   = c >= 0 ? xb[src + 1 + (o - tstart[lo])] : xb[src + 1];
            acc |= v << ((o - o_begin) * 8u);
        }
        if (o_end - o_begin == 4)
            *reinterpret_cast<uint32_t*>(ob + o_begin) = acc;
        else
            for (uint32_t o = o_begin; o < o_end; ++o) ob[o] = (uint8_t) (acc >> ((o - o_begin) * 8u));
    }
}

bool rle_decode_batch(const RleDecArgs& a, cudaStream_t st)
{
    if (a.nblk == 0 || a.max_r == 0) return true;
    const uint32_t tiles = bra_div_up(a.max_r, RD_TILE);
    const dim3     grid(tiles, a.nblk);
    const dim3 wgrid(bra_div_up(tiles, RD_WARPS), a.nblk);  // one warp per tile
    BRA_LAUNCH(P_RLE_DEC_EXIT, st, rle_dec_exit_kernel<<<wgrid, RD_WARPS * 32, 0, st>>>(a.d_in, a.stride, a.d_rlen, tiles, a.d_t_exit));
    BRA_LAUNCH(P_RLE_DEC_CHAIN, st, rle_dec_chain_kernel<<<a.nblk, 32, 0, st>>>(a.d_rlen, tiles, a.d_t_exit, a.d_t_entry, a.nblk));
    BRA_LAUNCH(P_RLE_DEC_MARK, st, rle_dec_mark_kernel<<<wgrid, RD_WARPS * 32, 0, st>>>(a.d_in, a.stride, a.d_rlen, tiles, a.d_t_entry, a.d_t_tok, a.d_t_ocnt, a.d_err));
    if (a.size_only)
        BRA_LAUNCH(P_RLE_DEC_EXPAND, st, rle_dec_expand_kernel<<<grid, RD_THREADS, 0, st>>>(a.d_in, a.stride, a.d_rlen, tiles, a.d_t_tok, a.d_t_ocnt, a.d_err, nullptr, 0, 0xFFFFFFFFu,
                                                           a.d_nlen));
    else
        BRA_LAUNCH(P_RLE_DEC_EXPAND, st, rle_dec_expand_kernel<<<grid, RD_THREADS, 0, st>>>(a.d_in, a.stride, a.d_rlen, tiles, a.d_t_tok, a.d_t_ocnt, a.d_err, a.d_out,
                                                           a.out_stride, a.out_cap, a.d_nlen));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

uint32_t rle_enc_tiles(uint32_t max_n) { return bra_div_up(max_n, RL_TILE) + 1; }  // (+1: the ticket word behind the per-tile status)
uint32_t rle_dec_tiles(uint32_t max_r) { return bra_div_up(max_r, RD_TILE); }
uint32_t rle_dec_entries() { return RD_ENTRIES; }

}  // namespace bra
