// crc32c.cu -- CRC-32C of every block of a batch (replaces reference bra_crc32c,
// src/utils/lib_bra_crc32c.c:102-179, for whole blocks).
//
// CRC is linear over GF(2): crc_raw(A||B) = crc_raw(A)*x^(8|B|) + crc_raw(B) (mod P), where
// crc_raw is the register run from 0 with no final inversion. Each thread runs a plain
// slicing-by-4 table CRC over its own 64 contiguous bytes, multiplies the result by the
// x^(8*bytes-that-follow-in-the-tile) constant ("fold"), the CTA XOR-reduces through warp
// shuffles, and thread 0 folds the tile value to the end of the block and XORs it into the
// block's accumulator. XOR is associative and commutative, so the atomic order is irrelevant and
// the result is deterministic. The init/final inversions are applied afterwards by
// crc_finalize_kernel: crc = ~(x^(8n)*~prev + raw).
//
// Roofline: reads n bytes once (HBM/L2 bound); tables live in shared memory.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"

#include <mutex>

namespace bra {

__device__ uint32_t       g_crc_tab[4][256];   // slicing-by-4 tables
__device__ uint32_t       g_crc_seg_pow[256];  // x^(8*64*j) mod P
__device__ bra_gf_pow_t   g_gf_pow;
static bra_gf_pow_t       h_gf_pow;
static std::once_flag     h_crc_once;

const bra_gf_pow_t* crc_host_pow()
{
    std::call_once(h_crc_once, [] { bra_gf_init_pow(&h_gf_pow); });  // reached from any thread (bra_crc32c_combine, the host paths)
    return &h_gf_pow;
}

// Uploads the tables to the CURRENT device once (module globals exist per device). Thread-safe: contexts on several
// GPUs may be created from several host threads.
bool crc_init_tables()
{
    static std::mutex mu;
    static bool       ready[64] = {false};
    static uint32_t   tab[4][256];
    static uint32_t   seg[256];
    static bool       host_ready = false;
    int               dev = 0;
    BRA_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && ready[dev]) return true;
    const bra_gf_pow_t* pw = crc_host_pow();
    if (!host_ready)
    {
        for (uint32_t b = 0; b < 256; ++b)
        {
            uint32_t r = b;
            for (int k = 0; k < 8; ++k) r = (r & 1u) ? (r >> 1) ^ BRA_CRC_POLY : (r >> 1);
            tab[0][b] = r;
        }
        for (int k = 1; k < 4; ++k)
            for (uint32_t b = 0; b < 256; ++b) tab[k][b] = tab[0][tab[k - 1][b] & 0xFFu] ^ (tab[k - 1][b] >> 8);
        for (uint32_t j = 0; j < 256; ++j) seg[j] = bra_gf_xpow8(pw, 64ull * j);
        host_ready = true;
    }
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_crc_tab, tab, sizeof(tab)));
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_crc_seg_pow, seg, sizeof(seg)));
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_gf_pow, pw, sizeof(bra_gf_pow_t)));
    if (dev >= 0 && dev < 64) ready[dev] = true;
    return true;
}

#define CRC_TILE 16384  // 256 threads x 64 bytes

__device__ __forceinline__ uint32_t crc_word(const uint32_t (*tab)[256], uint32_t c, uint32_t w)
{
    const uint32_t v = c ^ w;
    return tab[3][v & 0xFFu] ^ tab[2][(v >> 8) & 0xFFu] ^ tab[1][(v >> 16) & 0xFFu] ^ tab[0][v >> 24];
}

// acc[b] ^= raw CRC of block b (acc must be zeroed first).
// in: base pointer, block b at in + b*stride (stride multiple of 16), length len[b] (or fixed_len if len == nullptr)
__global__ void __launch_bounds__(256) crc_raw_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len,
                                                      uint32_t fixed_len, uint32_t* __restrict__ acc)
{
    __shared__ uint32_t tab[4][256];
    __shared__ uint32_t red[8];
    const uint32_t      b = blockIdx.y;
    const uint32_t      n = len ? len[b] : fixed_len;
    const uint64_t      tile0 = (uint64_t) blockIdx.x * CRC_TILE;
    if (tile0 >= n) return;
    for (int i = threadIdx.x; i < 1024; i += 256) (&tab[0][0])[i] = (&g_crc_tab[0][0])[i];
    __syncthreads();

    const uint32_t tile_len = (uint32_t) min((uint64_t) CRC_TILE, (uint64_t) n - tile0);
    const uint8_t* p        = in + (uint64_t) b * stride + tile0 + (uint64_t) threadIdx.x * 64;
    const uint32_t seg0     = threadIdx.x * 64;
    uint32_t       c        = 0;
    if (seg0 + 64 <= tile_len)
    {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4        v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = ld_stream_u4(q + i);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            c = crc_word(tab, c, v[i].x);
            c = crc_word(tab, c, v[i].y);
            c = crc_word(tab, c, v[i].z);
            c = crc_word(tab, c, v[i].w);
        }
        const uint32_t after = tile_len - (seg0 + 64);
        // full tiles hit the precomputed table; the ragged last tile computes its power
        const uint32_t mul = (tile_len == CRC_TILE) ? g_crc_seg_pow[255 - threadIdx.x] : bra_gf_xpow8(&g_gf_pow, after);
        c = bra_gf_mul(c, mul);
    }
    else if (seg0 < tile_len)
    {
        const uint32_t m = tile_len - seg0;  // last, partial segment of the block: nothing follows it in the tile
        for (uint32_t i = 0; i < m; ++i) c = tab[0][(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c ^= __shfl_xor_sync(BRA_FULL, c, d);
    if (lane_id() == 0) red[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t ^= red[i];
        const uint64_t after_tile = (uint64_t) n - (tile0 + tile_len);
        if (after_tile) t = bra_gf_mul(t, bra_gf_xpow8(&g_gf_pow, after_tile));
        atomicXor(&acc[b], t);
    }
}

// crc[b] = ~( x^(8 n) * ~prev[b]  +  raw[b] )
__global__ void crc_finalize_kernel(uint32_t* __restrict__ acc, const uint32_t* __restrict__ len, uint32_t fixed_len,
                                    const uint32_t* __restrict__ prev, uint32_t nblk)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint32_t n = len ? len[b] : fixed_len;
    const uint32_t p = prev ? prev[b] : 0u;
    acc[b]           = ~(bra_gf_mul(~p, bra_gf_xpow8(&g_gf_pow, n)) ^ acc[b]);
}

// CRC of the 268-byte in-memory chunk header of every block (reference chunks.c:248), prev = 0.
__global__ void crc_small_kernel(const uint8_t* __restrict__ data, uint32_t item_bytes, uint32_t* __restrict__ out, uint32_t nitems)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const uint8_t* p = data + (uint64_t) i * item_bytes;
    uint32_t       c = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < item_bytes; ++k) c = g_crc_tab[0][(c ^ p[k]) & 0xFFu] ^ (c >> 8);
    out[i] = ~c;
}

bool crc_blocks(const uint8_t* d_in, uint64_t stride, const uint32_t* d_len, uint32_t fixed_len, uint32_t max_len, uint32_t nblk,
                const uint32_t* d_prev, uint32_t* d_crc, cudaStream_t st)
{
    if (nblk == 0) return true;
    BRA_CUDA_TRY(cudaMemsetAsync(d_crc, 0, sizeof(uint32_t) * nblk, st));
    const uint32_t tiles = bra_div_up(max_len, CRC_TILE);
    if (tiles) BRA_LAUNCH(P_CRC, st, crc_raw_kernel<<<dim3(tiles, nblk), 256, 0, st>>>(d_in, stride, d_len, fixed_len, d_crc));
    BRA_LAUNCH(P_CRC, st, crc_finalize_kernel<<<bra_div_up(nblk, 128), 128, 0, st>>>(d_crc, d_len, fixed_len, d_prev, nblk));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

bool crc_headers(const uint8_t* d_hdr, uint32_t item_bytes, uint32_t nitems, uint32_t* d_out, cudaStream_t st)
{
    if (nitems == 0) return true;
    BRA_LAUNCH(P_CRC, st, crc_small_kernel<<<bra_div_up(nitems, 128), 128, 0, st>>>(d_hdr, item_bytes, d_out, nitems));
    BRA_CUDA_TRY(cudaGetLastError());
    return true;
}

}  // namespace bra
