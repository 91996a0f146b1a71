// crc32c.cu -- CRC-32C of every block of a batch (replaces reference bra_crc32c,
// src/utils/lib_bra_crc32c.c:102-179, for whole blocks).
//
// CRC is linear over GF(2): crc_raw(A||B) = crc_raw(A)*x^(8|B|) + crc_raw(B) (mod P), where
// crc_raw is the register run from 0 with no final inversion. Each thread runs a plain
// slicing-by-4 table CRC over its own 512 contiguous bytes, multiplies the result by the
// x^(8*bytes-that-follow-in-the-tile) constant ("fold"), the CTA XOR-reduces through warp
// shuffles, and thread 0 folds the tile value to the end of the block and XORs it into the
// block's accumulator. XOR is associative and commutative, so the atomic order is irrelevant and
// the result is deterministic. The init/final inversions are applied afterwards by
// crc_finalize_kernel: crc = ~(x^(8n)*~prev + raw).
//
// Roofline: reads n bytes once. One shared-memory lookup per byte is the real limit (32 lookups per clock and SM,
// conflict-free thanks to the per-lane table copies): about 9 TB/s of lookups against 6.5 TB/s of HBM.
#include "bra_common.cuh"
#include "bra_hd.h"
#include "bra_kernels.h"

#include <algorithm>
#include <mutex>

namespace bra {

__device__ uint32_t       g_crc_tab[4][256];   // slicing-by-4 tables
__device__ uint32_t       g_crc_seg_pow[512];  // x^(8*512*j) mod P: the bytes of a full tile that follow thread 511-j's segment
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
    static uint32_t   seg[512];
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
        for (uint32_t j = 0; j < 512; ++j) seg[j] = bra_gf_xpow8(pw, 512ull * j);
        host_ready = true;
    }
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_crc_tab, tab, sizeof(tab)));
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_crc_seg_pow, seg, sizeof(seg)));
    BRA_CUDA_TRY(cudaMemcpyToSymbol(g_gf_pow, pw, sizeof(bra_gf_pow_t)));
    if (dev >= 0 && dev < 64) ready[dev] = true;
    return true;
}

#define CRC_THREADS 512
#define CRC_SEG 512                          // contiguous bytes per thread
#define CRC_TILE (CRC_THREADS * CRC_SEG)     // 256 KiB per CTA step

// The four slicing tables are kept in shared memory once per LANE (entry e of lane l at word e*32 + l): every lane
// reads its own bank, so the 32 unrelated lookups of a warp are one conflict-free wavefront instead of 3-4 serialised
// ones. 128 KiB per CTA, one CTA per SM, filled once: the CTAs are persistent and walk over (block, tile) pairs.
struct CrcTables
{
    uint32_t t[4][256][32];
};

__device__ __forceinline__ uint32_t crc_word(const CrcTables& T, uint32_t l, uint32_t c, uint32_t w)
{
    const uint32_t v = c ^ w;
    return T.t[3][v & 0xFFu][l] ^ T.t[2][(v >> 8) & 0xFFu][l] ^ T.t[1][(v >> 16) & 0xFFu][l] ^ T.t[0][v >> 24][l];
}

// acc[b] ^= raw CRC of block b (acc must be zeroed first).
// in: base pointer, block b at in + b*stride, length len[b] (or fixed_len if len == nullptr)
__global__ void __launch_bounds__(CRC_THREADS, 1) crc_raw_kernel(const uint8_t* __restrict__ in, uint64_t stride, const uint32_t* __restrict__ len,
                                                                 uint32_t fixed_len, uint32_t tiles, uint32_t nblk, uint32_t* __restrict__ acc)
{
    extern __shared__ __align__(16) uint8_t crc_smem_raw[];
    CrcTables&          T = *reinterpret_cast<CrcTables*>(crc_smem_raw);
    __shared__ uint32_t red[CRC_THREADS / 32];
    const uint32_t      l = lane_id();
    for (uint32_t i = threadIdx.x; i < 4 * 256 * 32; i += CRC_THREADS) (&T.t[0][0][0])[i] = (&g_crc_tab[0][0])[i >> 5];
    __syncthreads();

    for (uint64_t work = blockIdx.x; work < (uint64_t) tiles * nblk; work += gridDim.x)
    {
        const uint32_t b = (uint32_t) (work / tiles), t = (uint32_t) (work % tiles);
        const uint32_t n = len ? len[b] : fixed_len;
        const uint64_t tile0 = (uint64_t) t * CRC_TILE;
        if (tile0 >= n) continue;  // uniform over the CTA
        const uint32_t tile_len = (uint32_t) min((uint64_t) CRC_TILE, (uint64_t) n - tile0);
        const uint32_t seg0     = threadIdx.x * CRC_SEG;
        const uint8_t* p        = in + (uint64_t) b * stride + tile0 + seg0;
        uint32_t       c        = 0;
        if (seg0 < tile_len)
        {
            const uint32_t m = min((uint32_t) CRC_SEG, tile_len - seg0);
            uint32_t       i = 0;
            if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)
            {
                const uint4* q = reinterpret_cast<const uint4*>(p);
                for (; i + 64 <= m; i += 64)
                {
                    uint4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __ldg(q + (i >> 4) + j);  // (allocating in L1: the four loads share two 32-byte sectors)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        c = crc_word(T, l, c, v[j].x);
                        c = crc_word(T, l, c, v[j].y);
                        c = crc_word(T, l, c, v[j].z);
                        c = crc_word(T, l, c, v[j].w);
                    }
                }
            }
            for (; i < m; ++i) c = T.t[0][(c ^ p[i]) & 0xFFu][l] ^ (c >> 8);  // unaligned block base, ragged end
            const uint32_t after = tile_len - (seg0 + m);
            // full tiles hit the precomputed table; the ragged last tile computes its power
            const uint32_t mul = (tile_len == CRC_TILE) ? g_crc_seg_pow[CRC_THREADS - 1 - threadIdx.x] : bra_gf_xpow8(&g_gf_pow, after);
            c = bra_gf_mul(c, mul);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c ^= __shfl_xor_sync(BRA_FULL, c, d);
        if (l == 0) red[warp_id()] = c;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            uint32_t x = 0;
            for (int i = 0; i < CRC_THREADS / 32; ++i) x ^= red[i];
            const uint64_t after_tile = (uint64_t) n - (tile0 + tile_len);
            if (after_tile) x = bra_gf_mul(x, bra_gf_xpow8(&g_gf_pow, after_tile));
            atomicXor(&acc[b], x);
        }
        __syncthreads();
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
    if (tiles)
    {
        int dev = 0, sms = 0;
        BRA_CUDA_TRY(cudaGetDevice(&dev));
        BRA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BRA_CUDA_TRY(cudaFuncSetAttribute(crc_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(CrcTables)));
        const uint32_t grid = (uint32_t) std::min<uint64_t>((uint64_t) tiles * nblk, (uint64_t) sms);
        BRA_LAUNCH(P_CRC, st, crc_raw_kernel<<<grid, CRC_THREADS, sizeof(CrcTables), st>>>(d_in, stride, d_len, fixed_len, tiles, nblk, d_crc));
    }
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
