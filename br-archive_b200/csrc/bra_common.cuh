// bra_common.cuh -- shared device/host helpers for the B200 block-compression kernels.
//
// Layout convention used by every batched kernel in this directory
// ----------------------------------------------------------------
// A *batch* is `nblk` independent blocks (the reference's 256 KiB "chunks",
// reference src/lib_bra_defs.h:93; here the size is a run-time parameter).
// Block b owns the slice [b*stride, b*stride + len[b]) of every per-byte array;
// `len` lives in device memory so that stages whose output size is data dependent
// (RLE, Huffman) can feed the next stage without a host round trip.
// Kernels are launched on a 2-D grid: blockIdx.y = block of the batch,
// blockIdx.x = tile inside that block; tiles past len[b] exit immediately.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define BRA_WARP 32
#define BRA_FULL 0xFFFFFFFFu

#define BRA_CUDA_TRY(expr)                                                                      \
    do                                                                                          \
    {                                                                                           \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
        {                                                                                       \
            bra_b200_log_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                               __LINE__, #expr);                                                \
            return false;                                                                       \
        }                                                                                       \
    } while (0)

// Host-side logger: forwards to the host program's bra_log_error when the reference's log
// module is linked in (reference src/log/bra_log.h), else prints to stderr.
void bra_b200_log_error(const char* fmt, ...);

static inline uint32_t bra_div_up(uint64_t a, uint64_t b) { return (uint32_t) ((a + b - 1) / b); }

// Makes `dev` the current device for the lifetime of the guard and restores the caller's device afterwards: library
// entry points must not leave a different current device behind in the calling thread.
struct BraDeviceGuard
{
    int  prev = -1;
    bool ok   = false;
    explicit BraDeviceGuard(int dev)
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) return;
        if (cur == dev)
        {
            ok = true;
            return;
        }
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (ok) prev = cur;
    }
    ~BraDeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
    BraDeviceGuard(const BraDeviceGuard&)            = delete;
    BraDeviceGuard& operator=(const BraDeviceGuard&) = delete;
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- warp scans -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_add(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        uint32_t t = __shfl_up_sync(BRA_FULL, v, d);
        if (lane_id() >= (uint32_t) d) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int t = __shfl_up_sync(BRA_FULL, v, d);
        if (lane_id() >= (uint32_t) d) v = max(v, t);
    }
    return v;
}
// suffix (right-to-left) inclusive min
__device__ __forceinline__ int warp_incl_min_rev(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        int t = __shfl_down_sync(BRA_FULL, v, d);
        if (lane_id() + d < 32u) v = min(v, t);
    }
    return v;
}

// ---- CTA-wide exclusive scans over one value per thread (blockDim.x multiple of 32, <= 1024)
// `red` is caller-provided shared scratch of at least 33 words. All threads must call.
__device__ __forceinline__ uint32_t block_excl_add(uint32_t v, uint32_t* red, uint32_t* total)
{
    const uint32_t inc = warp_incl_add(v);
    const uint32_t w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 31) red[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        uint32_t x = l < nw ? red[l] : 0u;
        uint32_t s = warp_incl_add(x);
        red[l]     = s - x;
        if (l == 31) red[32] = s;
    }
    __syncthreads();
    if (total) *total = red[32];
    return red[w] + inc - v;
}
// exclusive prefix max; identity = `ident`
__device__ __forceinline__ int block_excl_max(int v, int ident, int* red)
{
    const int      inc = warp_incl_max(v);
    const uint32_t w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 31) red[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        int x = l < nw ? red[l] : ident;
        int s = warp_incl_max(x);
        int e = __shfl_up_sync(BRA_FULL, s, 1);
        red[l] = l == 0 ? ident : e;
    }
    __syncthreads();
    int prev = __shfl_up_sync(BRA_FULL, inc, 1);
    if (l == 0) prev = ident;
    return max(red[w], prev);
}
// exclusive suffix min (over threads with larger index); identity = `ident`
__device__ __forceinline__ int block_excl_min_rev(int v, int ident, int* red)
{
    const int      inc = warp_incl_min_rev(v);
    const uint32_t w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        int x = l < nw ? red[l] : ident;
        int s = warp_incl_min_rev(x);
        int e = __shfl_down_sync(BRA_FULL, s, 1);
        red[l] = (l + 1 >= nw) ? ident : e;
    }
    __syncthreads();
    int nxt = __shfl_down_sync(BRA_FULL, inc, 1);
    if (l == 31) nxt = ident;
    return min(red[w], nxt);
}

// streaming (read-once) loads that do not allocate in L1
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

#endif  // __CUDACC__
