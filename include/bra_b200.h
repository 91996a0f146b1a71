/*
 * bra_b200.h -- batched C ABI of the B200 block-compression path.
 *
 * This is the seam UNDER the reference's chunk loop: what
 * bra_io_file_chunks_compress_file / bra_io_file_chunks_decompress_file
 * (reference src/io/lib_bra_io_file_chunks.c:199-266 and :338-414) do one 256 KiB chunk at a
 * time through bra_crc32c + bra_bwt_encode2 + bra_mtf_encode2 + bra_rle_encode +
 * bra_huffman_encode (and the inverse), these entry points do for many chunks at once on
 * the GPU. Results are byte-identical to looping the per-stage API (the headers under include/encoders).
 *
 * Plain C: pointers and sizes only. Pointers named d_* are device pointers on the context's
 * GPU; everything else is host memory. All functions return 0 on success, non-zero on failure
 * (and log through bra_log_error when the host program provides it); none of them aborts,
 * and none has a CPU fallback: without a usable CUDA device they fail.
 */
#ifndef BRA_B200_H
#define BRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRA_B200_HDR_BYTES 268u      /* in-memory chunk header (reference lib_bra_types.h:63-68); this is what is CRC'd */
#define BRA_B200_DISK_HDR_BYTES 267u /* on-disk chunk header: 3-byte index + 264 (reference lib_bra_defs.h:94,99) */
#define BRA_B200_MAX_BLOCK (1u << 24) /* the on-disk primary index has 3 bytes (reference bra_bwt.c:26) */

typedef struct bra_b200_ctx bra_b200_ctx_t;

/* Number of CUDA devices visible to the process, or -1. */
int bra_b200_device_count(void);

/* One context per GPU and calling thread. block_size: bytes per block (the reference's
 * BRA_MAX_CHUNK_SIZE, reference lib_bra_defs.h:93), multiple of 16, <= BRA_B200_MAX_BLOCK.
 * max_batch: blocks processed per internal batch (bounds device workspace: ~32 bytes per input byte). */
bra_b200_ctx_t* bra_b200_ctx_create(int device, uint32_t block_size, uint32_t max_batch);
void            bra_b200_ctx_destroy(bra_b200_ctx_t* ctx);

uint32_t bra_b200_block_size(const bra_b200_ctx_t* ctx);
uint32_t bra_b200_max_batch(const bra_b200_ctx_t* ctx);
/* bytes between consecutive per-block slots in d_payload; every payload fits (incompressible data expands) */
uint64_t bra_b200_payload_stride(const bra_b200_ctx_t* ctx);
/* device workspace the context holds, bytes */
uint64_t bra_b200_workspace_bytes(const bra_b200_ctx_t* ctx);
/* statistics of the last call: doubling rounds of the BWT sort, sweeps of the Huffman decode fixed point, kernel launches */
void bra_b200_last_stats(const bra_b200_ctx_t* ctx, uint32_t* bwt_rounds, uint32_t* huf_sweeps, uint64_t* launches);

/* ---- device-resident path (inputs and outputs stay in HBM) ---------------------------------
 * Encode nblk blocks: block b is d_in + b*block_size, block_size bytes long except the last one,
 * which is last_len bytes (1..block_size). Outputs, per block b:
 *   d_hdr     + b*268             the 268-byte in-memory chunk header
 *   d_payload + b*payload_stride  encoded_size bytes of Huffman payload
 *   d_crc_raw[b]                  CRC-32C of the uncompressed block (reference chunks.c:214)
 * d_in must be 16-byte aligned. stream is a cudaStream_t (NULL = default stream). nblk may exceed
 * max_batch; the call loops over batches. Returns after the work is enqueued and complete on `stream`. */
int bra_b200_encode_device(bra_b200_ctx_t* ctx, const uint8_t* d_in, uint32_t nblk, uint32_t last_len, uint8_t* d_hdr, uint8_t* d_payload,
                           uint32_t* d_crc_raw, void* stream);

/* Decode nblk blocks laid out as above. Per block b: d_out + b*block_size receives d_out_len[b]
 * bytes, d_crc_raw[b] their CRC-32C, d_status[b] is 0 on success and non-zero where the reference
 * chain would have failed (corrupt header/payload, truncated RLE token, primary index out of range,
 * decoded size above block_size). hint_max_r / hint_max_c: upper bounds of orig_size / encoded_size
 * over the batch when the caller knows them (they only trim empty CTAs), 0 otherwise. */
int bra_b200_decode_device(bra_b200_ctx_t* ctx, const uint8_t* d_hdr, const uint8_t* d_payload, uint32_t nblk, uint32_t hint_max_r,
                           uint32_t hint_max_c, uint8_t* d_out, uint32_t* d_out_len, uint32_t* d_crc_raw, uint32_t* d_status, void* stream);

/* ---- host-buffer path (what compress_file / decompress_file need) ---------------------------
 * Encode `total` bytes of host memory into the .BRa chunk stream the reference writes to its
 * temporary file (reference chunks.c:252-256): per chunk the 267-byte disk header followed by
 * the payload. *crc_chain is updated exactly as chunks.c:248-249 does:
 *   crc = crc32c(header268, crc); crc = combine(crc, crc32c(chunk), chunk_len)   for every chunk.
 * Copies host->device and device->host are pipelined against the kernels. */
uint64_t bra_b200_encode_bound(const bra_b200_ctx_t* ctx, uint64_t total);
int bra_b200_encode_host(bra_b200_ctx_t* ctx, const uint8_t* in, uint64_t total, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                         uint32_t* crc_chain);
/* Decode a chunk stream produced by the above (or by the reference). *crc_chain is updated as
 * reference chunks.c:396-397 does. out_cap must hold the decoded data. */
int bra_b200_decode_host(bra_b200_ctx_t* ctx, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                         uint32_t* crc_chain);
/* List mode (reference chunks.c:369-373, `decode == false`): Huffman-decode every chunk of the stream and
 * sum bra_rle_decode_compute_size over them into *plain_size. No MTF/BWT stage runs and nothing but the
 * sizes comes back. As in the reference, a broken Huffman stream fails the call; a truncated RLE token
 * makes its chunk count as 0 bytes. */
int bra_b200_list_host(bra_b200_ctx_t* ctx, const uint8_t* in, uint64_t in_size, uint64_t* plain_size);

/* ---- one job over several GPUs of ONE process (BASELINE config 5) ------------------------------------
 * A pool owns `workers_per_device` contexts on each of the listed devices (the same device may be listed more
 * than once). bra_b200_pool_encode_host / _decode_host behave exactly like the single-context calls above --
 * same chunk stream, same CRC chain -- but the block list of the input is cut into ranges of `range_blocks`
 * blocks that the workers take from a shared queue; every range lands at its final offset of the ordered
 * stream and the per-range CRC chains are folded with bra_crc32c_combine. There is no inter-GPU collective. */
typedef struct bra_b200_pool bra_b200_pool_t;
bra_b200_pool_t* bra_b200_pool_create(const int* devices, int ndev, uint32_t block_size, uint32_t range_blocks, int workers_per_device);
void             bra_b200_pool_destroy(bra_b200_pool_t* pool);
int              bra_b200_pool_workers(const bra_b200_pool_t* pool);
/* device and number of ranges worker `worker` processed in the last call */
int      bra_b200_pool_worker_ranges(const bra_b200_pool_t* pool, int worker, int* device, uint32_t* ranges);
uint64_t bra_b200_pool_encode_bound(const bra_b200_pool_t* pool, uint64_t total);
int bra_b200_pool_encode_host(bra_b200_pool_t* pool, const uint8_t* in, uint64_t total, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                              uint32_t* crc_chain);
int bra_b200_pool_decode_host(bra_b200_pool_t* pool, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint64_t* out_size,
                              uint32_t* crc_chain);

/* Streaming CRC-32C of host memory for the STORED path (reference chunks.c:114-167, which calls bra_crc32c per 256 KiB
 * piece while copying): _submit enqueues the host-to-device copy and the CRC kernels for `len` bytes (at most 1 GiB) and
 * returns at once, so the caller can write the piece out and read the next one meanwhile; `data` must stay valid until
 * _finish, which waits and updates *crc to bra_crc32c(data, len, *crc) (crc may be NULL to just drain). One submission
 * at a time per context. */
int bra_b200_crc32c_submit(bra_b200_ctx_t* ctx, const void* data, uint64_t len);
int bra_b200_crc32c_finish(bra_b200_ctx_t* ctx, uint32_t* crc);

/* Page-locked host memory for the buffers handed to the three calls above: copies to and from it run at the
 * full PCIe rate and overlap the kernels (ordinary memory works too, at a fraction of the rate).
 * bra_b200_host_alloc returns NULL when the memory cannot be locked. */
void* bra_b200_host_alloc(uint64_t bytes);
void  bra_b200_host_free(void* p);

/* ---- launch accounting (process-wide; used by bench.py for `gpu_launches` and the roofline) ----
 * Every kernel launch of the library is counted per kernel family. With timing enabled each launch is
 * also bracketed by CUDA events on its own stream; the accumulated device time is read back lazily. */
void bra_b200_prof_enable(int timing_on);
void bra_b200_prof_reset(void);
int  bra_b200_prof_count(void);
int  bra_b200_prof_read(int id, const char** name, uint64_t* launches, double* ms);

#ifdef __cplusplus
}
#endif
#endif
