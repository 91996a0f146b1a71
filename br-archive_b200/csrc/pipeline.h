// pipeline.h -- internals shared by pipeline.cu, hostpath.cu and capi.cu
#pragma once

#include <bra_b200.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bra {
uint64_t rle_stride_for(uint32_t block);
uint64_t pay_stride_for(uint32_t block);
bool encode_batch(bra_b200_ctx* c, const uint8_t* d_in, uint32_t nb, uint32_t last_len, uint8_t* d_hdr, uint8_t* d_payload, uint32_t* d_crc,
                  cudaStream_t st);
bool decode_batch(bra_b200_ctx* c, const uint8_t* d_hdr, const uint8_t* d_payload, uint32_t nb, uint32_t hint_r, uint32_t hint_c, uint8_t* d_out,
                  uint32_t* d_out_len, uint32_t* d_crc, uint32_t* d_status, cudaStream_t st, bool sizes_only);
// device staging buffer of the context for the host path (grown on demand)
uint8_t* ctx_io_buffer(bra_b200_ctx* c, uint64_t bytes);
cudaStream_t ctx_stream(bra_b200_ctx* c);
uint32_t* ctx_mail_host(bra_b200_ctx* c);  // pinned, device-visible words for the host path (8*max_batch + 16)
int ctx_device(const bra_b200_ctx* c);
// bra_b200_encode_host with the destination of every stage's stream bytes named by `place` (hostpath.cu; used by pool.cu)
int encode_host_impl(bra_b200_ctx_t* c, const uint8_t* in, uint64_t total, uint8_t* (*place)(void*, uint64_t), void* user, uint64_t* out_size,
                     uint32_t* crc_chain);
}  // namespace bra
