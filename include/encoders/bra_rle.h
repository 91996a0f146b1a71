/*
 * encoders/bra_rle.h -- PackBits-style run-length coding, drop-in for reference
 * src/encoders/bra_rle.h:24,33,47. Control byte c: 0..127 = c+1 literal bytes follow;
 * -1..-127 = next byte repeated 1-c times; -128 = ignored. GPU implementation:
 * br-archive_b200/csrc/rle.cu.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* *out_buf receives a malloc'd buffer (caller frees) of *out_buf_size bytes. false on failure. */
bool bra_rle_encode(const uint8_t* buf, const size_t buf_size, uint8_t** out_buf, size_t* out_buf_size);
/* Decoded size of an RLE stream, 0 when a token is truncated or buf_size is 0. */
size_t bra_rle_decode_compute_size(const uint8_t* buf, const size_t buf_size);
/* *out_buf receives a malloc'd buffer (caller frees). false on a truncated token or empty output. */
bool bra_rle_decode(const uint8_t* buf, const size_t buf_size, uint8_t** out_buf, size_t* out_buf_size);

#ifdef __cplusplus
}
#endif
