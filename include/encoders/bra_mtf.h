/*
 * encoders/bra_mtf.h -- move-to-front coding, drop-in for reference src/encoders/bra_mtf.h
 * (implementation reference src/encoders/bra_mtf.c:49-115). The list starts as 0..255 for every
 * call. GPU implementation: br-archive_b200/csrc/mtf.cu.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

uint8_t* bra_mtf_encode(const uint8_t* buf, const size_t buf_size);                 /* malloc'd, caller frees; NULL on failure */
bool     bra_mtf_encode2(const uint8_t* buf, const size_t buf_size, uint8_t* out_buf); /* out_buf: buf_size bytes */
uint8_t* bra_mtf_decode(const uint8_t* buf, const size_t buf_size);                 /* malloc'd, caller frees; NULL on failure */
void     bra_mtf_decode2(const uint8_t* buf, const size_t buf_size, uint8_t* out_buf);

#ifdef __cplusplus
}
#endif
