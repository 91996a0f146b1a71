/*
 * utils/lib_bra_crc32c.h -- CRC-32C (Castagnoli), drop-in for reference
 * src/utils/lib_bra_crc32c.h:7-67. The initial and final inversions happen inside, so 0 starts a
 * CRC and a returned CRC can be passed back as previous_crc to continue it.
 * bra_crc32c / _table / _sse42 all run the same GPU kernel here (br-archive_b200/csrc/crc32c.cu):
 * the two suffixed names are the reference's CPU strategies and exist so its callers and tests link.
 */
#pragma once

#include <stdbool.h>
#include <stdint.h>

#define BRA_CRC32C_INIT 0u

#ifdef __cplusplus
extern "C" {
#endif

uint32_t bra_crc32c_table(const void* data, const uint64_t length, const uint32_t previous_crc);
uint32_t bra_crc32c_sse42(const void* data, const uint64_t length, const uint32_t previous_crc);
uint32_t bra_crc32c(const void* data, const uint64_t length, const uint32_t previous_crc);
/* crc(A||B) from crc(A), crc(B) and |B| (32-bit length, like the reference). Host arithmetic. */
uint32_t bra_crc32c_combine(uint32_t crc32a, uint32_t crc32b, uint32_t len_b);
/* Accepted for source compatibility; the device implementation is always used. */
void bra_crc32c_use_sse42(const bool use_sse42);

#ifdef __cplusplus
}
#endif
