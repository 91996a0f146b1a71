/*
 * encoders/bra_bwt.h -- Burrows-Wheeler transform, drop-in for reference src/encoders/bra_bwt.h
 * (encode :42, encode2 :63, decode :93, decode2 :124). Same names, arguments, ownership and
 * failure values; the work runs on the GPU (br-archive_b200/csrc/bwt.cu).
 *
 * Order: all n cyclic rotations of buf sorted as unsigned bytes; identical rotations (periodic
 * input) in ascending start index, which is what the reference's stable qsort_r produces.
 * Limits of this implementation: 0 < n <= 2^24 (the on-disk index has 3 bytes); larger or
 * empty inputs return false / NULL instead of tripping the reference's assert.
 */
#pragma once

#include <lib_bra_types.h>

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Returns a malloc'd buffer of buf_size bytes (caller frees) holding the last column; *primary_index
 * receives the row of rotation 0. NULL on failure. */
uint8_t* bra_bwt_encode(const uint8_t* buf, const bra_bwt_index_t buf_size, bra_bwt_index_t* primary_index);
/* Same, into a caller buffer of buf_size bytes. */
bool bra_bwt_encode2(const uint8_t* buf, const bra_bwt_index_t buf_size, bra_bwt_index_t* primary_index, uint8_t* out_buf);
/* Inverse; returns a malloc'd buffer of buf_size bytes or NULL. primary_index must be < buf_size. */
uint8_t* bra_bwt_decode(const uint8_t* buf, const bra_bwt_index_t buf_size, const bra_bwt_index_t primary_index);
/* Inverse into out_buf. `transform` is the caller scratch of buf_size entries the reference fills with
 * its LF map (bra_bwt.c:155-159); it is filled with the same values here. */
void bra_bwt_decode2(const uint8_t* buf, const bra_bwt_index_t buf_size, const bra_bwt_index_t primary_index, bra_bwt_index_t* transform,
                     uint8_t* out_buf);

#ifdef __cplusplus
}
#endif
