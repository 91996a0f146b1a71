/*
 * lib_bra_defs.h -- constants the block-compression hot path shares with the rest of lib_bra.
 * Values (not text) mirror reference src/lib_bra_defs.h:93-100; only what the hot path needs.
 * BRA_MAX_CHUNK_SIZE is the reference's compile-time default; the B200 library takes the block
 * size at run time (bra_b200_ctx_create) and only uses this constant for the drop-in defaults.
 */
#pragma once

#include <stdint.h>

#if defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__)
#error "lib_bra is little-endian only (reference src/lib_bra_defs.h:17-19)"
#endif

#define BRA_ALPHABET_SIZE 256                 /* symbols of the byte alphabet */
#ifndef BRA_MAX_CHUNK_SIZE
#define BRA_MAX_CHUNK_SIZE (256 * 1024)       /* bytes per independently compressed chunk */
#endif
#define BRA_BWT_INDEX_BYTES 3                 /* bytes of the BWT primary index stored on disk */
#define BRA_RLE_MAX_RUNS 128                  /* longest run or literal one RLE token covers */
#define BRA_RLE_MIN_RUNS 3                    /* shortest run worth a run token */
#define BRA_RLE_CTL_RUNS -127                 /* control bytes in [-127,-1] are run tokens */
#define BRA_IO_CHUNK_HEADER_SIZE (BRA_BWT_INDEX_BYTES + sizeof(bra_huffman_t)) /* 267 on disk */
