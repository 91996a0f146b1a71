/*
 * lib_bra_types.h -- the types that cross the hot-path boundary, layout-identical to
 * reference src/lib_bra_types.h:11 (bra_bwt_index_t), :51-56 (bra_huffman_t, packed, 264 bytes)
 * and :63-68 (bra_io_chunk_header_t, 268 bytes in memory).
 */
#pragma once

#include "lib_bra_defs.h"

#include <stdbool.h>
#include <stdint.h>

typedef uint32_t bra_bwt_index_t; /* rotation index / block length; 3 bytes of it reach the disk */

/* Not used by the hot path, but other reference headers that test programs include next to the encoder
 * headers (src/fs/bra_fs.hpp) expect this header to provide them (reference lib_bra_types.h:10,16-21). */
typedef uint8_t bra_attr_t;
typedef enum bra_fs_overwrite_policy_e
{
    BRA_OVERWRITE_ASK        = 0,
    BRA_OVERWRITE_ALWAYS_YES = 1,
    BRA_OVERWRITE_ALWAYS_NO  = 2,
} bra_fs_overwrite_policy_e;

#pragma pack(push, 1)
typedef struct bra_huffman_t
{
    uint8_t  lengths[BRA_ALPHABET_SIZE]; /* canonical code length per symbol, 0 = absent */
    uint32_t orig_size;                  /* symbols encoded (= RLE output bytes) */
    uint32_t encoded_size;               /* payload bytes */
} bra_huffman_t;
#pragma pack(pop)

typedef struct bra_io_chunk_header_t
{
    bra_bwt_index_t primary_index; /* row of the original string in the sorted rotation matrix */
    bra_huffman_t   huffman;
} bra_io_chunk_header_t;

#ifdef __cplusplus
static_assert(sizeof(bra_huffman_t) == 264, "bra_huffman_t must be 264 packed bytes");
static_assert(sizeof(bra_io_chunk_header_t) == 268, "bra_io_chunk_header_t must be 268 bytes");
#else
_Static_assert(sizeof(bra_huffman_t) == 264, "bra_huffman_t must be 264 packed bytes");
_Static_assert(sizeof(bra_io_chunk_header_t) == 268, "bra_io_chunk_header_t must be 268 bytes");
#endif
