/*
 * encoders/bra_huffman.h -- canonical Huffman coding, drop-in for reference
 * src/encoders/bra_huffman.h:10-42. Code lengths replay the reference's sorted-list tree build
 * (src/encoders/bra_huffman.c:90-186), codes are canonical and packed MSB-first.
 * GPU implementation: br-archive_b200/csrc/huffman.cu.
 */
#pragma once

#include <lib_bra_defs.h>
#include <lib_bra_types.h>

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bra_huffman_chunk_t
{
    bra_huffman_t meta; /* lengths + sizes */
    uint8_t*      data; /* meta.encoded_size bytes, malloc'd */
} bra_huffman_chunk_t;

/* NULL when buf_size is 0 or on failure; otherwise free with bra_huffman_chunk_free. */
bra_huffman_chunk_t* bra_huffman_encode(const uint8_t* buf, const uint32_t buf_size);
/* malloc'd buffer of meta->orig_size bytes (caller frees), *out_size set; NULL on corrupt input. */
uint8_t* bra_huffman_decode(const bra_huffman_t* meta, const uint8_t* data, uint32_t* out_size);
/* NULL-safe. */
void bra_huffman_chunk_free(bra_huffman_chunk_t* chunk);

#ifdef __cplusplus
}
#endif
