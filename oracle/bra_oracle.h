/*
 * oracle/bra_oracle.h -- CPU restatement of BR-Archive's block-compression chain.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under br-archive_b200/ may include, link or
 * execute this. Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py (as the checker or the timed
 * CPU baseline, never as the product).
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function here against
 * (a) the golden vectors of the reference's own unit tests
 *     (reference test/test_bra_encoders.cpp, test/test_bra_crc32c.cpp), committed
 *     as tests/golden/reference_vectors.json, and
 * (b) the reference's own sources compiled in place into oracle/_ref/libbra_ref.so
 *     (oracle/Makefile), on seeded random inputs, whenever that library is present.
 *
 * Each function names the reference file:line it restates.
 */
#ifndef BRA_ORACLE_H
#define BRA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_HDR_BYTES 268u /* in-memory chunk header: u32 primary_index + 256 lengths + u32 orig + u32 enc */

/* CRC-32C. reference src/utils/lib_bra_crc32c.c:102-114 (table) / :176-179 */
uint32_t ora_crc32c(const void* data, uint64_t len, uint32_t prev);
/* reference src/utils/lib_bra_crc32c.c:181-231 (zlib-style combine; len_b is 32-bit there too) */
uint32_t ora_crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint32_t len_b);

/* BWT over cyclic rotations, ties by ascending rotation index.
 * reference src/encoders/bra_bwt.c:31-53 (comparator), :73-108 (encode2). */
int  ora_bwt_encode_naive(const uint8_t* in, uint32_t n, uint32_t* primary, uint8_t* out); /* merge sort + byte comparator, O(n^2 log n) worst */
int  ora_bwt_encode(const uint8_t* in, uint32_t n, uint32_t* primary, uint8_t* out);       /* prefix doubling, same result, O(n log^2 n) */
/* reference src/encoders/bra_bwt.c:133-168 */
int  ora_bwt_decode(const uint8_t* in, uint32_t n, uint32_t primary, uint8_t* out);

/* reference src/encoders/bra_mtf.c:67-82 and :98-115 */
void ora_mtf_encode(const uint8_t* in, size_t n, uint8_t* out);
void ora_mtf_decode(const uint8_t* in, size_t n, uint8_t* out);

/* reference src/encoders/bra_rle.c:20-56 (size), :60-120 (encode) */
size_t ora_rle_encode_size(const uint8_t* in, size_t n);
int    ora_rle_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n);
/* reference src/encoders/bra_rle.c:122-160 (size; 0 = error), :162-224 (decode) */
size_t ora_rle_decode_size(const uint8_t* in, size_t n);
int    ora_rle_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_n);

/* reference src/encoders/bra_huffman.c:90-118 (list insert), :132-186 (tree), :188-220 (depths) */
int ora_huffman_lengths(const uint32_t freq[256], uint8_t lengths[256]);
/* reference src/encoders/bra_huffman.c:227-261 (canonical codes, uint32 arithmetic) */
void ora_huffman_canonical(const uint8_t lengths[256], uint32_t codes[256]);
/* reference src/encoders/bra_huffman.c:352-432. Returns 0, or -1 when n == 0 (reference returns NULL). */
int ora_huffman_encode(const uint8_t* in, uint32_t n, uint8_t lengths[256], uint8_t* out, size_t cap, uint32_t* encoded_size);
/* reference src/encoders/bra_huffman.c:263-348 (tree from lengths), :434-498 (bit walk). 0 ok, -1 corrupt. */
int ora_huffman_decode(const uint8_t lengths[256], const uint8_t* data, uint32_t encoded_size, uint32_t orig_size, uint8_t* out);

/* Whole chain for one block, as reference src/io/lib_bra_io_file_chunks.c:214-249 (encode)
 * and :362-397 (decode) apply it. hdr is the 268-byte in-memory header. */
int ora_encode_block(const uint8_t* in, uint32_t n, uint8_t hdr[ORA_HDR_BYTES], uint8_t* payload, size_t cap, uint32_t* crc_raw, int naive_bwt);
int ora_decode_block(const uint8_t hdr[ORA_HDR_BYTES], const uint8_t* payload, uint8_t* out, size_t cap, uint32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
