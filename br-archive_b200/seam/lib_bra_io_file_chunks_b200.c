/*
 * lib_bra_io_file_chunks_b200.c -- batched replacement for the reference's chunk loop
 * (reference src/io/lib_bra_io_file_chunks.c, interface src/io/lib_bra_io_file_chunks.h:21-132).
 *
 * This is the "Level 2" integration of INTEGRATION.md: the same six functions with the same
 * signatures, error behaviour and on-disk bytes, but compress_file / decompress_file hand whole windows
 * of chunks to the GPU (bra_b200_encode_host / bra_b200_decode_host, include/bra_b200.h) instead of
 * calling the per-stage encoders once per 256 KiB chunk. It is written against the REFERENCE's headers
 * (it is a patch for the reference tree): build it with -I<reference>/src, drop the reference's own
 * lib_bra_io_file_chunks.c and the five hot-path sources from lib_bra, and link libbra_b200.so.
 * oracle/Makefile (target ref_cli) does exactly that to produce oracle/_ref/bra_gpu2 / unbra_gpu2.
 *
 * What is preserved from the reference, with the line it comes from:
 *   - tmpfile first, "not smaller => flip to STORED and return false"            chunks.c:192-197, 268-278
 *   - entry size / CRC composition: crc32c(u64 size) then combine with the chunk
 *     chain over data_size + num_chunks*268 bytes (32-bit length, quirk kept)     chunks.c:281-292
 *   - header validation on read, primary-index bound, "decoded <= stored" check   chunks.c:31-48, 385-389, 417-421
 *   - on any failure the files are closed, like every reference I/O helper        chunks.c:298-311, 428-440
 */
#include <io/lib_bra_io_file_chunks.h>

#include <io/lib_bra_io_file.h>
#include <io/lib_bra_io_file_meta_entries.h>
#include <lib_bra_defs.h>
#include <lib_bra_private.h>
#include <log/bra_log.h>
#include <utils/lib_bra_crc32c.h>

#include <bra_b200.h>

#include <assert.h>
#include <inttypes.h>
#include <stdlib.h>
#include <string.h>

#define B200_WINDOW_BYTES ((uint64_t) 256 << 20) /* plain data handed to the GPU per call */

static bra_b200_ctx_t* g_b200_ctx   = NULL;
static uint32_t        g_b200_chunk = 0; /* run-time chunk size; BRA_MAX_CHUNK_SIZE unless overridden */

/* Chunk size: the reference fixes it at compile time (BRA_MAX_CHUNK_SIZE, lib_bra_defs.h:93). Here it is a run-time
 * parameter of the library; BRA_B200_BLOCK_KIB=<KiB> (16 .. 16384, the 3-byte primary index bounds it) selects it
 * for the CLI, so that the 1 MiB / 8 MiB BASELINE configurations run through `bra` / `unbra`. Archives written with
 * a larger chunk size need the same setting to be read back (the reference reader rejects chunks above its
 * compile-time size, chunks.c:31-48, and so does this one above its run-time size). */
static uint32_t b200_chunk_size(void)
{
    if (g_b200_chunk == 0)
    {
        g_b200_chunk  = BRA_MAX_CHUNK_SIZE;
        const char* e = getenv("BRA_B200_BLOCK_KIB");
        if (e != NULL && *e != '\0')
        {
            char*               end = NULL;
            const unsigned long kib = strtoul(e, &end, 10);
            if (end != NULL && *end == '\0' && kib >= 16 && kib <= 16384)
                g_b200_chunk = (uint32_t) (kib * 1024u);
            else
                bra_log_warn("BRA_B200_BLOCK_KIB=%s ignored (expected 16..16384)", e);
        }
    }
    return g_b200_chunk;
}

static uint32_t b200_window_chunks(void)
{
    const uint64_t n = B200_WINDOW_BYTES / b200_chunk_size();
    return (uint32_t) (n < 4 ? 4 : n);
}

static bra_b200_ctx_t* b200_ctx(void)
{
    if (g_b200_ctx == NULL)
    {
        g_b200_ctx = bra_b200_ctx_create(0, b200_chunk_size(), b200_window_chunks());
        if (g_b200_ctx == NULL)
            bra_log_critical("unable to create the GPU compression context (no CPU fallback)");
    }
    return g_b200_ctx;
}

/* Window buffers: page-locked (the GPU copies run at the full PCIe rate from and to them) and kept between calls,
 * like the reference's own g_buf scratch (lib_bra.c:25-45). [0] = plain side, [1] = chunk-stream side, [2] = STORED copies.
 * If page-locking fails they are ordinary memory: slower copies, same results. */
typedef struct
{
    uint8_t* p;
    uint64_t cap;
    bool     pinned;
} b200_window_t;
static b200_window_t g_window[3]; /* [2] = the STORED path's double buffer */

static uint8_t* b200_window(const int which, const uint64_t bytes)
{
    b200_window_t* w = &g_window[which];
    if (w->p != NULL && w->cap >= bytes) return w->p;
    if (w->p != NULL)
    {
        if (w->pinned)
            bra_b200_host_free(w->p);
        else
            free(w->p);
    }
    w->p      = bra_b200_host_alloc(bytes);
    w->pinned = w->p != NULL;
    if (w->p == NULL) w->p = malloc(bytes);
    w->cap = w->p != NULL ? bytes : 0;
    return w->p;
}

static bool chunk_header_is_valid(const uint8_t* disk_hdr) /* chunks.c:31-48 on the 267-byte disk image */
{
    const uint32_t pi = (uint32_t) disk_hdr[0] | ((uint32_t) disk_hdr[1] << 8) | ((uint32_t) disk_hdr[2] << 16);
    uint32_t       orig, enc;
    memcpy(&orig, disk_hdr + 3 + BRA_ALPHABET_SIZE, 4);
    memcpy(&enc, disk_hdr + 3 + BRA_ALPHABET_SIZE + 4, 4);
    const uint32_t cs = b200_chunk_size();
    return pi < cs && enc <= cs && orig <= cs && enc != 0 && orig != 0;
}

/* ---- header I/O: 3-byte index + packed Huffman header (chunks.c:52-95) --------------------------------- */
bool bra_io_file_chunks_read_header(bra_io_file_t* src, bra_io_chunk_header_t* chunk_header)
{
    assert_bra_io_file_t(src);
    assert(chunk_header != NULL);
    uint8_t idx[4] = {0, 0, 0, 0};
    if (!bra_io_file_read(src, idx, BRA_BWT_INDEX_BYTES))
    {
        bra_log_error("unable to read chunk primary index from %s", src->fn);
        return false;
    }
    memcpy(&chunk_header->primary_index, idx, 4);
    if (!bra_io_file_read(src, &chunk_header->huffman, sizeof(bra_huffman_t)))
    {
        bra_log_error("unable to read chunk huffman header from %s", src->fn);
        return false;
    }
    return true;
}

bool bra_io_file_chunks_write_header(bra_io_file_t* dst, const bra_io_chunk_header_t* chunk_header)
{
    assert_bra_io_file_t(dst);
    assert(chunk_header != NULL);
    uint8_t idx[4];
    memcpy(idx, &chunk_header->primary_index, 4);
    if (!bra_io_file_write(dst, idx, BRA_BWT_INDEX_BYTES))
    {
        bra_log_error("unable to write chunk primary index to %s", dst->fn);
        return false;
    }
    if (!bra_io_file_write(dst, &chunk_header->huffman, sizeof(bra_huffman_t)))
    {
        bra_log_error("unable to write chunk huffman header to %s", dst->fn);
        return false;
    }
    return true;
}

bool bra_io_file_chunks_read_file(bra_io_file_t* src, const uint64_t data_size, bra_meta_entry_t* me, const bool decode)
{
    assert_bra_io_file_t(src);
    assert(me != NULL);
    switch (BRA_ATTR_COMP(me->attributes))
    {
    case BRA_ATTR_COMP_STORED:
        return bra_io_file_chunks_copy_file(NULL, src, data_size, me, decode);
    case BRA_ATTR_COMP_COMPRESSED:
        return bra_io_file_chunks_decompress_file(NULL, src, data_size, me, decode);
    default:
        bra_log_critical("invalid compression type for file: %u", BRA_ATTR_COMP(me->attributes));
        return false;
    }
}

/* ---- STORED path (chunks.c:114-167): double-buffered page-locked window; the GPU computes the CRC of piece k
 * (bra_b200_crc32c_submit: H2D copy + kernels, asynchronous) while piece k is written out and piece k+1 is read -------- */
bool bra_io_file_chunks_copy_file(bra_io_file_t* dst, bra_io_file_t* src, const uint64_t data_size, bra_meta_entry_t* me, const bool compute_crc32)
{
    assert_bra_io_file_t(src);
    const uint64_t  piece = (uint64_t) 64 << 20;
    bra_b200_ctx_t* ctx   = NULL;
    bool            busy  = false; /* a CRC submission is in flight */
    if (dst != NULL && (dst->f == NULL || dst->fn == NULL)) goto fail;
    if (compute_crc32 && me == NULL)
    {
        bra_log_critical("can't compute crc32: me is null");
        goto fail;
    }
    if (data_size == 0) return true;
    if (compute_crc32 && (ctx = b200_ctx()) == NULL) goto fail;
    const uint64_t half = data_size < piece ? data_size : piece;
    uint8_t*       buf  = b200_window(2, 2 * half);
    if (buf == NULL) goto fail;
    int k = 0;
    for (uint64_t i = 0; i < data_size; k ^= 1)
    {
        const uint64_t s = _bra_min(piece, data_size - i);
        uint8_t*       p = buf + (uint64_t) k * half;
        if (!bra_io_file_read(src, p, s)) goto fail;
        if (compute_crc32)
        {
            if (busy && bra_b200_crc32c_finish(ctx, &me->crc32) != 0) goto fail; /* piece k-1; its half is reused next round */
            busy = false;
            if (bra_b200_crc32c_submit(ctx, p, s) != 0) goto fail;
            busy = true;
        }
        if (dst != NULL && !bra_io_file_write(dst, p, s)) goto fail;
        i += s;
    }
    if (busy && bra_b200_crc32c_finish(ctx, &me->crc32) != 0)
    {
        busy = false;
        goto fail;
    }
    return true;
fail:
    if (busy) (void) bra_b200_crc32c_finish(ctx, NULL);
    if (dst != NULL) bra_io_file_close(dst);
    bra_io_file_close(src);
    return false;
}

/* ---- pack ---------------------------------------------------------------------------------------------------- */
bool bra_io_file_chunks_compress_file(bra_io_file_t* dst, bra_io_file_t* src, const uint64_t data_size, bra_meta_entry_t* me)
{
    assert_bra_io_file_t(dst);
    assert_bra_io_file_t(src);
    assert(me != NULL);

    bra_b200_ctx_t* ctx = b200_ctx();
    if (ctx == NULL) return false;

    const uint32_t chunk  = b200_chunk_size();
    const uint64_t window = (uint64_t) b200_window_chunks() * chunk;
    uint8_t*       in     = NULL;
    uint8_t*       out    = NULL;
    uint32_t       crc32  = BRA_CRC32C_INIT;
    bra_io_file_t  tmpfile;
    if (!bra_io_file_tmp_open(&tmpfile))
    {
        bra_log_error("unable to compress file: %s", src->fn);
        return false;
    }
    const uint64_t in_cap  = data_size < window ? (data_size ? data_size : 1) : window;
    const uint64_t out_cap = bra_b200_encode_bound(ctx, in_cap);
    in  = b200_window(0, in_cap);
    out = b200_window(1, out_cap ? out_cap : 1);
    if (in == NULL || out == NULL) goto fail;

    for (uint64_t i = 0; i < data_size;)
    {
        const uint64_t s = _bra_min(window, data_size - i);
        bra_log_printf("%3u%%", (unsigned int) (i * 100 / data_size));
        bra_log_printf("\b\b\b\b");
        if (!bra_io_file_read(src, in, s))
        {
            bra_io_file_close(&tmpfile);
            bra_io_file_close(dst);
            return false;
        }
        uint64_t out_size = 0;
        /* crc, bwt, mtf, rle, huffman, disk header + payload and the CRC chain of chunks.c:214-256 for every chunk of the window */
        if (bra_b200_encode_host(ctx, in, s, out, out_cap, &out_size, &crc32) != 0)
        {
            bra_log_error("GPU chunk encoding failed: %s (offset: %" PRIu64 ")", src->fn, i);
            goto fail;
        }
        if (!bra_io_file_write(&tmpfile, out, out_size)) goto fail;
        i += s;
    }

    const int64_t tmpfile_size = bra_io_file_tell(&tmpfile);
    if (tmpfile_size < 0) goto fail;
    bool res = true;
    if ((uint64_t) tmpfile_size >= data_size)
    {
        res            = false; /* not smaller: the caller rewinds and stores the entry instead */
        me->attributes = BRA_ATTR_SET_COMP(me->attributes, BRA_ATTR_COMP_STORED);
    }
    else
    {
        if (!bra_io_file_seek(&tmpfile, 0, SEEK_SET)) goto fail;
        uint64_t num_chunks = data_size / chunk;
        if (data_size % chunk > 0) ++num_chunks;
        bra_meta_entry_file_t* mef = (bra_meta_entry_file_t*) me->entry_data;
        mef->data_size             = tmpfile_size;
        me->crc32                  = bra_crc32c(&tmpfile_size, sizeof(tmpfile_size), me->crc32);
        me->crc32                  = bra_crc32c_combine(me->crc32, crc32, data_size + (num_chunks * sizeof(bra_io_chunk_header_t)));
        if (!bra_io_file_meta_entry_write_file_entry(dst, me)) goto fail;
        res = bra_io_file_chunks_copy_file(dst, &tmpfile, tmpfile_size, me, false);
    }
    bra_io_file_close(&tmpfile);
    return res;

fail:
    bra_io_file_close(&tmpfile);
    bra_io_file_close(dst);
    bra_io_file_close(src);
    return false;
}

/* ---- unpack / list / test ------------------------------------------------------------------------------------------ */
bool bra_io_file_chunks_decompress_file(bra_io_file_t* dst, bra_io_file_t* src, const uint64_t data_size, bra_meta_entry_t* me, const bool decode)
{
    assert_bra_io_file_t(src);
    assert(me != NULL);

    bra_b200_ctx_t* ctx = b200_ctx();
    if (ctx == NULL) return false;

    const uint32_t chunk      = b200_chunk_size();
    const uint32_t wchunks    = b200_window_chunks();
    const uint64_t max_stream = (uint64_t) wchunks * (BRA_IO_CHUNK_HEADER_SIZE + chunk);
    uint8_t*       stream     = NULL;
    uint8_t*       plain      = NULL;
    uint64_t       file_orig_size = 0;

    if (dst != NULL && (dst->f == NULL || dst->fn == NULL)) goto fail;
    const uint64_t stream_cap = data_size < max_stream ? (data_size ? data_size : 1) : max_stream;
    stream = b200_window(1, stream_cap);
    if (decode) plain = b200_window(0, _bra_min((uint64_t) wchunks, data_size / BRA_IO_CHUNK_HEADER_SIZE + 1) * chunk);
    if (stream == NULL || (decode && plain == NULL)) goto fail;

    for (uint64_t i = 0; i < data_size;)
    {
        /* collect a window of whole chunks: each header names the size of its payload (chunks.c:344-357) */
        uint64_t w = 0;
        uint32_t n = 0;
        while (n < wchunks && i + w < data_size)
        {
            uint8_t* h = stream + w;
            if (w + BRA_IO_CHUNK_HEADER_SIZE > stream_cap)
            {
                bra_log_error("truncated chunk header in %s", src->fn);
                goto fail;
            }
            if (!bra_io_file_read(src, h, BRA_IO_CHUNK_HEADER_SIZE))
            {
                bra_log_error("unable to read chunk header from %s", src->fn);
                goto fail;
            }
            if (!chunk_header_is_valid(h))
            {
                bra_log_error("chunk header not valid in %s", src->fn);
                goto fail;
            }
            uint32_t enc;
            memcpy(&enc, h + 3 + BRA_ALPHABET_SIZE + 4, 4);
            if (w + BRA_IO_CHUNK_HEADER_SIZE + enc > stream_cap || i + w + BRA_IO_CHUNK_HEADER_SIZE + enc > data_size)
            {
                bra_log_error("chunk payload overruns its entry in %s", src->fn);
                goto fail;
            }
            if (!bra_io_file_read(src, h + BRA_IO_CHUNK_HEADER_SIZE, enc)) goto fail;
            w += BRA_IO_CHUNK_HEADER_SIZE + enc;
            ++n;
        }
        uint64_t plain_size = 0;
        uint32_t crc        = decode ? me->crc32 : 0;
        /* list mode (chunks.c:369-373): huffman decode + rle size only;
         * otherwise huffman, rle, mtf, bwt decode of every chunk + CRC chain of chunks.c:396-397 */
        const int rc = decode ? bra_b200_decode_host(ctx, stream, w, plain, (uint64_t) n * chunk, &plain_size, &crc)
                              : bra_b200_list_host(ctx, stream, w, &plain_size);
        if (rc != 0)
        {
            bra_log_error("unable to decode chunks of file: %s ", src->fn);
            goto fail;
        }
        file_orig_size += plain_size;
        if (decode)
        {
            me->crc32 = crc;
            if (dst != NULL && !bra_io_file_write(dst, plain, plain_size)) goto fail;
        }
        i += w;
    }

    if (file_orig_size <= data_size) /* chunks.c:417-421 */
    {
        bra_log_error("corrupted file entry: %s", me->name);
        goto fail;
    }
    me->_compression_ratio = (float) ((double) data_size / (double) file_orig_size);
    return true;

fail:
    if (dst != NULL) bra_io_file_close(dst);
    bra_io_file_close(src);
    return false;
}
