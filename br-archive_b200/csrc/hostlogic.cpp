// hostlogic.cpp -- host-only build of the scalar logic the kernels share with the host (bra_hd.h),
// exported for the CPU test-suite (tests/test_host_logic.py). Not part of the product library.
#include "bra_hd.h"

#include <stdlib.h>
#include <string.h>

extern "C" {

uint32_t hl_gf_mul(uint32_t a, uint32_t b) { return bra_gf_mul(a, b); }

uint32_t hl_crc_combine(uint32_t a, uint32_t b, uint64_t len_b)
{
    bra_gf_pow_t t;
    bra_gf_init_pow(&t);
    return bra_crc_combine(&t, a, b, len_b);
}

uint32_t hl_huf_lengths(const uint32_t* freq, uint8_t* lengths)
{
    bra_huf_build_ws_t* ws = (bra_huf_build_ws_t*) malloc(sizeof(bra_huf_build_ws_t));
    const uint32_t      k  = bra_huf_build_lengths(freq, lengths, ws);
    free(ws);
    return k;
}

void hl_huf_canonical(const uint8_t* lengths, uint32_t* codes) { bra_huf_canonical(lengths, codes); }

// sequential decode through the device decode tables: returns 0 ok, -1 tables rejected, -2 invalid code, -3 out of data
int hl_huf_decode(const uint8_t* lengths, const uint8_t* data, uint32_t nbytes, uint32_t nsym, uint8_t* out)
{
    bra_huf_dec_t d;
    if (!bra_huf_make_dec(lengths, &d)) return -1;
    uint64_t pos = 0;
    const uint64_t end = (uint64_t) nbytes * 8;
    for (uint32_t i = 0; i < nsym; ++i)
    {
        uint32_t w = 0;
        for (int k = 0; k < 5; ++k)
        {
            const uint64_t byte = (pos >> 3) + k;
            const uint64_t v    = byte < nbytes ? data[byte] : 0;
            // assemble 40 bits then take 32 from the bit offset
            if (k == 0) w = 0;
            (void) v;
        }
        uint64_t acc = 0;
        for (int k = 0; k < 5; ++k)
        {
            const uint64_t byte = (pos >> 3) + k;
            acc = (acc << 8) | (byte < nbytes ? data[byte] : 0);
        }
        w = (uint32_t) ((acc >> (8 - (pos & 7))) & 0xFFFFFFFFull);
        uint8_t        sym;
        const uint32_t l = bra_huf_decode_one(&d, w, &sym);
        if (l == 0) return -2;
        if (pos + l > end) return -3;
        out[i] = sym;
        pos += l;
    }
    return 0;
}

// move-to-front over one 256-entry list held as sixteen 128-bit chunks (bra_mtf_list_encode / _decode), identity list at the start
void hl_mtf(const uint8_t* in, uint8_t* out, uint64_t n, int decode)
{
    uint4 Q[16];
    for (uint32_t q = 0; q < 16; ++q)
    {
        const uint32_t w0 = 0x03020100u + (q * 16) * 0x01010101u;
        Q[q]              = make_uint4(w0, w0 + 0x04040404u, w0 + 0x08080808u, w0 + 0x0C0C0C0Cu);
    }
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t) (decode ? bra_mtf_list_decode(Q, in[i]) : bra_mtf_list_encode(Q, in[i]));
}

// BWT finisher comparison (bra_rot_cmp_window); T must be 4-byte aligned and readable up to p rounded up to 4
int hl_rot_cmp(const uint8_t* T, uint32_t p, uint32_t a, uint32_t c, uint32_t from, uint32_t depth) { return bra_rot_cmp_window(T, p, a, c, from, depth); }

// host-path pipeline plan (bra_stage_plan); `plan` must hold nblk / hb + 4 entries
uint32_t hl_stage_plan(uint64_t nblk, uint32_t hb, uint32_t* plan) { return bra_stage_plan(nblk, hb, plan); }
uint32_t hl_stage_plan_encode(uint64_t nblk, uint32_t hb, uint32_t head_div, uint32_t* plan) { return bra_stage_plan_encode(nblk, hb, head_div, plan); }

}  // extern "C"
