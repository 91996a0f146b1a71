/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's CRC module asks lib_bra.c for bra_has_sse42() (reference
 * src/utils/lib_bra_crc32c.c:233-239 calling src/lib_bra.c:70-89). lib_bra.c
 * drags in the whole file-I/O layer, so the oracle/_ref build links this
 * one-function stand-in instead. The reference sources themselves are compiled
 * in place from /root/reference by oracle/Makefile and never copied.
 */
#include <stdbool.h>

bool bra_has_sse42(void)
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_cpu_init();
    return __builtin_cpu_supports("sse4.2") != 0;
#else
    return false;
#endif
}
