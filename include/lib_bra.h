/*
 * lib_bra.h -- the three library-level entry points the hot path owns, mirroring
 * reference src/lib_bra.h:32,40,48. bra_init() brings up the CUDA device and fails (returns
 * false) when there is none: there is no CPU fallback. The per-stage functions also initialise
 * lazily, because the reference's unit tests call them without bra_init().
 * These three are exported as WEAK symbols so that a host program that still links the
 * reference's own lib_bra.c keeps its definitions.
 */
#pragma once

#include <lib_bra_defs.h>
#include <lib_bra_types.h>
#include <utils/lib_bra_crc32c.h>

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

bool bra_init(void);
bool bra_quit(void);
bool bra_has_sse42(void);

#ifdef __cplusplus
}
#endif
