/* ORACLE -- TEST INFRASTRUCTURE. Stand-in for /root/reference/tg/src/platform/tg_platform.h (Win32 window, file, thread and
 * timer API) when the reference's portable C files are compiled under gcc for oracle/_ref (oracle/Makefile). Only the
 * allocation entry points the compiled files call are declared (tg_platform.h:245-248); they are defined in
 * oracle/ref_shim.c with the one property the SVO builder relies on: tgp_malloc zero-fills (tg_platform_win32.c:133). */
#ifndef TG_PLATFORM_H
#define TG_PLATFORM_H

#include "tg_common.h"

void* tgp_malloc(tg_size size);
void* tgp_realloc(tg_size size, void* p_memory);
void  tgp_free(void* p_memory);

#endif
