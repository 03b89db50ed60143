/* oracle/shim/intrin.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Force-included (-include) replacement for MSVC's <intrin.h> so that the UNMODIFIED
 * reference sources under /root/reference/rate compile with gcc on Linux (the reference
 * is MSVC/Win32-only: rate/sox_i.h:19, rate/rate_uni.c:18 include <intrin.h>).
 *
 * What it supplies:
 *   _BitScanReverse            (rate/sox_i.h:21-26, dlog2)
 *   __cpuid(info, leaf)        (rate/rate_uni.c:114-132) with a test-controlled mask so the
 *                              library-global FFT tables can be built for the *generic* engines
 *                              (gotcha 1)
 *   _aligned_malloc/_aligned_free/_aligned_realloc
 *                              (rate/xmalloc.c:60-71, rate/fft-float/fft_ffmpeg.h:33-41,
 *                               rate/fft-double/fft4g_dbl.c:100-126)
 *   _MM_ALIGN16, __cdecl       (rate/rate_filters_generic.h:31, rate/rate_base.h:194)
 */
#ifndef ORACLE_SHIM_INTRIN_H
#define ORACLE_SHIM_INTRIN_H

#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <malloc.h>
#include <cpuid.h>
#include <x86intrin.h>

#define __cdecl
#ifndef _MM_ALIGN16
#define _MM_ALIGN16 __attribute__((aligned(16)))
#endif

/* bit0: hide SSE (leaf 1, EDX bit 25); bit1: hide SSE3 (leaf 1, ECX bit 0).
 * Defined in oracle/ref_glue.c; set through ref_set_cpu_mask(). */
extern int ref_cpu_mask;

static inline unsigned char _BitScanReverse(unsigned long *index, unsigned long mask)
{
  if (!mask) return 0;
  *index = (unsigned long)(63 - __builtin_clzl(mask));
  return 1;
}

static inline void shim_cpuid(int info[4], int leaf)
{
  unsigned a = 0, b = 0, c = 0, d = 0;
  __get_cpuid((unsigned)leaf, &a, &b, &c, &d);
  if (leaf == 1) {
    if (ref_cpu_mask & 1) d &= ~(1u << 25);
    if (ref_cpu_mask & 2) c &= ~1u;
  }
  info[0] = (int)a; info[1] = (int)b; info[2] = (int)c; info[3] = (int)d;
}
#undef __cpuid
#define __cpuid(info, leaf) shim_cpuid(info, leaf)

static inline void *_aligned_malloc(size_t size, size_t align)
{
  void *p = NULL;
  if (align < sizeof(void *)) align = sizeof(void *);
  if (posix_memalign(&p, align, size ? size : 1)) return NULL;
  return p;
}

static inline void _aligned_free(void *p) { free(p); }

static inline void *_aligned_realloc(void *p, size_t size, size_t align)
{
  void *q = _aligned_malloc(size, align);
  if (q && p) {
    size_t old = malloc_usable_size(p);
    memcpy(q, p, old < size ? old : size);
    free(p);
  }
  return q;
}

#endif
