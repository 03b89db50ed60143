/* oracle/ref_glue.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Glue needed to link the unmodified reference `rate/` sources on Linux without an assembler:
 *  - ref_cpu_mask / ref_set_cpu_mask: lets a test build the library-global FFT tables for the
 *    generic engines (rate/rate_uni.c:134-189 picks table flavours from cpuid).
 *  - ff_fft_permute_sse / ff_fft_calc_sse: the reference implements these in yasm
 *    (rate/fft-float/fft_asm.asm, x86-32 only); no assembler exists in this image, so the C
 *    versions (rate/fft-float/fft.c:169-177,343-346) stand in. Consequently the SSE float engine
 *    must be run with tables built for sse=0 (ref_set_cpu_mask(1, x)); it is used ONLY as a speed
 *    baseline, never as a parity oracle (SURVEY.md 8c gotcha 3).
 */
#include "fft_ffmpeg.h"

int ref_cpu_mask = 0;

void ref_set_cpu_mask(int hide_sse, int hide_sse3)
{
  ref_cpu_mask = (hide_sse ? 1 : 0) | (hide_sse3 ? 2 : 0);
}

void ff_fft_permute_sse(FFTContext *s, FFTComplex *z, FFTComplex *tmp_buf)
{
  ff_fft_permute_c(s, z, tmp_buf);
}

void ff_fft_calc_sse(FFTContext *s, FFTComplex *z)
{
  ff_fft_calc_c(s, z);
}

/* rate/rate_uni.c:225 defines close_ratelib() but ratelib.h does not declare it. */
void close_ratelib(void);
void ref_reinit(int hide_sse, int hide_sse3, void (*oom)(void))
{
  extern int init_ratelib(void (*)(void));
  extern int initialized;
  if (initialized) close_ratelib();
  ref_set_cpu_mask(hide_sse, hide_sse3);
  init_ratelib(oom);
}
