/* oracle/ref_fft_tap.c -- TEST INFRASTRUCTURE. Exposes the reference's two real-FFT entry points
 * (rate/sox_i.h:42-64: lsx_safe_rdft_generic -> lsx_rdft_generic, ff_rdft_generic -> ff_rdft_calc_c)
 * with plain signatures so the oracle's transforms can be unit-tested bit for bit.
 * Requires init_ratelib() with both SSE flavours hidden (ref_reinit(1, 1, ...)). */
#include "sox_i.h"

void ref_rdft_f64(int n, int inverse, double *d)
{
  lsx_safe_rdft_generic(n, inverse ? -1 : 1, d, NULL);
}

void ref_rdft_f32(int n, int inverse, float *d)
{
  FFTComplex *tmp = (FFTComplex *)_aligned_malloc((size_t)n * sizeof(float), 32);
  ff_rdft_generic(n, inverse ? -1 : 1, d, tmp);
  _aligned_free(tmp);
}
