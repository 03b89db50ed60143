/* examples/rate_harness.c -- headless C harness over the reference's own entry points
 * (rate/ratelib.h:72-81, rate/rate_i.h:39-43). It includes only a ratelib-compatible header and links against
 * EITHER libb200rate.so (the B200 engine) or oracle/_ref/libref_rate.so (the compiled reference): the same
 * binary logic drives both, which is what "drop-in" means for this path.
 *
 *   make -C examples                     # builds harness_b200 (and harness_ref when oracle/_ref exists)
 *   ./examples/harness_b200 44100 48000 2 60 out.f32
 *
 * Synthetic input: sweep + noise. Prints frames in/out, wall time, an FNV-1a checksum of the
 * output bytes (identical between the two libraries for the fp32 engine).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200_ratelib.h"

static void oom(void) { fprintf(stderr, "out of memory\n"); exit(2); }

/* Only the compiled-reference checker exports this (oracle/ref_glue.c): its library-global FFT tables must be
 * built for the generic engines before RR_ctor_float is used (gotcha 1). */
extern void ref_set_cpu_mask(int hide_sse, int hide_sse3) __attribute__((weak));

static double now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
  size_t in_rate = argc > 1 ? (size_t)atol(argv[1]) : 44100, out_rate = argc > 2 ? (size_t)atol(argv[2]) : 48000;
  int nch = argc > 3 ? atoi(argv[3]) : 2;
  double seconds = argc > 4 ? atof(argv[4]) : 10.0;
  const char *dump = argc > 5 ? argv[5] : NULL;
  const size_t frames = (size_t)(in_rate * seconds), chunk = 65536;
  RR_config cfg = {in_rate, out_rate, 50.0, 95.0, 0, RR_best};
  float *x = malloc(frames * nch * sizeof(float));
  float *y = malloc((chunk * 4 + 65536) * nch * sizeof(float));
  uint64_t rng = 0x9E3779B97F4A7C15ull, hash = 1469598103934665603ull;
  size_t i, total = 0, got;
  RR_handle *h;
  FILE *f = dump ? fopen(dump, "wb") : NULL;
  double t0;
  int c, rc;

  for (i = 0; i < frames; ++i) {
    double t = (double)i / in_rate;
    double ph = 2 * M_PI * (20 * t + (0.45 * in_rate - 20) * t * t / (2 * seconds));
    for (c = 0; c < nch; ++c) {
      rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
      x[i * nch + c] = (float)(0.5 * sin(ph + 0.3 * c) + 0.05 * ((double)(rng >> 11) / 4503599627370496.0 - 1.0));
    }
  }
  if (ref_set_cpu_mask) ref_set_cpu_mask(1, 1);
  if (init_ratelib(oom)) return 1;
  h = RR_ctor_float(&cfg, nch);            /* fp32 engine with Best quality: BASELINE config 1 */
  if (!h) { fprintf(stderr, "RR_ctor_float failed\n"); return 1; }
  t0 = now();
  for (i = 0; i <= frames; i += chunk) {
    size_t n = frames - i < chunk ? frames - i : chunk;
    if (n && (rc = RR_push(h, x + i * nch, n))) { fprintf(stderr, "RR_push: %s\n", RR_strerror(rc)); return 1; }
    if (!n || i + chunk >= frames) RR_drain(h);
    for (;;) {
      if ((rc = RR_pull(h, y, chunk * 4, &got))) { fprintf(stderr, "RR_pull: %s\n", RR_strerror(rc)); return 1; }
      if (!got) break;
      total += got;
      { const unsigned char *p = (const unsigned char *)y; size_t k, nb = got * nch * sizeof(float);
        for (k = 0; k < nb; ++k) { hash ^= p[k]; hash *= 1099511628211ull; } }
      if (f) fwrite(y, sizeof(float), got * nch, f);
    }
    if (!n) break;
  }
  printf("%zu -> %zu Hz, %d ch: %zu frames in, %zu frames out, %.3f s, fnv1a %016llx\n", in_rate, out_rate, nch, frames,
         total, now() - t0, (unsigned long long)hash);
  RR_close(&h);
  if (f) fclose(f);
  free(x); free(y);
  return 0;
}
