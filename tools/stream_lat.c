/* tools/stream_lat.c -- per-call latency of the drop-in streaming path: RR_push(chunk) / RR_pull in a loop on one handle,
 * page-locked caller buffers (RRX_host_alloc), median / minimum of the two calls in microseconds.
 *   gcc -O2 -Iinclude -o tools/bin/stream_lat tools/stream_lat.c -Lfoo_dsp_resampler_b200 -lb200rate -Wl,-rpath,$PWD/foo_dsp_resampler_b200 -lm
 *   tools/bin/stream_lat 44100 48000 2 65536 float */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b200_ratelib.h"
static void oom(void) { exit(2); }
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
static int cmp(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }
int main(int argc, char **argv)
{
  size_t in_rate = argc > 1 ? atol(argv[1]) : 44100, out_rate = argc > 2 ? atol(argv[2]) : 48000;
  int nch = argc > 3 ? atoi(argv[3]) : 2;
  size_t chunk = argc > 4 ? atol(argv[4]) : 65536;
  int use_double = argc > 5 && !strcmp(argv[5], "double");
  const int reps = 200;
  RR_config cfg = {in_rate, out_rate, 50.0, 95.0, 0, RR_best};
  float *x = RRX_host_alloc(chunk * nch * sizeof(float)), *y = RRX_host_alloc(chunk * 8 * nch * sizeof(float));
  double tp[200], tl[200], t0, t1, t2, total = 0;
  size_t got, frames_out = 0;
  RR_handle *h;
  int i;
  if (!x || !y || init_ratelib(oom)) return 1;
  for (i = 0; i < (int)(chunk * nch); ++i) x[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
  h = use_double ? RR_ctor_double(&cfg, nch) : RR_ctor_float(&cfg, nch);
  if (!h) return 1;
  for (i = 0; i < 20; ++i) { RR_push(h, x, chunk); RR_pull(h, y, chunk * 8, &got); }     /* warm-up: rings grown, kernels loaded */
  for (i = 0; i < reps; ++i) {
    t0 = now_us();
    RR_push(h, x, chunk);
    t1 = now_us();
    RR_pull(h, y, chunk * 8, &got);
    t2 = now_us();
    tp[i] = t1 - t0; tl[i] = t2 - t1; total += t2 - t0; frames_out += got;
  }
  qsort(tp, reps, sizeof(double), cmp); qsort(tl, reps, sizeof(double), cmp);
  printf("%zu->%zu %dch %s chunk %zu: push median %.1f us (min %.1f), pull median %.1f us (min %.1f), %.1f Msamples/s out\n", in_rate, out_rate,
         nch, use_double ? "double" : "float", chunk, tp[reps / 2], tp[0], tl[reps / 2], tl[0], frames_out * nch / total);
  RR_close(&h);
  return 0;
}
