/* examples/multi_harness.c -- headless C program over the multi-GPU entry points of libb200rate.so (RRX_multi_*):
 * what a C host (the plugin's dsp adapter through chain.h:22-43, or any batch converter) does to use every GPU of
 * the box. No Python, no torch: plain C, host buffers.
 *
 *   make -C examples harness_multi
 *   ./examples/harness_multi batch  48000 44100 2 512 10 8     # 512 stereo streams of 10 s sharded over 8 GPUs
 *   ./examples/harness_multi stream 384000 48000 8 1 600 8     # one 10-minute 8-channel stream, time-chunked over 8 GPUs
 *
 * Prints frames in / out, wall time of the conversion (host buffers to host buffers), output Msamples/s and an FNV-1a
 * checksum of the result, which does not depend on the number of GPUs (the results are bit-identical). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200_ratelib.h"

static double now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
  const char *mode = argc > 1 ? argv[1] : "batch";
  size_t in_rate = argc > 2 ? (size_t)atol(argv[2]) : 48000, out_rate = argc > 3 ? (size_t)atol(argv[3]) : 44100;
  int nch = argc > 4 ? atoi(argv[4]) : 2;
  size_t nstreams = argc > 5 ? (size_t)atol(argv[5]) : 64;
  double seconds = argc > 6 ? atof(argv[6]) : 10.0;
  int ngpus = argc > 7 ? atoi(argv[7]) : 1, reps = argc > 8 ? atoi(argv[8]) : 2;
  const int stream_mode = !strcmp(mode, "stream");
  const size_t frames = (size_t)(in_rate * seconds);
  RR_config cfg = {in_rate, out_rate, 50.0, 95.0, 0, RR_best};
  int devices[64], k, rc, r;
  RRX_multi *m = NULL;
  size_t nout = 0, i, total_in, total_out;
  float *x, *y;
  uint64_t rng = 0x9E3779B97F4A7C15ull, hash = 1469598103934665603ull;
  double best = 1e30;

  if (ngpus < 1 || ngpus > 64) return 1;
  if (stream_mode) nstreams = 1;
  for (k = 0; k < ngpus; ++k) devices[k] = k;
  /* one long stream: a call converts at most 60 s of input */
  rc = RRX_multi_open(&cfg, 4, nch, nstreams, stream_mode ? (size_t)(in_rate * 60) + 65536 : frames, devices, ngpus, &m);
  if (rc) { fprintf(stderr, "RRX_multi_open: %s (%s)\n", RR_strerror(rc), RRX_last_error()); return 1; }
  nout = RRX_multi_frames_out(m, frames);
  total_in = nstreams * frames * (size_t)nch;
  total_out = nstreams * nout * (size_t)nch;
  x = RRX_host_alloc(total_in * sizeof(float));           /* page-locked: transfers run at the speed of the host link */
  y = RRX_host_alloc((total_out + 16) * sizeof(float));
  if (!x || !y) { fprintf(stderr, "host allocation failed\n"); return 1; }
  for (i = 0; i < total_in; ++i) {                 /* noise is enough here: parity is the test-suite's job */
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    x[i] = (float)(0.5 * ((double)(rng >> 11) / 4503599627370496.0 - 1.0));
  }
  for (r = 0; r < reps; ++r) {
    const double t0 = now();
    size_t got = nout;
    rc = stream_mode ? RRX_multi_process_stream_host(m, x, frames, y, &got) : RRX_multi_process_host(m, x, frames, y);
    if (rc) { fprintf(stderr, "conversion failed: %s (%s)\n", RR_strerror(rc), RRX_last_error()); return 1; }
    if (now() - t0 < best) best = now() - t0;
    nout = got;
  }
  { const unsigned char *p = (const unsigned char *)y; size_t nb = total_out * sizeof(float);
    for (i = 0; i < nb; i += 97) { hash ^= p[i]; hash *= 1099511628211ull; } }
  printf("{\"mode\": \"%s\", \"in_rate\": %zu, \"out_rate\": %zu, \"channels\": %d, \"streams\": %zu, \"gpus\": %d, \"frames_in\": %zu, "
         "\"frames_out\": %zu, \"seconds\": %.4f, \"out_Msamples_per_s\": %.1f, \"fnv1a_sampled\": \"%016llx\"}\n",
         mode, in_rate, out_rate, nch, nstreams, ngpus, frames, nout, best, total_out / best / 1e6, (unsigned long long)hash);
  RRX_multi_close(&m);
  RRX_host_free(x); RRX_host_free(y);
  return 0;
}
