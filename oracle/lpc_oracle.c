/* oracle/lpc_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the track-edge extrapolation the plugin runs on the samples it feeds to the rate engine
 * (SURVEY.md 8f rank 4): linear prediction from the first / last `prime` frames of a track, used to synthesise the
 * frames "before the beginning" and "after the end" so that the resampler's filters do not ring on a hard edge.
 *   reference: lpc/lpc.cpp:25-191 (lpc_extrapolate2 and its four helpers), lpc/lpc.h:26-38 (the two inline
 *   wrappers), foo_dsp_rate.cpp:96-101 + util.h:38-49 (how many frames are added / dropped, prime length).
 * Pinned bit for bit against the reference itself (oracle/_ref/libref_lpc.so = lpc/lpc.cpp and util.h compiled
 * where they lie, tests/test_lpc_oracle.py) and against tests/golden/lpc_reference.npz generated from it.
 *
 * Arithmetic contract (what "bit for bit" rests on; all of it IEEE round-to-nearest, no contraction):
 *   window      float:  n2 = (len+1)/2, k = ((i+1) - n2) / n2, x[i] *= 1 - k*k            lpc.cpp:84-91
 *   autocorr    double: r[j] = sum_{i=j}^{len-1} (double)x[i] * x[i-j], i ascending          lpc.cpp:101-108
 *   Levinson    double, in place, with the early exit on a vanishing error                   lpc.cpp:111-146
 *   damping     lpc[j] *= 0.999^(j+1)                                                        lpc.cpp:148-157
 *   recursion   float:  s = 0; s -= x[.] * (float)lpc[.] oldest sample first; clamp +-10     lpc.cpp:168-191
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* Welch window applied to one channel's base segment (lpc.cpp:84-91). */
static void welch(float *x, size_t len)
{
  const float half = (float)(len + 1) / 2.0f;
  for (size_t i = 0; i < len; ++i) {
    float k = ((float)((int)i + 1) - half) / half;
    float kk = k * k;
    x[i] = x[i] * (1.0f - kk);
  }
}

/* Lags 0..order of the windowed segment, each one a sequential double sum (lpc.cpp:94-109). */
static void lags(const float *x, size_t len, int order, double *r)
{
  for (int j = order; j >= 0; --j) {
    double acc = 0;
    for (size_t i = (size_t)j; i < len; ++i) {
      double p = (double)x[i] * (double)x[i - j];
      acc = acc + p;
    }
    r[j] = acc;
  }
}

/* Levinson-Durbin with the reference's regularisation, early exit and damping; returns the usable order
 * (lpc.cpp:111-166). a[] receives `order` coefficients (zeros above the returned order). */
static int levinson(const double *r, int order, double *a)
{
  double err = r[0] * (1. + 1e-10);
  const double floor_ = 1e-9 * r[0] + 1e-10;
  int used = order;
  for (int i = 0; i < order; ++i) {
    if (err < floor_) {
      for (int j = i; j < order; ++j) a[j] = 0;
      used = i;
      break;
    }
    double k = -r[i + 1];
    for (int j = 0; j < i; ++j) k -= a[j] * r[i - j];
    k /= err;
    a[i] = k;
    int j = 0;
    for (; j < i / 2; ++j) {
      double lo = a[j], hi = a[i - 1 - j];
      a[j] = lo + k * hi;
      a[i - 1 - j] = hi + k * lo;
    }
    if (i & 1) a[j] += a[j] * k;
    err *= 1.0 - k * k;
  }
  double damp = 0.999;
  for (int j = 0; j < used; ++j) {
    a[j] *= damp;
    damp *= 0.999;
  }
  if (used == 0) {
    used = 1;
    a[0] = -1;
  }
  return used;
}

static float clamp10(float s)
{
  if (s > 10.f) return 10.f;
  if (s < -10.f) return -10.f;
  return s;
}

/* x[0..len) is the (un-windowed) base; writes x[len .. len+extra) (lpc.cpp:170-181). */
static void run_forward(float *x, size_t len, size_t extra, const double *a, int used)
{
  float *w = x + len - used;
  for (size_t i = 0; i < extra; ++i) {
    float s = 0;
    for (int j = 0; j < used; ++j) {
      float p = w[i + j] * (float)a[used - 1 - j];
      s = s - p;
    }
    w[used + i] = clamp10(s);
  }
}

/* writes x[-extra .. 0) from x[0..len) (lpc.cpp:182-190): the same recursion on the reversed time axis. */
static void run_backward(float *x, size_t extra, const double *a, int used)
{
  float *w = x - 1 + used;
  for (size_t i = 0; i < extra; ++i) {
    float s = 0;
    for (int j = 0; j < used; ++j) {
      float p = *(w - (ptrdiff_t)i - j) * (float)a[used - 1 - j];
      s = s - p;
    }
    *(w - used - (ptrdiff_t)i) = clamp10(s);
  }
}

/* Analysis of one channel: r[order+1], a[order]; returns the usable order. */
int orc_lpc_analyse(const float *data, size_t data_len, int nch, int ch, int order, double *r, double *a)
{
  float *x = (float *)malloc(sizeof(float) * (data_len ? data_len : 1));
  for (size_t i = 0; i < data_len; ++i) x[i] = data[i * (size_t)nch + ch];
  welch(x, data_len);
  lags(x, data_len, order, r);
  int used = levinson(r, order, a);
  free(x);
  return used;
}

/* Same arguments and memory layout as lpc_extrapolate2 (lpc/lpc.h:4-26): `data` points at frame 0 of the base
 * segment of interleaved frames; frames [-extra_bkwd, 0) and [data_len, data_len + extra_fwd) are written. */
void orc_lpc_extrapolate2(float *data, size_t data_len, int nch, int order, size_t extra_bkwd, size_t extra_fwd)
{
  size_t total = extra_bkwd + data_len + extra_fwd;
  float *buf = (float *)malloc(sizeof(float) * (total ? total : 1));
  double *r = (double *)malloc(sizeof(double) * (order + 1));
  double *a = (double *)malloc(sizeof(double) * (order > 0 ? order : 1));
  for (int c = 0; c < nch; ++c) {
    float *x = buf + extra_bkwd;
    int used = orc_lpc_analyse(data, data_len, nch, c, order, r, a);
    for (size_t i = 0; i < data_len; ++i) x[i] = data[i * (size_t)nch + c];
    if (extra_fwd) {
      run_forward(x, data_len, extra_fwd, a, used);
      for (size_t i = data_len; i < data_len + extra_fwd; ++i) data[i * (size_t)nch + c] = x[i];
    }
    if (extra_bkwd) {
      run_backward(x, extra_bkwd, a, used);
      for (ptrdiff_t i = -(ptrdiff_t)extra_bkwd; i < 0; ++i) data[i * (ptrdiff_t)nch + c] = x[i];
    }
  }
  free(a);
  free(r);
  free(buf);
}

static unsigned gcd_u(unsigned a, unsigned b)
{
  if (!a || !b) return 0;
  while (b) {
    unsigned t = a % b;
    a = b;
    b = t;
  }
  return a;
}

/* Frames the plugin adds at each track edge (input rate) and drops from each end of the result (output rate),
 * the prime length and the size of its input block: foo_dsp_rate.cpp:96-101 with util.h:38-49 (samples_len with
 * N = 20, M = 8192: both counts span the same duration, at most 1/20 s, at most 8192 frames). */
void orc_track_edge_lengths(unsigned in_rate, unsigned out_rate, int lpc_order, unsigned *add, unsigned *drop,
                            unsigned *prime_len, unsigned *inbuf)
{
  unsigned a = in_rate, d = out_rate;
  unsigned g = gcd_u(a, d);
  if (g) {
    a /= g;
    d /= g;
    unsigned n = (g + 20 - 1) / 20;
    unsigned big = a > d ? a : d;
    if (big * n > 8192u) n = 8192u / big;
    if (n < 1) n = 1;
    a *= n;
    d *= n;
  }
  unsigned ib = in_rate / 10;
  if (ib < 2048u) ib = 2048u;
  if (ib > 65536u) ib = 65536u;
  unsigned p = in_rate / 20;
  if (p < 1024u) p = 1024u;
  if (p > 16384u) p = 16384u;
  if (p < 2u * (unsigned)lpc_order + 1) p = 2u * (unsigned)lpc_order + 1;
  *add = a;
  *drop = d;
  *prime_len = p;
  *inbuf = ib;
}
