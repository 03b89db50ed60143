/* oracle/rate_oracle.c -- TEST INFRASTRUCTURE, not product code. See rate_oracle.h.
 *
 * Layout of this file
 *   1. fp32 real FFT   : FFmpeg conjugate-pair split-radix DAG, evaluated level by level
 *   2. fp64 real FFT   : Ooura radix-4 DAG
 *   3. filter designer : Kaiser-sinc low-pass, cepstral phase transform, polyphase bank layout
 *   4. planner         : stage decomposition (the integers of the parity contract)
 *   5. engine          : rate_oracle_engine.inc, instantiated for float and for double
 *   6. C API
 *
 * Reference citations are `path:line` under /root/reference/.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rate_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif
#ifndef M_PI_4
#define M_PI_4 0.78539816339744830962
#endif
#ifndef M_SQRT1_2
#define M_SQRT1_2 0.70710678118654752440
#endif

#define ORC_MAX(a, b) ((a) >= (b) ? (a) : (b))
#define ORC_MIN(a, b) ((a) <= (b) ? (a) : (b))

static void *xcalloc(size_t n, size_t s)
{
  void *p = calloc(n ? n : 1, s ? s : 1);
  if (!p) { fprintf(stderr, "rate_oracle: out of memory\n"); abort(); }
  return p;
}

static int ilog2(unsigned n) { int l = 0; while (n >>= 1) ++l; return l; }

/* ===================================================================================== */
/* 1. fp32 real FFT -- same expression DAG as rate/fft-float/fft.c + rdft.c              */
/* ===================================================================================== */

/* cos(2*pi*k/2^b), k = 0..2^b/4, rounded to float from a double cosine (fft.c:50-62). */
static float *g_costab[18];

static const float *costab(int bits)
{
  if (!g_costab[bits]) {
    int m = 1 << bits, i;
    const double freq = 2 * M_PI / m;
    float *t = xcalloc((size_t)m / 4 + 1, sizeof(float));
    for (i = 0; i <= m / 4; ++i) t[i] = (float)cos(i * freq);
    g_costab[bits] = t;
  }
  return g_costab[bits];
}

/* Index map of the conjugate-pair split-radix decomposition (fft.c:82-91). */
static int sr_index(int i, int n, int inverse)
{
  int half = n >> 1, quarter = n >> 2, sub, upper;
  if (n <= 2) return i & 1;
  if (!(i & half)) return 2 * sr_index(i, half, inverse);
  sub = sr_index(i, quarter, inverse);
  upper = (i & quarter) != 0;
  return 4 * sub + (inverse != upper ? 1 : -1);
}

typedef struct {
  int bits;           /* complex length M = 1 << bits                                    */
  int *gather[2];     /* [inverse]: permuted[i] = natural[gather[i]]  (fft.c:156-177)    */
  int *level_off[17]; /* [log2 S]: base offsets of the sub-transforms of size S          */
  int level_cnt[17];
} cfft_plan;

static cfft_plan *g_cfft[17];

static void collect_nodes(cfft_plan *pl, int size, int off, int pass)
{
  int lg = ilog2((unsigned)size);
  if (size >= 8) {                       /* DECL_FFT, fft.c:265-272 */
    collect_nodes(pl, size >> 1, off, pass);
    collect_nodes(pl, size >> 2, off + (size >> 1), pass);
    collect_nodes(pl, size >> 2, off + 3 * (size >> 2), pass);
  }
  if (pass) pl->level_off[lg][pl->level_cnt[lg]] = off;
  pl->level_cnt[lg]++;
}

static const cfft_plan *cfft_get(int bits)
{
  if (!g_cfft[bits]) {
    cfft_plan *pl = xcalloc(1, sizeof(*pl));
    int m = 1 << bits, inv, i, lg;
    pl->bits = bits;
    for (inv = 0; inv < 2; ++inv) {
      pl->gather[inv] = xcalloc((size_t)m, sizeof(int));
      for (i = 0; i < m; ++i) pl->gather[inv][i] = (-sr_index(i, m, inv)) & (m - 1);
    }
    collect_nodes(pl, m, 0, 0);
    for (lg = 0; lg <= bits; ++lg) {
      pl->level_off[lg] = xcalloc((size_t)pl->level_cnt[lg], sizeof(int));
      pl->level_cnt[lg] = 0;
    }
    collect_nodes(pl, m, 0, 1);
    g_cfft[bits] = pl;
  }
  return g_cfft[bits];
}

/* One conjugate-pair butterfly of a combining pass of size S at quarter index k
 * (TRANSFORM / TRANSFORM_ZERO + BUTTERFLIES, fft.c:200-235). z is interleaved re,im. */
static inline void sr_butterfly(float *z, int i0, int q, int k, float wre, float wim)
{
  float *a0 = z + 2 * i0, *a1 = a0 + 2 * q, *a2 = a1 + 2 * q, *a3 = a2 + 2 * q;
  float t1, t2, t3, t4, t5, t6;
  if (k == 0) {
    t1 = a2[0]; t2 = a2[1]; t5 = a3[0]; t6 = a3[1];
  } else {
    t1 = a2[0] * wre + a2[1] * wim;
    t2 = a2[1] * wre - a2[0] * wim;
    t5 = a3[0] * wre - a3[1] * wim;
    t6 = a3[0] * wim + a3[1] * wre;
  }
  t3 = t5 - t1; t5 = t5 + t1;
  a2[0] = a0[0] - t5; a0[0] = a0[0] + t5;
  a3[1] = a1[1] - t3; a1[1] = a1[1] + t3;
  t4 = t2 - t6; t6 = t2 + t6;
  a3[0] = a1[0] - t4; a1[0] = a1[0] + t4;
  a2[1] = a0[1] - t6; a0[1] = a0[1] + t6;
}

/* In-place complex FFT of permuted data; every level is a set of independent butterflies. */
static void cfft_f32(const cfft_plan *pl, float *z)
{
  int lg, j, k;
  /* size-2 leaves: only the two quarter-children of size-8 nodes (fft8, fft.c:288-301) */
  for (j = 0; j < pl->level_cnt[1]; ++j) {
    float *p = z + 2 * pl->level_off[1][j];
    float ar = p[0], ai = p[1], br = p[2], bi = p[3];
    p[0] = ar + br; p[1] = ai + bi; p[2] = ar - br; p[3] = ai - bi;
  }
  /* size-4 leaves (fft4, fft.c:274-286) */
  for (j = 0; j < pl->level_cnt[2]; ++j) {
    float *p = z + 2 * pl->level_off[2][j];
    float s01r = p[0] + p[2], d01r = p[0] - p[2], s01i = p[1] + p[3], d01i = p[1] - p[3];
    float s32r = p[6] + p[4], d32r = p[6] - p[4], s23i = p[5] + p[7], d23i = p[5] - p[7];
    p[0] = s01r + s32r; p[4] = s01r - s32r;
    p[1] = s01i + s23i; p[5] = s01i - s23i;
    p[3] = d01i + d32r; p[7] = d01i - d32r;
    p[2] = d01r + d23i; p[6] = d01r - d23i;
  }
  for (lg = 3; lg <= pl->bits; ++lg) {
    int q = 1 << (lg - 2);
    const float *tw = lg == 3 ? NULL : costab(lg);
    for (j = 0; j < pl->level_cnt[lg]; ++j)
      for (k = 0; k < q; ++k) {
        float wre, wim;
        if (lg == 3) wre = wim = (float)M_SQRT1_2;   /* fft8 uses sqrthalf, fft.c:300 */
        else if (lg == 4 && k == 2) wre = wim = (float)M_SQRT1_2; /* fft16, fft.c:315 */
        else { wre = tw[k]; wim = tw[q - k]; }
        sr_butterfly(z, pl->level_off[lg][j] + k, q, k, wre, wim);
      }
  }
}

void orc_rdft_f32(int n, int inverse, float *d)
{
  const int bits = ilog2((unsigned)n);
  const cfft_plan *pl = cfft_get(bits - 1);
  const int m = n >> 1;
  const float *tcos = costab(bits);        /* tsin[i] == tcos[n/4 - i] (mirrored table) */
  const float k1 = 0.5f, k2 = inverse ? -0.5f : 0.5f;
  float *tmp = xcalloc((size_t)n, sizeof(float));
  int i;

  assert(n >= 16 && (n & (n - 1)) == 0 && n <= 131072);
  if (!inverse) {
    for (i = 0; i < m; ++i) { tmp[2 * i] = d[2 * pl->gather[0][i]]; tmp[2 * i + 1] = d[2 * pl->gather[0][i] + 1]; }
    memcpy(d, tmp, (size_t)n * sizeof(float));
    cfft_f32(pl, d);
  }
  { float e = d[0]; d[0] = e + d[1]; d[1] = e - d[1]; }
  for (i = 1; i < (n >> 2); ++i) {             /* rdft.c:49-71 */
    int i1 = 2 * i, i2 = n - i1;
    float c = tcos[i], s = tcos[(n >> 2) - i];
    float evr = k1 * (d[i1] + d[i2]);
    float odi = k2 * (d[i2] - d[i1]);
    float evi = k1 * (d[i1 + 1] - d[i2 + 1]);
    float odr = k2 * (d[i1 + 1] + d[i2 + 1]);
    float sr, si;
    if (!inverse) { sr = odr * c + odi * s; si = odi * c - odr * s; }
    else          { sr = odr * c - odi * s; si = odi * c + odr * s; }
    d[i1] = evr + sr; d[i1 + 1] = evi + si;
    d[i2] = evr - sr; d[i2 + 1] = si - evi;
  }
  d[m + 1] = -d[m + 1];
  if (inverse) {
    d[0] *= k1; d[1] *= k1;
    for (i = 0; i < m; ++i) { tmp[2 * i] = d[2 * pl->gather[1][i]]; tmp[2 * i + 1] = d[2 * pl->gather[1][i] + 1]; }
    memcpy(d, tmp, (size_t)n * sizeof(float));
    cfft_f32(pl, d);
  }
  free(tmp);
}

/* ===================================================================================== */
/* 2. fp64 real FFT -- same expression DAG as rate/fft-double/fft4g_dbl.c                */
/* ===================================================================================== */

typedef struct { int n; double *w, *c; int *rev; } ooura_plan;
static ooura_plan *g_ooura[18];

/* bitrv2 (fft4g_dbl.c:210-305) is a plain bit reversal of the complex index. */
static void bitrev_pairs(int ncomplex, double *a)
{
  int bits = ilog2((unsigned)ncomplex), i;
  for (i = 0; i < ncomplex; ++i) {
    int r = 0, b;
    for (b = 0; b < bits; ++b) if (i >> b & 1) r |= 1 << (bits - 1 - b);
    if (r > i) {
      double tr = a[2 * i], ti = a[2 * i + 1];
      a[2 * i] = a[2 * r]; a[2 * i + 1] = a[2 * r + 1];
      a[2 * r] = tr; a[2 * r + 1] = ti;
    }
  }
}

static const ooura_plan *ooura_get(int n)
{
  int lg = ilog2((unsigned)n);
  if (!g_ooura[lg]) {
    ooura_plan *pl = xcalloc(1, sizeof(*pl));
    int nw = n >> 2, j;
    pl->n = n;
    pl->w = xcalloc((size_t)ORC_MAX(nw, 4), sizeof(double));
    pl->c = xcalloc((size_t)ORC_MAX(nw, 4), sizeof(double));
    if (nw > 2) {                              /* makewt, fft4g_dbl.c:159-185 */
      int nwh = nw >> 1;
      double delta = M_PI_2 / nw;
      pl->w[0] = 1; pl->w[1] = 0;
      pl->w[nwh] = cos(M_PI_4); pl->w[nwh + 1] = pl->w[nwh];
      if (nwh > 2) {
        for (j = 2; j < nwh; j += 2) {
          double x = cos(delta * j), y = sin(delta * j);
          pl->w[j] = x; pl->w[j + 1] = y; pl->w[nw - j] = y; pl->w[nw - j + 1] = x;
        }
        bitrev_pairs(nw >> 1, pl->w);
      }
    }
    if (nw > 1) {                              /* makect, fft4g_dbl.c:188-205 */
      int nch = nw >> 1;
      double delta = M_PI_2 / nw;
      pl->c[0] = cos(M_PI_4); pl->c[nch] = 0.5 * pl->c[0];
      for (j = 1; j < nch; ++j) {
        pl->c[j] = 0.5 * cos(delta * j);
        pl->c[nw - j] = 0.5 * sin(delta * j);
      }
    }
    g_ooura[lg] = pl;
  }
  return g_ooura[lg];
}

/* One radix-4 stage with quarter-span l doubles (cft1st is the l == 2 instance of cftmdl,
 * fft4g_dbl.c:462-686). */
static void ooura_r4_stage(int n, int l, double *a, const double *w)
{
  const int m = l << 2, m2 = 2 * m;
  int j, k, k1 = 0;
  double x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i;
#define LOAD4(j0)                                                              \
  { const int j1 = (j0) + l, j2 = j1 + l, j3 = j2 + l;                         \
    x0r = a[j0] + a[j1]; x0i = a[(j0) + 1] + a[j1 + 1];                        \
    x1r = a[j0] - a[j1]; x1i = a[(j0) + 1] - a[j1 + 1];                        \
    x2r = a[j2] + a[j3]; x2i = a[j2 + 1] + a[j3 + 1];                          \
    x3r = a[j2] - a[j3]; x3i = a[j2 + 1] - a[j3 + 1]; }
  for (j = 0; j < l; j += 2) {                 /* twiddle 1 */
    const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
    LOAD4(j)
    a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
    a[j2] = x0r - x2r; a[j2 + 1] = x0i - x2i;
    a[j1] = x1r - x3i; a[j1 + 1] = x1i + x3r;
    a[j3] = x1r + x3i; a[j3 + 1] = x1i - x3r;
  }
  {
    const double wq = w[2];                    /* twiddles on the pi/4 diagonal */
    for (j = m; j < l + m; j += 2) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      double ur, ui;
      LOAD4(j)
      a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
      a[j2] = x2i - x0i; a[j2 + 1] = x0r - x2r;
      ur = x1r - x3i; ui = x1i + x3r;
      a[j1] = wq * (ur - ui); a[j1 + 1] = wq * (ur + ui);
      ur = x3i + x1r; ui = x3r - x1i;
      a[j3] = wq * (ui - ur); a[j3 + 1] = wq * (ui + ur);
    }
  }
  for (k = m2; k < n; k += m2) {
    double wk1r, wk1i, wk2r, wk2i, wk3r, wk3i, ur, ui;
    int k2;
    k1 += 2; k2 = 2 * k1;
    wk2r = w[k1]; wk2i = w[k1 + 1];
    wk1r = w[k2]; wk1i = w[k2 + 1];
    wk3r = wk1r - 2 * wk2i * wk1i;
    wk3i = 2 * wk2i * wk1r - wk1i;
    for (j = k; j < l + k; j += 2) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      LOAD4(j)
      a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
      ur = x0r - x2r; ui = x0i - x2i;
      a[j2] = wk2r * ur - wk2i * ui; a[j2 + 1] = wk2r * ui + wk2i * ur;
      ur = x1r - x3i; ui = x1i + x3r;
      a[j1] = wk1r * ur - wk1i * ui; a[j1 + 1] = wk1r * ui + wk1i * ur;
      ur = x1r + x3i; ui = x1i - x3r;
      a[j3] = wk3r * ur - wk3i * ui; a[j3 + 1] = wk3r * ui + wk3i * ur;
    }
    wk1r = w[k2 + 2]; wk1i = w[k2 + 3];
    wk3r = wk1r - 2 * wk2r * wk1i;
    wk3i = 2 * wk2r * wk1r - wk1i;
    for (j = k + m; j < l + (k + m); j += 2) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      LOAD4(j)
      a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
      ur = x0r - x2r; ui = x0i - x2i;
      a[j2] = -wk2i * ur - wk2r * ui; a[j2 + 1] = -wk2i * ui + wk2r * ur;
      ur = x1r - x3i; ui = x1i + x3r;
      a[j1] = wk1r * ur - wk1i * ui; a[j1 + 1] = wk1r * ui + wk1i * ur;
      ur = x1r + x3i; ui = x1i - x3r;
      a[j3] = wk3r * ur - wk3i * ui; a[j3 + 1] = wk3r * ui + wk3i * ur;
    }
  }
#undef LOAD4
}

/* cftfsub / cftbsub (fft4g_dbl.c:308-412): radix-4 stages, then one radix-4 or radix-2 finish;
 * the backward flavour conjugates inside the finishing stage. */
static void ooura_cft(int n, double *a, const double *w, int backward)
{
  int j, l = 2;
  const double sg = backward ? -1. : 1.;
  if (n > 8) {
    ooura_r4_stage(n, 2, a, w);
    l = 8;
    while ((l << 2) < n) { ooura_r4_stage(n, l, a, w); l <<= 2; }
  }
  if ((l << 2) == n) {
    for (j = 0; j < l; j += 2) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      double x0r = a[j] + a[j1], x1r = a[j] - a[j1];
      double x0i, x1i;
      double x2r = a[j2] + a[j3], x2i = a[j2 + 1] + a[j3 + 1];
      double x3r = a[j2] - a[j3], x3i = a[j2 + 1] - a[j3 + 1];
      if (!backward) {
        x0i = a[j + 1] + a[j1 + 1]; x1i = a[j + 1] - a[j1 + 1];
        a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
        a[j2] = x0r - x2r; a[j2 + 1] = x0i - x2i;
        a[j1] = x1r - x3i; a[j1 + 1] = x1i + x3r;
        a[j3] = x1r + x3i; a[j3 + 1] = x1i - x3r;
      } else {
        x0i = -a[j + 1] - a[j1 + 1]; x1i = -a[j + 1] + a[j1 + 1];
        a[j] = x0r + x2r; a[j + 1] = x0i - x2i;
        a[j2] = x0r - x2r; a[j2 + 1] = x0i + x2i;
        a[j1] = x1r - x3i; a[j1 + 1] = x1i - x3r;
        a[j3] = x1r + x3i; a[j3 + 1] = x1i + x3r;
      }
    }
  } else {
    for (j = 0; j < l; j += 2) {
      const int j1 = j + l;
      double x0r = a[j] - a[j1], x0i;
      if (!backward) {
        x0i = a[j + 1] - a[j1 + 1];
        a[j] += a[j1]; a[j + 1] += a[j1 + 1];
      } else {
        x0i = -a[j + 1] + a[j1 + 1];
        a[j] += a[j1]; a[j + 1] = -a[j + 1] - a[j1 + 1];
      }
      a[j1] = x0r; a[j1 + 1] = x0i;
    }
  }
  (void)sg;
}

void orc_rdft_f64(int n, int inverse, double *a)
{
  const ooura_plan *pl = ooura_get(n);
  const int nc = n >> 2, m = n >> 1;
  int j, kk = 0;
  assert(n >= 8 && (n & (n - 1)) == 0);
  if (!inverse) {                              /* lsx_rdft_generic isgn >= 0, fft4g_dbl.c:31-45 */
    double xi;
    bitrev_pairs(m, a);
    ooura_cft(n, a, pl->w, 0);
    for (j = 2; j < m; j += 2) {               /* rftfsub, fft4g_dbl.c:415-436 */
      int k = n - j;
      double wkr, wki, xr, xim, yr, yi;
      ++kk;
      wkr = 0.5 - pl->c[nc - kk]; wki = pl->c[kk];
      xr = a[j] - a[k]; xim = a[j + 1] + a[k + 1];
      yr = wkr * xr - wki * xim; yi = wkr * xim + wki * xr;
      a[j] -= yr; a[j + 1] -= yi; a[k] += yr; a[k + 1] -= yi;
    }
    xi = a[0] - a[1]; a[0] += a[1]; a[1] = xi;
  } else {                                     /* isgn < 0, fft4g_dbl.c:47-60 */
    a[1] = 0.5 * (a[0] - a[1]); a[0] -= a[1];
    a[1] = -a[1];                              /* rftbsub, fft4g_dbl.c:439-459 */
    for (j = 2; j < m; j += 2) {
      int k = n - j;
      double wkr, wki, xr, xim, yr, yi;
      ++kk;
      wkr = 0.5 - pl->c[nc - kk]; wki = pl->c[kk];
      xr = a[j] - a[k]; xim = a[j + 1] + a[k + 1];
      yr = wkr * xr + wki * xim; yi = wkr * xim - wki * xr;
      a[j] -= yr; a[j + 1] = yi - a[j + 1]; a[k] += yr; a[k + 1] = yi - a[k + 1];
    }
    a[m + 1] = -a[m + 1];
    bitrev_pairs(m, a);
    ooura_cft(n, a, pl->w, 1);
  }
}

/* ===================================================================================== */
/* 3. filter designer -- rate/effects_i_dsp.c, rate/prepare_coefs.h                      */
/* ===================================================================================== */

void orc_free(void *p) { free(p); }

static double bessel_i0(double x)              /* effects_i_dsp.c:46-55 */
{
  double term = 1, sum = 1, last, half = x / 2;
  int i = 1;
  do {
    double y = half / i++;
    last = sum; sum += term *= y * y;
  } while (sum != last);
  return sum;
}

static int pick_dft_length(int num_taps)       /* lsx_set_dft_length, effects_i_dsp.c:64-73 */
{
  int len = 8, n = num_taps;
  for (; n > 2; n >>= 1) len <<= 1;
  if (len < 65536) len *= 2;
  if (len < 2048) len = 2048;
  if (len > 131072) len = 131072;
  assert(num_taps * 2 < len);
  return len;
}

static double kaiser_beta(double att, double tr_bw)   /* effects_i_dsp.c:83-108 */
{
  static const double fit[10][4] = {
    {-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + .001},
    {-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + .002},
    {-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + .003},
    {-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + .006},
    {8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + .015},
    {9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + .025},
    {-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + .05},
    {-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + .085},
    {1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + .1},
    {-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + .18},
  };
  if (att >= 60) {
    double realm = log(tr_bw / .0005) / log(2.);
    int r0 = (int)realm, r1 = 1 + (int)realm;
    const double *c0, *c1;
    double b0, b1;
    r0 = ORC_MIN(ORC_MAX(r0, 0), 9); r1 = ORC_MIN(ORC_MAX(r1, 0), 9);
    c0 = fit[r0]; c1 = fit[r1];
    b0 = ((c0[0] * att + c0[1]) * att + c0[2]) * att + c0[3];
    b1 = ((c1[0] * att + c1[1]) * att + c1[2]) * att + c1[3];
    return b0 + (b1 - b0) * (realm - (int)realm);
  }
  if (att > 50) return .1102 * (att - 8.7);
  if (att > 20.96) return .58417 * pow(att - 20.96, .4) + .07886 * (att - 20.96);
  return 0;
}

static double *make_lpf(int num_taps, double Fc, double beta, double rho, double scale)
{                                              /* lsx_make_lpf (dc_norm = false), effects_i_dsp.c:110-127 */
  int i, m = num_taps - 1;
  double *h = xcalloc((size_t)num_taps, sizeof(*h));
  double mult = scale / bessel_i0(beta), mult1 = 1 / (.5 * m + rho);
  assert(Fc >= 0 && Fc <= 1);
  for (i = 0; i <= m / 2; ++i) {
    double z = i - .5 * m, x = z * M_PI, y = z * mult1;
    h[i] = x ? sin(Fc * x) / x : Fc;
    h[i] *= bessel_i0(beta * sqrt(1 - y * y)) * mult;
    if (m - i != i) h[m - i] = h[i];
  }
  return h;
}

double *orc_design_lpf(double Fp, double Fs, double Fn, double att, int *num_taps, int k, double beta)
{                                              /* lsx_design_lpf, effects_i_dsp.c:137-171 */
  int n = *num_taps, phases = ORC_MAX(k, 1), modulo = ORC_MAX(-k, 1);
  double tr_bw, Fc, rho = phases == 1 ? .5 : att < 120 ? .63 : .75;
  double att_k;

  Fp /= fabs(Fn); Fs /= fabs(Fn);
  tr_bw = .5 * (Fs - Fp);
  tr_bw /= phases; Fs /= phases;
  tr_bw = ORC_MIN(tr_bw, .5 * Fs);
  Fc = Fs - tr_bw;
  assert(Fc - tr_bw >= 0);
  /* lsx_kaiser_params, effects_i_dsp.c:129-135 */
  if (beta < 0) beta = kaiser_beta(att, tr_bw * .5 / Fc);
  att_k = att < 60 ? (att - 7.95) / (2.285 * M_PI * 2)
                   : ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902;
  if (!*num_taps) *num_taps = (int)ceil(att_k / tr_bw + 1);
  if (!n) {                                    /* CREATE_4X_NUMTAPS branch, effects_i_dsp.c:156-165 */
    if (phases > 1) {
      int per_phase = *num_taps / phases + 1;
      per_phase = (per_phase + 3) & ~3;
      *num_taps = per_phase * phases - 1;
    } else
      *num_taps = (*num_taps + modulo - 2) / modulo * modulo + 1;
  }
  return Fn < 0 ? NULL : make_lpf(*num_taps, Fc, beta, rho, (double)phases);
}

void orc_fir_to_phase(double **h, int *len, int *post_len, double phase)
{                                              /* lsx_fir_to_phase, effects_i_dsp.c:181-278 */
  double *wraps, *work, phase1 = (phase > 50 ? 100 - phase : phase) / 50;
  int i, work_len, begin, end, peak = 0;
  double imp_sum = 0, peak_imp_sum = 0;
  double prev_angle2 = 0, cum_2pi = 0, prev_angle1 = 0, cum_1pi = 0;

  for (i = *len, work_len = 2 * 2 * 8; i > 1; work_len <<= 1, i >>= 1);
  assert(work_len <= 131072);                  /* larger sizes use a private table in the reference */
  work = xcalloc((size_t)work_len + 2, sizeof(*work));
  wraps = xcalloc(((size_t)work_len + 2) / 2, sizeof(*wraps));

  memcpy(work, *h, (size_t)*len * sizeof(*work));
  orc_rdft_f64(work_len, 0, work);
  work[work_len] = work[1]; work[work_len + 1] = work[1] = 0;       /* LSX_UNPACK */

  for (i = 0; i <= work_len; i += 2) {
    double angle = atan2(work[i + 1], work[i]);
    double detect = 2 * M_PI;
    double delta = angle - prev_angle2;
    double adjust = detect * ((delta < -detect * .7) - (delta > detect * .7));
    prev_angle2 = angle;
    cum_2pi += adjust;
    angle += cum_2pi;
    detect = M_PI;
    delta = angle - prev_angle1;
    adjust = detect * ((delta < -detect * .7) - (delta > detect * .7));
    prev_angle1 = angle;
    cum_1pi += fabs(adjust);
    wraps[i >> 1] = cum_1pi;
    {
      double mag = sqrt(work[i] * work[i] + work[i + 1] * work[i + 1]);
      work[i] = mag ? log(mag) : -26;
    }
    work[i + 1] = 0;
  }
  work[1] = work[work_len];                                          /* LSX_PACK */
  orc_rdft_f64(work_len, 1, work);
  for (i = 0; i < work_len; ++i) work[i] *= 2. / work_len;

  for (i = 1; i < work_len / 2; ++i) {
    work[i] *= 2;
    work[i + work_len / 2] = 0;
  }
  orc_rdft_f64(work_len, 0, work);

  for (i = 2; i < work_len; i += 2)
    work[i + 1] = phase1 * i / work_len * wraps[work_len >> 1] +
                  (1 - phase1) * (work[i + 1] + wraps[i >> 1]) - wraps[i >> 1];

  work[0] = exp(work[0]); work[1] = exp(work[1]);
  for (i = 2; i < work_len; i += 2) {
    double x = exp(work[i]);
    work[i] = x * cos(work[i + 1]);
    work[i + 1] = x * sin(work[i + 1]);
  }
  orc_rdft_f64(work_len, 1, work);
  for (i = 0; i < work_len; ++i) work[i] *= 2. / work_len;

  for (i = 0; i <= (int)(wraps[work_len >> 1] / M_PI + .5); ++i) {
    imp_sum += work[i];
    if (fabs(imp_sum) > fabs(peak_imp_sum)) { peak_imp_sum = imp_sum; peak = i; }
  }
  while (peak && fabs(work[peak - 1]) > fabs(work[peak]) && work[peak - 1] * work[peak] > 0) --peak;

  if (!phase1) begin = 0;
  else if (phase1 == 1) begin = peak - *len / 2;
  else {
    begin = (int)((.997 - (2 - phase1) * .22) * *len + .5);
    end = (int)((.997 + (0 - phase1) * .22) * *len + .5);
    begin = peak - (begin & ~3);
    end = peak + 1 + ((end + 3) & ~3);
    *len = end - begin;
    *h = realloc(*h, (size_t)*len * sizeof(**h));
  }
  for (i = 0; i < *len; ++i)
    (*h)[i] = work[(begin + (phase > 50 ? *len - 1 - i : i) + work_len) & (work_len - 1)];
  *post_len = phase > 50 ? peak - begin : begin + *len - (peak + 1);
  free(wraps); free(work);
}

/* Polyphase bank as doubles: bank[(phase * n + tap) * (order + 1) + (order - d)] holds the d-th
 * interpolation coefficient (prepare_coefs.h:18-46, multiplier = 1). */
static double *layout_poly_bank(const double *h, int n, int phases, int order)
{
  double *bank = xcalloc((size_t)n * phases * (order + 1), sizeof(*bank));
  double fm1 = h[0], f1 = 0, f2 = 0;
  int i, j;
  for (i = n - 1; i >= 0; --i)
    for (j = phases - 1; j >= 0; --j) {
      double f0 = fm1, b = 0, c = 0, d = 0;
      int pos = i * phases + j - 1;
      double *slot = bank + ((size_t)j * n + (n - 1 - i)) * (order + 1);
      fm1 = pos > 0 ? h[pos - 1] : 0;
      switch (order) {
        case 1: b = f1 - f0; break;
        case 2: b = f1 - (.5 * (f2 + f0) - f1) - f0; c = .5 * (f2 + f0) - f1; break;
        case 3: c = .5 * (f1 + fm1) - f0; d = (1 / 6.) * (f2 - f1 + fm1 - f0 - 4 * c); b = f1 - f0 - d - c; break;
        default: break;
      }
      slot[order] = f0;
      if (order > 0) slot[order - 1] = b;
      if (order > 1) slot[order - 2] = c;
      if (order > 2) slot[order - 3] = d;
      f2 = f1; f1 = f0;
    }
  return bank;
}

/* ===================================================================================== */
/* 4. planner -- rate/rate_base.h:247-423 (rate_init), :156-192 (dft_stage_init),        */
/*               :674-704 (convert_settings)                                             */
/* ===================================================================================== */

typedef struct {
  int dft_length, num_taps, post_peak;
  double *taps;        /* raw prototype (after the phase transform), num_taps doubles */
  double *coefs_time;  /* dft_length doubles: wrapped, scaled, before the forward transform */
} dft_design;

typedef struct {
  rr_plan plan;
  dft_design dft[2];
  double *poly_bank;   /* doubles, layout_poly_bank order */
  int poly_count;      /* number of elements in poly_bank */
} orc_design;

/* {att threshold (a float in the reference), one-sided coefficient count}; rate_filters_generic.h:255-262 */
static const struct { int num_coefs; float att; } k_half_firs[6] = {
  {8, 136.51f}, {9, 152.32f}, {10, 168.07f}, {11, 183.78f}, {12, 199.44f}, {13, 212.75f},
};

/* poly_firs, rate_filters_generic.h:724-746: {beta, {scalar, interpolation order of the kernel}}.
 * order -1 = no kernel. Rows 12/13 (fixed-length U100 kernels) are unreachable from RR_config. */
typedef struct { float beta; struct { float scalar; int order; } interp[3]; } poly_row;
static const poly_row k_poly_firs[19] = {
  {-1, {{0, 0}, {7.2f, 1}, {5.0f, 2}}},
  {-1, {{0, 0}, {9.4f, 1}, {6.7f, 2}}},
  {-1, {{0, 0}, {12.4f, 1}, {7.8f, 2}}},
  {-1, {{0, 0}, {13.6f, 1}, {9.3f, 2}}},
  {-1, {{0, 0}, {10.5f, 2}, {8.4f, 3}}},
  {-1, {{0, 0}, {11.85f, 2}, {9.0f, 3}}},
  {-1, {{0, 0}, {8.0f, 1}, {5.3f, 2}}},
  {-1, {{0, 0}, {8.6f, 1}, {5.7f, 2}}},
  {-1, {{0, 0}, {10.6f, 1}, {6.75f, 2}}},
  {-1, {{0, 0}, {12.6f, 1}, {8.6f, 2}}},
  {-1, {{0, 0}, {9.6f, 2}, {7.6f, 3}}},
  {-1, {{0, 0}, {11.4f, 2}, {8.65f, 3}}},
  {10.62f, {{44, 0}, {0, -1}, {0, -1}}},
  {11.28f, {{12, 0}, {8, 1}, {6, 2}}},
  {-1, {{0, 0}, {9, 1}, {6, 2}}},
  {-1, {{0, 0}, {11, 1}, {7, 2}}},
  {-1, {{0, 0}, {13, 1}, {8, 2}}},
  {-1, {{0, 0}, {10, 2}, {8, 3}}},
  {-1, {{0, 0}, {12, 2}, {9, 3}}},
};

static int is_pow2_ge2(int x) { return !(x < 2 || (x & (x - 1))); }

static void design_dft_stage(orc_design *D, unsigned instance, double Fp, double Fs, double Fn, double att,
                             double phase, rr_stage_plan *st, int L, int M)
{                                              /* dft_stage_init, rate_base.h:156-192 */
  dft_design *f = &D->dft[instance];
  if (!f->num_taps) {
    int num_taps = 0, i;
    int k = phase == 50 && is_pow2_ge2(L) && Fn == L ? L << 1 : 4;
    double *h = orc_design_lpf(Fp, Fs, Fn, att, &num_taps, -k, -1.);
    if (phase != 50) orc_fir_to_phase(&h, &num_taps, &f->post_peak, phase);
    else f->post_peak = num_taps / 2;
    f->dft_length = pick_dft_length(num_taps);
    f->coefs_time = xcalloc((size_t)f->dft_length, sizeof(double));
    for (i = 0; i < num_taps; ++i)
      f->coefs_time[(i + f->dft_length - num_taps + 1) & (f->dft_length - 1)] = h[i] / f->dft_length * 2 * L;
    f->taps = h;
    f->num_taps = num_taps;
  }
  st->kind = RR_STAGE_DFT;
  st->interp_order = -1;
  st->preload = f->post_peak / L;
  st->remL = f->post_peak % L;
  st->L = L;
  st->step_int = abs(3 - M) == 1 && Fs == 1 ? -M / 2 : M;
  st->dft_filter_num = (int)instance;
  st->dft_length = f->dft_length; st->num_taps = f->num_taps; st->post_peak = f->post_peak;
}

static int plan_build(orc_design *D, const orc_config *cfg, int sample_bytes)
{
  /* convert_settings, rate_base.h:674-704 */
  const int quality = cfg->quality == 0 ? 6 : 4;
  const int rolloff = cfg->quality == 0 ? 0 : 1;          /* rolloff_none / rolloff_small */
  const double bits = 16 + 4 * ORC_MAX(quality - 3, 0);
  const double rej = bits * (log10(2.) * 20);
  const double to3dB = (1.6e-6 * rej - 7.5e-4) * rej + .646;
  const double bw_pc = 100 - (100 - cfg->bandwidth) / to3dB;
  const double anti_aliasing_pc = cfg->allow_aliasing ? cfg->bandwidth : 100;
  const double phase = cfg->phase;
  const int interpolator = -1, max_coefs_size = 400, iOpt = 1;
  const int maintain_3dB_pt = 1;
  const double factor = (double)cfg->in_rate / (double)cfg->out_rate;
  const float mult32f = 65536.f * 65536.f;
  const double MULT32 = sample_bytes == 4 ? (double)mult32f : 65536. * 65536.;

  /* rate_init, rate_base.h:267-310 */
  double att = (bits + 1) * (log10(2.) * 20), attArb = att;
  double tbw0 = 1 - bw_pc / 100, Fs_a = 2 - anti_aliasing_pc / 100;
  double arbM = factor, tbw_tighten = 1;
  int n = 0, i, preL = 1, preM = 1, shift = 0, arbL = 1, postL = 1, postM = 1;
  int upsample = 0, rational = 0;
  int mode = rolloff > 1 ? (factor > 1 || bw_pc > (67 + 5 / 8.)) : (int)ceil(2 + (bits - 17) / 4);
  int have_pre, have_arb, have_post, num_stages, hb;
  rr_plan *P = &D->plan;
  rr_stage_plan *s;

  if (factor > 5644.8 || factor < 1.0 / 5644.8) return -1;   /* rate_base.h:528 */
  memset(P, 0, sizeof(*P));
  P->factor = factor;
  P->sample_bytes = sample_bytes;
  P->isamp_max = 1048576;
  if (factor < 1) P->isamp_max = (uint64_t)(P->isamp_max * factor);   /* rate_base.h:531 */

  while (!n++) {
    int try_i, L, M, x, maxL = interpolator > 0 ? 1 : mode ? 2048 :
        (int)ceil(max_coefs_size * 1000. / (44 * sample_bytes));
    double d, epsilon = 0, frac;
    upsample = arbM < 1;
    for (i = (int)(arbM * .5), shift = 0; i >>= 1; arbM *= .5, ++shift);
    preM = upsample || (arbM > 1.5 && arbM < 2);
    postM = 1 + (arbM > 1 && preM); arbM /= postM;
    preL = 1 + (!preM && arbM < 2) + (upsample && mode); arbM *= preL;
    if ((frac = arbM - (int)arbM) != 0)
      epsilon = fabs(floor(frac * MULT32 + .5) / (frac * MULT32) - 1);
    for (i = 1, rational = !frac; i <= maxL && !rational; ++i) {
      d = frac * i; try_i = (int)(d + .5);
      if ((rational = fabs(try_i / d - 1) <= epsilon)) {
        if (try_i == i) {
          arbM = ceil(arbM); x = arbM > 3; shift += x; arbM /= 1 + x;
        } else { arbM = i * (int)arbM + try_i; arbL = i; }
      }
    }
    L = preL * arbL; M = (int)(arbM * postM); x = (L | M) & 1; L >>= !x; M >>= !x;
    if (iOpt && postL == 1 && (d = preL * arbL / arbM) > 4 && d != 5) {
      for (postL = 4, i = (int)(d / 16); i >>= 1; postL <<= 1);
      arbM = arbM * postL / arbL / preL; arbL = 1; n = 0;
    } else if (rational && (ORC_MAX(L, M) < 3 + 2 * iOpt || L * M < 6 * iOpt)) {
      preL = L; preM = M; arbM = arbL = postM = 1;
    }
    if (!mode && (!rational || !n)) { ++mode; n = 0; }
  }

  have_pre = preM * preL != 1;
  have_arb = arbM * arbL != 1;
  have_post = postM * postL != 1;
  num_stages = shift + have_pre + have_arb + have_post;
  if (num_stages > RR_MAX_STAGES) return -1;
  P->num_stages = num_stages;

  if ((n = num_stages) > 1) {                  /* attenuation budget, rate_base.h:317-321 */
    if (have_arb) { att += log10(2.) * 20; attArb = att; --n; }
    att += log10((double)n) * 20;
  }

  for (hb = 0; hb + 1 < 6 && att > k_half_firs[hb].att; ++hb);
  for (i = 0, s = P->st; i < shift; ++i, ++s) {
    s->kind = RR_STAGE_HALFBAND;
    s->hb_coefs = k_half_firs[hb].num_coefs;
    s->pre_post = 4 * s->hb_coefs;
    s->preload = s->pre = s->pre_post >> 1;
    s->interp_order = -1;
  }

  if (have_pre) {                              /* rate_base.h:330-341 */
    if (maintain_3dB_pt && have_post) {
      double tbw3 = tbw0 * ((1.6e-6 * att - 7.5e-4) * att + .646);
      double x = ((2.1429e-4 - 5.2083e-7 * att) * att - .015863) * att + 3.95;
      x = att * pow((tbw0 - tbw3) / (postM / (factor * postL) - 1 + tbw0), x);
      if (x > .035) tbw_tighten = ((4.3074e-3 - 3.9121e-4 * x) * x - .040009) * x + 1.0014;
    }
    design_dft_stage(D, 0, 1 - tbw0 * tbw_tighten, Fs_a, preM ? ORC_MAX(preL, preM) : arbM / arbL, att, phase,
                     &P->st[shift], preL, ORC_MAX(preM, 1));
  }

  if (have_arb) {                              /* rate_base.h:350-410 */
    const poly_row *f = &k_poly_firs[6 * (upsample + !!preM) + mode - !upsample];
    int order, num_coefs = (int)f->interp[0].scalar, phase_bits, phases, coefs_size;
    double x = .5, at, Fp, Fs, Fn, mult = upsample ? 1 : arbL / arbM;
    rr_stage_plan *a = &P->st[shift + have_pre];

    Fn = !upsample && preM ? x = arbM / arbL : 1;
    Fp = !preM ? mult : mode ? .5 : 1;
    Fs = 2 - Fp;
    Fp *= 1 - tbw0;
    if (rolloff > 1 && mode) Fp = !preM ? mult * .5 - .125 : mult * .05 + .1;
    else if (rolloff == 1) Fp = Fs - (Fs - .148 * x - Fp * .852) * (.00813 * bits + .973);

    i = (interpolator < 0 ? !rational : ORC_MAX(interpolator, !rational)) - 1;
    do {
      ++i;
      assert(f->interp[i].order >= 0);
      if (i) { arbM /= arbL; arbL = 1; rational = 0; }
      phase_bits = (int)ceil(f->interp[i].scalar + log(mult) / log(2.));
      phases = !rational ? (1 << phase_bits) : arbL;
      if (!f->interp[0].scalar) {
        int phases0 = ORC_MAX(phases, 19), n0 = 0;
        orc_design_lpf(Fp, Fs, -Fn, attArb, &n0, phases0, f->beta);
        num_coefs = n0 / phases0 + 1; num_coefs += num_coefs & !preM;
      }
      if ((num_coefs & 1) && rational && (arbL & 1)) { phases <<= 1; arbL <<= 1; arbM *= 2; }
      at = arbL * .5 * (num_coefs & 1);
      order = i + (i && mode > 4);
      coefs_size = num_coefs * phases * (order + 1) * sample_bytes;
    } while (interpolator < 0 && i < 2 && f->interp[i + 1].order >= 0 && coefs_size / 1000 > max_coefs_size);

    {
      int num_taps = num_coefs * phases - 1;
      double *h = orc_design_lpf(Fp, Fs, Fn, attArb, &num_taps, phases, f->beta);
      D->poly_bank = layout_poly_bank(h, num_coefs, phases, order);
      D->poly_count = num_coefs * phases * (order + 1);
      free(h);
    }
    a->kind = RR_STAGE_POLY;
    a->interp_order = f->interp[i].order;
    assert(a->interp_order == order);
    a->pre_post = num_coefs - 1;
    a->preload = (num_coefs - 1) >> 1;
    a->n = num_coefs; assert(a->n % 4 == 0);
    a->phase_bits = phase_bits;
    a->L = arbL;
    a->at = (int64_t)(at * MULT32 + .5);
    a->step = (int64_t)(arbM * MULT32 + .5);
  }

  if (have_post)                               /* rate_base.h:412-415 */
    design_dft_stage(D, 1, 1 - (1 - (1 - tbw0) * (upsample ? factor * postL / postM : 1)) * tbw_tighten, Fs_a,
                     (double)ORC_MAX(postL, postM), att, phase, &P->st[shift + have_pre + have_arb], postL, postM);
  return 0;
}

/* Half-band prototype coefficients (data of the algorithm; rate_filters_generic.h:31-70). */
static const double k_half_fir_coefs[6][13] = {
  {0.3115465451887802, -0.08734497241282892, 0.03681452335604365, -0.01518925831569441,
   0.005454118437408876, -0.001564400922162005, 0.0003181701445034203, -3.48001341225749e-5},
  {0.3122703613711853, -0.08922155288172305, 0.03913974805854332, -0.01725059723447163,
   0.006858970092378141, -0.002304518467568703, 0.0006096426006051062, -0.0001132393923815236,
   1.119795386287666e-5},
  {0.3128545521327376, -0.09075671986104322, 0.04109637155154835, -0.01906629512749895,
   0.008184039342054333, -0.0030766775017262, 0.0009639607022414314, -0.0002358552746579827,
   4.025184282444155e-5, -3.629779111541012e-6},
  {0.3133358837508807, -0.09203588680609488, 0.04276515428384758, -0.02067356614745591,
   0.00942253142371517, -0.003856330993895144, 0.001363470684892284, -0.0003987400965541919,
   9.058629923971627e-5, -1.428553070915318e-5, 1.183455238783835e-6},
  {0.3137392991811407, -0.0931182192961332, 0.0442050575271454, -0.02210391200618091,
   0.01057473015666001, -0.00462766983973885, 0.001793630226239453, -0.0005961819959665878,
   0.0001631475979359577, -3.45557865639653e-5, 5.06188341942088e-6, -3.877010943315563e-7},
  {0.3140822554324578, -0.0940458550886253, 0.04545990399121566, -0.02338339450796002,
   0.01164429409071052, -0.005380686021429845, 0.002242915773871009, -0.000822047600000082,
   0.0002572510962395222, -6.607320708956279e-5, 1.309926399120154e-5, -1.790719575255006e-6,
   1.27504961098836e-7},
};

/* ===================================================================================== */
/* 5. engine, once per sample type                                                       */
/* ===================================================================================== */

#define SAMPLE float
#define SFX(name) name##_f32
#define RDFT(n, inv, d) orc_rdft_f32(n, inv, d)
#define SAMPLE_IS_FLOAT 1
#include "rate_oracle_engine.inc"
#undef SAMPLE
#undef SFX
#undef RDFT
#undef SAMPLE_IS_FLOAT

#define SAMPLE double
#define SFX(name) name##_f64
#define RDFT(n, inv, d) orc_rdft_f64(n, inv, d)
#define SAMPLE_IS_FLOAT 0
#include "rate_oracle_engine.inc"
#undef SAMPLE
#undef SFX
#undef RDFT
#undef SAMPLE_IS_FLOAT

/* ===================================================================================== */
/* 6. C API                                                                              */
/* ===================================================================================== */

struct orc_handle {
  int sample_bytes;
  void *eng;
  orc_design design;
};

orc_handle *orc_open(const orc_config *cfg, int nchannels, int sample_bytes)
{
  orc_handle *h = xcalloc(1, sizeof(*h));
  h->sample_bytes = sample_bytes;
  if (plan_build(&h->design, cfg, sample_bytes)) { free(h); return NULL; }
  h->eng = sample_bytes == 4 ? (void *)engine_open_f32(&h->design, cfg, nchannels)
                             : (void *)engine_open_f64(&h->design, cfg, nchannels);
  return h;
}

void orc_close(orc_handle *h)
{
  int i;
  if (!h) return;
  if (h->sample_bytes == 4) engine_close_f32(h->eng); else engine_close_f64(h->eng);
  for (i = 0; i < 2; ++i) { free(h->design.dft[i].taps); free(h->design.dft[i].coefs_time); }
  free(h->design.poly_bank);
  free(h);
}

#define DISPATCH(call32, call64) (h->sample_bytes == 4 ? call32 : call64)

void orc_keep_history(orc_handle *h, int keep) { DISPATCH(engine_keep_f32(h->eng, keep), engine_keep_f64(h->eng, keep)); }
size_t orc_push(orc_handle *h, const float *x, size_t n) { return DISPATCH(engine_push_f32(h->eng, x, n), engine_push_f64(h->eng, x, n)); }
size_t orc_pull(orc_handle *h, float *y, size_t n) { return DISPATCH(engine_pull_f32(h->eng, y, NULL, n), engine_pull_f64(h->eng, y, NULL, n)); }
size_t orc_pull_native(orc_handle *h, void *y, size_t n) { return DISPATCH(engine_pull_f32(h->eng, NULL, y, n), engine_pull_f64(h->eng, NULL, y, n)); }
void orc_drain(orc_handle *h) { DISPATCH(engine_drain_f32(h->eng), engine_drain_f64(h->eng)); }
int orc_plan_dump(const orc_handle *h, rr_plan *out) { *out = h->design.plan; return 0; }
int orc_dft_coefs(const orc_handle *h, int inst, void *out, int max_n) { return DISPATCH(engine_dft_coefs_f32(h->eng, inst, out, max_n), engine_dft_coefs_f64(h->eng, inst, out, max_n)); }
int orc_poly_coefs(const orc_handle *h, void *out, int max_n) { return DISPATCH(engine_poly_coefs_f32(h->eng, out, max_n), engine_poly_coefs_f64(h->eng, out, max_n)); }
uint64_t orc_fifo_written(const orc_handle *h, int i) { return DISPATCH(engine_written_f32(h->eng, i), engine_written_f64(h->eng, i)); }
uint64_t orc_fifo_consumed(const orc_handle *h, int i) { return DISPATCH(engine_consumed_f32(h->eng, i), engine_consumed_f64(h->eng, i)); }
size_t orc_fifo_read(const orc_handle *h, int ch, int i, uint64_t start, size_t count, void *out)
{ return DISPATCH(engine_fifo_read_f32(h->eng, ch, i, start, count, out), engine_fifo_read_f64(h->eng, ch, i, start, count, out)); }

int orc_dft_taps(const orc_handle *h, int inst, double *out, int max_n)
{
  const dft_design *f = &h->design.dft[inst];
  int n = ORC_MIN(f->num_taps, max_n);
  if (n > 0) memcpy(out, f->taps, (size_t)n * sizeof(double));
  return f->num_taps;
}
