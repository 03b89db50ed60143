// host_design.cpp -- see host_design.hpp. Everything here is double-precision host code that runs once per
// open. The arithmetic expressions deliberately keep the reference's evaluation order: the fp32 engine is
// only bit-faithful downstream if the designed taps are bit-identical.
#include "host_design.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace b200rate {

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kPi2 = 1.57079632679489661923;
constexpr double kPi4 = 0.78539816339744830962;

inline int floor_log2(unsigned v) { int l = 0; while (v >>= 1) ++l; return l; }
inline bool pow2_ge2(int x) { return x >= 2 && (x & (x - 1)) == 0; }
inline double to_dB(double x) { return std::log10(x) * 20; }          // linear_to_dB, rate/util.h
inline double to_3dB(double a) { return (1.6e-6 * a - 7.5e-4) * a + .646; }  // TO_3dB, rate_base.h:240

// ---------------------------------------------------------------------------------------------------
// fp64 packed real FFT, Ooura radix-4 decomposition (rate/fft-double/fft4g_dbl.c)
// ---------------------------------------------------------------------------------------------------

struct OouraTables { std::vector<double> w, c; };

void bit_reverse_complex(int count, double *a)   // bitrv2 == plain bit reversal of complex indices
{
  const int bits = floor_log2(static_cast<unsigned>(count));
  for (int i = 0; i < count; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    if (r > i) { std::swap(a[2 * i], a[2 * r]); std::swap(a[2 * i + 1], a[2 * r + 1]); }
  }
}

const OouraTables &ooura_tables(int n)
{
  static std::map<int, OouraTables> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(n);
  if (it != cache.end()) return it->second;
  OouraTables t;
  const int nw = n >> 2;
  t.w.assign(std::max(nw, 4), 0.0);
  t.c.assign(std::max(nw, 4), 0.0);
  if (nw > 2) {                                   // makewt, fft4g_dbl.c:159-185
    const int nwh = nw >> 1;
    const double delta = kPi2 / nw;
    t.w[0] = 1; t.w[1] = 0;
    t.w[nwh] = std::cos(kPi4); t.w[nwh + 1] = t.w[nwh];
    if (nwh > 2) {
      for (int j = 2; j < nwh; j += 2) {
        const double x = std::cos(delta * j), y = std::sin(delta * j);
        t.w[j] = x; t.w[j + 1] = y; t.w[nw - j] = y; t.w[nw - j + 1] = x;
      }
      bit_reverse_complex(nw >> 1, t.w.data());
    }
  }
  if (nw > 1) {                                   // makect, fft4g_dbl.c:188-205
    const int nch = nw >> 1;
    const double delta = kPi2 / nw;
    t.c[0] = std::cos(kPi4); t.c[nch] = 0.5 * t.c[0];
    for (int j = 1; j < nch; ++j) {
      t.c[j] = 0.5 * std::cos(delta * j);
      t.c[nw - j] = 0.5 * std::sin(delta * j);
    }
  }
  return cache.emplace(n, std::move(t)).first->second;
}

struct Quad { double x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i; };

inline Quad quad_sums(const double *a, int j, int l)
{
  const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
  return Quad{a[j] + a[j1], a[j + 1] + a[j1 + 1], a[j] - a[j1], a[j + 1] - a[j1 + 1],
              a[j2] + a[j3], a[j2 + 1] + a[j3 + 1], a[j2] - a[j3], a[j2 + 1] - a[j3 + 1]};
}

// One radix-4 layer with quarter-span l (cft1st is the l == 2 case of cftmdl; fft4g_dbl.c:462-686).
void radix4_layer(int n, int l, double *a, const double *w)
{
  const int m = l << 2, m2 = 2 * m;
  for (int j = 0; j < l; j += 2) {
    const Quad s = quad_sums(a, j, l);
    const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
    a[j] = s.x0r + s.x2r; a[j + 1] = s.x0i + s.x2i;
    a[j2] = s.x0r - s.x2r; a[j2 + 1] = s.x0i - s.x2i;
    a[j1] = s.x1r - s.x3i; a[j1 + 1] = s.x1i + s.x3r;
    a[j3] = s.x1r + s.x3i; a[j3 + 1] = s.x1i - s.x3r;
  }
  {
    const double wd = w[2];
    for (int j = m; j < l + m; j += 2) {
      const Quad s = quad_sums(a, j, l);
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      a[j] = s.x0r + s.x2r; a[j + 1] = s.x0i + s.x2i;
      a[j2] = s.x2i - s.x0i; a[j2 + 1] = s.x0r - s.x2r;
      double pr = s.x1r - s.x3i, pi = s.x1i + s.x3r;
      a[j1] = wd * (pr - pi); a[j1 + 1] = wd * (pr + pi);
      pr = s.x3i + s.x1r; pi = s.x3r - s.x1i;
      a[j3] = wd * (pi - pr); a[j3 + 1] = wd * (pi + pr);
    }
  }
  int k1 = 0;
  for (int k = m2; k < n; k += m2) {
    k1 += 2;
    const int k2 = 2 * k1;
    const double wk2r = w[k1], wk2i = w[k1 + 1];
    double wk1r = w[k2], wk1i = w[k2 + 1];
    double wk3r = wk1r - 2 * wk2i * wk1i;
    double wk3i = 2 * wk2i * wk1r - wk1i;
    for (int j = k; j < l + k; j += 2) {
      const Quad s = quad_sums(a, j, l);
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      a[j] = s.x0r + s.x2r; a[j + 1] = s.x0i + s.x2i;
      double pr = s.x0r - s.x2r, pi = s.x0i - s.x2i;
      a[j2] = wk2r * pr - wk2i * pi; a[j2 + 1] = wk2r * pi + wk2i * pr;
      pr = s.x1r - s.x3i; pi = s.x1i + s.x3r;
      a[j1] = wk1r * pr - wk1i * pi; a[j1 + 1] = wk1r * pi + wk1i * pr;
      pr = s.x1r + s.x3i; pi = s.x1i - s.x3r;
      a[j3] = wk3r * pr - wk3i * pi; a[j3 + 1] = wk3r * pi + wk3i * pr;
    }
    wk1r = w[k2 + 2]; wk1i = w[k2 + 3];
    wk3r = wk1r - 2 * wk2r * wk1i;
    wk3i = 2 * wk2r * wk1r - wk1i;
    for (int j = k + m; j < l + (k + m); j += 2) {
      const Quad s = quad_sums(a, j, l);
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      a[j] = s.x0r + s.x2r; a[j + 1] = s.x0i + s.x2i;
      double pr = s.x0r - s.x2r, pi = s.x0i - s.x2i;
      a[j2] = -wk2i * pr - wk2r * pi; a[j2 + 1] = -wk2i * pi + wk2r * pr;
      pr = s.x1r - s.x3i; pi = s.x1i + s.x3r;
      a[j1] = wk1r * pr - wk1i * pi; a[j1 + 1] = wk1r * pi + wk1i * pr;
      pr = s.x1r + s.x3i; pi = s.x1i - s.x3r;
      a[j3] = wk3r * pr - wk3i * pi; a[j3 + 1] = wk3r * pi + wk3i * pr;
    }
  }
}

// cftfsub / cftbsub, fft4g_dbl.c:308-412.
void complex_core(int n, double *a, const double *w, bool backward)
{
  int l = 2;
  if (n > 8) {
    radix4_layer(n, 2, a, w);
    for (l = 8; (l << 2) < n; l <<= 2) radix4_layer(n, l, a, w);
  }
  if ((l << 2) == n) {
    for (int j = 0; j < l; j += 2) {
      const int j1 = j + l, j2 = j1 + l, j3 = j2 + l;
      const double x0r = a[j] + a[j1], x1r = a[j] - a[j1];
      const double x2r = a[j2] + a[j3], x2i = a[j2 + 1] + a[j3 + 1];
      const double x3r = a[j2] - a[j3], x3i = a[j2 + 1] - a[j3 + 1];
      if (!backward) {
        const double x0i = a[j + 1] + a[j1 + 1], x1i = a[j + 1] - a[j1 + 1];
        a[j] = x0r + x2r; a[j + 1] = x0i + x2i;
        a[j2] = x0r - x2r; a[j2 + 1] = x0i - x2i;
        a[j1] = x1r - x3i; a[j1 + 1] = x1i + x3r;
        a[j3] = x1r + x3i; a[j3 + 1] = x1i - x3r;
      } else {
        const double x0i = -a[j + 1] - a[j1 + 1], x1i = -a[j + 1] + a[j1 + 1];
        a[j] = x0r + x2r; a[j + 1] = x0i - x2i;
        a[j2] = x0r - x2r; a[j2 + 1] = x0i + x2i;
        a[j1] = x1r - x3i; a[j1 + 1] = x1i - x3r;
        a[j3] = x1r + x3i; a[j3 + 1] = x1i + x3r;
      }
    }
  } else {
    for (int j = 0; j < l; j += 2) {
      const int j1 = j + l;
      const double x0r = a[j] - a[j1];
      double x0i;
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
}

}  // namespace

void rdft_f64_host(int n, bool inverse, double *a)   // lsx_rdft_generic, fft4g_dbl.c:26-62
{
  assert(n >= 8 && (n & (n - 1)) == 0);
  const OouraTables &t = ooura_tables(n);
  const int nc = n >> 2, m = n >> 1;
  if (!inverse) {
    bit_reverse_complex(m, a);
    complex_core(n, a, t.w.data(), false);
    for (int j = 2, kk = 1; j < m; j += 2, ++kk) {     // rftfsub
      const int k = n - j;
      const double wkr = 0.5 - t.c[nc - kk], wki = t.c[kk];
      const double xr = a[j] - a[k], xi = a[j + 1] + a[k + 1];
      const double yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
      a[j] -= yr; a[j + 1] -= yi; a[k] += yr; a[k + 1] -= yi;
    }
    const double xi = a[0] - a[1];
    a[0] += a[1]; a[1] = xi;
  } else {
    a[1] = 0.5 * (a[0] - a[1]); a[0] -= a[1];
    a[1] = -a[1];                                      // rftbsub
    for (int j = 2, kk = 1; j < m; j += 2, ++kk) {
      const int k = n - j;
      const double wkr = 0.5 - t.c[nc - kk], wki = t.c[kk];
      const double xr = a[j] - a[k], xi = a[j + 1] + a[k + 1];
      const double yr = wkr * xr + wki * xi, yi = wkr * xi - wki * xr;
      a[j] -= yr; a[j + 1] = yi - a[j + 1]; a[k] += yr; a[k + 1] = yi - a[k + 1];
    }
    a[m + 1] = -a[m + 1];
    bit_reverse_complex(m, a);
    complex_core(n, a, t.w.data(), true);
  }
}

// ---------------------------------------------------------------------------------------------------
// Kaiser-windowed sinc designer (rate/effects_i_dsp.c:46-171)
// ---------------------------------------------------------------------------------------------------
namespace {

double bessel_I0(double x)
{
  double term = 1, sum = 1, last_sum, x2 = x / 2;
  int i = 1;
  do {
    const double y = x2 / i++;
    last_sum = sum; sum += term *= y * y;
  } while (sum != last_sum);
  return sum;
}

double kaiser_beta(double att, double tr_bw)
{
  if (att >= 60) {
    static const double fit[][4] = {
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
    const int rows = static_cast<int>(sizeof(fit) / sizeof(fit[0]));
    const double realm = std::log(tr_bw / .0005) / std::log(2.);
    const double *c0 = fit[std::clamp(static_cast<int>(realm), 0, rows - 1)];
    const double *c1 = fit[std::clamp(1 + static_cast<int>(realm), 0, rows - 1)];
    const double b0 = ((c0[0] * att + c0[1]) * att + c0[2]) * att + c0[3];
    const double b1 = ((c1[0] * att + c1[1]) * att + c1[2]) * att + c1[3];
    return b0 + (b1 - b0) * (realm - static_cast<int>(realm));
  }
  if (att > 50) return .1102 * (att - 8.7);
  if (att > 20.96) return .58417 * std::pow(att - 20.96, .4) + .07886 * (att - 20.96);
  return 0;
}

std::vector<double> windowed_sinc(int num_taps, double Fc, double beta, double rho, double scale)
{
  const int m = num_taps - 1;
  std::vector<double> h(static_cast<size_t>(num_taps));
  const double mult = scale / bessel_I0(beta), mult1 = 1 / (.5 * m + rho);
  for (int i = 0; i <= m / 2; ++i) {
    const double z = i - .5 * m, x = z * kPi, y = z * mult1;
    h[i] = x ? std::sin(Fc * x) / x : Fc;
    h[i] *= bessel_I0(beta * std::sqrt(1 - y * y)) * mult;
    if (m - i != i) h[m - i] = h[i];
  }
  return h;
}

int choose_dft_length(int num_taps)   // lsx_set_dft_length, effects_i_dsp.c:64-73
{
  int result = 8;
  for (int n = num_taps; n > 2; n >>= 1) result <<= 1;
  if (result < 65536) result *= 2;
  return std::clamp(result, 2048, 131072);
}

}  // namespace

std::vector<double> design_lpf(double Fp, double Fs, double Fn, double att, int &num_taps, int k, double beta)
{
  const int n = num_taps, phases = std::max(k, 1), modulo = std::max(-k, 1);
  const double rho = phases == 1 ? .5 : att < 120 ? .63 : .75;
  Fp /= std::fabs(Fn); Fs /= std::fabs(Fn);
  double tr_bw = .5 * (Fs - Fp);
  tr_bw /= phases; Fs /= phases;
  tr_bw = std::min(tr_bw, .5 * Fs);
  const double Fc = Fs - tr_bw;
  if (beta < 0) beta = kaiser_beta(att, tr_bw * .5 / Fc);
  const double att_k = att < 60 ? (att - 7.95) / (2.285 * kPi * 2)
                                : ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902;
  if (!num_taps) num_taps = static_cast<int>(std::ceil(att_k / tr_bw + 1));
  if (!n) {
    if (phases > 1) {                               // taps per phase rounded up to a multiple of 4
      const int per_phase = ((num_taps / phases + 1) + 3) & ~3;
      num_taps = per_phase * phases - 1;
    } else
      num_taps = (num_taps + modulo - 2) / modulo * modulo + 1;
  }
  if (Fn < 0) return {};
  return windowed_sinc(num_taps, Fc, beta, rho, static_cast<double>(phases));
}

// ---------------------------------------------------------------------------------------------------
// Phase response transform of a linear-phase prototype (lsx_fir_to_phase, effects_i_dsp.c:181-278): the filter
// keeps its magnitude response and gets a phase response between minimum phase (phase 0), linear phase (50) and
// maximum phase (100). Homomorphic method: log-magnitude spectrum -> real cepstrum -> fold onto the causal side ->
// back to a (log-magnitude, phase) spectrum of the minimum-phase filter -> blend that phase with the linear one ->
// impulse response, cropped around its peak. Every expression keeps the reference's operand order (the taps are
// part of the parity contract); the transforms are the Ooura-ordered host FFT above.
// ---------------------------------------------------------------------------------------------------
namespace {

struct PhaseWork {
  int n;                           // transform length (power of two, >= 32 and > 2 * taps)
  std::vector<double> v;           // packed spectrum / sequence, n + 2 values (explicit Nyquist pair at [n], [n + 1])
  std::vector<double> half_turns;  // per bin: accumulated phase discontinuities of the prototype, in units of pi
};

void scaled_inverse(PhaseWork &w)
{
  rdft_f64_host(w.n, true, w.v.data());
  for (int i = 0; i < w.n; ++i) w.v[i] *= 2. / w.n;
}

// Spectrum of the prototype -> log magnitudes (imaginary parts cleared), counting on the way how often its phase
// jumps by a whole and by half a turn: the half-turn count at Nyquist is the prototype's group delay in samples.
void log_magnitude(PhaseWork &w)
{
  rdft_f64_host(w.n, false, w.v.data());
  w.v[w.n] = w.v[1]; w.v[w.n + 1] = w.v[1] = 0;                        // unpack Nyquist
  double last_raw = 0, turns = 0, last_unwrapped = 0, half_turns = 0;
  for (int i = 0; i <= w.n; i += 2) {
    double angle = std::atan2(w.v[i + 1], w.v[i]);
    double unit = 2 * kPi;
    double jump = angle - last_raw;
    double fix = unit * ((jump < -unit * .7) - (jump > unit * .7));
    last_raw = angle;
    turns += fix;
    angle += turns;
    unit = kPi;
    jump = angle - last_unwrapped;
    fix = unit * ((jump < -unit * .7) - (jump > unit * .7));
    last_unwrapped = angle;
    half_turns += std::fabs(fix);
    w.half_turns[i >> 1] = half_turns;
    const double mag = std::sqrt(w.v[i] * w.v[i] + w.v[i + 1] * w.v[i + 1]);
    w.v[i] = mag ? std::log(mag) : -26;
    w.v[i + 1] = 0;
  }
  w.v[1] = w.v[w.n];                                                   // pack Nyquist
}

// Real cepstrum folded onto the causal side, transformed back: v = (log magnitude, minimum phase) per bin.
void minimum_phase_spectrum(PhaseWork &w)
{
  scaled_inverse(w);
  for (int i = 1; i < w.n / 2; ++i) {
    w.v[i] *= 2;
    w.v[i + w.n / 2] = 0;
  }
  rdft_f64_host(w.n, false, w.v.data());
}

// Blend the minimum phase with the prototype's linear phase (mix = 0: minimum, 1: linear) and leave the impulse
// response of the result in v.
void blended_impulse(PhaseWork &w, double mix)
{
  const double total = w.half_turns[w.n >> 1];
  for (int i = 2; i < w.n; i += 2)
    w.v[i + 1] = mix * i / w.n * total + (1 - mix) * (w.v[i + 1] + w.half_turns[i >> 1]) - w.half_turns[i >> 1];
  w.v[0] = std::exp(w.v[0]); w.v[1] = std::exp(w.v[1]);
  for (int i = 2; i < w.n; i += 2) {
    const double mag = std::exp(w.v[i]);
    w.v[i] = mag * std::cos(w.v[i + 1]);
    w.v[i + 1] = mag * std::sin(w.v[i + 1]);
  }
  scaled_inverse(w);
}

// Index of the main lobe: where the running sum of the response peaks within the prototype's group delay, walked
// back to the local extremum.
int impulse_peak(const PhaseWork &w)
{
  int peak = 0;
  double running = 0, best = 0;
  const int last = static_cast<int>(w.half_turns[w.n >> 1] / kPi + .5);
  for (int i = 0; i <= last; ++i) {
    running += w.v[i];
    if (std::fabs(running) > std::fabs(best)) { best = running; peak = i; }
  }
  while (peak && std::fabs(w.v[peak - 1]) > std::fabs(w.v[peak]) && w.v[peak - 1] * w.v[peak] > 0) --peak;
  return peak;
}

}  // namespace

// Replaces h by the transformed filter (its length may change) and returns the number of taps behind the peak.
int fir_to_phase(std::vector<double> &h, double phase)
{
  const bool mirrored = phase > 50;                      // maximum-phase side: the minimum-phase side reversed in time
  const double mix = (mirrored ? 100 - phase : phase) / 50;
  int len = static_cast<int>(h.size());
  PhaseWork w;
  w.n = 2 * 2 * 8;
  for (int i = len; i > 1; w.n <<= 1, i >>= 1) {}
  w.v.assign(static_cast<size_t>(w.n) + 2, 0.0);
  w.half_turns.assign((static_cast<size_t>(w.n) + 2) / 2, 0.0);
  std::copy(h.begin(), h.end(), w.v.begin());

  log_magnitude(w);
  minimum_phase_spectrum(w);
  blended_impulse(w, mix);
  const int peak = impulse_peak(w);

  // crop: everything from the start (minimum phase), centred on the peak (linear), or a window around the peak whose
  // two sides grow with their share of the phase
  int begin;
  if (!mix) begin = 0;
  else if (mix == 1) begin = peak - len / 2;
  else {
    begin = static_cast<int>((.997 - (2 - mix) * .22) * len + .5);
    int end = static_cast<int>((.997 + (0 - mix) * .22) * len + .5);
    begin = peak - (begin & ~3);
    end = peak + 1 + ((end + 3) & ~3);
    len = end - begin;
    h.resize(static_cast<size_t>(len));
  }
  for (int i = 0; i < len; ++i) h[i] = w.v[(begin + (mirrored ? len - 1 - i : i) + w.n) & (w.n - 1)];
  return mirrored ? peak - begin : begin + len - (peak + 1);
}

const double *half_band_coefs(int num_coefs)
{
  static const double table[6][13] = {
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
  assert(num_coefs >= 8 && num_coefs <= 13);
  return table[num_coefs - 8];
}

// ---------------------------------------------------------------------------------------------------
// Planner
// ---------------------------------------------------------------------------------------------------
namespace {

// Attenuation a half-band family member is good for; floats in the reference (rate_filters_generic.h:255-262).
struct HalfBandChoice { int num_coefs; float att; };
const HalfBandChoice kHalfBands[] = {{8, 136.51f}, {9, 152.32f}, {10, 168.07f},
                                     {11, 183.78f}, {12, 199.44f}, {13, 212.75f}};

// Polyphase family table (rate_filters_generic.h:724-746): per candidate interpolation order the
// phase-bits scalar (a float there) and the kernel's interpolation order (-1: none).
struct PolyCandidate { float scalar; int order; };
struct PolyFamily { float beta; PolyCandidate cand[3]; };
const PolyFamily kPolyFamilies[19] = {
  {-1, {{0, 0}, {7.2f, 1}, {5.0f, 2}}},    {-1, {{0, 0}, {9.4f, 1}, {6.7f, 2}}},
  {-1, {{0, 0}, {12.4f, 1}, {7.8f, 2}}},   {-1, {{0, 0}, {13.6f, 1}, {9.3f, 2}}},
  {-1, {{0, 0}, {10.5f, 2}, {8.4f, 3}}},   {-1, {{0, 0}, {11.85f, 2}, {9.0f, 3}}},
  {-1, {{0, 0}, {8.0f, 1}, {5.3f, 2}}},    {-1, {{0, 0}, {8.6f, 1}, {5.7f, 2}}},
  {-1, {{0, 0}, {10.6f, 1}, {6.75f, 2}}},  {-1, {{0, 0}, {12.6f, 1}, {8.6f, 2}}},
  {-1, {{0, 0}, {9.6f, 2}, {7.6f, 3}}},    {-1, {{0, 0}, {11.4f, 2}, {8.65f, 3}}},
  {10.62f, {{44, 0}, {0, -1}, {0, -1}}},   {11.28f, {{12, 0}, {8, 1}, {6, 2}}},
  {-1, {{0, 0}, {9, 1}, {6, 2}}},          {-1, {{0, 0}, {11, 1}, {7, 2}}},
  {-1, {{0, 0}, {13, 1}, {8, 2}}},         {-1, {{0, 0}, {10, 2}, {8, 3}}},
  {-1, {{0, 0}, {12, 2}, {9, 3}}},
};

void plan_dft_stage(Design &D, int instance, double Fp, double Fs, double Fn, double att, double phase,
                    rr_stage_plan &st, int L, int M)
{
  DftFilterDesign &f = D.dft[instance];
  if (!f.num_taps) {
    int num_taps = 0;
    const int k = phase == 50 && pow2_ge2(L) && Fn == L ? L << 1 : 4;
    std::vector<double> h = design_lpf(Fp, Fs, Fn, att, num_taps, -k, -1.);
    if (phase != 50) { f.post_peak = fir_to_phase(h, phase); num_taps = static_cast<int>(h.size()); }
    else f.post_peak = num_taps / 2;
    f.dft_length = choose_dft_length(num_taps);
    f.coefs_time.assign(static_cast<size_t>(f.dft_length), 0.0);
    for (int i = 0; i < num_taps; ++i)
      f.coefs_time[(i + f.dft_length - num_taps + 1) & (f.dft_length - 1)] = h[i] / f.dft_length * 2 * L;
    f.num_taps = num_taps;
    f.taps = std::move(h);
  }
  st.kind = RR_STAGE_DFT;
  st.interp_order = -1;
  st.preload = f.post_peak / L;
  st.remL = f.post_peak % L;
  st.L = L;
  st.step_int = std::abs(3 - M) == 1 && Fs == 1 ? -M / 2 : M;
  st.dft_filter_num = instance;
  st.dft_length = f.dft_length; st.num_taps = f.num_taps; st.post_peak = f.post_peak;
}

// Polyphase bank [phase][tap][order..0] with finite-difference interpolation terms.
std::vector<double> poly_bank_layout(const std::vector<double> &h, int taps, int phases, int order)
{
  std::vector<double> bank(static_cast<size_t>(taps) * phases * (order + 1), 0.0);
  double fm1 = h[0], f1 = 0, f2 = 0;
  for (int i = taps - 1; i >= 0; --i)
    for (int j = phases - 1; j >= 0; --j) {
      const double f0 = fm1;
      double b = 0, c = 0, d = 0;
      const int pos = i * phases + j - 1;
      fm1 = pos > 0 ? h[pos - 1] : 0;
      if (order == 1) b = f1 - f0;
      else if (order == 2) { b = f1 - (.5 * (f2 + f0) - f1) - f0; c = .5 * (f2 + f0) - f1; }
      else if (order == 3) { c = .5 * (f1 + fm1) - f0; d = (1 / 6.) * (f2 - f1 + fm1 - f0 - 4 * c); b = f1 - f0 - d - c; }
      double *slot = &bank[(static_cast<size_t>(j) * taps + (taps - 1 - i)) * (order + 1)];
      slot[order] = f0;
      if (order > 0) slot[order - 1] = b;
      if (order > 1) slot[order - 2] = c;
      if (order > 2) slot[order - 3] = d;
      f2 = f1; f1 = f0;
    }
  return bank;
}

// ---------------------------------------------------------------------------------------------------
// Planner (rate_base.h:247-423,674-704), in four steps: what the configuration asks for (QualitySpec), how the
// rate ratio is split over stage kinds (RatioSplit), how much stop-band attenuation each stage must deliver
// (attenuation budget), and the stages themselves. The floating-point expressions keep the reference's operand
// order: the plan integers and the filter taps are part of the parity contract.
// ---------------------------------------------------------------------------------------------------

// Step 1: the switches of RR_config in numbers (convert_settings, rate_base.h:674-704).
struct QualitySpec {
  int rolloff;                 // 0: none (Best), 1: small (Normal)
  double precision_bits;       // 28 (Best) / 20 (Normal)
  double passband_pc;          // 0 dB point of the pass band, % of Nyquist
  double antialias_pc;         // 100 unless aliasing above the pass band is allowed
  double phase_pc;             // 50 = linear
};
bool quality_spec(const RR_config &cfg, QualitySpec &q)
{
  const bool best = cfg.quality == RR_best;
  q.rolloff = best ? 0 : 1;
  q.precision_bits = 16 + 4 * std::max((best ? 6 : 4) - 3, 0);
  const double rejection = q.precision_bits * to_dB(2.);
  q.passband_pc = 100 - (100 - cfg.bandwidth) / to_3dB(rejection);
  q.antialias_pc = cfg.allow_aliasing ? cfg.bandwidth : 100;
  q.phase_pc = cfg.phase;
  // the reference asserts these ranges (rate_base.h:276-280)
  return q.phase_pc >= 0 && q.phase_pc <= 100 && q.passband_pc >= 53 && q.passband_pc <= 100 && q.antialias_pc >= 85 &&
         q.antialias_pc <= 100;
}

// Step 2: in_rate / out_rate = 2^halvings * (pre_down / pre_up) * (poly_ratio / poly_up) * (post_down / post_up)
// (rate_base.h:283-310). `mode` picks the polyphase filter family; a ratio that is not rational within 2^-32 leaves
// poly_up == 1 and a fractional poly_ratio for the interpolated polyphase stages.
struct RatioSplit {
  int halvings = 0, pre_up = 1, pre_down = 1, poly_up = 1, post_up = 1, post_down = 1, mode = 0;
  double poly_ratio = 1;
  bool upsampling = false, rational = false;
};
RatioSplit split_ratio(double factor, const QualitySpec &q, int sample_bytes)
{
  constexpr double kTwo32 = 65536. * 65536.;
  constexpr int kMaxBankKB = 400;                        // max_coefs_size
  const bool allow_post_stage = true;                    // iOpt
  RatioSplit r;
  r.poly_ratio = factor;
  r.mode = q.rolloff > 1 ? (factor > 1 || q.passband_pc > (67 + 5 / 8.)) : static_cast<int>(std::ceil(2 + (q.precision_bits - 17) / 4));
  for (bool again = true; again;) {
    again = false;
    const int max_phases = r.mode ? 2048 : static_cast<int>(std::ceil(kMaxBankKB * 1000. / (44 * sample_bytes)));
    r.upsampling = r.poly_ratio < 1;
    // whole octaves of decimation become half-band stages
    r.halvings = 0;
    for (int i = static_cast<int>(r.poly_ratio * .5); i >>= 1; r.poly_ratio *= .5, ++r.halvings) {}
    r.pre_down = r.upsampling || (r.poly_ratio > 1.5 && r.poly_ratio < 2);
    r.post_down = 1 + (r.poly_ratio > 1 && r.pre_down);
    r.poly_ratio /= r.post_down;
    r.pre_up = 1 + (!r.pre_down && r.poly_ratio < 2) + (r.upsampling && r.mode);
    r.poly_ratio *= r.pre_up;
    // is what is left a ratio of small integers (to the precision of the 32.32 phase accumulator)?
    const double frac = r.poly_ratio - static_cast<int>(r.poly_ratio);
    double epsilon = 0;
    if (frac != 0) epsilon = std::fabs(std::floor(frac * kTwo32 + .5) / (frac * kTwo32) - 1);
    r.rational = !frac;
    for (int i = 1; i <= max_phases && !r.rational; ++i) {
      const double d = frac * i;
      const int nearest = static_cast<int>(d + .5);
      if ((r.rational = std::fabs(nearest / d - 1) <= epsilon)) {
        if (nearest == i) {
          r.poly_ratio = std::ceil(r.poly_ratio);
          const int extra = r.poly_ratio > 3;
          r.halvings += extra;
          r.poly_ratio /= 1 + extra;
        } else {
          r.poly_ratio = i * static_cast<int>(r.poly_ratio) + nearest;
          r.poly_up = i;
        }
      }
    }
    int up = r.pre_up * r.poly_up, down = static_cast<int>(r.poly_ratio * r.post_down);
    const int odd = (up | down) & 1;
    up >>= !odd; down >>= !odd;
    const double gain = r.pre_up * r.poly_up / r.poly_ratio;
    bool restarted = false;
    if (allow_post_stage && r.post_up == 1 && gain > 4 && gain != 5) {
      // steep up-sampling: move a power of two >= 4 into a post stage and split what is left again
      r.post_up = 4;
      for (int i = static_cast<int>(gain / 16); i >>= 1; r.post_up <<= 1) {}
      r.poly_ratio = r.poly_ratio * r.post_up / r.poly_up / r.pre_up;
      r.poly_up = 1;
      again = restarted = true;
    } else if (r.rational && (std::max(up, down) < 3 + 2 * allow_post_stage || up * down < 6 * allow_post_stage)) {
      // tiny integer ratios: one DFT stage does it all
      r.pre_up = up; r.pre_down = down;
      r.poly_ratio = 1; r.poly_up = 1; r.post_down = 1;
    }
    if (!r.mode && (!r.rational || restarted)) { ++r.mode; again = true; }
  }
  return r;
}

}  // namespace

int build_design(const RR_config &cfg, int sample_bytes, Design &D)
{
  constexpr double kTwo32 = 65536. * 65536.;
  constexpr int kMaxBankKB = 400;
  if (!cfg.in_rate || !cfg.out_rate) return RR_INVPARAM;
  const double factor = static_cast<double>(cfg.in_rate) / static_cast<double>(cfg.out_rate);
  if (factor > 5644.8 || factor < 1.0 / 5644.8) return RR_INVPARAM;
  QualitySpec q;
  if (!quality_spec(cfg, q)) return RR_INVPARAM;

  D = Design{};
  rr_plan &P = D.plan;
  P.factor = factor;
  P.sample_bytes = sample_bytes;
  P.isamp_max = 1048576;
  if (factor < 1) P.isamp_max = static_cast<uint64_t>(P.isamp_max * factor);

  RatioSplit r = split_ratio(factor, q, sample_bytes);
  const bool have_pre = r.pre_down * r.pre_up != 1, have_poly = r.poly_ratio * r.poly_up != 1,
             have_post = r.post_down * r.post_up != 1;
  const int num_stages = r.halvings + have_pre + have_poly + have_post;
  if (num_stages > RR_MAX_STAGES) return RR_INVPARAM;
  P.num_stages = num_stages;

  // Step 3: attenuation budget. Every stage adds its stop-band leakage, so each must reject a little more than the
  // target; the polyphase stage is budgeted separately (rate_base.h:321-327).
  double atten = (q.precision_bits + 1) * to_dB(2.), atten_poly = atten;
  if (num_stages > 1) {
    int others = num_stages;
    if (have_poly) { atten += to_dB(2.); atten_poly = atten; --others; }
    atten += to_dB(static_cast<double>(others));
  }
  const double trans0 = 1 - q.passband_pc / 100;         // transition band of the whole conversion, fraction of Nyquist
  const double stop_alias = 2 - q.antialias_pc / 100;
  double tighten = 1;

  // Step 4a: the half-band stages, the shortest table entry that meets the budget (rate_base.h:329-341)
  int hb = 0;
  while (hb + 1 < 6 && atten > kHalfBands[hb].att) ++hb;
  for (int i = 0; i < r.halvings; ++i) {
    rr_stage_plan &s = P.st[i];
    s.kind = RR_STAGE_HALFBAND;
    s.hb_coefs = kHalfBands[hb].num_coefs;
    s.pre_post = 4 * s.hb_coefs;
    s.preload = s.pre = s.pre_post >> 1;
    s.interp_order = -1;
  }

  // Step 4b: the DFT stage in front of the polyphase stage (rate_base.h:343-358). When a post stage follows, its
  // transition band is tightened so that the cascade keeps the 3 dB point.
  if (have_pre) {
    if (have_post) {
      const double trans3 = trans0 * to_3dB(atten);
      double x = ((2.1429e-4 - 5.2083e-7 * atten) * atten - .015863) * atten + 3.95;
      x = atten * std::pow((trans0 - trans3) / (r.post_down / (factor * r.post_up) - 1 + trans0), x);
      if (x > .035) tighten = ((4.3074e-3 - 3.9121e-4 * x) * x - .040009) * x + 1.0014;
    }
    plan_dft_stage(D, 0, 1 - trans0 * tighten, stop_alias, r.pre_down ? std::max(r.pre_up, r.pre_down) : r.poly_ratio / r.poly_up,
                   atten, q.phase_pc, P.st[r.halvings], r.pre_up, std::max(r.pre_down, 1));
  }

  // Step 4c: the polyphase stage (rate_base.h:360-411): filter family by direction and mode, band edges, then the
  // cheapest interpolation order whose coefficient bank stays under the size cap.
  if (have_poly) {
    const PolyFamily &fam = kPolyFamilies[6 * (r.upsampling + !!r.pre_down) + r.mode - !r.upsampling];
    rr_stage_plan &a = P.st[r.halvings + have_pre];
    const double mult = r.upsampling ? 1 : r.poly_up / r.poly_ratio;
    double x = .5;
    const double Fn = !r.upsampling && r.pre_down ? x = r.poly_ratio / r.poly_up : 1;
    double Fp = !r.pre_down ? mult : r.mode ? .5 : 1;
    const double Fs = 2 - Fp;
    Fp *= 1 - trans0;
    if (q.rolloff > 1 && r.mode) Fp = !r.pre_down ? mult * .5 - .125 : mult * .05 + .1;
    else if (q.rolloff == 1) Fp = Fs - (Fs - .148 * x - Fp * .852) * (.00813 * q.precision_bits + .973);

    int order = 0, taps_per_phase = static_cast<int>(fam.cand[0].scalar), phase_bits = 0, phases = 0;
    double start_phase = 0;
    int cand = (r.rational ? 0 : 1) - 1;                 // rational ratios start with the un-interpolated bank
    for (;;) {
      ++cand;
      if (fam.cand[cand].order < 0) return RR_INTERNAL;
      if (cand) { r.poly_ratio /= r.poly_up; r.poly_up = 1; r.rational = false; }
      phase_bits = static_cast<int>(std::ceil(fam.cand[cand].scalar + std::log(mult) / std::log(2.)));
      phases = !r.rational ? (1 << phase_bits) : r.poly_up;
      if (!fam.cand[0].scalar) {                           // tap count by design rather than from the table
        const int phases0 = std::max(phases, 19);
        int n0 = 0;
        design_lpf(Fp, Fs, -Fn, atten_poly, n0, phases0, fam.beta);
        taps_per_phase = n0 / phases0 + 1;
        taps_per_phase += taps_per_phase & !r.pre_down;
      }
      if ((taps_per_phase & 1) && r.rational && (r.poly_up & 1)) { phases <<= 1; r.poly_up <<= 1; r.poly_ratio *= 2; }
      start_phase = r.poly_up * .5 * (taps_per_phase & 1);
      order = cand + (cand && r.mode > 4);
      const int bank_bytes = taps_per_phase * phases * (order + 1) * sample_bytes;
      if (!(cand < 2 && fam.cand[cand + 1].order >= 0 && bank_bytes / 1000 > kMaxBankKB)) break;
    }
    int num_taps = taps_per_phase * phases - 1;
    const std::vector<double> h = design_lpf(Fp, Fs, Fn, atten_poly, num_taps, phases, fam.beta);
    D.poly_bank = poly_bank_layout(h, taps_per_phase, phases, order);
    D.poly_phases = phases; D.poly_order = order;

    a.kind = RR_STAGE_POLY;
    a.interp_order = order;
    a.pre_post = taps_per_phase - 1;
    a.preload = (taps_per_phase - 1) >> 1;
    a.n = taps_per_phase;
    a.phase_bits = phase_bits;
    a.L = r.poly_up;
    a.at = static_cast<int64_t>(start_phase * kTwo32 + .5);
    a.step = static_cast<int64_t>(r.poly_ratio * kTwo32 + .5);
  }

  // Step 4d: the DFT stage behind it (rate_base.h:412-415)
  if (have_post)
    plan_dft_stage(D, 1, 1 - (1 - (1 - trans0) * (r.upsampling ? factor * r.post_up / r.post_down : 1)) * tighten, stop_alias,
                   static_cast<double>(std::max(r.post_up, r.post_down)), atten, q.phase_pc,
                   P.st[r.halvings + have_pre + have_poly], r.post_up, r.post_down);
  return RR_OK;
}

}  // namespace b200rate
