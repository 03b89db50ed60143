// host_design.hpp -- host-side planner and filter designer of the B200 rate engine.
//
// Runs once per open, in double precision, on the CPU (north_star: "coefficient design runs once on the
// host and is uploaded"). Reference behaviour followed (paths under /root/reference/):
//   convert_settings   rate/rate_base.h:674-704
//   rate_init          rate/rate_base.h:247-423   (stage decomposition: every integer of the plan)
//   dft_stage_init     rate/rate_base.h:156-192
//   lsx_design_lpf & co rate/effects_i_dsp.c:46-171
//   lsx_fir_to_phase   rate/effects_i_dsp.c:181-278
//   prepare_coefs      rate/prepare_coefs.h:18-46
#pragma once

#include <cstdint>
#include <vector>

#include "b200_ratelib.h"

namespace b200rate {

struct DftFilterDesign {
  int dft_length = 0, num_taps = 0, post_peak = 0;
  std::vector<double> taps;        // prototype after the optional phase transform
  std::vector<double> coefs_time;  // dft_length values: wrapped + scaled, before the forward transform
};

struct Design {
  rr_plan plan{};
  DftFilterDesign dft[2];
  std::vector<double> poly_bank;   // [phase][tap][order..0]
  int poly_phases = 0, poly_order = 0;
};

// RR_OK, or RR_INVPARAM for a ratio outside [1/5644.8, 5644.8] (rate/rate_base.h:528).
int build_design(const RR_config &cfg, int sample_bytes, Design &out);

// Kaiser-windowed-sinc low-pass; empty vector for a sizing-only run (Fn < 0).
std::vector<double> design_lpf(double Fp, double Fs, double Fn, double att, int &num_taps, int k, double beta);
// Cepstral linear->intermediate/minimum phase transform; may change h.size(); returns post_peak.
int fir_to_phase(std::vector<double> &h, double phase);

// Packed real FFT with the reference's fp64 rounding behaviour (Ooura radix-4 split, fft4g_dbl.c).
// Only used inside fir_to_phase, where the taps must match the reference bit for bit.
void rdft_f64_host(int n, bool inverse, double *a);

// Half-band prototype (one-sided, 8..13 coefficients); rate/rate_filters_generic.h:31-70.
const double *half_band_coefs(int num_coefs);

}  // namespace b200rate
