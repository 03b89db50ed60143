// fft_tables.hpp -- host-side construction of the FFT schedule and constant tables that the kernels read.
//
// The tables are part of the fp32 bit-exactness contract: cosines are evaluated in double
// by the host libm and then rounded to the sample type exactly as ff_init_ff_cos_tabs does
// (/root/reference/rate/fft-float/fft.c:50-62), and the permutation is FFmpeg's split-radix index map
// (fft.c:82-91,156-161 with sse == 0).
#pragma once

#include <cstdint>
#include <vector>

namespace b200rate {

// Schedule of a complex FFT of M = 1 << bits points (5 <= bits <= 16: positions are 16-bit; the kernels add the padding).
struct CfftHostSched {
  int bits = 0;
  std::vector<uint16_t> leaf16_off, leaf8_off;   // permuted offsets of the register-resident leaves
  std::vector<uint16_t> gather16[2], gather8[2]; // [inverse]: transposed [element][leaf] natural indices feeding the leaf
  std::vector<uint16_t> node_off;                // node offsets of sizes 32..M, concatenated
  int level_begin[17] = {0}, level_cnt[17] = {0};
  // nodes of each size that are QUARTER children of a node four times their size (the others are the first
  // halves of a node twice their size): the fused two-level passes treat the two kinds differently
  int qchild_begin[17] = {0}, qchild_cnt[17] = {0};
  int pyr_off[17] = {0};                         // row offsets inside twiddle_pyramid()
  int pyr_len = 0;
};

CfftHostSched build_cfft_sched(int bits);

// Tables of the lane-pair kernel (rate_kernels_pk.cuh) for the same transform: the permutation folded into
// the buffer-filling writes and the task table in the layout pk_plan.hpp defines (leaf lists ordered for
// bank-conflict-free 16-byte accesses, then per phase the positions o = node offset + k of its tasks).
struct PkHostSched {
  std::vector<uint16_t> tasks;                   // pk_task_entries(bits) entries
  std::vector<uint16_t> local;                   // pk_local_entries(bits) entries: per-warp lists of the barrier-free part
  std::vector<uint16_t> perm[2];                 // [inverse][natural index] -> pk_slot(permuted position)
};
int pk_slot(int p);                              // == pslot() of rate_kernels_pk.cuh
PkHostSched build_pk_sched(const CfftHostSched &h);

// Split-radix position of natural index i (FFmpeg's split_radix_permutation).
int split_radix_index(int i, int n, int inverse);

// (T)cos(2*pi*i / 2^bits) for i = 0 .. 2^bits/4, cosine evaluated in double.
template <class T> std::vector<T> cos_quarter_table(int bits);

// Rows k = 0..S/4 of cos(2*pi*k/S) for S = 32 .. 2^bits, laid out at sched.pyr_off[log2 S].
template <class T> std::vector<T> twiddle_pyramid(const CfftHostSched &sched);

// Constants of the size-8/16 leaves: sqrt(1/2), cos(2 pi/16), cos(6 pi/16) (fft.c:300,304-318).
template <class T> void leaf_constants(T &sqrthalf, T &c16_1, T &c16_3);

// ---- tables of the fp64 DFT-stage kernel (rate_kernels_f64.cuh), all in double precision ----
// exp(sign * 2 pi i k / n), k = 0 .. count-1, as (re, im) pairs; the angle is reduced to an octant first so that
// the values carry the symmetries of the circle exactly.
std::vector<double> unit_circle(int n, int count, int sign);
// The filter spectrum the kernel multiplies with: scale * sum_j taps[j] exp(-2 pi i j k / n), k = 0 .. n/2,
// as (re, im) pairs. n is a power of two.
std::vector<double> real_spectrum(const std::vector<double> &taps, double scale);

}  // namespace b200rate
