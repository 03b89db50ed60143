// pk_plan.hpp -- compile-time shape of the lane-pair FFT schedule (rate_kernels_pk.cuh): how the combining
// levels of an M = 1 << bits point split-radix FFT (fft.c:265-272) are grouped into phases, how many tasks
// each phase has, and where each phase's task list and twiddle rows start. Shared by the host table builder
// (fft_tables.cpp) and the device code, so both agree by construction.
#pragma once

#if defined(__CUDACC__)
#define RR_PLAN_HD __host__ __device__
#else
#define RR_PLAN_HD
#endif

namespace b200rate {

constexpr int kPkMaxPhases = 4;

// Number of nodes of size 1 << lgs in the split-radix tree of size 1 << lgm
// (fft(S) = fft(S/2) + 2 x fft(S/4), recursion stops at the size-16 / size-8 leaves).
RR_PLAN_HD constexpr int pk_nodes(int lgs, int lgm)
{
  return lgs > lgm ? 0 : lgs == lgm ? 1 : pk_nodes(lgs, lgm - 1) + 2 * pk_nodes(lgs, lgm - 2);
}
// ... of which quarter children (every node of size 2S has exactly one first-half child of size S)
RR_PLAN_HD constexpr int pk_qchildren(int lgs, int lgm) { return lgs >= lgm ? 0 : pk_nodes(lgs, lgm) - pk_nodes(lgs + 1, lgm); }

RR_PLAN_HD constexpr int pk_n16(int bits) { return pk_nodes(4, bits); }
RR_PLAN_HD constexpr int pk_n8(int bits) { return 2 * pk_nodes(5, bits); }      // size-8 leaves hang under size-32 nodes only
RR_PLAN_HD constexpr int pk_n8p(int bits) { return (pk_n8(bits) + 1) / 2; }

// Phases cover levels 5 .. bits, bottom up: three-level phases first while the transform is large enough to
// keep a 128-thread group busy with 16-value tasks, two-level phases otherwise, one level only if it must.
struct PkPhaseList { int n; int lg[kPkMaxPhases]; int depth[kPkMaxPhases]; };
RR_PLAN_HD constexpr PkPhaseList pk_phase_list(int bits)
{
  PkPhaseList p{0, {0, 0, 0, 0}, {0, 0, 0, 0}};
  int lg = 5, left = bits - 4;
  if (bits >= 11) {
    while (left >= 3 && left != 4) { p.lg[p.n] = lg; p.depth[p.n] = 3; ++p.n; lg += 3; left -= 3; }
  } else if (left & 1) {
    const int d = left >= 3 ? 3 : 1;
    p.lg[p.n] = lg; p.depth[p.n] = d; ++p.n; lg += d; left -= d;
  }
  while (left >= 2) { p.lg[p.n] = lg; p.depth[p.n] = 2; ++p.n; lg += 2; left -= 2; }
  if (left == 1) { p.lg[p.n] = lg; p.depth[p.n] = 1; ++p.n; }
  return p;
}

// Tasks of a phase: `main` tasks of the phase's full depth and `light` tasks one level shallower (nodes that
// only become part of a full-depth node in a later phase); task (node, k) with k < (1 << lg) / 4.
RR_PLAN_HD constexpr int pk_phase_main(int bits, int lg, int depth) { return pk_nodes(lg + depth - 1, bits) << (lg - 2); }
RR_PLAN_HD constexpr int pk_phase_light(int bits, int lg, int depth)
{
  return depth == 1 ? 0 : pk_qchildren(lg + depth - 2, bits) << (lg - 2);
}

// Task table of one transform (uint16 entries): leaf16 offsets, leaf8 offsets in pairs, then per phase the
// main tasks followed by the light tasks; an entry is the position o = node offset + k.
RR_PLAN_HD constexpr int pk_leaf8_base(int bits) { return pk_n16(bits); }
RR_PLAN_HD constexpr int pk_phase_base(int bits, int phase)
{
  int b = pk_n16(bits) + 2 * pk_n8p(bits);
  const PkPhaseList p = pk_phase_list(bits);
  for (int i = 0; i < phase; ++i) b += pk_phase_main(bits, p.lg[i], p.depth[i]) + pk_phase_light(bits, p.lg[i], p.depth[i]);
  return b;
}
RR_PLAN_HD constexpr int pk_task_entries(int bits) { return pk_phase_base(bits, pk_phase_list(bits).n); }

// Warp-local part of the schedule. A group has kPkWarps warps; warp w owns positions [w M/4, (w+1) M/4) --
// the subtrees of size <= M/4 of the split-radix tree -- so the leaves and every phase whose top size is
// <= M/4 run on warp-private data and need no group barrier, only __syncwarp(). The local task table starts
// with a header of 4 uint16 per (stage, warp) -- {main begin, main count, light begin, light count}, stage 0 =
// leaves (main: size-16, light: size-8), stage s = phase s-1 -- followed by the lists.
constexpr int kPkWarps = 4;
constexpr int kPkHole = 0xffff;      // entry of a leaf list that stands for an idle lane (bank-conflict padding, fft_tables.cpp)
RR_PLAN_HD constexpr int pk_local_phases(int bits)
{
  const PkPhaseList p = pk_phase_list(bits);
  int n = 0;
  for (int i = 0; i < p.n; ++i)
    if (p.lg[i] + p.depth[i] - 1 <= bits - 2) ++n;
  return n;
}
RR_PLAN_HD constexpr int pk_local_header(int bits) { return 4 * kPkWarps * (1 + pk_local_phases(bits)); }
// entries without the holes of the leaf lists; the table's real length travels with it (PkHostSched::local.size())
RR_PLAN_HD constexpr int pk_local_entries(int bits)
{
  int e = pk_local_header(bits) + pk_n16(bits) + pk_n8(bits);
  const PkPhaseList p = pk_phase_list(bits);
  for (int i = 0; i < pk_local_phases(bits); ++i) e += pk_phase_main(bits, p.lg[i], p.depth[i]) + pk_phase_light(bits, p.lg[i], p.depth[i]);
  return e;
}

// Twiddle pyramid: row of size S = 1 << lg holds cos(2 pi k / S), k = 0 .. S/4 (same layout as CfftHostSched).
RR_PLAN_HD constexpr int pk_pyr_off(int lg)
{
  int o = 0;
  for (int l = 5; l < lg; ++l) o += (1 << (l - 2)) + 1;
  return o;
}
RR_PLAN_HD constexpr int pk_pyr_len(int bits) { return pk_pyr_off(bits + 1); }

}  // namespace b200rate
