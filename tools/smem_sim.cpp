// tools/smem_sim.cpp -- offline model of the shared-memory wavefronts of the lane-pair DFT kernel.
//
// Shared memory moves 128 bytes per wavefront; a 16-byte access of a warp is served a quarter-warp (8 lanes) at a
// time and costs as many wavefronts as the most loaded 16-byte bank group (slot mod 8) among the active lanes; an
// 8-byte access is served a half-warp at a time (16 bank pairs). The access patterns are pure functions of the
// host-built schedule tables (fft_tables.cpp), so the count can be reproduced without a GPU and compared with
// ncu's l1tex__data_pipe_lsu_wavefronts_mem_shared (profiles/). Build: see tools/Makefile-less one-liner in
// profiles/README.md.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "fft_tables.hpp"
#include "pk_plan.hpp"

using namespace b200rate;

static std::function<int(int)> g_slot;

// wavefronts of one warp-wide 16-byte access; slot < 0: lane inactive
static int wf16(const int *slot)
{
  int total = 0;
  for (int q = 0; q < 4; ++q) {
    int cnt[8] = {0}, mx = 0;
    for (int l = 0; l < 8; ++l) {
      const int s = slot[q * 8 + l];
      if (s >= 0) mx = std::max(mx, ++cnt[s & 7]);
    }
    total += mx;
  }
  return total;
}

struct Tally { long long wf = 0, ideal_x8 = 0; void add(const int *slot) { wf += wf16(slot); for (int l = 0; l < 32; ++l) if (slot[l] >= 0) ++ideal_x8; }
               double ideal() const { return ideal_x8 / 8.0; } };

static void report(const char *name, const Tally &t) { printf("  %-34s %8lld wavefronts (ideal %8.1f, x%.2f)\n", name, t.wf, t.ideal(), t.ideal() > 0 ? t.wf / t.ideal() : 0.0); }

// leaves + local phases + top phase of one transform (reads and writes have the same pattern)
static long long transform(int bits, const PkHostSched &ps, bool verbose, const char *tag)
{
  long long total = 0;
  const uint16_t *ltab = ps.local.data();
  const PkPhaseList pl = pk_phase_list(bits);
  const int nlocal = pk_local_phases(bits);
  Tally leaves, local, top;
  for (int w = 0; w < kPkWarps; ++w) {
    const uint16_t *hd = ltab + 4 * w;
    for (int kind = 0; kind < 2; ++kind) {
      const int b0 = hd[2 * kind], cnt = hd[2 * kind + 1], nv = kind ? 8 : 16;
      for (int t0 = 0; t0 < cnt; t0 += 32)
        for (int e = 0; e < nv; ++e) {
          int slot[32];
          for (int l = 0; l < 32; ++l) { const int t = t0 + l; const int off = t < cnt ? ltab[b0 + t] : 0xffff; slot[l] = off == 0xffff ? -1 : g_slot(off) + e; }
          leaves.add(slot);
        }
    }
    for (int ph = 0; ph < nlocal; ++ph) {
      const uint16_t *h2 = ltab + 4 * ((1 + ph) * kPkWarps + w);
      const int lg = pl.lg[ph], q = 1 << (lg - 2);
      for (int kind = 0; kind < 2; ++kind) {
        const int b0 = h2[2 * kind], cnt = h2[2 * kind + 1];
        const int depth = kind ? (pl.depth[ph] > 1 ? pl.depth[ph] - 1 : 1) : pl.depth[ph], nv = 4 << (depth - 1);
        for (int t0 = 0; t0 < cnt; t0 += 32)
          for (int j = 0; j < nv; ++j) {
            int slot[32];
            for (int l = 0; l < 32; ++l) { const int t = t0 + l; slot[l] = t < cnt ? g_slot(ltab[b0 + t] + j * q) : -1; }
            local.add(slot);
          }
      }
    }
  }
  {
    const int ph = pl.n - 1, lg = pl.lg[ph], d = pl.depth[ph], q = 1 << (lg - 2), nv = 4 << (d - 1);
    const int nmain = pk_phase_main(bits, lg, d);
    for (int t0 = 0; t0 < nmain; t0 += 32)
      for (int j = 0; j < nv; ++j) {
        int slot[32];
        for (int l = 0; l < 32; ++l) { const int t = t0 + l; slot[l] = t < nmain ? g_slot(t + j * q) : -1; }
        top.add(slot);
      }
  }
  if (verbose) {
    printf(" %s transform, %d complex points: per direction (load or store)\n", tag, 1 << bits);
    report("leaves", leaves); report("warp-local phases", local); report("top phase", top);
  }
  total = 2 * (leaves.wf + local.wf) + top.wf;     // top: loads only when it sinks; caller adds stores
  return total;
}

int main(int argc, char **argv)
{
  const int fb = argc > 1 ? atoi(argv[1]) : 10, ib = argc > 2 ? atoi(argv[2]) : 11;
  const int variant = argc > 3 ? atoi(argv[3]) : 0;
  if (variant == 0) g_slot = [](int p) { return p + (p >> 4) + (p >> 8); };
  else if (variant == 1) g_slot = [](int p) { return p + (p >> 4) + (p >> 7); };
  else if (variant == 2) g_slot = [](int p) { return p + (p >> 4) + (p >> 7) + (p >> 10); };
  else g_slot = [](int p) { return p + (p >> 3); };
  const CfftHostSched hf = build_cfft_sched(fb), hi = build_cfft_sched(ib);
  const PkHostSched pf = build_pk_sched(hf), pi = build_pk_sched(hi);
  long long f = transform(fb, pf, true, "forward");
  long long i = transform(ib, pi, true, "inverse");
  // tile copy through the permutation (16-byte copies), forward
  Tally tile;
  const int M = 1 << fb;
  std::vector<int> permslot(M);
  for (int p = 0; p < M; ++p) { const int nat = (-split_radix_index(p, M, 0)) & (M - 1); permslot[nat] = g_slot(p); }
  for (int j0 = 0; j0 < M; j0 += 32) { int slot[32]; for (int l = 0; l < 32; ++l) slot[l] = permslot[j0 + l]; tile.add(slot); }
  report("tile copy as STS.128 (model)", tile);
  {
    long long best = 0;
    for (int j0 = 0; j0 < M; j0 += 32) { int c[8] = {0}, mx = 4; for (int l = 0; l < 32; ++l) mx = std::max(mx, ++c[permslot[j0 + l] & 7]); best += mx; }
    printf("  tile copy, lanes of each 32-chunk reordered: %lld wavefronts\n", best);
  }
  // spectrum phase (UP2): reads F[i], F[M-i]; writes 4 permuted slots of B
  Tally sr, sw;
  const int Mi = 1 << ib, n = M >> 1;
  std::vector<int> pinv(Mi);
  for (int p = 0; p < Mi; ++p) { const int nat = (-split_radix_index(p, Mi, 1)) & (Mi - 1); pinv[nat] = g_slot(p); }
  for (int i0 = 0; i0 < n; i0 += 32) {
    int a[32], b[32], s0[32], s1[32], s2[32], s3[32];
    for (int l = 0; l < 32; ++l) {
      const int ii = i0 + l ? i0 + l : n;
      a[l] = g_slot(ii); b[l] = g_slot(M - ii);
      s0[l] = pinv[ii]; s1[l] = pinv[Mi - ii]; s2[l] = pinv[M - ii]; s3[l] = pinv[M + ii];
    }
    sr.add(a); sr.add(b); sw.add(s0); sw.add(s1); sw.add(s2); sw.add(s3);
  }
  report("spectrum loads", sr); report("spectrum stores", sw);
  const long long top_store_f = 0;
  printf(" total per block pair (tile as model STS, inverse top sinks): %lld\n", f + i + tile.wf + sr.wf + sw.wf + top_store_f +
         /* forward top stores */ (long long)(M / 8));
  return 0;
}
