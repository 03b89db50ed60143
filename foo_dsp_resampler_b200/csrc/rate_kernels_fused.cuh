// rate_kernels_fused.cuh -- the DFT stage and the rational polyphase stage behind it as ONE kernel for fp32 lane
// pairs: stage chaining (rate/rate_base.h:425-432: stage i runs into the FIFO of stage i + 1) without the FIFO ever
// leaving the SM.
//
// Reference behaviour: dft_stage_fn (dft_filter.h:60-190) followed by vpoly0 (rate_filters_generic.h:272-305); same
// expression DAGs as the two separate kernels of rate_kernels_pk.cuh, hence the same bits.
//
// How it is laid on the SM. A work item is a RUN of consecutive DFT blocks of one lane pair, processed by one group
// (128 threads) of a CTA. For every block the group runs the forward transform, the spectrum phase and the inverse
// transform exactly like dftp_program; the inverse transform's top pass, however, keeps its results in registers
// across a group barrier and writes the block's V valid samples back IN PLACE in natural order: the inverse buffer B
// then is the polyphase stage's input window, sample j of the block at B[j] (8 bytes: both lanes). The `n - 1`
// samples of look-ahead history the next block's first outputs need are copied in front of B (halo H) before B is
// overwritten, so windows that straddle a block boundary are contiguous too. The polyphase phase is the
// phase-stationary two-slots-per-thread scheme of poly0_pair2_tile: a thread owns two adjacent output slots of the
// period (output index mod L), its two coefficient rows come from a transposed per-thread table, consecutive
// periods advance the window by exactly `step` samples (exact integer phase arithmetic), taps are summed in tap
// order with FMUL2 / FFMA2(x, 1, y). Which outputs a block produces is a closed form of absolute coordinates
// (pk_fused_block): those whose window ends inside the block.
//
// A run that does not start at the first block of the launch begins one block early: that block only provides the
// halo (its polyphase phase is skipped), so runs are independent and there is no exchange between CTAs. The host
// picks the run length so that the recomputation stays below a few percent (engine.cu).
//
// HBM traffic of the pair of stages drops from (in + 2 x intermediate + out) to (in + out): for 48 -> 44.1 kHz the
// 2x-rate intermediate FIFO was 2/3 of all bytes (SURVEY.md 8d counts intermediates as on-chip).
#pragma once

#include "rate_kernels_pk.cuh"

namespace b200rate {

constexpr int kFusedHalo = 32;           // Pk slots in front of B: >= n - 1 + 3 for n <= 24 ... 28 taps
constexpr int kFusedPolyThreads = 128;   // threads of a group (all may take part in the polyphase phase)

struct DftPolyParams {
  DftPkParams dft;               // DFT stage (views: in = the stage's input; out unused)
  // polyphase stage (vpoly0): output i <-> at = at0 + i * step, (q, r) = divmod(at, L); y = sum_k c[r][k] x[q + k]
  int L, n;
  long long at0, step;
  long long poly_preload;        // FIFO coordinate of DFT output 0 in the polyphase stage's input FIFO
  long long out0, nout;          // outputs [out0, out0 + nout) of every lane
  LaneView out;                  // final output view
  long long out_preload;
  // per-thread tables, transposed: coef[j * kFusedPolyThreads + t], j < 2 n + 1 (row of the first slot, then the shifted
  // row of the second slot, see Poly0Pair2Thread); slot[t] = first slot of the thread's slot pair or 0xffff (idle),
  // qs[t] = first window sample of that slot relative to the period, flags[t]: bit 0 = d_lo, bit 1 = second slot exists
  const float *coef;
  const uint16_t *slot;
  const uint16_t *qs;
  const uint8_t *flags;
  int tile_t0;                   // first thread of a group without a slot pair (multiple of 16, < kPkGroupThreads)
  // runs
  long long block0;              // first block of the launch
  int nblocks;                   // blocks of the launch per lane pair
  int run_len;                   // blocks per run (the last run of a pair may be shorter)
  int runs_per_pair;
};

// Geometry of one block of a run, computed by one thread (64-bit coordinate arithmetic).
struct PkFusedBlock {
  PkItem it;                     // tile description (output fields unused)
  float *d_base;                 // output pointer of (period m_lo, slot 0) of the pair's first lane
  long long lane1;               // elements from the pair's first lane to its second
  int m_lo;                      // (unused on the device beyond debugging) first period touched, low bits
  int base_x;                    // window index (relative to B[0]) of (period m_lo, qs = 0)
  int i_lo, i_hi;                // outputs of this block, relative to period m_lo's slot 0: [i_lo, i_hi)
  int halo_only;                 // first block of a run that starts inside the launch: no outputs
  int zero_halo;                 // the stream starts here: the halo is the FIFO's preload zeros
};

RR_HD long long fused_poly_ready(const DftPolyParams &fp, long long W)   // outputs whose window ends at or below coordinate W
{
  const long long avail = W - (fp.n - 1);
  if (avail <= 0) return 0;
  const long long num = avail * fp.L - fp.at0;
  return num <= 0 ? 0 : (num + fp.step - 1) / fp.step;
}

// work = pair * runs_per_pair + run; k = block of the run (0 may be the halo block)
RR_PROG bool pk_fused_run(const DftPolyParams &fp, long long work, long long *b_first, int *count, int *pair, int *halo_first)
{
  long long pr; int run;
  divmod_ll(work, fp.runs_per_pair, pr, run);
  *pair = (int)pr;
  const long long r0 = (long long)run * fp.run_len;
  const int len = (fp.nblocks - r0) < fp.run_len ? (int)(fp.nblocks - r0) : fp.run_len;
  *halo_first = run > 0 ? 1 : 0;
  *b_first = fp.block0 + r0 - (run > 0 ? 1 : 0);
  *count = len + (run > 0 ? 1 : 0);
  return len > 0;
}

RR_PROG PkFusedBlock pk_fused_block(const DftPolyParams &fp, int pair, long long b, bool halo_only)
{
  const DftPkParams &pp = fp.dft;
  const DftParams<float> &p = pp.base;
  PkFusedBlock fb;
  // tile description: work index of (block, pair) in dftp_program's numbering
  fb.it = pk_make_item(pp, (b - p.block0) * (long long)(p.nlanes >> 1) + pair);
  const int V = p.N - p.overlap;
  const long long c_b = fp.poly_preload + b * (long long)V;          // coordinate of B[0]
  fb.halo_only = halo_only ? 1 : 0;
  fb.zero_halo = b == 0 ? 1 : 0;
  const long long out_end = fp.out0 + fp.nout;
  long long ilo = fused_poly_ready(fp, c_b), ihi = fused_poly_ready(fp, c_b + V);
  if (ilo < fp.out0) ilo = fp.out0;
  if (ihi > out_end) ihi = out_end;
  if (ihi < ilo || halo_only) ihi = ilo;
  const long long m_lo = ilo / fp.L;
  fb.m_lo = (int)m_lo;
  fb.i_lo = (int)(ilo - m_lo * fp.L);
  fb.i_hi = (int)(ihi - m_lo * fp.L);
  fb.base_x = (int)(m_lo * fp.step - c_b);
  const int lane0 = 2 * pair;
  const long long off0 = lane_offset(fp.out, lane0);
  fb.d_base = view_ptr<float>(fp.out, off0, fp.out_preload + m_lo * fp.L);
  fb.lane1 = lane_offset(fp.out, lane0 + 1) - off0;
  return fb;
}

// The top pass of the inverse transform with its results written back in place in natural order: element c (samples
// 2c, 2c + 1 of the block) at nat[c] for c < half. All of a thread's tasks are loaded before the group barrier that
// separates the last padded-layout read from the first natural-layout write.
template <int BITS>
RR_PROG void pk_fft_top_natural(const Grp &g, CPk *buf, const float *pyr, int half)
{
  constexpr PkPhaseList pl = pk_phase_list(BITS);
  constexpr int PH = pl.n - 1, LG = pl.lg[PH], D = pl.depth[PH];
  constexpr int q = 1 << (LG - 2), NV = 4 << (D - 1), ntask = pk_phase_main(BITS, LG, D);
  static_assert(pk_phase_light(BITS, LG, D) == 0, "the top phase has no light tasks");
  typedef PkGeo<LG> G;
#if defined(__CUDA_ARCH__)
  constexpr int T = (ntask + kPkGroupThreads - 1) / kPkGroupThreads;
  CPk e[T][NV];
#pragma unroll
  for (int tt = 0; tt < T; ++tt) {
    const int o = g.tid + tt * g.size;
    if (o < ntask) pk_item_regs<LG, D>(o, buf, pyr, e[tt]);
  }
  grp_sync(g);
#pragma unroll
  for (int tt = 0; tt < T; ++tt) {
    const int o = g.tid + tt * g.size;
    if (o < ntask) {
#pragma unroll
      for (int j = 0; j < NV; ++j) if (o + j * q < half) buf[o + j * q] = e[tt][j];
    }
  }
  grp_sync(g);
#else
  (void)g;
  CPk *tmp = new CPk[(size_t)1 << BITS];
  for (int o = 0; o < ntask; ++o) {
    CPk e[NV];
    pk_item_regs<LG, D>(o, buf, pyr, e);
    for (int j = 0; j < NV; ++j) tmp[o + j * q] = e[j];
  }
  for (int c = 0; c < half; ++c) buf[c] = tmp[c];
  delete[] tmp;
  (void)sizeof(G);
#endif
}

// Polyphase phase of one block: x = B viewed as Pk samples (x[j] = sample j of the block, x[-h] = halo).
template <int NT, int DLO>
RR_PROG void pk_fused_poly(const DftPolyParams &fp, const Grp &g, const PkFusedBlock &fb, const Pk *x)
{
  typedef Arith<Pk> A;
  const int L = fp.L, step = (int)fp.step;
  const int es = fp.out.elem_stride;
  for (int t = g.tid; t < kFusedPolyThreads; t += g.size) {
    const int s0 = ldg(fp.slot + t);
    if (s0 == 0xffff) continue;
    // periods this thread has work in: at least one of its two outputs inside [i_lo, i_hi)
    int m_first = 0;
    if (s0 + 1 < fb.i_lo) m_first = 1;                                // both slots of period m_lo precede i_lo
    const int last = fb.i_hi - 1 - s0;                                // (m - m_lo) * L <= last
    if (last < 0) continue;
    const int m_last = last / L;
    if (m_last < m_first) continue;
    const unsigned fl = ldg(fp.flags + t);
    const bool dlo = fl & 1, two = (fl & 2) != 0;
    float c0[NT], c1[NT + 1];
#pragma unroll
    for (int k = 0; k < NT; ++k) c0[k] = ldg(fp.coef + k * kFusedPolyThreads + t);
#pragma unroll
    for (int k = 0; k <= NT; ++k) c1[k] = ldg(fp.coef + (NT + k) * kFusedPolyThreads + t);
    const Pk *xw = x + fb.base_x + (int)ldg(fp.qs + t) + m_first * step;
    float *d0 = fb.d_base + (long long)(m_first * L + s0) * es, *d1 = d0 + fb.lane1;
    const int dstep = L * es;
    const bool packed_out = fb.lane1 == 1 && !((size_t)d0 & 7) && !(dstep & 1) && !(es & 1);
    // one tap of both slots: window sample j carries tap j of the first slot and tap j - DLO (d == DLO) or
    // j - DLO - 1 (d == DLO + 1) of the second
    auto tap = [&](int j, Pk xv, Pk &a0, Pk &a1) {
      if (j < NT) a0 = A::addp(a0, A::mul(pk_bcast(c0[j]), xv));
      if (j >= DLO) {
        const int jj = j - DLO;                           // 0 .. NT
        if (jj == 0) { if (dlo) a1 = A::addp(a1, A::mul(pk_bcast(c1[0]), xv)); }
        else if (jj == NT) { if (!dlo) a1 = A::addp(a1, A::mul(pk_bcast(c1[NT]), xv)); }
        else a1 = A::addp(a1, A::mul(pk_bcast(c1[jj]), xv));
      }
    };
    auto emit = [&](int m, Pk a0, Pk a1, float *e0, float *e1) {
      const int i = m * L + s0;                           // relative to period m_lo's slot 0
      if (i >= fb.i_lo && i < fb.i_hi) {
        if (packed_out) *reinterpret_cast<Pk *>(e0) = a0;
        else { *e0 = a0.a; *e1 = a0.b; }
      }
      if (two && i + 1 >= fb.i_lo && i + 1 < fb.i_hi) {
        if (packed_out) *reinterpret_cast<Pk *>(e0 + es) = a1;
        else { e0[es] = a1.a; e1[es] = a1.b; }
      }
    };
    constexpr int NW = NT + DLO + 1, KB = 9;              // window samples per period; loads issued KB at a time
    int m = m_first;
    // Two periods at a time: four independent accumulation chains and 2 x KB shared-memory loads in flight. Few warps
    // run this phase (the group's others wait or fetch the next tile), so it must not depend on other warps to hide
    // the shared-memory latency; its FMUL2 / FFMA2 stream fills issue slots the FFT phases of the other groups leave idle.
    for (; m + 1 <= m_last; m += 2, xw += 2 * step, d0 += 2 * dstep, d1 += 2 * dstep) {
      const Pk *xb = xw + step;
      Pk a0 = pk_bcast(0.0f), a1 = pk_bcast(0.0f), b0 = pk_bcast(0.0f), b1 = pk_bcast(0.0f);
#pragma unroll
      for (int j0 = 0; j0 < NW; j0 += KB) {
        Pk va[KB], vb[KB];
#pragma unroll
        for (int j = 0; j < KB; ++j) if (j0 + j < NW) { va[j] = pk_load8(xw + j0 + j); vb[j] = pk_load8(xb + j0 + j); }
#pragma unroll
        for (int j = 0; j < KB; ++j) if (j0 + j < NW) { tap(j0 + j, va[j], a0, a1); tap(j0 + j, vb[j], b0, b1); }
      }
      emit(m, a0, a1, d0, d1);
      emit(m + 1, b0, b1, d0 + dstep, d1 + dstep);
    }
    if (m <= m_last) {
      Pk a0 = pk_bcast(0.0f), a1 = pk_bcast(0.0f);
#pragma unroll
      for (int j0 = 0; j0 < NW; j0 += KB) {
        Pk va[KB];
#pragma unroll
        for (int j = 0; j < KB; ++j) if (j0 + j < NW) va[j] = pk_load8(xw + j0 + j);
#pragma unroll
        for (int j = 0; j < KB; ++j) if (j0 + j < NW) tap(j0 + j, va[j], a0, a1);
      }
      emit(m, a0, a1, d0, d1);
    }
  }
}

// One block of a run. blk[slot] describes it; blk[slot ^ 1] is filled for the next block of the run (next_b >= 0).
// H | B are contiguous: Pk index -kFusedHalo .. of B. The caller has put the tile of the run's first block into F;
// every later tile is brought in during the polyphase phase of the block before it by the threads that have no
// polyphase work (F is idle from the spectrum phase on; the host guarantees at least sixteen such threads).
template <int MODE, int FB, int IB, int NT, int DLO>
RR_PROG void dft_poly_program(const DftPolyParams &fp, const Grp &g, const PkTables &tb, PkFusedBlock *blk, int slot, int pair,
                              long long next_b, bool first_of_run, CPk *F, CPk *B)
{
  static_assert(FB > 0 && IB > 0, "the fused kernel exists for compile-time transform sizes only");
  static_assert(NT + DLO + 2 <= kFusedHalo, "halo too small for this tap count");
  const DftPkParams &pp = fp.dft;
  const DftParams<float> &p = pp.base;
  const int V = p.N - p.overlap;
  Pk *x = reinterpret_cast<Pk *>(B);
  grp_sync(g);                                            // blk[slot] and F are visible; the previous polyphase phase is over
  const PkFusedBlock &fb = blk[slot];
  if (next_b >= 0 && g.tid == 0) blk[slot ^ 1] = pk_fused_block(fp, pair, next_b, false);
  // history for this block's first outputs: the tail of the previous block's samples (or the preload zeros)
  if (!first_of_run) { for (int h = g.tid; h < kFusedHalo; h += g.size) x[h - kFusedHalo] = x[V - kFusedHalo + h]; }
  else if (fb.zero_halo) { for (int h = g.tid; h < kFusedHalo; h += g.size) x[h - kFusedHalo] = pk_bcast(0.0f); }
  grp_sync(g);

  PkSink sink{nullptr, nullptr, 0, 0};
  pk_fft_lower_any<FB>(pp.fb, g, F, tb.ltab_f, tb.pyr_f, p.sqrthalf, p.c16_1, p.c16_3);
  {
    PkSpecRegs pre;
    if (kPkPrefetchAcrossTop) pk_spec_prefetch<MODE>(pp, g, pre);
    pk_fft_top_any<FB>(pp.fb, g, F, nullptr, tb.pyr_f, false, sink);
    if (!kPkPrefetchAcrossTop) pk_spec_prefetch<MODE>(pp, g, pre);
    pk_spectrum<MODE, false>(pp, g, pre, F, B);           // ends with a barrier: F is free from here on
  }
  pk_fft_lower_any<IB>(pp.ib, g, B, tb.ltab_i, tb.pyr_i, p.sqrthalf, p.c16_1, p.c16_3);
  pk_fft_top_natural<IB>(g, B, tb.pyr_i, V >> 1);        // ends with a barrier: x[0 .. V) are the block's samples

#if defined(__CUDA_ARCH__)
  // The threads from fp.tile_t0 on own no slot pair (the deal fills whole half-warps from thread 0): while the others
  // run the polyphase phase they bring the next block's tile into F -- neither phase waits for the other.
  const int t0 = fp.tile_t0;
  if (g.tid < t0) { if (fb.i_hi > fb.i_lo) pk_fused_poly<NT, DLO>(fp, g, fb, x); }
  else if (next_b >= 0) pk_tile_now<FB, true>(pp, Grp{g.tid - t0, kPkGroupThreads - t0, g.bar}, blk[slot ^ 1].it, F, tb.perm_f);
#else
  if (fb.i_hi > fb.i_lo) pk_fused_poly<NT, DLO>(fp, g, fb, x);
  if (next_b >= 0) pk_tile_now<FB, true>(pp, g, blk[slot ^ 1].it, F, tb.perm_f);
#endif
}

}  // namespace b200rate
