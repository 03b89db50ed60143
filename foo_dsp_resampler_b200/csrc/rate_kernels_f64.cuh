// rate_kernels_f64.cuh -- the overlap-save DFT stage (dft_filter.h:60-190) of the fp64 engine, the engine RR_open
// selects for Best quality (rate_uni.c:38-51).
//
// The fp64 contract is max |err| <= 1e-12 against rate_double.c, not bit equality, so this kernel is free to use
// its own factorisation of the transforms (the reference's Ooura DAG is only restated in the oracle) and fused
// multiply-adds. What it computes per block is the reference's arithmetic in exact terms: y = the valid part of
// the circular convolution of the (zero-stuffed / spectrally replicated) input with the stage's taps, times the
// planner's scale -- see d64_spectrum for the bin-by-bin correspondence with dft_filter.h:86-188.
//
// Layout of the work on an SM:
//  * one work item = one block of one lane; a CTA holds several GROUPS of 64 or 128 threads (named barriers), each
//    a persistent worker with its own buffer behind shared twiddle tables, exactly like the fp32 lane-pair kernel;
//  * a complex double is one 16-byte shared-memory slot (LDS.128 / STS.128);
//  * the N-point real transforms are M = N/2 point complex transforms plus a split step. The forward transform is
//    an in-place decimation-in-frequency radix-16 (16 values = 64 registers per thread) network: natural order in,
//    BIT-REVERSED order out; the spectrum phase works on bit-reversed positions; the inverse transform is the
//    transposed (decimation-in-time) network: bit-reversed in, natural out. No permutation pass, no second buffer:
//    a record of the spectrum phase reads the slots of bins k and M - k and writes the same slots;
//  * the first forward pass reads its inputs straight from global memory and the last inverse pass stores the
//    block's valid samples straight to global memory, so a 1024 + 2048 point block costs 18.4 k slot accesses
//    where the split-radix DAG of the generic kernel needs 43 k;
//  * x2 frequency-domain up-sampling (dft_filter.h:86-104): the inverse transform of 2M points is split by one
//    radix-2 decimation-in-frequency step that is fused into the spectrum phase (bins k and k + M meet in one
//    thread), leaving two independent M-point transforms in the two halves of the buffer whose outputs are the
//    even and the odd complex samples.
//
// Like the other kernel headers this text also compiles as plain C++ (tests/emu: a group is one serial thread);
// that build is test infrastructure only.
#pragma once

#include "rate_kernels_pk.cuh"

namespace b200rate {

typedef C2<double> CD;

RR_HD CD cd_add(const CD &a, const CD &b) { return CD{a.x + b.x, a.y + b.y}; }
RR_HD CD cd_sub(const CD &a, const CD &b) { return CD{a.x - b.x, a.y - b.y}; }
RR_HD CD cd_mul(const CD &a, const CD &b) { return CD{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
RR_HD CD cd_conj(const CD &a) { return CD{a.x, -a.y}; }
RR_HD CD cd_muli(const CD &a) { return CD{-a.y, a.x}; }       // i a
RR_HD CD cd_mulmi(const CD &a) { return CD{a.y, -a.x}; }      // -i a
RR_HD CD cd_scale(const CD &a, double s) { return CD{a.x * s, a.y * s}; }

// v * exp(-/+ 2 pi i t / 16), t a compile-time constant after unrolling (INV: +).
template <bool INV> RR_HD CD cd_rot16(const CD &v, int t)
{
  constexpr double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
  switch (t & 7) {
    case 0: return v;
    case 4: return INV ? cd_muli(v) : cd_mulmi(v);
    case 2: return INV ? CD{(v.x - v.y) * h, (v.x + v.y) * h} : CD{(v.x + v.y) * h, (v.y - v.x) * h};
    case 6: return INV ? CD{(-v.x - v.y) * h, (v.x - v.y) * h} : CD{(v.y - v.x) * h, (-v.x - v.y) * h};
    default: {
      const double c = (t & 7) == 1 ? c1 : (t & 7) == 3 ? s1 : (t & 7) == 5 ? -s1 : -c1;
      const double s = (t & 7) == 1 || (t & 7) == 7 ? s1 : c1;
      return cd_mul(v, CD{c, INV ? s : -s});
    }
  }
}

RR_HD constexpr int d64_brev(int m, int bits)
{
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((m >> b) & 1) << (bits - 1 - b);
  return r;
}

// 2^LR point transform in registers. DIF: natural order in, bit-reversed out (e[brev(m)] = sum_a e[a] W^(a m)).
template <int LR, bool INV> RR_HD void d64_dif(CD (&e)[1 << LR])
{
  constexpr int R = 1 << LR;
#pragma unroll
  for (int l = 0; l < LR; ++l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const CD u = e[base + i], v = e[base + i + half];
        e[base + i] = cd_add(u, v);
        e[base + i + half] = cd_rot16<INV>(cd_sub(u, v), i * (8 / half));
      }
    }
  }
}
// DIT, the transposed network: bit-reversed in (e[brev(m)] = x_m), natural out.
template <int LR, bool INV> RR_HD void d64_dit(CD (&e)[1 << LR])
{
  constexpr int R = 1 << LR;
#pragma unroll
  for (int l = LR - 1; l >= 0; --l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const CD u = e[base + i], v = cd_rot16<INV>(e[base + i + half], i * (8 / half));
        e[base + i] = cd_add(u, v);
        e[base + i + half] = cd_sub(u, v);
      }
    }
  }
}

// e[brev(m)] *= w^m, m = 1 .. R-1: two multiplication chains (odd and even powers) so that neither the latency
// nor the rounding of a power grows with more than R/2 steps.
template <int LR> RR_HD void d64_twiddle(CD (&e)[1 << LR], const CD &w)
{
  constexpr int R = 1 << LR;
  if (R == 2) { e[1] = cd_mul(e[1], w); return; }
  const CD w2 = cd_mul(w, w);
  CD po = w, pe = w2;
#pragma unroll
  for (int m = 1; m < R; m += 2) {
    e[d64_brev(m, LR)] = cd_mul(e[d64_brev(m, LR)], po);
    if (m + 1 < R) e[d64_brev(m + 1, LR)] = cd_mul(e[d64_brev(m + 1, LR)], pe);
    if (m + 2 < R) po = cd_mul(po, w2);
    if (m + 3 < R) pe = cd_mul(pe, w2);
  }
}

// Slot of position p of a 2^bits point buffer (hb = bits - 3). Consecutive threads touch (a) consecutive positions,
// (b) sixteen-apart positions (the stride-1 radix-16 pass) and (c) positions whose TOP three bits differ (bit-reversed
// neighbours in the spectrum phase); each term keeps one of these on eight different 16-byte bank groups. The map is
// additive over bit-disjoint summands, which is all the passes need.
RR_HD int dslot(int p, int hb) { return p + (p >> 4) + (p >> hb); }
RR_HD int d64_buf_slots(int bits) { return ((dslot((1 << bits) - 1, bits - 3) + 1 + 7) / 8) * 8; }

enum D64Mode {
  D64_UP2 = 0,     // F-domain up-sampling by 2 (44.1 <-> 48 family), 4 or 8: inverse = 2 / 4 / 8 independent M-point transforms
  D64_SAME = 1,    // Pf == Ni: plain / zero-stuffed input, step >= 1
  D64_DECIM = 2    // F-domain decimation by 2^m (step -m): separate inverse buffer
};
constexpr int kD64MaxPasses = 4;
constexpr int kD64MaxGroups = 6, kD64MaxThreads = 384;   // groups / threads per CTA of dft64_kernel

struct Dft64Params {
  DftParams<double> base;        // geometry and views (schedules / split-radix tables unused)
  const CD *H;                   // N/2 + 1 bins of 0.25 * DFT(coefs_time), natural order
  const CD *ta;                  // exp(-2 pi i k / Pf), k <= Pf/4
  const CD *tb;                  // exp(+2 pi i k / Ni), k <= Ni/4
  const CD *tw;                  // pass twiddles: forward passes, then inverse passes (global; staged into shared memory)
  int ntw;                       // entries of tw
  int mode;                      // D64Mode
  int fb, ib;                    // log2 of the forward transform and of each inverse (sub-)transform
  int npf, npi;                  // passes of the forward / inverse transform
  int lr_f[kD64MaxPasses], lr_i[kD64MaxPasses];      // log2 radix per pass, in execution order
  int tw_f[kD64MaxPasses], tw_i[kD64MaxPasses];      // offset of the pass's twiddle row in tw (-1: none)
  int in_f32, out_f32;           // sample types of the views
  int groups, gthreads;
  int fslots, bslots, hstride;   // slots of the forward buffer, of the inverse buffer (D64_DECIM), of one half (D64_UP2)
  int group_slots;
  int up_bits;                   // D64_UP2: log2 of the F-domain up-sampling factor (1, 2 or 3); the inverse is 2^up_bits transforms
  int lane_major;                // work order: consecutive items are consecutive blocks of one lane (else consecutive lanes of one block)
};

// Pass plan of a 2^bits point transform, DIF order: the last pass is the stride-1 radix-16 pass, every earlier
// stride is a multiple of 16; only the first pass is smaller than radix 16.
struct D64Plan { int n; int lr[kD64MaxPasses]; };
RR_HD D64Plan d64_plan(int bits)
{
  D64Plan p{0, {0, 0, 0, 0}};
  const int rest = bits % 4;
  if (rest) p.lr[p.n++] = rest;
  for (int k = 0; k < bits / 4; ++k) p.lr[p.n++] = 4;
  return p;
}

// Everything about a work item that needs 64-bit coordinate arithmetic, computed by one thread per item.
struct D64Item {
  DftItem<double> d;
  const void *src;               // lane pointer at the tile's first input sample (in_kind != 0)
  void *dst;                     // lane pointer at the block's first kept output sample (out_kind != 0)
  long long c0;                  // its coordinate
  int in_kind, out_kind;         // 0: through the view (zero-stuffed / clipped / wrapping), 1: 16-byte complex accesses, 2: strided
  int first, stride, count;      // kept samples: block sample first + j * stride, j < count
};

RR_PROG D64Item d64_make_item(const Dft64Params &dp, long long work)
{
  const DftParams<double> &p = dp.base;
  D64Item it;
  if (dp.lane_major) {                                    // item = lane * nblocks + block -> the block-major index dft_item expects
    const long long lane = work / p.nblocks, bq = work - lane * p.nblocks;
    work = bq * p.nlanes + lane;
  }
  it.d = dft_item<double, 1>(p, work);
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N;
  it.in_kind = 0;
  it.src = nullptr;
  if (p.in_mode != DFT_IN_ZERO_STUFF && view_range_direct(p.in, it.d.Rb, it.d.Rb + span)) {
    if (dp.in_f32) it.src = view_ptr<const float>(p.in, it.d.in_off0, it.d.Rb);
    else it.src = view_ptr<const double>(p.in, it.d.in_off0, it.d.Rb);
    it.in_kind = (!dp.in_f32 && p.in.elem_stride == 1 && !((size_t)it.src & 15)) ? 1 : 2;
  }
  const long long b = it.d.b;
  const int V = p.N - p.overlap;
  long long k0;
  it.first = 0; it.stride = 1;
  if (p.step == 1) { it.count = V; k0 = b * (long long)V; }
  else if (p.step > 1) {
    const long long v0 = b * (long long)V;
    const int Mq = p.step;
    it.first = (int)((Mq - v0 % Mq) % Mq); it.stride = Mq;
    k0 = (v0 + Mq - 1) / Mq;
    it.count = it.first < V ? (V - it.first + Mq - 1) / Mq : 0;
  } else { it.count = p.kept; k0 = b * (long long)p.kept; }
  it.c0 = p.out_preload + k0;
  it.out_kind = 0;
  it.dst = nullptr;
  if (view_range_direct(p.out, it.c0, it.c0 + it.count)) {
    if (dp.out_f32) it.dst = view_ptr<float>(p.out, it.d.out_off0, it.c0);
    else it.dst = view_ptr<double>(p.out, it.d.out_off0, it.c0);
    if (it.stride == 1) it.out_kind = (!dp.out_f32 && p.out.elem_stride == 1 && !((size_t)it.dst & 15)) ? 1 : 2;
  }
  return it;
}

// How the first pass reads the tile / the last pass stores the kept samples. kind 1: 16-byte complex accesses;
// kind 2: direct element accesses `es` apart (float or double); kind 0: the slow path (zero-stuffed, clipped or
// wrapping tiles, decimated outputs) goes through shared memory. The passes are instantiated per kind: a run-time
// choice inside the sixteen unrolled accesses of a thread makes ptxas keep the values in local memory.

// The slow path: the whole tile through the view into F (natural order), one element per thread and round.
RR_PROG void d64_stage_tile(const Dft64Params &dp, const Grp &g, const D64Item &it, CD *F)
{
  const DftParams<double> &p = dp.base;
  const int M = 1 << dp.fb, hf = dp.fb - 3;
  for (int w = g.tid; w < 2 * M; w += g.size) {
    const int j = w;
    long long coord = it.d.Rb + j;
    bool on_grid = true;
    if (p.in_mode == DFT_IN_ZERO_STUFF) {
      const int d = j - it.d.remLb;
      on_grid = d >= 0 && d % p.L == 0;
      coord = it.d.Rb + (on_grid ? d / p.L : 0);
    }
    const double v = !on_grid ? 0.0 : dp.in_f32 ? view_read<float, double>(p.in, it.d.in_off0, coord) : view_read<double, double>(p.in, it.d.in_off0, coord);
    reinterpret_cast<double *>(F + dslot(j >> 1, hf))[j & 1] = v;
  }
  grp_sync(g);
}

// The slow path: kept samples first + j * stride out of the inverse result in B (natural order), through the view.
RR_PROG void d64_emit(const Dft64Params &dp, const Grp &g, const D64Item &it, const CD *B, int lgnt, int hs)
{
  const DftParams<double> &p = dp.base;
  const int hi = dp.ib - 3;
  grp_sync(g);
  for (int j = g.tid; j < it.count; j += g.size) {
    const int t = it.first + j * it.stride, n = t >> 1;
    const int h = n & ((1 << lgnt) - 1), pos = n >> lgnt;
    const double v = reinterpret_cast<const double *>(B + h * hs + dslot(pos, hi))[t & 1];
    if (dp.out_f32) view_write<float, double>(p.out, it.d.out_off0, it.c0 + j, v);
    else view_write<double, double>(p.out, it.d.out_off0, it.c0 + j, v);
  }
}

// One pass over `ntrans` transforms of 2^bits points each (buffers `tstride` slots apart): radix 2^LR butterflies on
// sub-blocks of 2^lgS points. DIT = false: decimation in frequency (butterfly, then twiddles, results in bit-reversed
// digit order); DIT = true: the transposed pass. LD(h, pos) / ST(h, pos, v) move element `pos` of transform h.
// HFAST: consecutive threads take the same butterfly of the `ntrans` = 2^lgnt transforms, so that the interleaved
// complex samples the two halves of an up-sampling block produce are stored by neighbouring threads.
template <int LR, bool INV, bool DIT, bool HFAST = false, class LD, class ST>
RR_PROG void d64_pass(const Grp &g, int lgnt, int bits, int lgS, const CD *tw, LD ld, ST st)
{
  const int ntrans = 1 << lgnt;
  constexpr int R = 1 << LR, U = 16 / R;                 // sixteen values in flight per thread whatever the radix
  const int lgs = lgS - LR, s = 1 << lgs, lgper = bits - LR, per = 1 << lgper, total = ntrans * per;
  for (int t0 = g.tid; t0 < total; t0 += U * g.size) {
    // every value of e[][] is defined on every path (an `if (t < total)` around the accesses makes ptxas keep the
    // arrays in local memory): butterflies beyond the end load zeros and store nothing
    CD e[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * g.size;
      const bool live = t < total;
      const int h = HFAST ? (t & (ntrans - 1)) : (t >> lgper), tt = HFAST ? (t >> lgnt) : (t & (per - 1));
      const int pos0 = ((tt >> lgs) << lgS) | (tt & (s - 1));
#pragma unroll
      for (int a = 0; a < R; ++a) e[u][a] = live ? ld(h, pos0 + (a << lgs)) : CD{0.0, 0.0};
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * g.size;
      const bool live = t < total;
      const int h = HFAST ? (t & (ntrans - 1)) : (t >> lgper), tt = HFAST ? (t >> lgnt) : (t & (per - 1));
      const int j = tt & (s - 1), pos0 = ((tt >> lgs) << lgS) | j;
      if (!DIT) {
        d64_dif<LR, INV>(e[u]);
        if (tw) d64_twiddle<LR>(e[u], tw[j]);
      } else {
        if (tw) d64_twiddle<LR>(e[u], tw[j]);
        d64_dit<LR, INV>(e[u]);
      }
#pragma unroll
      for (int a = 0; a < R; ++a) if (live) st(h, pos0 + (a << lgs), e[u][a]);
    }
  }
}

template <bool INV, bool DIT, bool HFAST, class LD, class ST>
RR_PROG void d64_pass_any(int lr, const Grp &g, int lgnt, int bits, int lgS, const CD *tw, LD ld, ST st)
{
  switch (lr) {
    case 1: d64_pass<1, INV, DIT, HFAST>(g, lgnt, bits, lgS, tw, ld, st); break;
    case 2: d64_pass<2, INV, DIT, HFAST>(g, lgnt, bits, lgS, tw, ld, st); break;
    case 3: d64_pass<3, INV, DIT, HFAST>(g, lgnt, bits, lgS, tw, ld, st); break;
    default: d64_pass<4, INV, DIT, HFAST>(g, lgnt, bits, lgS, tw, ld, st); break;
  }
}

// 2 X[k] of the real transform from the forward result Z (bit-reversed slots of F): X[k] = E + a O with
// E = (Z[k] + conj Z[M-k]) / 2, O = -i (Z[k] - conj Z[M-k]) / 2, a = exp(-2 pi i k / Pf); X[M-k] = conj(E - a O).
RR_HD void d64_split(const CD &za, const CD &zb, const CD &a, CD &x1, CD &x2)
{
  const CD zc = cd_conj(zb);
  const CD e = cd_add(za, zc), o = cd_mulmi(cd_sub(za, zc));
  const CD t = cd_mul(a, o);
  x1 = cd_add(e, t);
  x2 = cd_conj(cd_sub(e, t));
}
// Inverse pre-processing of the bin pair (k, Mi - k): w[k] = A + i b D, w[Mi-k] = conj(A - i b D) with
// A = Y[k] + conj Y[Mi-k], D = Y[k] - conj Y[Mi-k], b = exp(+2 pi i k / Ni).
RR_HD void d64_merge(const CD &ya, const CD &yb, const CD &b, CD &wa, CD &wb)
{
  const CD yc = cd_conj(yb);
  const CD A = cd_add(ya, yc), G = cd_muli(cd_mul(b, cd_sub(ya, yc)));
  wa = cd_add(A, G);
  wb = cd_conj(cd_sub(A, G));
}

RR_HD int d64_rev(int k, int bits)
{
#if defined(__CUDA_ARCH__)
  return (int)(__brev((unsigned)k) >> (32 - bits));
#else
  return d64_brev(k, bits);
#endif
}

// Spectrum phase: real-transform split of the forward result (bit-reversed in F), spectral replication
// (dft_filter.h:86-104) or truncation (:157-188), filter multiply (:118-146), inverse pre-processing and -- for
// D64_UP2 -- the first radix-2 step of the inverse transform; results at the bit-reversed slots of B.
// H carries the factor 1/4 that undoes the two factors 2 of this formulation.
template <int MODE>
RR_PROG void d64_spectrum(const Dft64Params &dp, const Grp &g, CD *F, CD *B)
{
  const int fb = dp.fb, M = 1 << fb, hf = fb - 3;
  const CD *H = dp.H;
  auto Fz = [&](int k) -> CD { return F[dslot(d64_rev(k & (M - 1), fb), hf)]; };
  if (MODE == D64_UP2) {
    const int Mi = 2 * M, hs = dp.hstride, half = M >> 1;
    // the constants of a thread's next index are requested before the current one is computed
    struct Rec { CD b, h0, h1, h2, h3; };
    auto fetch = [&](int k) -> Rec {
      const int kk = k > half ? half : k;                  // beyond the end: any valid address
      return Rec{ldg(dp.tb + kk), ldg(H + kk), ldg(H + Mi - kk), ldg(H + M - kk), ldg(H + M + kk)};
    };
    Rec nx = fetch(g.tid);
    for (int k = g.tid; k <= half; k += g.size) {
      const Rec r = nx;
      nx = fetch(k + g.size);
      if (k == 0) {
        const CD z = F[0];
        const double x0 = 2.0 * (z.x + z.y), xm = 2.0 * (z.x - z.y);       // 2 X[0], 2 X[M] (both real)
        const double y0 = r.h0.x * x0, yn = r.h1.x * x0;                     // bins 0 and Ni/2 see X[0]
        const CD w0 = CD{y0 + yn, y0 - yn};
        const CD wm = cd_scale(cd_conj(cd_scale(r.h2, xm)), 2.0);            // bin M is its own partner: b = i
        B[0] = cd_add(w0, wm);
        B[hs] = cd_sub(w0, wm);
        continue;
      }
      const CD b = r.b, b2 = cd_mul(b, b);                                   // b2 = exp(2 pi i k / Pf) = conj(a)
      CD x1, x2;
      d64_split(Fz(k), Fz(M - k), cd_conj(b2), x1, x2);
      const int ra = dslot(d64_rev(k, fb), hf), rb = dslot(d64_rev(M - k, fb), hf);
      CD wa, wb;                                                             // bins k and Mi - k
      d64_merge(cd_mul(r.h0, x1), cd_mul(r.h1, cd_conj(x1)), b, wa, wb);
      if (k == half) {                                                       // M - k == k: one pair, partner of k is Mi - k
        B[ra] = cd_add(wa, wb);
        B[hs + ra] = cd_muli(cd_sub(wa, wb));
        continue;
      }
      CD wc, wd;                                                             // bins M - k and M + k: b' = i conj(b)
      d64_merge(cd_mul(r.h2, x2), cd_mul(r.h3, cd_conj(x2)), cd_muli(cd_conj(b)), wc, wd);
      // first radix-2 step of the inverse transform: (k, k + M) and (M - k, Mi - k), twiddles b2 and -conj(b2)
      B[ra] = cd_add(wa, wd);
      B[hs + ra] = cd_mul(cd_sub(wa, wd), b2);
      B[rb] = cd_add(wc, wb);
      B[hs + rb] = cd_mul(cd_sub(wb, wc), cd_conj(b2));
    }
  } else if (MODE == D64_SAME) {
    const int half = M >> 1;
    struct Rec { CD a, h0, h1; };
    auto fetch = [&](int k) -> Rec {
      const int kk = k > half ? half : k;
      return Rec{ldg(dp.ta + kk), ldg(H + kk), ldg(H + M - kk)};
    };
    Rec nx = fetch(g.tid);
    for (int k = g.tid; k <= half; k += g.size) {
      const Rec r = nx;
      nx = fetch(k + g.size);
      if (k == 0) {
        const CD z = F[0];
        const double y0 = r.h0.x * 2.0 * (z.x + z.y), yn = r.h1.x * 2.0 * (z.x - z.y);
        B[0] = CD{y0 + yn, y0 - yn};
        continue;
      }
      CD x1, x2;
      d64_split(Fz(k), Fz(M - k), r.a, x1, x2);
      const int ra = dslot(d64_rev(k, fb), hf), rb = dslot(d64_rev(M - k, fb), hf);
      CD wa, wb;
      d64_merge(cd_mul(r.h0, x1), cd_mul(r.h1, x2), cd_conj(r.a), wa, wb);
      B[ra] = wa;
      if (k != half) B[rb] = wb;
    }
  } else {
    const int ibits = dp.ib, Mi = 1 << ibits, hi = ibits - 3;
    auto X2 = [&](int k) -> CD {                         // 2 X[k], 0 < k <= M/2
      CD x1, x2;
      d64_split(Fz(k), Fz(M - k), ldg(dp.ta + k), x1, x2);
      return x1;
    };
    for (int k = g.tid; k <= (Mi >> 1); k += g.size) {
      if (k == 0) {
        const CD z = F[0];
        const double y0 = ldg(H).x * 2.0 * (z.x + z.y);
        const double yn = cd_mul(ldg(H + Mi), X2(Mi)).x;                     // new Nyquist bin: real part, dft_filter.h:185
        B[0] = CD{y0 + yn, y0 - yn};
        continue;
      }
      CD wa, wb;
      d64_merge(cd_mul(ldg(H + k), X2(k)), cd_mul(ldg(H + Mi - k), X2(Mi - k)), ldg(dp.tb + k), wa, wb);
      B[dslot(d64_rev(k, ibits), hi)] = wa;
      if (k != (Mi >> 1)) B[dslot(d64_rev(Mi - k, ibits), hi)] = wb;
    }
  }
  grp_sync(g);
}

// Spectrum phase for F-domain up-sampling by L = 2^LR >= 4 (dft_filter.h:86-104 with its doubling copies): the Pf-point
// spectrum repeats L/2 times below Nyquist, bin k + cM sees X[k] for even c and conj X[M-k] for odd c. A record (k, M-k)
// of the forward result therefore determines the 2L bins k + cM and (M-k) + cM; bin k + cM pairs with (M-k) + (L-1-c)M
// in the inverse pre-processing, and the L bins kappa + cM of one residue kappa meet in the first radix-L
// decimation-in-frequency step of the inverse transform, which is done here: what is stored are the inputs of L
// independent M-point transforms (transform h at B + h * hstride) whose outputs are the complex samples L n' + h.
template <int LR>
RR_PROG void d64_spectrum_up(const Dft64Params &dp, const Grp &g, CD *F, CD *B)
{
  constexpr int L = 1 << LR;
  const int fb = dp.fb, M = 1 << fb, hf = fb - 3, Mi = L * M, hs = dp.hstride, half = M >> 1;
  const CD *H = dp.H;
  auto Fz = [&](int k) -> CD { return F[dslot(d64_rev(k & (M - 1), fb), hf)]; };
  // radix-L step over c for residue kappa: e[c] = w[kappa + cM] -> transform h gets sum_c e[c] exp(2 pi i c h / L),
  // times exp(2 pi i kappa h / Mi) = tw^h
  auto finish = [&](CD (&e)[L], const CD &tw, bool twiddled, int slot) {
    d64_dif<LR, true>(e);
    if (twiddled) d64_twiddle<LR>(e, tw);
#pragma unroll
    for (int a = 0; a < L; ++a) B[d64_brev(a, LR) * hs + slot] = e[a];       // e[a] is output h = brev(a)
  };
  for (int k = g.tid; k <= half; k += g.size) {
    if (k == 0) {
      const CD z = F[0];
      const double x0 = 2.0 * (z.x + z.y), xm = 2.0 * (z.x - z.y);           // 2 X[0], 2 X[M]: bins cM see them for even / odd c
      CD e[L];
      {
        const double y0 = ldg(H).x * x0, yn = ldg(H + Mi).x * ((L & 1) ? xm : x0);
        e[0] = CD{y0 + yn, y0 - yn};
      }
#pragma unroll
      for (int c = 1; c < L / 2; ++c) {                                        // pairs (cM, (L-c)M), b = exp(i pi c / L)
        const double xa = (c & 1) ? xm : x0, xb = ((L - c) & 1) ? xm : x0;
        d64_merge(cd_scale(ldg(H + c * M), xa), cd_scale(ldg(H + (L - c) * M), xb), cd_rot16<true>(CD{1.0, 0.0}, 8 * c / L), e[c], e[L - c]);
      }
      e[L / 2] = cd_scale(cd_conj(cd_scale(ldg(H + (L / 2) * M), ((L / 2) & 1) ? xm : x0)), 2.0);   // its own partner: b = i
      finish(e, CD{1.0, 0.0}, false, 0);
      continue;
    }
    const CD b = ldg(dp.tb + k);                                               // exp(2 pi i k / Ni), Ni = 2 L M
    CD b2 = cd_mul(b, b);                                                      // exp(2 pi i k / Mi)
    CD a = b2;                                                                 // -> exp(2 pi i k / Pf) = b^(2L/2)... squared LR - 1 more times
#pragma unroll
    for (int r = 1; r < LR; ++r) a = cd_mul(a, a);
    CD x1, x2;
    d64_split(Fz(k), Fz(M - k), cd_conj(a), x1, x2);
    const int ra = dslot(d64_rev(k, fb), hf), rb = dslot(d64_rev(M - k, fb), hf);
    CD ea[L], eb[L];                                                           // w[k + cM], w[(M-k) + cM]
#pragma unroll
    for (int c = 0; c < L; ++c) {
      const CD ya = cd_mul(ldg(H + k + c * M), (c & 1) ? cd_conj(x2) : x1);
      const int cb = L - 1 - c;
      const CD yb = cd_mul(ldg(H + (M - k) + cb * M), (cb & 1) ? cd_conj(x1) : x2);
      d64_merge(ya, yb, cd_rot16<true>(b, 8 * c / L), ea[c], eb[cb]);          // b_{k + cM} = b exp(i pi c / L)
    }
    finish(ea, b2, true, ra);
    if (k != half) finish(eb, cd_rot16<true>(cd_conj(b2), 16 / L), true, rb);  // exp(2 pi i (M-k) / Mi) = exp(2 pi i / L) conj(b2)
  }
  grp_sync(g);
}

// Ask L2 for the next item's tile while this item's inverse transform runs.
RR_PROG void d64_tile_prefetch(const Dft64Params &dp, const Grp &g, const D64Item &it)
{
#if defined(__CUDA_ARCH__)
  const DftParams<double> &p = dp.base;
  if (it.in_kind == 0) return;
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N;
  const size_t bytes = (size_t)span * p.in.elem_stride * (dp.in_f32 ? 4 : 8);
  const int lines = (int)((bytes + 127) >> 7);
  const char *base = static_cast<const char *>(it.src);
  for (int k = g.tid; k < lines; k += g.size) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)k * 128));
#else
  (void)dp; (void)g; (void)it;
#endif
}

// One work item (block b of one lane). items[slot] describes it; items[slot ^ 1] is filled for the next one.
// `twt`: the pass twiddle rows (shared memory on the device), indexed by dp.tw_f / dp.tw_i.
// UPL: the kernel instance for F-domain up-sampling by 4 / 8 (kept apart: the 2L values a thread holds in that spectrum
// phase would otherwise perturb the register allocation of the x2 / 1:1 / decimating paths).
template <bool UPL>
RR_PROG void dft64_program(const Dft64Params &dp, const Grp &g, const CD *twt, D64Item *items, int slot, long long work_next, CD *buf)
{
  const int MODE = dp.mode;
  grp_sync(g);                                            // items[slot] is visible; the buffer is free
  const D64Item &it = items[slot];
  if (work_next >= 0 && g.tid == 0) items[slot ^ 1] = d64_make_item(dp, work_next);
  const int fb = dp.fb, hf = fb - 3;
  CD *F = buf, *B = MODE == D64_DECIM ? buf + dp.fslots : buf;
  auto ldF = [&](int, int pos) -> CD { return F[dslot(pos, hf)]; };
  auto stF = [&](int, int pos, const CD &v) { F[dslot(pos, hf)] = v; };

  // ---- forward transform: DIF, the first pass reads the input tile from global memory ----
  int lgS = fb;
  {
    const CD *tw0 = dp.tw_f[0] >= 0 ? twt + dp.tw_f[0] : nullptr;
    const int lr0 = dp.lr_f[0], kind = it.in_kind;
    const void *src = it.src;
    const long long es = dp.base.in.elem_stride;
    // sample pair `pos` of the tile: samples 2 pos and 2 pos + 1 (dft_filter.h:86-116)
    if (kind == 1)
      d64_pass_any<false, false, false>(lr0, g, 0, fb, lgS, tw0, [&](int, int pos) -> CD { return ldg(static_cast<const CD *>(src) + pos); }, stF);
    else if (kind == 2 && dp.in_f32)
      d64_pass_any<false, false, false>(lr0, g, 0, fb, lgS, tw0, [&](int, int pos) -> CD {
        const float *q = static_cast<const float *>(src) + 2ll * pos * es;
        return CD{(double)ldg(q), (double)ldg(q + es)};
      }, stF);
    else if (kind == 2)
      d64_pass_any<false, false, false>(lr0, g, 0, fb, lgS, tw0, [&](int, int pos) -> CD {
        const double *q = static_cast<const double *>(src) + 2ll * pos * es;
        return CD{ldg(q), ldg(q + es)};
      }, stF);
    else {
      d64_stage_tile(dp, g, it, F);
      d64_pass_any<false, false, false>(lr0, g, 0, fb, lgS, tw0, ldF, stF);
    }
  }
  grp_sync(g);
  lgS -= dp.lr_f[0];
  for (int ps = 1; ps < dp.npf; ++ps) {
    d64_pass<4, false, false>(g, 0, fb, lgS, dp.tw_f[ps] >= 0 ? twt + dp.tw_f[ps] : nullptr, ldF, stF);
    grp_sync(g);
    lgS -= 4;
  }

  if constexpr (UPL) {
    if (dp.up_bits == 2) d64_spectrum_up<2>(dp, g, F, B);
    else d64_spectrum_up<3>(dp, g, F, B);
  } else if (MODE == D64_UP2) d64_spectrum<D64_UP2>(dp, g, F, B);
  else if (MODE == D64_SAME) d64_spectrum<D64_SAME>(dp, g, F, B);
  else d64_spectrum<D64_DECIM>(dp, g, F, B);
  if (work_next >= 0) d64_tile_prefetch(dp, g, items[slot ^ 1]);   // published before the barriers of the forward transform

  // ---- inverse transform(s): DIT, the last pass stores the kept samples to global memory ----
  const int ib = dp.ib, hi = ib - 3, nt = MODE == D64_UP2 ? dp.up_bits : 0, hs = dp.hstride;   // nt: log2 of the transforms
  auto ldB = [&](int h, int pos) -> CD { return B[h * hs + dslot(pos, hi)]; };
  auto stB = [&](int h, int pos, const CD &v) { B[h * hs + dslot(pos, hi)] = v; };
  lgS = 0;
  for (int ps = 0; ps + 1 < dp.npi; ++ps) {
    lgS += 4;
    d64_pass<4, true, true>(g, nt, ib, lgS, dp.tw_i[ps] >= 0 ? twt + dp.tw_i[ps] : nullptr, ldB, stB);
    grp_sync(g);
  }
  {
    const int lrl = dp.lr_i[dp.npi - 1], kind = it.out_kind, count = it.count, up = nt;
    const CD *twl = dp.tw_i[dp.npi - 1] >= 0 ? twt + dp.tw_i[dp.npi - 1] : nullptr;
    void *dst = it.dst;
    const long long es = dp.base.out.elem_stride;
    lgS += lrl;
    // complex output element n = real samples 2n and 2n + 1 of the inverse transform; the first `count` are kept
    if (kind == 1)
      d64_pass_any<true, true, true>(lrl, g, nt, ib, lgS, twl, ldB, [&](int h, int pos, const CD &v) {
        const int n = (pos << up) + h;
        if (2 * n + 1 < count) static_cast<CD *>(dst)[n] = v;
        else if (2 * n < count) static_cast<double *>(dst)[2 * n] = v.x;
      });
    else if (kind == 2 && dp.out_f32)
      d64_pass_any<true, true, true>(lrl, g, nt, ib, lgS, twl, ldB, [&](int h, int pos, const CD &v) {
        const int n = (pos << up) + h;
        float *q = static_cast<float *>(dst) + 2ll * n * es;
        if (2 * n < count) q[0] = (float)v.x;
        if (2 * n + 1 < count) q[es] = (float)v.y;
      });
    else if (kind == 2)
      d64_pass_any<true, true, true>(lrl, g, nt, ib, lgS, twl, ldB, [&](int h, int pos, const CD &v) {
        const int n = (pos << up) + h;
        double *q = static_cast<double *>(dst) + 2ll * n * es;
        if (2 * n < count) q[0] = v.x;
        if (2 * n + 1 < count) q[es] = v.y;
      });
    else {
      d64_pass_any<true, true, true>(lrl, g, nt, ib, lgS, twl, ldB, stB);
      d64_emit(dp, g, it, B, nt, hs);
    }
  }
}

}  // namespace b200rate
