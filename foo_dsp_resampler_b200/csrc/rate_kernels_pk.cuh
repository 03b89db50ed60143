// rate_kernels_pk.cuh -- the overlap-save DFT stage for fp32 LANE PAIRS (two channels of one stream), the
// dominant kernel of the fp32 engine on B200.
//
// Same reference behaviour as dft_stage_program in rate_kernels.cuh (dft_filter.h:60-190 over
// fft-float/{fft.c,rdft.c}); what changes is how the work is laid on the SM:
//
//  * The two lanes of a pair travel together as one 64-bit register pair `Pk` and every arithmetic
//    instruction is a packed sm_100 FMUL2 / FADD2 (PTX mul/add.rn.f32x2): one issue slot for both channels,
//    which doubles the un-fused fp32 rate (profiles/fp_peaks.json: 36 -> 66 TFLOP/s). Each half is an IEEE
//    round-to-nearest fp32 operation, so the bits are those of the scalar expression DAG.
//    ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry .rn, so every add whose
//    operand is a product is written as fma(x, one, y) with `one` read from constant memory (not foldable):
//    one rounding of x + y, no contraction possible. The GPU parity tests (bit equality) guard this.
//  * A complex value of the pair is one 16-byte shared-memory slot {re.a, re.b, im.a, im.b}: one LDS.128 /
//    STS.128 per element for both lanes; two interleaved stereo frames are exactly one slot, so the input
//    tile is brought in by 16-byte (or 8-byte) LDGSTS copies.
//  * The out-of-place permutation of ff_fft_permute_c (fft.c:169-177) is folded into the WRITES that fill an
//    FFT buffer (input tile copy, spectrum stage), so the leaves work in place on 16 contiguous slots.
//  * Combining passes are fused two or three levels deep (8 or 16 values per task in registers) and
//    instantiated per size, so every shared-memory address of a task is base + immediate.
//  * Real-FFT post-processing, spectrum assembly, filter multiply and inverse pre-processing are one phase.
//  * A CTA holds several independent GROUPS of 128 threads (named barriers), each with its own work item
//    and buffers; they share the twiddle tables and hide each other's barrier stalls.
//
// Like rate_kernels.cuh this text also compiles as plain C++ (tests/emu): a group is then one serial
// thread. That build is test infrastructure only.
#pragma once

#include "pk_plan.hpp"
#include "rate_kernels.cuh"

namespace b200rate {

// ---------------------------------------------------------------------------------------------------
// Packed pair arithmetic
// ---------------------------------------------------------------------------------------------------
struct alignas(8) Pk { float a, b; };                     // lane 0, lane 1

#if defined(__CUDACC__)
__constant__ float rr_pk_ones[2] = {1.0f, -1.0f};          // opaque to ptxas (see header)
#endif

RR_HD Pk pk_make(float a, float b) { Pk r; r.a = a; r.b = b; return r; }
RR_HD Pk pk_bcast(float w) { return pk_make(w, w); }
RR_HD Pk pk_neg(Pk x) { return pk_make(-x.a, -x.b); }

template <> struct Arith<Pk> {
#if defined(__CUDA_ARCH__)
#define RR_PK_BIN(NAME, OP)                                                                                        \
  static __device__ __forceinline__ Pk NAME(Pk x, Pk y)                                                            \
  {                                                                                                                \
    Pk r;                                                                                                          \
    asm("{\n .reg .b64 pa, pb, pc;\n mov.b64 pa, {%2, %3};\n mov.b64 pb, {%4, %5};\n " OP                         \
        ".rn.f32x2 pc, pa, pb;\n mov.b64 {%0, %1}, pc;\n}"                                                         \
        : "=f"(r.a), "=f"(r.b)                                                                                     \
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));                                                                 \
    return r;                                                                                                      \
  }
  RR_PK_BIN(mul, "mul")
  RR_PK_BIN(add, "add")
  RR_PK_BIN(sub, "sub")
#undef RR_PK_BIN
  // x + y / x - y where an operand is a product: fma(x, 1, y) / fma(y, -1, x), see header
  static __device__ __forceinline__ Pk addp(Pk x, Pk y)
  {
    Pk r;
    const float one = rr_pk_ones[0];
    asm("{\n .reg .b64 pa, pb, pc, po;\n mov.b64 pa, {%2, %3};\n mov.b64 pb, {%4, %5};\n mov.b64 po, {%6, %6};\n"
        " fma.rn.f32x2 pc, pa, po, pb;\n mov.b64 {%0, %1}, pc;\n}"
        : "=f"(r.a), "=f"(r.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b), "f"(one));
    return r;
  }
  static __device__ __forceinline__ Pk subp(Pk x, Pk y)
  {
    Pk r;
    const float mone = rr_pk_ones[1];
    asm("{\n .reg .b64 pa, pb, pc, po;\n mov.b64 pa, {%2, %3};\n mov.b64 pb, {%4, %5};\n mov.b64 po, {%6, %6};\n"
        " fma.rn.f32x2 pc, pb, po, pa;\n mov.b64 {%0, %1}, pc;\n}"
        : "=f"(r.a), "=f"(r.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b), "f"(mone));
    return r;
  }
#else
  static RR_HD Pk mul(Pk x, Pk y) { return pk_make(x.a * y.a, x.b * y.b); }   // host build: -ffp-contract=off
  static RR_HD Pk add(Pk x, Pk y) { return pk_make(x.a + y.a, x.b + y.b); }
  static RR_HD Pk sub(Pk x, Pk y) { return pk_make(x.a - y.a, x.b - y.b); }
  static RR_HD Pk addp(Pk x, Pk y) { return add(x, y); }
  static RR_HD Pk subp(Pk x, Pk y) { return sub(x, y); }
#endif
};

typedef C2<Pk> CPk;                                        // one 16-byte slot: {re.a, re.b, im.a, im.b}

RR_PROG Pk pk_load8(const Pk *p)                         // one 8-byte shared / global load
{
#if defined(__CUDA_ARCH__)
  const float2 v = *reinterpret_cast<const float2 *>(p);
  return pk_make(v.x, v.y);
#else
  return *p;
#endif
}

// ---------------------------------------------------------------------------------------------------
// Thread groups
// ---------------------------------------------------------------------------------------------------
struct Grp { int tid, size, bar; };                        // index inside the group, threads, named barrier id
constexpr int kPkGroupThreads = 128, kPkMaxGroups = 4;    // threads per group, groups per CTA of the lane-pair DFT kernels
constexpr int kPkInplaceGroups = 5;                        // ... of the single-buffer variant (pk_spectrum_inplace)

#if defined(__CUDACC__)
RR_PROG void grp_sync(const Grp &g) { asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.size) : "memory"); }
#else
inline void grp_sync(const Grp &) {}
#endif
template <class F> RR_PROG void grp_for(const Grp &g, int count, F f)
{
  for (int i = g.tid; i < count; i += g.size) f(i);
  grp_sync(g);
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory slot of FFT position p. Consecutive threads touch (a) consecutive positions, (b) positions
// 16 apart (leaves) and (c) positions whose high bits differ (writes through the split-radix permutation);
// the three terms keep all of these at most two-way conflicting for 16-byte accesses. The map is additive,
// so slot(o + x) = slot(o) + slot(x) whenever o is a multiple of a power of two > x (node offsets).
// ---------------------------------------------------------------------------------------------------
RR_HD int pslot(int p) { return p + (p >> 4) + (p >> 8); }
RR_HD int pk_buf_slots(int m) { return ((pslot(m - 1) + 1 + 7) / 8) * 8; }

// Where the top pass may put the block's valid samples instead of shared memory.
struct PkSink {                  // passed by value: lives in registers
  float *d0, *d1;                // planar: lane pointers at the block's first output; interleaved: d0 only
  int es;                        // interleaved: elements between consecutive samples of a lane; 0 = planar
  int half;                      // complex elements to store (valid samples / 2)
};

RR_PROG void pk_sink_store(const PkSink &k, int c, const CPk &v)
{
  if (c >= k.half) return;
  if (k.es == 0) {                                       // planar: (sample 2c, 2c+1) of each lane
    reinterpret_cast<C2<float> *>(k.d0)[c] = C2<float>{v.x.a, v.y.a};
    reinterpret_cast<C2<float> *>(k.d1)[c] = C2<float>{v.x.b, v.y.b};
  } else {                                               // adjacent interleaved lanes: one pair per frame
    float *f = k.d0 + (long long)(2 * c) * k.es;
    if (k.es == 2 && !((size_t)k.d0 & 15)) *reinterpret_cast<CPk *>(f) = v;   // two stereo frames = one slot
    else {
      *reinterpret_cast<Pk *>(f) = v.x;
      *reinterpret_cast<Pk *>(f + k.es) = v.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// The FFT of M = 1 << BITS points over FFmpeg's split-radix DAG, everything about its shape a compile-time
// constant (pk_plan.hpp); `tasks` is the task table of the transform (shared memory), `pyr` its twiddle pyramid.
// ---------------------------------------------------------------------------------------------------
// Leaves (fft16 / fft8, fft.c:288-318), in place at position `off`.
template <int NV>
RR_PROG void pk_leaf(CPk *buf, int off, Pk sh, Pk c1, Pk c3)
{
  CPk *b = buf + pslot(off);                             // multiple of NV: the NV slots are contiguous
  Pk re[NV], im[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const CPk v = b[e];
    re[e] = v.x; im[e] = v.y;
  }
  if constexpr (NV == 16) leaf_fft16<Pk>(re, im, sh, c1, c3);
  else leaf_fft8<Pk>(re, im, sh);
#pragma unroll
  for (int e = 0; e < NV; ++e) b[e] = CPk{re[e], im[e]};
}

#if defined(__CUDA_ARCH__)
RR_PROG void pk_warp_sync() { __syncwarp(); }
#else
inline void pk_warp_sync() {}
#endif

// Combining passes (pass(), fft.c:237-256), fused 1, 2 or 3 levels deep. S = 1 << LG is the smallest size of
// the task, q = S/4; the task at position o = node offset + k (k < q) owns the values at o + j*q. pslot() is
// additive over the node offset, so every address is `base + constant`.
template <int LG> struct PkGeo {
  static constexpr int q = 1 << (LG - 2);
  static RR_HD constexpr int rel(int j) { return j * q + ((j * q) >> 4) + ((j * q) >> 8); }   // pslot(j*q)
};

RR_PROG void pk_bfly(CPk &a0, CPk &a1, CPk &a2, CPk &a3, float wre, float wim, bool zero)
{
  sr_bfly<Pk>(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y, pk_bcast(wre), pk_bcast(wim), zero);
}

// DEPTH 1: one butterfly of size S.  DEPTH 2: sizes S and 2S on a node of size 2S (8 values): butterfly k of the
// S-pass on the first half, then butterflies k and k + S/4 of the 2S-pass (the quarters of size S/2 are
// finished).  DEPTH 3: sizes S, 2S, 4S on a node of size 4S (16 values): the S-pass on the first quarter of the
// first half and on both quarter children, the 2S-pass on the first half, four butterflies of the 4S-pass.
// Same butterflies on the same operands as the level-by-level order of fft.c:265-272, hence the same bits.
// Loads and butterflies of one task; the results stay in e[] (value j belongs at position o + j * q).
template <int LG, int DEPTH>
RR_PROG void pk_item_regs(int o, const CPk *buf, const float *pyr, CPk (&e)[4 << (DEPTH - 1)])
{
  typedef PkGeo<LG> G;
  constexpr int q = G::q, NV = 4 << (DEPTH - 1);
  const int k = o & (q - 1);
  const CPk *b = buf + pslot(o);
  const float *twa = pyr + pk_pyr_off(LG);
#pragma unroll
  for (int j = 0; j < NV; ++j) e[j] = b[G::rel(j)];
  {
    const float wr = twa[k], wi = twa[q - k];
    pk_bfly(e[0], e[1], e[2], e[3], wr, wi, k == 0);
    if constexpr (DEPTH == 3) {
      pk_bfly(e[8], e[9], e[10], e[11], wr, wi, k == 0);
      pk_bfly(e[12], e[13], e[14], e[15], wr, wi, k == 0);
    }
  }
  if constexpr (DEPTH >= 2) {
    const float *twb = pyr + pk_pyr_off(LG + 1);
    pk_bfly(e[0], e[2], e[4], e[6], twb[k], twb[2 * q - k], k == 0);
    pk_bfly(e[1], e[3], e[5], e[7], twb[k + q], twb[q - k], false);
  }
  if constexpr (DEPTH == 3) {
    const float *twc = pyr + pk_pyr_off(LG + 2);
#pragma unroll
    for (int m = 0; m < 4; ++m)
      pk_bfly(e[m], e[m + 4], e[m + 8], e[m + 12], twc[k + m * q], twc[(4 - m) * q - k], m == 0 && k == 0);
  }
}

template <int LG, int DEPTH, bool SINK>
RR_PROG void pk_item(int o, CPk *buf, const float *pyr, const PkSink &sink)
{
  typedef PkGeo<LG> G;
  constexpr int q = G::q, NV = 4 << (DEPTH - 1);
  CPk e[NV];
  pk_item_regs<LG, DEPTH>(o, buf, pyr, e);
  if (SINK) {
#pragma unroll
    for (int j = 0; j < NV; ++j) pk_sink_store(sink, o + j * q, e[j]);
  } else {
    CPk *b = buf + pslot(o);
#pragma unroll
    for (int j = 0; j < NV; ++j) b[G::rel(j)] = e[j];
  }
}

// Phase PH of the transform. Light tasks (one level shallower) are bundled in threes so that every task of a
// phase costs about the same.
template <int BITS, int PH, bool SINK>
RR_PROG void pk_fft_phase(const Grp &g, CPk *buf, const uint16_t *tasks, const float *pyr, const PkSink &sink)
{
  constexpr PkPhaseList pl = pk_phase_list(BITS);
  constexpr int LG = pl.lg[PH], D = pl.depth[PH];
  constexpr int base = pk_phase_base(BITS, PH), nmain = pk_phase_main(BITS, LG, D), nlight = pk_phase_light(BITS, LG, D);
  constexpr int nb = (nlight + 2) / 3, DL = D > 1 ? D - 1 : 1;
  constexpr bool top = PH == pl.n - 1;                   // one node at offset 0: task t is position t, no table needed
  for (int t = g.tid; t < nmain + nb; t += g.size) {
    if (t < nmain) pk_item<LG, D, SINK>(top ? t : tasks[base + t], buf, pyr, sink);
    else
      for (int j = t - nmain; j < nlight; j += nb) pk_item<LG, DL, false>(tasks[base + nmain + j], buf, pyr, sink);
  }
  if (!SINK) grp_sync(g);                                // a sinking top phase leaves nothing in shared memory
}

// The barrier-free part: warp w of the group runs the leaves and the local phases of the quarter of the
// transform it owns (pk_plan.hpp), synchronising only with itself. `ltab`: the transform's local task table.
template <int BITS, int PH>
RR_PROG void pk_local_phase(int w, int lane, int nl, CPk *buf, const uint16_t *ltab, const float *pyr)
{
  constexpr PkPhaseList pl = pk_phase_list(BITS);
  constexpr int LG = pl.lg[PH], D = pl.depth[PH], DL = D > 1 ? D - 1 : 1;
  const uint16_t *hd = ltab + 4 * ((1 + PH) * kPkWarps + w);
  const PkSink none{nullptr, nullptr, 0, 0};
  const int mb = hd[0], mc = hd[1], lb = hd[2], lc = hd[3];
  for (int t = lane; t < mc; t += nl) pk_item<LG, D, false>(ltab[mb + t], buf, pyr, none);
  for (int t = lane; t < lc; t += nl) pk_item<LG, DL, false>(ltab[lb + t], buf, pyr, none);
  pk_warp_sync();
}

template <int BITS>
RR_PROG void pk_fft_local(const Grp &g, CPk *buf, const uint16_t *ltab, const float *pyr, float sqrthalf, float c16_1,
                          float c16_3)
{
  constexpr int nlocal = pk_local_phases(BITS);
  const Pk sh = pk_bcast(sqrthalf), c1 = pk_bcast(c16_1), c3 = pk_bcast(c16_3);
#if defined(__CUDA_ARCH__)
  const int w0 = g.tid >> 5, w1 = w0 + 1, lane = g.tid & 31, nl = 32;
#else
  const int w0 = 0, w1 = kPkWarps, lane = g.tid, nl = g.size;   // emulation: one thread plays the warps in turn
#endif
  for (int w = w0; w < w1; ++w) {
    const uint16_t *hd = ltab + 4 * w;
    const int mb = hd[0], mc = hd[1], lb = hd[2], lc = hd[3];
    for (int t = lane; t < mc; t += nl) { const int off = ltab[mb + t]; if (off != kPkHole) pk_leaf<16>(buf, off, sh, c1, c3); }
    for (int t = lane; t < lc; t += nl) { const int off = ltab[lb + t]; if (off != kPkHole) pk_leaf<8>(buf, off, sh, c1, c3); }
    pk_warp_sync();
    if constexpr (nlocal > 0) pk_local_phase<BITS, 0>(w, lane, nl, buf, ltab, pyr);
    if constexpr (nlocal > 1) pk_local_phase<BITS, 1>(w, lane, nl, buf, ltab, pyr);
    if constexpr (nlocal > 2) pk_local_phase<BITS, 2>(w, lane, nl, buf, ltab, pyr);
  }
}

// The warp-local part followed by the group barrier / the top phase, inlined: for kernels specialised on the size.
template <int BITS>
RR_PROG void pk_fft_lower_impl(const Grp &g, CPk *buf, const uint16_t *ltab, const float *pyr, float sqrthalf, float c16_1,
                               float c16_3)
{
  // every supported size (6 <= BITS <= 13) has exactly one phase above the warp-local ones: the top one
  static_assert(BITS < 6 || pk_local_phases(BITS) + 1 == pk_phase_list(BITS).n, "one global phase expected");
  pk_fft_local<BITS>(g, buf, ltab, pyr, sqrthalf, c16_1, c16_3);
  grp_sync(g);
}
template <int BITS, bool SINK>
RR_PROG void pk_fft_top_impl(const Grp &g, CPk *buf, const uint16_t *tasks, const float *pyr, const PkSink &sink)
{
  pk_fft_phase<BITS, pk_phase_list(BITS).n - 1, SINK>(g, buf, tasks, pyr, sink);
}

#if defined(__CUDACC__)
#define RR_PK_CALL __device__ __noinline__
#else
#define RR_PK_CALL inline
#endif

// The same as separate, not inlined functions for the size-generic kernels: one copy per size serves the
// forward and the inverse transform.
template <int BITS>
RR_PK_CALL void pk_fft_lower(Grp g, CPk *buf, const uint16_t *ltab, const float *pyr, float sqrthalf, float c16_1, float c16_3)
{
  pk_fft_lower_impl<BITS>(g, buf, ltab, pyr, sqrthalf, c16_1, c16_3);
}
template <int BITS, bool SINK>
RR_PK_CALL void pk_fft_top(Grp g, CPk *buf, const uint16_t *tasks, const float *pyr, PkSink sink)
{
  pk_fft_top_impl<BITS, SINK>(g, buf, tasks, pyr, sink);
}

#define RR_PK_BITS_SWITCH(BITS_EXPR, CALL)                                                                   \
  switch (BITS_EXPR) {                                                                                       \
    case 6: CALL(6); break;   case 7: CALL(7); break;                                                        \
    case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break;                            \
    case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break;                            \
    default: break;                                                                                          \
  }

// BITS > 0: size known at compile time, everything inlined; BITS == 0: dispatch on the run-time size.
template <int BITS>
RR_PROG void pk_fft_lower_any(int bits, const Grp &g, CPk *buf, const uint16_t *ltab, const float *pyr, float sqrthalf,
                              float c16_1, float c16_3)
{
  if constexpr (BITS > 0) pk_fft_lower_impl<BITS>(g, buf, ltab, pyr, sqrthalf, c16_1, c16_3);
  else {
#define RR_PK_LOWER(B) pk_fft_lower<B>(g, buf, ltab, pyr, sqrthalf, c16_1, c16_3)
    RR_PK_BITS_SWITCH(bits, RR_PK_LOWER)
#undef RR_PK_LOWER
  }
}
template <int BITS>
RR_PROG void pk_fft_top_any(int bits, const Grp &g, CPk *buf, const uint16_t *tasks, const float *pyr, bool use_sink, const PkSink &sink)
{
  if constexpr (BITS > 0) {
    if (use_sink) pk_fft_top_impl<BITS, true>(g, buf, tasks, pyr, sink);
    else pk_fft_top_impl<BITS, false>(g, buf, tasks, pyr, sink);
  } else {
#define RR_PK_TOP_SINK(B) pk_fft_top<B, true>(g, buf, tasks, pyr, sink)
#define RR_PK_TOP(B) pk_fft_top<B, false>(g, buf, tasks, pyr, sink)
    if (use_sink) { RR_PK_BITS_SWITCH(bits, RR_PK_TOP_SINK) }
    else { RR_PK_BITS_SWITCH(bits, RR_PK_TOP) }
#undef RR_PK_TOP_SINK
#undef RR_PK_TOP
  }
}

// ---------------------------------------------------------------------------------------------------
// The stage
// ---------------------------------------------------------------------------------------------------
enum PkSpecMode {
  PK_SPEC_UP2 = 0,    // F-domain up-sampling by 2, step 1: Ni = 2 Pf (44.1->48, 48->44.1, 44.1->96, ...)
  PK_SPEC_SAME = 1,   // Ni = Pf, step >= 1 (plain / zero-stuffed input)
  PK_SPEC_GEN = 2     // everything else: separate post-processing and spectrum phases
};

// Constants of one index of the fused spectrum phase (modes UP2 / SAME): filter spectrum values, real-FFT
// cosines and the inverse transform's slots. 64 bytes, fetched with four 16-byte loads.
struct alignas(16) PkSpecConst {
  C2<float> c0, c1, c2, c3;
  float tfc, tfs, tic, tis;
  unsigned s01, s23;             // two 16-bit slots each: where the results go in the inverse buffer
  unsigned fab;                  // two 16-bit slots: the record's inputs F[i], F[M - i] in the forward buffer
  unsigned first;                // 1: the record of index 0 (the two real bins and the self-paired bin M/2)
};

struct DftPkParams {
  DftParams<float> base;         // geometry, views, cosine tables, filter spectrum (schedules unused)
  const uint16_t *ltab_f, *ltab_i, *perm_f, *perm_i;   // local task tables (pk_plan.hpp), permutations
  const PkSpecConst *spec;       // [M/2] records (modes UP2 / SAME); index 0 holds those of M/2 and of the two real bins
  int fb, ib;                    // log2 of the forward / inverse complex transform sizes
  int fslots, bslots;            // slots of the forward / inverse buffer
  int halo_slots;                // slots between them (fused DFT + polyphase kernel: history in front of B), else 0
  int groups, gthreads;          // groups per CTA, threads per group
  int spec_mode;
  int stereo;                    // every lane pair is the two channels of adjacent stereo frames (16-byte tile loads)
  int lay_pyr_f, lay_pyr_i, lay_ltab_f, lay_ltab_i, lay_perm_f, lay_data;     // shared-memory byte offsets (host: pk_smem_layout)
  int n_pyr_f, n_pyr_i, n_ltab_f, n_ltab_i;                                   // table lengths
};

// Shared memory: tables (twiddle pyramids, task tables, forward permutation), then per group F and B.
struct PkSmemLayout { int pyr_f, pyr_i, ltab_f, ltab_i, perm_f, data, group_slots; size_t total; };
RR_HD PkSmemLayout pk_smem_layout(const DftPkParams &pp)
{
  PkSmemLayout l;
  int o = 0;
  l.pyr_f = o; o += 4 * pk_pyr_len(pp.fb);
  l.pyr_i = o; o += 4 * pk_pyr_len(pp.ib);
  l.ltab_f = o; o += 2 * pp.n_ltab_f;
  l.ltab_i = o; o += 2 * pp.n_ltab_i;
  o = (o + 3) & ~3;
  l.perm_f = o; o += 2 << pp.fb;
  l.data = (o + 15) & ~15;
  l.group_slots = pp.fslots + pp.halo_slots + pp.bslots;
  l.total = (size_t)l.data + (size_t)pp.groups * l.group_slots * sizeof(CPk);
  return l;
}

// How the input tile of an item is copied: whole slots when the lanes are adjacent in an interleaved buffer,
// 8 bytes per lane from planar lanes (re-paired by the leaves), else scalar copies with zero fill.
enum PkTileMode { PK_TILE_SCALAR = 0, PK_TILE_INTERLEAVED = 1, PK_TILE_PLANAR = 2 };

// Everything about a work item that needs 64-bit coordinate arithmetic, computed by one thread per item.
struct PkItem {
  DftItem<float> d;
  const float *s0, *s1;          // lane pointers at the tile's first input sample (direct tiles)
  float *d0, *d1;                // lane pointers at the block's first output sample
  long long c0;                  // its coordinate
  int tile_mode;
  int first, stride, count;      // kept samples: block sample first + j * stride, j < count
  int direct;                    // outputs are stored contiguously (no ring wrap / clipping)
  int sink;                      // 0: via shared memory, 1: top pass stores planar pairs, 2: interleaved frames
};

RR_PROG PkItem pk_make_item(const DftPkParams &pp, long long work)
{
  const DftParams<float> &p = pp.base;
  PkItem it;
  it.d = dft_item<float, 2>(p, work);
  it.tile_mode = PK_TILE_SCALAR;
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N;
  it.s0 = view_ptr<const float>(p.in, it.d.in_off0, it.d.Rb); it.s1 = view_ptr<const float>(p.in, it.d.in_off1, it.d.Rb);
  if (p.in_mode != DFT_IN_ZERO_STUFF && view_range_direct(p.in, it.d.Rb, it.d.Rb + span)) {
    const int es = p.in.elem_stride;
    if (it.s1 == it.s0 + 1 && !(es & 1) && !((size_t)it.s0 & 7)) it.tile_mode = PK_TILE_INTERLEAVED;
    else if (es == 1 && !(((size_t)it.s0 | (size_t)it.s1) & 7)) it.tile_mode = PK_TILE_PLANAR;
  }
  // output geometry of the block (as in dft_stage_program)
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
  it.direct = view_range_direct(p.out, it.c0, it.c0 + it.count);
  it.d0 = view_ptr<float>(p.out, it.d.out_off0, it.c0); it.d1 = view_ptr<float>(p.out, it.d.out_off1, it.c0);
  const int es = p.out.elem_stride;
  it.sink = 0;
  if (it.direct && it.stride == 1 && !(it.count & 1)) {
    if (es == 1 && !(((size_t)it.d0 | (size_t)it.d1) & 7)) it.sink = 1;
    else if (it.d1 == it.d0 + 1 && !(es & 1) && !((size_t)it.d0 & 7)) it.sink = 2;
  }
  return it;
}

#if defined(__CUDA_ARCH__)
RR_PROG void pk_async_copy8(void *smem_dst, const void *gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
RR_PROG void pk_async_copy16(void *smem_dst, const void *gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
#else
inline void pk_async_copy8(void *dst, const void *src) { memcpy(dst, src, 8); }
inline void pk_async_copy16(void *dst, const void *src) { memcpy(dst, src, 16); }
#endif

// Brings an item's input tile into `F`, element j at the slot of its permuted position (perm: the forward
// transform's table, shared memory): coalesced global loads into registers (eight elements in flight per thread),
// then 16-byte shared stores. Round 1 prefetched the tile with LDGSTS during the previous item's inverse transform;
// a scattered LDGSTS, however, costs one shared-memory wavefront per 16-byte element (32 per warp copy: 20 % of the
// kernel's wavefronts by ncu) where the same scatter as STS.128 is served eight lanes per wavefront. The other
// groups of the CTA cover the exposed load latency, and the lines are requested into L2 ahead of time
// (pk_tile_prefetch). Planar lanes are re-paired in registers, so the leaves always see {re.a, re.b, im.a, im.b}.
// STEREO: the launch reads adjacent stereo frames (two frames = one slot, 16-byte loads); items that are not
// stored that way (clipped at a stream edge, wrapping in a ring) go element by element. !STEREO: any layout.
// The two variants are separate kernels: with every path inlined into one, ptxas spills inside the spectrum phase.
template <int FB, bool STEREO>
RR_PROG void pk_tile_now(const DftPkParams &pp, const Grp &g, const PkItem &it, CPk *F, const uint16_t *perm)
{
  const DftParams<float> &p = pp.base;
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N;
  const int m = FB > 0 ? (1 << FB) : (span >> 1);
  if (STEREO) {
    if (it.tile_mode == PK_TILE_INTERLEAVED && p.in.elem_stride == 2) {
      if (!((size_t)it.s0 & 15)) {
        const CPk *src = reinterpret_cast<const CPk *>(it.s0);
        for (int j0 = g.tid; j0 < m; j0 += 8 * g.size) {
          CPk v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { const int j = j0 + k * g.size; if (j < m) v[k] = ldg(src + j); }
#pragma unroll
          for (int k = 0; k < 8; ++k) { const int j = j0 + k * g.size; if (j < m) F[perm[j]] = v[k]; }
        }
      } else {                                            // the stream starts on an odd frame: 8-byte loads
        const Pk *src = reinterpret_cast<const Pk *>(it.s0);
        for (int j0 = g.tid; j0 < m; j0 += 4 * g.size) {
          Pk a[4], b[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { const int j = j0 + k * g.size; if (j < m) { a[k] = pk_load8(src + 2 * j); b[k] = pk_load8(src + 2 * j + 1); } }
#pragma unroll
          for (int k = 0; k < 4; ++k) { const int j = j0 + k * g.size; if (j < m) F[perm[j]] = CPk{a[k], b[k]}; }
        }
      }
      return;
    }
  } else {
    if (it.tile_mode == PK_TILE_INTERLEAVED) {            // the pair's 8 bytes of every frame, frames es floats apart
      const int es = p.in.elem_stride;
      const float *s0 = it.s0;
      for (int j0 = g.tid; j0 < m; j0 += 4 * g.size) {
        Pk a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + k * g.size;
          if (j < m) { const float *f = s0 + (long long)(2 * j) * es; a[k] = pk_load8(reinterpret_cast<const Pk *>(f)); b[k] = pk_load8(reinterpret_cast<const Pk *>(f + es)); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = j0 + k * g.size; if (j < m) F[perm[j]] = CPk{a[k], b[k]}; }
      }
      return;
    }
    if (it.tile_mode == PK_TILE_PLANAR) {                 // {x[2j], x[2j+1]} of each lane -> {re.a, re.b, im.a, im.b}
      const Pk *s0 = reinterpret_cast<const Pk *>(it.s0), *s1 = reinterpret_cast<const Pk *>(it.s1);
      for (int j0 = g.tid; j0 < m; j0 += 4 * g.size) {
        Pk a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = j0 + k * g.size; if (j < m) { a[k] = pk_load8(s0 + j); b[k] = pk_load8(s1 + j); } }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + k * g.size;
          if (j < m) F[perm[j]] = CPk{pk_make(a[k].a, b[k].a), pk_make(a[k].b, b[k].b)};
        }
      }
      return;
    }
  }
  // zero-stuffed, clipped or wrapping tiles: element by element through the view
  const int L = p.L;
  for (int w = g.tid; w < 2 * span; w += g.size) {
    const int l = w & 1, j = w >> 1;
    long long coord = it.d.Rb + j;
    bool on_grid = true;
    if (p.in_mode == DFT_IN_ZERO_STUFF) {
      const int d = j - it.d.remLb;
      on_grid = d >= 0 && d % L == 0;
      coord = it.d.Rb + (on_grid ? d / L : 0);
    }
    float *dst = reinterpret_cast<float *>(F + perm[j >> 1]) + 2 * (j & 1) + l;
    *dst = on_grid ? view_read<float, float>(p.in, l ? it.d.in_off1 : it.d.in_off0, coord) : 0.0f;
  }
}

// Ask L2 for the next item's tile (one 128-byte line per thread and round) while this item's inverse transform runs.
RR_PROG void pk_tile_prefetch(const DftPkParams &pp, const Grp &g, const PkItem &it)
{
#if defined(__CUDA_ARCH__)
  const DftParams<float> &p = pp.base;
  if (it.tile_mode == PK_TILE_SCALAR) return;
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N;
  if (it.tile_mode == PK_TILE_PLANAR) {
    const int lines = (span * 4 + 127) >> 7;
    for (int k = g.tid; k < 2 * lines; k += g.size) {
      const char *base = reinterpret_cast<const char *>(k & 1 ? it.s1 : it.s0);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(k >> 1) * 128));
    }
  } else {
    const size_t bytes = (size_t)span * p.in.elem_stride * 4;     // frames of all channels between the pair's samples
    const int lines = (int)((bytes + 127) >> 7);
    const char *base = reinterpret_cast<const char *>(it.s0);
    for (int k = g.tid; k < lines; k += g.size) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)k * 128));
  }
#else
  (void)pp; (void)g; (void)it;
#endif
}

RR_PROG PkSpecConst pk_load_spec(const PkSpecConst *p)
{
#if defined(__CUDA_ARCH__)
  // field-by-field from three 16-byte and one 8-byte load (no memcpy: the record must stay in registers)
  const float4 *q = reinterpret_cast<const float4 *>(p);
  const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  const uint4 d = __ldg(reinterpret_cast<const uint4 *>(p) + 3);
  PkSpecConst r;
  r.c0 = C2<float>{a.x, a.y}; r.c1 = C2<float>{a.z, a.w};
  r.c2 = C2<float>{b.x, b.y}; r.c3 = C2<float>{b.z, b.w};
  r.tfc = c.x; r.tfs = c.y; r.tic = c.z; r.tis = c.w;
  r.s01 = d.x; r.s23 = d.y; r.fab = d.z; r.first = d.w;
  return r;
#else
  return *p;
#endif
}

// rdft.c:46-77, forward transform: (F[i], F[M-i]) -> (X[i], X[M-i])
RR_HD void pk_post_pair(const CPk &za, const CPk &zb, float c, float s, CPk &xa, CPk &xb)
{
  typedef Arith<Pk> A;
  const Pk half = pk_bcast(0.5f), pc = pk_bcast(c), ps = pk_bcast(s);
  const Pk evr = A::mul(half, A::add(za.x, zb.x));
  const Pk odi = A::mul(half, A::sub(zb.x, za.x));
  const Pk evi = A::mul(half, A::sub(za.y, zb.y));
  const Pk odr = A::mul(half, A::add(za.y, zb.y));
  const Pk sr = A::addp(A::mul(odr, pc), A::mul(odi, ps));
  const Pk si = A::subp(A::mul(odi, pc), A::mul(odr, ps));
  xa = CPk{A::addp(evr, sr), A::addp(evi, si)};
  xb = CPk{A::subp(evr, sr), A::subp(si, evi)};
}

RR_HD CPk pk_cmul(const C2<float> &cf, const CPk &v)      // dft_filter.h:140-145
{
  typedef Arith<Pk> A;
  const Pk cx = pk_bcast(cf.x), cy = pk_bcast(cf.y);
  return CPk{A::subp(A::mul(cx, v.x), A::mul(cy, v.y)), A::addp(A::mul(cy, v.x), A::mul(cx, v.y))};
}

// filter multiply of the bin pair (va = spectrum[i], vb = spectrum[Ni/2 - i]) and the inverse transform's
// pre-processing (rdft.c:44-80 with inverse = 1) -> (d[i], d[Ni/2 - i])
RR_HD void pk_mul_pre_pair(const CPk &va, const CPk &vb, const C2<float> &ca, const C2<float> &cb, float c, float s,
                           CPk &da, CPk &db)
{
  typedef Arith<Pk> A;
  const Pk half = pk_bcast(0.5f), mhalf = pk_bcast(-0.5f), pc = pk_bcast(c), ps = pk_bcast(s);
  const CPk a = pk_cmul(ca, va), bb = pk_cmul(cb, vb);
  const Pk evr = A::mul(half, A::add(a.x, bb.x));
  const Pk odi = A::mul(mhalf, A::sub(bb.x, a.x));
  const Pk evi = A::mul(half, A::sub(a.y, bb.y));
  const Pk odr = A::mul(mhalf, A::add(a.y, bb.y));
  const Pk sr = A::subp(A::mul(odr, pc), A::mul(odi, ps));
  const Pk si = A::addp(A::mul(odi, pc), A::mul(odr, ps));
  da = CPk{A::addp(evr, sr), A::addp(evi, si)};
  db = CPk{A::subp(evr, sr), A::subp(si, evi)};
}

RR_HD CPk pk_conj(const CPk &v) { return CPk{v.x, pk_neg(v.y)}; }

// Spectrum phase, F (forward FFT result, natural order) -> B (input of the inverse FFT, permuted order).
// Round r of thread t handles index i = t + r * g.size, i < M/2 (i = 0 also does M/2 and the two real bins);
// the first kPkSpecRounds rounds find their constants in `pre` (loaded before the previous phase), the next
// kPkSpecLate rounds load theirs when the phase starts, ahead of the arithmetic of the first rounds.
constexpr int kPkSpecRounds = 1, kPkSpecLate = 3;
constexpr bool kPkPrefetchAcrossTop = true;
struct PkSpecRegs { PkSpecConst r[kPkSpecRounds]; };

template <int MODE>
RR_PROG void pk_spec_prefetch(const DftPkParams &pp, const Grp &g, PkSpecRegs &pre)
{
  const int n = pp.base.Pf >> 2;
#pragma unroll
  for (int r = 0; r < kPkSpecRounds; ++r) {
    const int i = g.tid + r * g.size;
    if (i < n) pre.r[r] = pk_load_spec(pp.spec + i);
  }
}

// One record of the spectrum phase: (za, zb) = (F[i], F[M - i]) -- for i == 0: (F[0], F[M/2]) -- and its constants k;
// writes the two or four bins of the inverse transform's input they determine. The records are stored in an order
// chosen for bank-conflict-free accesses (build_pk_spec), so a record carries the slots of its own inputs.
template <int MODE>
RR_PROG void pk_spec_index(const PkSpecConst &k, const CPk &za, const CPk &zb, CPk *B)
{
  typedef Arith<Pk> A;
  const int sl0 = k.s01 & 0xffff, sl1 = k.s01 >> 16, sl2 = k.s23 & 0xffff, sl3 = k.s23 >> 16;
  if (k.first) {
    // bins 0 and Pf/2 (packed in F[0]) and the self-paired bin M/2
    // record 0 carries the constants of M/2 in c0, c1, tic, tis, s01 and those of the real bins in the
    // fields index 0 has no use for: c2 = coef[0], c3 = coef[M], s23 = slots of d[0] and d[M]
    const CPk x0 = CPk{A::add(za.x, za.y), A::sub(za.x, za.y)};                // rdft.c:46-48
    CPk zm = zb;
    zm.y = pk_neg(zm.y);                                                       // rdft.c:77 (forward)
    if (MODE == PK_SPEC_UP2) {
      // spectrum[0] = (X0.re, X0.re); d[0] = (.5 (d0 + d1), .5 (d0 - d1)), dft_filter.h:96-98,118-119, rdft.c:44-46,79-80
      const Pk d0 = A::mul(x0.x, pk_bcast(k.c2.x)), d1 = A::mul(x0.x, pk_bcast(k.c2.y));
      B[sl2] = CPk{A::mul(A::addp(d0, d1), pk_bcast(0.5f)), A::mul(A::subp(d0, d1), pk_bcast(0.5f))};
      // bin Ni/4 = M: spectrum = (X0.im, 0); d[M] = conj(coef[M] * spectrum)
      B[sl3] = pk_conj(pk_cmul(k.c3, CPk{x0.y, pk_bcast(0.0f)}));
      // bins M/2 and Ni/2 - M/2: spectrum[M/2] = X[M/2], spectrum[Ni/2 - M/2] = conj(X[M/2])
      CPk da, db;
      pk_mul_pre_pair(zm, pk_conj(zm), k.c0, k.c1, k.tic, k.tis, da, db);
      B[sl0] = da; B[sl1] = db;
    } else {
      const Pk d0 = A::mul(x0.x, pk_bcast(k.c2.x)), d1 = A::mul(x0.y, pk_bcast(k.c2.y));
      B[sl2] = CPk{A::mul(A::addp(d0, d1), pk_bcast(0.5f)), A::mul(A::subp(d0, d1), pk_bcast(0.5f))};
      // bin Ni/4 = M/2: d = conj(coef * X[M/2])
      B[sl0] = pk_conj(pk_cmul(k.c0, zm));
    }
    return;
  }
  CPk xa, xb;
  pk_post_pair(za, zb, k.tfc, k.tfs, xa, xb);
  if (MODE == PK_SPEC_UP2) {
    // bins i / Ni/2 - i see X[i] / conj(X[i]); bins M - i / M + i see X[M-i] / conj(X[M-i])
    CPk da, db;
    pk_mul_pre_pair(xa, pk_conj(xa), k.c0, k.c1, k.tic, k.tis, da, db);
    B[sl0] = da; B[sl1] = db;
    pk_mul_pre_pair(xb, pk_conj(xb), k.c2, k.c3, k.tis, k.tic, da, db);
    B[sl2] = da; B[sl3] = db;
  } else {
    CPk da, db;
    pk_mul_pre_pair(xa, xb, k.c0, k.c1, k.tic, k.tis, da, db);
    B[sl0] = da; B[sl1] = db;
  }
}

template <int MODE, bool PIPE = true>
RR_PROG void pk_spectrum(const DftPkParams &pp, const Grp &g, const PkSpecRegs &pre, const CPk *F, CPk *B)
{
  const DftParams<float> &p = pp.base;
  const int M = p.Pf >> 1;                               // forward transform: M complex points
  const int n = M >> 1;
  auto body = [&](int, const PkSpecConst &k) { pk_spec_index<MODE>(k, F[k.fab & 0xffff], F[k.fab >> 16], B); };
  (void)M;
  // software pipeline, two records live: the record of the next round is requested before a round is computed
  static_assert(kPkSpecRounds == 1 && kPkSpecLate == 3, "pipeline below is written for 1 + 3 rounds");
  const int i0 = g.tid, i1 = i0 + g.size, i2 = i1 + g.size, i3 = i2 + g.size;
  if (!PIPE) {                                            // one record live at a time (kernels short of registers)
    if (i0 < n) body(i0, pre.r[0]);
    for (int i = i1; i < n; i += g.size) body(i, pk_load_spec(pp.spec + i));
    grp_sync(g);
    return;
  }
  PkSpecConst ra = pre.r[0], rb;
  if (i1 < n) rb = pk_load_spec(pp.spec + i1);
  if (i0 < n) body(i0, ra);
  if (i2 < n) ra = pk_load_spec(pp.spec + i2);
  if (i1 < n) body(i1, rb);
  if (i3 < n) rb = pk_load_spec(pp.spec + i3);
  if (i2 < n) body(i2, ra);
  if (i3 < n) body(i3, rb);
  for (int i = g.tid + (kPkSpecRounds + kPkSpecLate) * g.size; i < n; i += g.size) body(i, pk_load_spec(pp.spec + i));
  grp_sync(g);
}

// The same phase IN PLACE: the forward transform's result and the inverse transform's input share one buffer W (the
// forward result occupies its first slots). Every thread first takes its pairs (F[i], F[M - i]) into registers, the
// group synchronises, then the bins are written -- so a group needs one buffer instead of two and five groups fit an
// SM where four did. Compile-time size only (the register arrays must be static).
template <int MODE, int FB>
RR_PROG void pk_spectrum_inplace(const DftPkParams &pp, const Grp &g, CPk *W)
{
  constexpr int n = 1 << (FB - 1);
#if defined(__CUDA_ARCH__)
  constexpr int R = (n + kPkGroupThreads - 1) / kPkGroupThreads;
  CPk za[R], zb[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = g.tid + r * kPkGroupThreads;
    if (e < n) { const unsigned fab = ldg(&pp.spec[e].fab); za[r] = W[fab & 0xffff]; zb[r] = W[fab >> 16]; }
  }
  grp_sync(g);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = g.tid + r * kPkGroupThreads;
    if (e < n) pk_spec_index<MODE>(pk_load_spec(pp.spec + e), za[r], zb[r], W);
  }
  grp_sync(g);
#else
  (void)g;
  CPk *za = new CPk[n], *zb = new CPk[n];
  for (int e = 0; e < n; ++e) { const unsigned fab = pp.spec[e].fab; za[e] = W[fab & 0xffff]; zb[e] = W[fab >> 16]; }
  for (int e = 0; e < n; ++e) pk_spec_index<MODE>(pk_load_spec(pp.spec + e), za[e], zb[e], W);
  delete[] za; delete[] zb;
#endif
}

// Generic spectrum path (any L, F-domain decimation): forward post-processing in place, then the bins of
// the inverse transform one pair at a time through the reference's up-sampling index map (dft_filter.h:86-104).
RR_PROG CPk pk_spec_freq_up(const CPk *X, int Pf, int idx)
{
  const int twoP = Pf << 1, r = idx & (twoP - 1);
  if (r == 0) { const Pk a0 = X[0].x; return CPk{a0, idx == 0 ? a0 : pk_bcast(0.0f)}; }
  if (r < Pf) return X[pslot(r >> 1)];
  if (r == Pf) return CPk{X[0].y, pk_bcast(0.0f)};
  return pk_conj(X[pslot((twoP - r) >> 1)]);
}

RR_PROG void pk_spectrum_generic(const DftPkParams &pp, const Grp &g, CPk *F, CPk *B)
{
  typedef Arith<Pk> A;
  const DftParams<float> &p = pp.base;
  const int Pf = p.Pf, Ni = p.Ni;
  grp_for(g, (Pf >> 2) + 1, [&](int i) {
    if (i == 0) {
      const CPk z = F[0];
      F[0] = CPk{A::add(z.x, z.y), A::sub(z.x, z.y)};
    } else if (i == (Pf >> 2)) {
      CPk *z = F + pslot(Pf >> 2);
      z->y = pk_neg(z->y);
    } else {
      const int ia = pslot(i), ib = pslot((Pf >> 1) - i);
      CPk xa, xb;
      pk_post_pair(F[ia], F[ib], ldg(p.tcos_f + i), ldg(p.tcos_f + (Pf >> 2) - i), xa, xb);
      F[ia] = xa; F[ib] = xb;
    }
  });
  const bool freq_up = p.in_mode == DFT_IN_FREQ_UP;
  const C2<float> *coef = reinterpret_cast<const C2<float> *>(p.coef);
  const uint16_t *perm = pp.perm_i;
  grp_for(g, (Ni >> 2) + 1, [&](int i) {
    auto spec = [&](int bin) -> CPk { return freq_up ? pk_spec_freq_up(F, Pf, 2 * bin) : F[pslot(bin)]; };
    if (i > 0 && i < (Ni >> 2)) {
      CPk da, db;
      pk_mul_pre_pair(spec(i), spec((Ni >> 1) - i), ldg(coef + i), ldg(coef + (Ni >> 1) - i), ldg(p.tcos_i + i),
                      ldg(p.tcos_i + (Ni >> 2) - i), da, db);
      B[ldg(perm + i)] = da; B[ldg(perm + (Ni >> 1) - i)] = db;
    } else if (i == 0) {
      const CPk v0 = spec(0);
      const C2<float> ca = ldg(coef);
      const Pk d0 = A::mul(v0.x, pk_bcast(ca.x));
      Pk d1;
      if (p.step > 0) d1 = A::mul(v0.y, pk_bcast(ca.y));
      else {                                              // new Nyquist bin of the decimated spectrum, dft_filter.h:185
        const CPk vn = spec(Ni >> 1);
        const C2<float> cb = ldg(coef + (Ni >> 1));
        d1 = A::subp(A::mul(pk_bcast(cb.x), vn.x), A::mul(pk_bcast(cb.y), vn.y));
      }
      B[ldg(perm)] = CPk{A::mul(A::addp(d0, d1), pk_bcast(0.5f)), A::mul(A::subp(d0, d1), pk_bcast(0.5f))};
    } else {
      B[ldg(perm + (Ni >> 2))] = pk_conj(pk_cmul(ldg(coef + (Ni >> 2)), spec(Ni >> 2)));
    }
  });
}

// Pointers into the CTA's shared tables.
struct PkTables { const float *pyr_f, *pyr_i; const uint16_t *ltab_f, *ltab_i, *perm_f; };

// One work item (block b, lane pair). F holds (or is receiving) the item's input tile; items[slot] describes
// this item, items[slot ^ 1] is filled for the next one, whose tile is requested as soon as F is free.
template <int MODE, int FB, int IB, bool STEREO, bool INPLACE = false>
RR_PROG void dftp_program(const DftPkParams &pp, const Grp &g, const PkTables &tb, PkItem *items, int slot, long long work_next,
                          CPk *F, CPk *B)
{
  const DftParams<float> &p = pp.base;
  grp_sync(g);                                            // items[slot] is visible; F and B are free
  const PkItem &it = items[slot];                        // stays valid for the whole item; fields are read where needed
  if (work_next >= 0 && g.tid == 0) items[slot ^ 1] = pk_make_item(pp, work_next);
  pk_tile_now<FB, STEREO>(pp, g, it, F, tb.perm_f);
  grp_sync(g);

  PkSink sink{nullptr, nullptr, 0, 0};
  pk_fft_lower_any<FB>(pp.fb, g, F, tb.ltab_f, tb.pyr_f, p.sqrthalf, p.c16_1, p.c16_3);
  if (MODE == PK_SPEC_GEN) {
    pk_fft_top_any<FB>(pp.fb, g, F, nullptr, tb.pyr_f, false, sink);
    pk_spectrum_generic(pp, g, F, B);
  } else {
    if constexpr (INPLACE) {
      static_assert(FB > 0, "in-place spectrum phase needs a compile-time size");
      pk_fft_top_any<FB>(pp.fb, g, F, nullptr, tb.pyr_f, false, sink);
      pk_spectrum_inplace<MODE, FB>(pp, g, B);                                  // F == B
    } else {
      PkSpecRegs pre;
      if (FB > 0 && kPkPrefetchAcrossTop) pk_spec_prefetch<MODE>(pp, g, pre);   // in flight while the (inlined) top forward phase runs
      pk_fft_top_any<FB>(pp.fb, g, F, nullptr, tb.pyr_f, false, sink);
      if (FB == 0 || !kPkPrefetchAcrossTop) pk_spec_prefetch<MODE>(pp, g, pre);   // not across a call: the registers would be spilled
      pk_spectrum<MODE>(pp, g, pre, F, B);
    }
  }
  if (work_next >= 0) pk_tile_prefetch(pp, g, items[slot ^ 1]);   // published before the barriers of the forward transform

  pk_fft_lower_any<IB>(pp.ib, g, B, tb.ltab_i, tb.pyr_i, p.sqrthalf, p.c16_1, p.c16_3);
  sink.d0 = it.d0; sink.d1 = it.d1; sink.half = it.count >> 1; sink.es = it.sink == 2 ? p.out.elem_stride : 0;
  pk_fft_top_any<IB>(pp.ib, g, B, nullptr, tb.pyr_i, it.sink != 0, sink);
  if (it.sink) return;

  const int first = it.first, stride = it.stride, count = it.count, es = p.out.elem_stride;
  const bool direct = it.direct != 0;
  float *d0 = it.d0, *d1 = it.d1;
  const long long c0 = it.c0;
  const float *Br = reinterpret_cast<const float *>(B);
  grp_for(g, 2 * count, [&](int w) {
    const int l = w & 1, j = w >> 1;
    const int t = first + j * stride;
    const float v = Br[4 * pslot(t >> 1) + 2 * (t & 1) + l];
    if (direct) (l ? d1 : d0)[(long long)j * es] = v;
    else view_write<float, float>(p.out, l ? it.d.out_off1 : it.d.out_off0, c0 + j, v);
  });
}


// ---------------------------------------------------------------------------------------------------
// vpoly0 for lane pairs (rate_filters_generic.h:272-305): the phase-stationary scheme of poly0_fast_kernel with
// both channels of a pair in one thread -- the input window is staged pair-interleaved (one LDS.64 per tap
// feeds both channels), products and sums are FMUL2 / FFMA2(x, 1, y) -- and with the slots of a period dealt to
// the threads so that the sixteen lanes of a half-warp read sixteen different 8-byte banks.
// ---------------------------------------------------------------------------------------------------
struct Poly0PairParams {
  Poly0FastParams<float> fast;   // tile geometry (CH = lanes per CTA, even), window, tiling
  int P;                         // pairs per CTA = CH / 2
  int PG;                        // period groups: thread (slot, pair, g) takes periods g, g + PG, ...
  int tslots;                    // threads along the slot dimension (multiple of 16, >= slots per column)
  int spread;                    // deal the slots over the banks (needs one column per period)
  int CL;                        // slots per thread: 1, or 2 adjacent slots sharing their input window (poly0_pair2_*)
  // Shifted windows (poly0_pair2_*): the first input samples of a period's slot clusters do not fall evenly into the 16
  // eight-byte bank pairs (cfg4: 2 ... 7 clusters per bank for 5 rows of threads). A cluster of an overfull bank b is
  // moved to bank b - 1 by starting its window ONE SAMPLE EARLIER (that extra sample carries no tap), which makes every
  // half-warp read 16 different banks in every step. keep[b]: clusters that stay in bank b (rows 0 .. keep[b] - 1); the
  // others go to bank b - 1, rows keep[b - 1] ... Solved on the host from the same bank counts the device deal sees.
  int shift;
  unsigned char keep[16];
};

// thread slot ts -> slot of the column (or 0xffff): slot_of[j * 16 + b] = the j-th slot (cluster) whose first
// input sample falls into bank pair b, as long as row j exists; what does not fit (a bank with more entries than
// rows) goes to an overflow list and is put into the remaining holes by poly0_pair_deal_overflow after a barrier
// (a few two-way conflicts instead of idle threads). cnt: 16 bank counters + 1 overflow counter, zero on entry.
constexpr int kPolyDealOverflow = 128;
constexpr int kPolyShifted = 0x8000;                       // flag in slot_of: this cluster's window starts one sample early
RR_PROG void poly0_pair_deal(const Poly0PairParams &pp, const Poly0Tile &t, uint16_t *slot_of, int *cnt, uint16_t *ovf, int tid,
                             int nthreads)
{
  const PolyParams<float> &p = pp.fast.base;
  for (int fs = tid * pp.CL; fs < t.nslots; fs += nthreads * pp.CL) {        // first slot of each cluster
    const unsigned at_rel = (unsigned)t.r_first + (unsigned)fs * (unsigned)p.step;
    const int b = (int)((at_rel / (unsigned)p.L) & 15);
#if defined(__CUDA_ARCH__)
    const int j = atomicAdd(cnt + b, 1);
#else
    const int j = cnt[b]++;
#endif
    if (pp.shift) {                                        // every cluster has a conflict-free place (host-solved)
      const int b2 = (b + 15) & 15;
      if (j < pp.keep[b]) slot_of[j * 16 + b] = (uint16_t)fs;
      else slot_of[(pp.keep[b2] + j - pp.keep[b]) * 16 + b2] = (uint16_t)(fs | kPolyShifted);
      continue;
    }
    if (j * 16 + b < pp.tslots) slot_of[j * 16 + b] = (uint16_t)fs;
    else {
#if defined(__CUDA_ARCH__)
      const int k = atomicAdd(cnt + 16, 1);
#else
      const int k = cnt[16]++;
#endif
      if (k < kPolyDealOverflow) ovf[k] = (uint16_t)fs;
    }
  }
}
RR_PROG void poly0_pair_deal_overflow(const Poly0PairParams &pp, const Poly0Tile &t, uint16_t *slot_of, const int *cnt, const uint16_t *ovf, int tid)
{
  if (tid != 0 || pp.shift) return;
  (void)t;
  const int n = cnt[16] < kPolyDealOverflow ? cnt[16] : kPolyDealOverflow;
  int hole = 0;
  for (int k = 0; k < n; ++k) {
    while (hole < pp.tslots && slot_of[hole] != 0xffff) ++hole;
    if (hole < pp.tslots) slot_of[hole] = ovf[k];
  }
}

// Stage the pair-interleaved input windows of a tile. Does not wait. tma != 0 (decided per tile by
// poly0_pair_make_tile): the windows are contiguous 16-byte aligned ranges of a pair-interleaved FIFO and thread 0
// moves each of them with one bulk copy (TMA) that completes on `bar`; `head` samples in front of the window are
// copied along so that source and destination share their 16-byte phase.
RR_PROG void poly0_pair_load(const Poly0PairParams &pp, const Poly0Tile &t, int tma, int head, Pk *buf, unsigned long long *bar, int tid,
                             int nthreads)
{
  const Poly0FastParams<float> &fp = pp.fast;
  const PolyParams<float> &p = fp.base;
  const long long c0 = t.q_first + p.pre;
  if (tma) {
    if (tid != 0) return;
    const unsigned bytes = (unsigned)(((t.win + head + 1) & ~1) * (int)sizeof(Pk));
    tma_bar_expect(bar, bytes * (unsigned)pp.P);
    for (int pr = 0; pr < pp.P; ++pr) {
      const float *s0 = view_ptr<const float>(p.in, lane_offset(p.in, t.lane0 + 2 * pr), c0) - 2 * head;
      tma_load_1d(buf + pr * fp.win, s0, bytes, bar);
    }
    return;
  }
  const bool direct = view_range_direct(p.in, c0, c0 + t.win);
  const int es = p.in.elem_stride;
  for (int pr = 0; pr < pp.P; ++pr) {
    const long long off0 = lane_offset(p.in, t.lane0 + 2 * pr), off1 = lane_offset(p.in, t.lane0 + 2 * pr + 1);
    Pk *dst = buf + pr * fp.win;
    const float *s0 = view_ptr<const float>(p.in, off0, c0), *s1 = view_ptr<const float>(p.in, off1, c0);
    if (direct && s1 == s0 + 1 && !(es & 1) && !((size_t)s0 & 7)) {
      for (int j = tid; j < t.win; j += nthreads) pk_async_copy8(dst + j, s0 + (long long)j * es);
    } else if (direct) {
      for (int w = tid; w < 2 * t.win; w += nthreads) {
        const int l = w & 1, j = w >> 1;
        async_copy_elem<float>(&dst[j].a + l, (l ? s1 : s0) + (long long)j * es, true);
      }
    } else {
      for (int w = tid; w < 2 * t.win; w += nthreads) {
        const int l = w & 1, j = w >> 1;
        bool valid;
        const float *src = view_addr<float>(p.in, l ? off1 : off0, c0 + j, &valid);
        async_copy_elem<float>(&dst[j].a + l, src, valid);
      }
    }
  }
  async_copy_commit();
}

// What a thread keeps for the whole launch: its slot (hence phase r and first-sample offset q: one column per
// period, so neither depends on the tile), pair, period group and the coefficient row of the phase.
template <int NT> struct Poly0PairThread {
  int fs, q, pr, g;              // fs < 0: no work (hole of the deal / padding)
  float c[NT];
};

template <int NT>
RR_PROG Poly0PairThread<NT> poly0_pair_setup(const Poly0PairParams &pp, const Poly0Tile &t, const uint16_t *slot_of, int w)
{
  const PolyParams<float> &p = pp.fast.base;
  Poly0PairThread<NT> st;
  const int per_group = pp.tslots * pp.P;
  st.g = w / per_group;
  const int rest = w - st.g * per_group;
  st.pr = rest / pp.tslots;
  const int ts = rest - st.pr * pp.tslots;
  const int raw = st.g < pp.PG ? (pp.spread ? (int)slot_of[ts] : ts) : 0xffff;
  const int fs = raw;
  st.fs = fs < t.nslots ? fs : -1;
  const unsigned at_rel = (unsigned)t.r_first + (unsigned)(st.fs < 0 ? 0 : st.fs) * (unsigned)p.step;
  st.q = (int)(at_rel / (unsigned)p.L);
  const int r = (int)(at_rel - (unsigned)st.q * (unsigned)p.L);
  const float *row = p.coefs + (long long)r * NT;
#pragma unroll
  for (int k = 0; k < NT; ++k) st.c[k] = ldg(row + k);
  return st;
}


// Tile description for the pair kernel: the generic tile plus everything about its output that needs 64-bit
// arithmetic, computed by one thread per tile.
struct Poly0PairTile {
  Poly0Tile t;
  float *d_base;                 // first output sample of the tile (slot 0, period 0) of lane t.lane0
  long long lane1;               // elements from a pair's first lane to its second (the two need not share a stream)
  long long pair_stride;         // elements from one pair of the CTA to the next in the output (channels of one stream, or
                                 // stereo pairs of consecutive streams)
  long long i_end;               // outputs at or beyond this index do not exist
  int direct;                    // every output of the tile exists and is stored contiguously
  int tma, head;                 // the windows are staged by bulk copies; window sample j then sits at buffer index head + j
};

RR_PROG Poly0PairTile poly0_pair_make_tile(const Poly0PairParams &pp, long long work)
{
  const PolyParams<float> &p = pp.fast.base;
  Poly0PairTile pt;
  pt.t = poly0_tile(pp.fast, work);
  pt.i_end = p.out0 + p.nout;
  const long long i_tile_end = pt.t.i_first + (long long)pt.t.mcount * p.L;
  pt.direct = i_tile_end <= pt.i_end && view_range_direct(p.out, p.out_preload + pt.t.i_first, p.out_preload + i_tile_end);
  pt.d_base = view_ptr<float>(p.out, lane_offset(p.out, pt.t.lane0), p.out_preload + pt.t.i_first);
  pt.lane1 = lane_offset(p.out, pt.t.lane0 + 1) - lane_offset(p.out, pt.t.lane0);
  pt.pair_stride = pp.P > 1 ? lane_offset(p.out, pt.t.lane0 + 2) - lane_offset(p.out, pt.t.lane0) : 0;
  // bulk-copy staging: every window of the tile a contiguous range of adjacent stereo frames, one sample of slack
  // either side inside the view (the copies are rounded out to 16-byte boundaries)
  const long long c0 = pt.t.q_first + p.pre;
  pt.tma = 0; pt.head = 0;
  if (p.in.elem_stride == 2 && view_range_direct(p.in, c0 - 1, c0 + pt.t.win + 1) && pt.t.win + 2 <= pp.fast.win) {
    const float *s0 = view_ptr<const float>(p.in, lane_offset(p.in, pt.t.lane0), c0);
    pt.head = (int)(((size_t)s0 >> 3) & 1);
    pt.tma = 1;
    for (int pr = 0; pr < pp.P; ++pr) {
      const long long off0 = lane_offset(p.in, pt.t.lane0 + 2 * pr), off1 = lane_offset(p.in, pt.t.lane0 + 2 * pr + 1);
      const float *q0 = view_ptr<const float>(p.in, off0, c0) - 2 * pt.head;
      if (off1 != off0 + 1 || ((size_t)q0 & 15)) pt.tma = 0;
    }
    if (!pt.tma) pt.head = 0;
  }
  return pt;
}

template <int NT>
RR_PROG void poly0_pair_tile(const Poly0PairParams &pp, const Poly0PairTile &pt, const Pk *buf, const Poly0PairThread<NT> &st)
{
  typedef Arith<Pk> A;
  if (st.fs < 0) return;
  const Poly0Tile &t = pt.t;
  const Poly0FastParams<float> &fp = pp.fast;
  const PolyParams<float> &p = fp.base;
  const int L = p.L, PG = pp.PG, es = p.out.elem_stride;
  // the P pairs of a CTA are equally spaced in the output: channels of one stream, or stereo pairs of consecutive streams
  const long long rel = (long long)(st.fs + st.g * L) * es + st.pr * pt.pair_stride;
  float *d0 = pt.d_base + rel, *d1 = d0 + pt.lane1;
  const int dstep = PG * L * es;                          // between this thread's consecutive outputs
  const bool direct = pt.direct != 0;
  const bool packed_out = direct && d1 == d0 + 1 && !((size_t)d0 & 7) && !(dstep & 1);
  const int xstep = PG * (int)p.step;
  const Pk *x = buf + st.pr * fp.win + pt.head + st.q + st.g * (int)p.step;
  int m = st.g;
  auto emit = [&](int mm, Pk s) {
    if (packed_out) *reinterpret_cast<Pk *>(d0) = s;
    else if (direct) { *d0 = s.a; *d1 = s.b; }
    else {
      const long long i = t.i_first + st.fs + (long long)mm * L;
      if (i < pt.i_end) {
        const int lane_a = t.lane0 + 2 * st.pr;
        view_write<float, float>(p.out, lane_offset(p.out, lane_a), p.out_preload + i, s.a);
        view_write<float, float>(p.out, lane_offset(p.out, lane_a + 1), p.out_preload + i, s.b);
      }
    }
    d0 += dstep; d1 += dstep;
  };
  for (; m + PG < t.mcount; m += 2 * PG, x += 2 * xstep) {                    // two periods at a time: two independent chains
    const Pk *xb = x + xstep;
    Pk sa = pk_bcast(0.0f), sb = pk_bcast(0.0f);
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const Pk ck = pk_bcast(st.c[k]);
      sa = A::addp(sa, A::mul(ck, pk_load8(x + k)));
      sb = A::addp(sb, A::mul(ck, pk_load8(xb + k)));
    }
    emit(m, sa); emit(m + PG, sb);
  }
  if (m < t.mcount) {
    Pk sa = pk_bcast(0.0f);
#pragma unroll
    for (int k = 0; k < NT; ++k) sa = A::addp(sa, A::mul(pk_bcast(st.c[k]), pk_load8(x + k)));
    emit(m, sa);
  }
}


// Two adjacent slots per thread: outputs i and i + 1 read windows that start d = q(i+1) - q(i) samples apart,
// d = DLO or DLO + 1 with DLO = floor(step / L), so one pass over NT + DLO + 1 window samples feeds both: 1.8x
// fewer shared-memory reads per output (the kernel's bound). The second row is kept shifted by d in registers;
// only its first and last window positions depend on d (one predicate). Same products, same order per output.
// A thread whose cluster was moved one bank down by the deal (pp.shift) starts its pass one sample early: window
// position j then carries tap j - 1 of the first output, so the pass covers NT + DLO + 2 positions for everybody and
// each thread skips the one or two positions at either end that carry no tap of its outputs (predicates, no zero taps:
// a sample outside an output's window never touches it, whatever it holds).
template <int NT, int DLO> struct Poly0Pair2Thread {
  int fs, q, pr;                 // fs < 0: no work; q: first window sample of the pass (already moved by sh)
  int sh, f1;                    // the first output's taps start at position sh (0 / 1), the second's at DLO + f1 (0 / 1 / 2)
  bool two;                      // the second slot exists
  float c0[NT + 1], c1[NT + 2];  // by window position: c0[j] at position j, c1[j] at position DLO + j
};

template <int NT, int DLO>
RR_PROG Poly0Pair2Thread<NT, DLO> poly0_pair2_setup(const Poly0PairParams &pp, const Poly0Tile &t, const uint16_t *slot_of, int w)
{
  const PolyParams<float> &p = pp.fast.base;
  Poly0Pair2Thread<NT, DLO> st;
  st.pr = w / pp.tslots;
  const int ts = w - st.pr * pp.tslots;
  const int raw = st.pr < pp.P ? (pp.spread ? (int)slot_of[ts] : 2 * ts) : 0xffff;
  const int fs = raw == 0xffff ? raw : (raw & (kPolyShifted - 1));
  st.sh = (raw != 0xffff && (raw & kPolyShifted)) ? 1 : 0;
  st.fs = fs < t.nslots ? fs : -1;
  const unsigned at0 = (unsigned)t.r_first + (unsigned)(st.fs < 0 ? 0 : st.fs) * (unsigned)p.step;
  const int q0 = (int)(at0 / (unsigned)p.L);
  const int r0 = (int)(at0 - (unsigned)q0 * (unsigned)p.L);
  const unsigned at1 = at0 + (unsigned)p.step;
  const int q1 = (int)(at1 / (unsigned)p.L), r1 = (int)(at1 - (unsigned)q1 * (unsigned)p.L);
  st.q = q0 - st.sh;
  st.two = st.fs >= 0 && st.fs + 1 < t.nslots;
  st.f1 = q1 - q0 - DLO + st.sh;                          // d = DLO or DLO + 1
  const float *row0 = p.coefs + (long long)r0 * NT, *row1 = p.coefs + (long long)(st.two ? r1 : r0) * NT;
#pragma unroll
  for (int j = 0; j <= NT; ++j) {
    const int k = j - st.sh;
    st.c0[j] = (k >= 0 && k < NT) ? ldg(row0 + k) : 0.f;
  }
#pragma unroll
  for (int j = 0; j <= NT + 1; ++j) {
    const int k = j - st.f1;
    st.c1[j] = (k >= 0 && k < NT) ? ldg(row1 + k) : 0.f;
  }
  return st;
}

template <int NT, int DLO>
RR_PROG void poly0_pair2_tile(const Poly0PairParams &pp, const Poly0PairTile &pt, const Pk *buf, const Poly0Pair2Thread<NT, DLO> &st)
{
  typedef Arith<Pk> A;
  if (st.fs < 0) return;
  const Poly0Tile &t = pt.t;
  const Poly0FastParams<float> &fp = pp.fast;
  const PolyParams<float> &p = fp.base;
  const int L = p.L, es = p.out.elem_stride;
  const long long rel = (long long)st.fs * es + st.pr * pt.pair_stride;   // the P pairs of a CTA are equally spaced in the output
  float *d0 = pt.d_base + rel, *d1 = d0 + pt.lane1;
  const int dstep = L * es;
  const bool direct = pt.direct != 0;
  const bool packed_out = direct && d1 == d0 + 1 && !((size_t)d0 & 7) && !(dstep & 1) && !(es & 1);
  const int xstep = (int)p.step;
  const Pk *x = buf + st.pr * fp.win + pt.head + st.q;       // st.q may be -1: the buffer has two elements of slack in front
  const bool two = st.two, sh = st.sh != 0;
  const int f1 = st.f1;
  auto emit = [&](int mm, int which, int ahead, Pk s) {  // output of slot fs + which in period mm (= `ahead` periods past d0)
    float *e0 = d0 + which * es + ahead * dstep, *e1 = d1 + which * es + ahead * dstep;
    if (packed_out) *reinterpret_cast<Pk *>(e0) = s;
    else if (direct) { *e0 = s.a; *e1 = s.b; }
    else {
      const long long i = t.i_first + st.fs + which + (long long)mm * L;
      if (i < pt.i_end) {
        const int lane_a = t.lane0 + 2 * st.pr;
        view_write<float, float>(p.out, lane_offset(p.out, lane_a), p.out_preload + i, s.a);
        view_write<float, float>(p.out, lane_offset(p.out, lane_a + 1), p.out_preload + i, s.b);
      }
    }
  };
  // which window positions carry a tap of the first / second output (see Poly0Pair2Thread)
  auto on0 = [&](int j) { return j == 0 ? !sh : (j == NT ? sh : true); };
  auto on1 = [&](int jj) { return jj == 0 ? f1 == 0 : (jj == 1 ? f1 <= 1 : (jj == NT ? f1 >= 1 : (jj == NT + 1 ? f1 == 2 : true))); };
  int m = 0;
  // two periods at a time: four independent sums per thread (the pass is bound by the latency of its LDS -> FMUL2 ->
  // FFMA2 chains once the bank conflicts are gone), same products in the same order per output
  for (; m + 1 < t.mcount; m += 2, x += 2 * xstep, d0 += 2 * dstep, d1 += 2 * dstep) {
    const Pk *xb = x + xstep;
    Pk s0 = pk_bcast(0.0f), s1 = pk_bcast(0.0f), u0 = pk_bcast(0.0f), u1 = pk_bcast(0.0f);
#pragma unroll
    for (int j = 0; j < NT + DLO + 2; ++j) {
      const Pk xv = pk_load8(x + j), xw = pk_load8(xb + j);
      if (j <= NT && on0(j)) {
        const Pk c = pk_bcast(st.c0[j]);
        s0 = A::addp(s0, A::mul(c, xv));
        u0 = A::addp(u0, A::mul(c, xw));
      }
      if (j >= DLO && on1(j - DLO)) {
        const Pk c = pk_bcast(st.c1[j - DLO]);
        s1 = A::addp(s1, A::mul(c, xv));
        u1 = A::addp(u1, A::mul(c, xw));
      }
    }
    emit(m, 0, 0, s0);
    if (two) emit(m, 1, 0, s1);
    emit(m + 1, 0, 1, u0);
    if (two) emit(m + 1, 1, 1, u1);
  }
  if (m < t.mcount) {
    Pk s0 = pk_bcast(0.0f), s1 = pk_bcast(0.0f);
#pragma unroll
    for (int j = 0; j < NT + DLO + 2; ++j) {
      const Pk xv = pk_load8(x + j);
      if (j <= NT && on0(j)) s0 = A::addp(s0, A::mul(pk_bcast(st.c0[j]), xv));
      if (j >= DLO && on1(j - DLO)) s1 = A::addp(s1, A::mul(pk_bcast(st.c1[j - DLO]), xv));
    }
    emit(m, 0, 0, s0);
    if (two) emit(m, 1, 0, s1);
  }
}

// ---------------------------------------------------------------------------------------------------
// Half-band 2:1 decimator (h8..h13, rate_filters_generic.h:80-249) for lane pairs: the scheme of
// halfband_program (window split by sample parity, four consecutive outputs per thread from 16-byte shared
// loads) with both channels of a pair in every value, so each add / multiply is one FADD2 / FMUL2.
// y[k] = 0.5 x[2k+pre] + sum_t c[t] (x[2k+pre-(2t+1)] + x[2k+pre+(2t+1)]), summed in that order.
// ---------------------------------------------------------------------------------------------------
struct HalfbandPairParams {
  HalfbandParams<float> base;    // coefficients, views, ranges; tile = outputs per pair per CTA (multiple of 4), half, qbits
  int G;                         // pairs per CTA: all channels of a stream when the input is interleaved, else 1
};

// Geometry of one tile -- `tile` outputs of the G pairs of one stream (or of one pair) -- with everything that
// needs divisions or 64-bit coordinate arithmetic, computed by one thread per tile.
struct HalfbandPairTile {
  int lane0, cnt, win;
  long long k0, x0;              // first output / first input coordinate (window index u = coord - x0)
  const float *src0;             // input sample x0 of lane0 (valid when in_direct)
  float *dst0;                   // output sample k0 of lane0 (valid when out_direct)
  long long in_lane1, out_lane1; // elements from a pair's first lane to its second
  long long in_off0, out_off0;   // lane offsets of lane0 (for the clipped / ring paths)
  int in_direct, out_direct;     // the whole window / output range of the tile is stored contiguously
};
template <int NC>
RR_PROG HalfbandPairTile halfband_pair_tile(const HalfbandPairParams &hp, long long work)
{
  const HalfbandParams<float> &p = hp.base;
  HalfbandPairTile t;
  const long long tiles = (p.nout + p.tile - 1) / p.tile;
  long long group; int tix_i;
  divmod_ll(work, (int)tiles, group, tix_i);
  t.lane0 = (int)group * 2 * hp.G;
  t.k0 = p.out0 + (long long)tix_i * p.tile;
  const long long rest = p.out0 + p.nout - t.k0;
  t.cnt = rest < p.tile ? (int)rest : p.tile;
  const int reach = 2 * NC - 1;
  t.x0 = 2 * t.k0 + p.pre - reach;
  t.win = 2 * (t.cnt - 1) + 2 * reach + 1;
  t.in_off0 = lane_offset(p.in, t.lane0); t.out_off0 = lane_offset(p.out, t.lane0);
  t.in_lane1 = lane_offset(p.in, t.lane0 + 1) - t.in_off0;
  t.out_lane1 = lane_offset(p.out, t.lane0 + 1) - t.out_off0;
  t.in_direct = view_range_direct(p.in, t.x0, t.x0 + t.win);
  const long long c0 = p.out_preload + t.k0;
  t.out_direct = view_range_direct(p.out, c0, c0 + t.cnt);
  t.src0 = view_ptr<const float>(p.in, t.in_off0, t.x0);
  t.dst0 = view_ptr<float>(p.out, t.out_off0, c0);
  return t;
}

// Stage the window of a tile, split by sample parity: even u -> P0[g][u/2]; odd u -> P1[g][(u+1)/2 + shift], so
// that the centre tap of output j (u = 2j + reach) sits at P1[j + 4]; odd samples below the first centre tap
// (index < 4) are never read and not stored. Asynchronous (LDGSTS): returns after committing the copies.
// Shared-memory index of window element i (8-byte units from the start of the buffer). A thread reads its window as
// 16-byte chunks and neighbouring threads start 32 bytes apart, so lanes t and t + 4 of a quarter-warp would meet in
// the same banks (two wavefronts per LDS.128: 42 % of the kernel's wavefronts were such conflicts while l1tex was 86 %
// busy). Swapping the two 16-byte chunk pairs of every second 128-byte line moves lane t + 4 one chunk over.
RR_HD int hb_swz(int i) { return i ^ ((i >> 3) & 2); }

template <int NC>
RR_PROG void halfband_pair_load(const HalfbandPairParams &hp, const HalfbandPairTile &t, Pk *smem, int tid, int nthreads)
{
  const HalfbandParams<float> &p = hp.base;
  const int G = hp.G, gbits = G == 4 ? 2 : G == 2 ? 1 : 0;      // G is 1, 2 or 4
  // the odd array starts 8 values (16 banks) further so that the even and the odd sample of a frame pair never
  // share a bank
  const int o1 = G * p.half + 8;                         // first element of the odd array
  const int shift = 4 - NC, win = t.win;
  const int ics = p.in.ch_stride, ies = p.in.elem_stride;
  // lane l of the tile relative to its first lane: channels of one stream when G > 1, any two lanes when G == 1
  auto lane_rel = [&](int l) -> long long { return G > 1 ? (long long)l * ics : (long long)l * t.in_lane1; };
  const bool direct = t.in_direct != 0;
  const float *src0 = t.src0;
  const int npairs_u = (win + 1) >> 1;                   // frame pairs (u = 2f, 2f + 1)
  if (direct && lane_rel(1) == 1 && !(ies & 1) && !((size_t)src0 & 7)) {
    // interleaved frames: the pairs of a frame are consecutive 8-byte words; one thread takes both frames of a
    // frame pair for one lane pair
    for (int w = tid; w < (npairs_u << gbits); w += nthreads) {
      const int g = w & (G - 1), f = w >> gbits;
      const float *sp = src0 + (2 * f) * ies + 2 * g;
      pk_async_copy8(smem + hb_swz(g * p.half + f), sp);
      const int io = f + 1 + shift;
      if (2 * f + 1 < win && io >= 4) pk_async_copy8(smem + hb_swz(o1 + g * p.half + io), sp + ies);
    }
  } else if (direct) {
    // planar lanes (or any regular strides): lane by lane, frame pairs fastest
    for (int l = 0; l < 2 * G; ++l) {
      const float *sl = src0 + lane_rel(l);
      const int b0 = (l >> 1) * p.half, b1 = o1 + b0;
      for (int f = tid; f < npairs_u; f += nthreads) {
        async_copy_elem<float>(&smem[hb_swz(b0 + f)].a + (l & 1), sl + (long long)(2 * f) * ies, true);
        const int io = f + 1 + shift;
        if (2 * f + 1 < win && io >= 4) async_copy_elem<float>(&smem[hb_swz(b1 + io)].a + (l & 1), sl + (long long)(2 * f + 1) * ies, true);
      }
    }
  } else {
    for (int w = tid; w < 2 * win * G; w += nthreads) {
      const int u = w % win, l = w / win, g = l >> 1;
      bool valid;
      const float *src = view_addr<float>(p.in, t.in_off0 + lane_rel(l), t.x0 + u, &valid);
      if (u & 1) {
        const int io = ((u + 1) >> 1) + shift;
        if (io >= 4) async_copy_elem<float>(&smem[hb_swz(o1 + g * p.half + io)].a + (l & 1), src, valid);
      } else async_copy_elem<float>(&smem[hb_swz(g * p.half + (u >> 1))].a + (l & 1), src, valid);
    }
  }
  async_copy_commit();
}

template <int NC>
RR_PROG void halfband_pair_compute(const HalfbandPairParams &hp, const float (&cf)[NC], const HalfbandPairTile &t, const Pk *smem,
                                   int tid, int nthreads)
{
  typedef Arith<Pk> A;
  const HalfbandParams<float> &p = hp.base;
  constexpr int c = NC;
  const int G = hp.G, cnt = t.cnt;
  const int o1 = G * p.half + 8;                         // first element of the odd array
  const int qbits = p.qbits;                             // log2(tile / 4)
  const int ocs = p.out.ch_stride, oes = p.out.elem_stride;
  const long long lane1 = G > 1 ? ocs : t.out_lane1;
  for (int w = tid; w < (G << qbits); w += nthreads) {
    const int g = w >> qbits, j = 4 * (w & ((1 << qbits) - 1));
    if (j >= cnt) continue;
    const int e = g * p.half + j, o = o1 + g * p.half + j + 4;
    // outputs j..j+3 use even-array elements j .. j+2c+2 and the centres, odd-array elements j+4 .. j+7; rows are 16-byte
    // aligned, element index -> shared-memory position through hb_swz (whole 16-byte chunks move)
    constexpr int kVecs = (2 * c + 3 + 1) / 2;
    Pk x[2 * kVecs], ctr[4];
#pragma unroll
    for (int i = 0; i < kVecs; ++i) { const CPk v = *reinterpret_cast<const CPk *>(smem + hb_swz(e + 2 * i)); x[2 * i] = v.x; x[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 2; ++i) { const CPk v = *reinterpret_cast<const CPk *>(smem + hb_swz(o + 2 * i)); ctr[2 * i] = v.x; ctr[2 * i + 1] = v.y; }
    Pk y[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      Pk sum = A::mul(ctr[r], pk_bcast(0.5f));
#pragma unroll
      for (int tt = 0; tt < c; ++tt) sum = A::addp(sum, A::mul(A::add(x[r + c - 1 - tt], x[r + c + tt]), pk_bcast(cf[tt])));
      y[r] = sum;
    }
    if (j + 4 <= cnt && t.out_direct) {
      float *da = t.dst0 + (G > 1 ? (2 * g) * ocs : 0) + j * oes, *db = da + lane1;
      if (lane1 == 1 && !(oes & 1) && !((size_t)da & 7)) {                // interleaved: one pair per frame
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<Pk *>(da + r * oes) = y[r];
      } else if (oes == 1 && !(((size_t)da | (size_t)db) & 15)) {         // planar, aligned: one vector store per lane
        struct alignas(16) O4 { float a, b, c, d; };
        *reinterpret_cast<O4 *>(da) = O4{y[0].a, y[1].a, y[2].a, y[3].a};
        *reinterpret_cast<O4 *>(db) = O4{y[0].b, y[1].b, y[2].b, y[3].b};
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) { da[r * oes] = y[r].a; db[r * oes] = y[r].b; }
      }
    } else {
      const long long off_a = t.out_off0 + (G > 1 ? (long long)(2 * g) * ocs : 0), off_b = off_a + lane1;
      const long long cbase = p.out_preload + t.k0 + j;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j + r < cnt) {
          view_write<float, float>(p.out, off_a, cbase + r, y[r].a);
          view_write<float, float>(p.out, off_b, cbase + r, y[r].b);
        }
    }
  }
}

}  // namespace b200rate
