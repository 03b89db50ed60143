// rate_kernels.cuh -- the stage kernels of the B200 rate engine, written once as "CTA programs".
//
// A CTA program is a sequence of phases; each phase is a set of independent work items executed by
// cta_for() (a strided loop over the CTA's threads followed by a barrier). Compiled by nvcc for sm_100a
// this is the product. The same text also compiles as plain C++ (tests/emu), where cta_for() is a serial
// loop: that build exists only so the index arithmetic can be checked against the oracle on a machine
// without a GPU; it is test infrastructure and is never linked into libb200rate.so.
//
// Reference behaviour per stage (paths under /root/reference/rate/):
//   dft_stage_program   dft_filter.h:60-190 with fft-float/{fft.c,rdft.c} (fp32) -- see below
//   poly0_program       rate_filters_generic.h:272-305 (vpoly0)
//   polyN_program       rate_filters_generic.h:311-504 (vpoly1..3)
//   halfband_program    rate_filters_generic.h:80-249  (h8..h13)
//
// fp32 bit-faithfulness: the reference's fp32 output is a function of its
// expression DAG, so every fp32 operation below goes through Arith<float>, which maps to
// __fmul_rn/__fadd_rn/__fsub_rn (never contracted into FFMA), and the FFT evaluates FFmpeg's
// conjugate-pair split-radix DAG -- leaves of size 16/8 in registers, then one combining pass per
// power of two -- with the reference's float-rounded cosine tables. The fp64 engine shares the code with
// Arith<double> (contraction allowed; its contract is 1e-12, not bit equality).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RR_HD __host__ __device__ __forceinline__   // small arithmetic / addressing helpers
#define RR_PROG __device__ __forceinline__          // CTA programs and their phases (device only under nvcc)
#else
#define RR_HD inline
#define RR_PROG inline
#endif

namespace b200rate {

// ---------------------------------------------------------------------------------------------------
// Arithmetic policy
// ---------------------------------------------------------------------------------------------------
template <class T> struct Arith;
template <> struct Arith<float> {
#if defined(__CUDA_ARCH__)
  static RR_HD float mul(float a, float b) { return __fmul_rn(a, b); }
  static RR_HD float add(float a, float b) { return __fadd_rn(a, b); }
  static RR_HD float sub(float a, float b) { return __fsub_rn(a, b); }
#else
  static RR_HD float mul(float a, float b) { return a * b; }   // host build: -ffp-contract=off
  static RR_HD float add(float a, float b) { return a + b; }
  static RR_HD float sub(float a, float b) { return a - b; }
#endif
};
template <> struct Arith<double> {
  static RR_HD double mul(double a, double b) { return a * b; }
  static RR_HD double add(double a, double b) { return a + b; }
  static RR_HD double sub(double a, double b) { return a - b; }
};

// ---------------------------------------------------------------------------------------------------
// CTA execution context
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
template <class F> RR_PROG void cta_for(int count, F f)
{
  for (int i = threadIdx.x; i < count; i += blockDim.x) f(i);
  __syncthreads();
}
RR_PROG void cta_sync() { __syncthreads(); }
#else
template <class F> inline void cta_for(int count, F f)
{
  for (int i = 0; i < count; ++i) f(i);
}
inline void cta_sync() {}
#endif

// ---------------------------------------------------------------------------------------------------
// Addressing of a FIFO (device ring buffer / linear intermediate) or of a caller's interleaved buffer.
// A lane is one channel of one stream. FIFO coordinates are ABSOLUTE stream positions.
// ---------------------------------------------------------------------------------------------------
struct LaneView {
  void *base;
  long long origin;              // FIFO coordinate stored at physical sample index 0
  unsigned long long mask;       // ring mask (capacity-1), ~0ull for a linear buffer
  long long lo, hi;              // coordinates outside [lo, hi) read as zero (preload / zero feed); writes there are dropped
  long long stream_stride;       // elements between consecutive streams
  int ch_stride;                 // elements between channels of one stream
  int elem_stride;               // elements between consecutive samples of one lane
  int nch;                       // lanes per stream
};

RR_HD long long lane_offset(const LaneView &v, int lane)
{
  return (long long)(lane / v.nch) * v.stream_stride + (long long)(lane % v.nch) * v.ch_stride;
}
template <class E, class T> RR_HD T view_read(const LaneView &v, long long lane_off, long long coord)
{
  if (coord < v.lo || coord >= v.hi) return (T)0;
  const unsigned long long phys = (unsigned long long)(coord - v.origin) & v.mask;
  return (T) static_cast<const E *>(v.base)[lane_off + (long long)phys * v.elem_stride];
}
template <class E, class T> RR_HD void view_write(const LaneView &v, long long lane_off, long long coord, T value)
{
  if (coord < v.lo || coord >= v.hi) return;            // clipping of whole-block producers
  const unsigned long long phys = (unsigned long long)(coord - v.origin) & v.mask;
  static_cast<E *>(v.base)[lane_off + (long long)phys * v.elem_stride] = (E)value;
}

// ---------------------------------------------------------------------------------------------------
// Complex FFT of size M = 1 << bits over FFmpeg's split-radix DAG (fft.c:186-346)
// ---------------------------------------------------------------------------------------------------
// Device-side schedule. Node lists are per power-of-two size; the gather tables are the out-of-place
// permutation of ff_fft_permute_c (fft.c:169-177) folded into the leaf loads and stored transposed
// ([element][leaf]) so consecutive threads read consecutive entries.
struct CfftSched {
  int bits;
  int n16, n8;                   // leaves: nodes of size 16, and size-8 quarter-children of size-32 nodes
  const uint16_t *leaf16_off;    // [n16] offset of the leaf in the permuted array
  const uint16_t *leaf8_off;     // [n8]
  const uint16_t *gather16;      // [16][n16] natural index feeding permuted element off+e
  const uint16_t *gather8;       // [8][n8]
  const uint16_t *node_off;      // concatenated node offsets for sizes 32 .. M
  int level_begin[17], level_cnt[17];
  int pyr_off[17];               // start of the twiddle row of each size inside the pyramid
};

// Bank-conflict padding of the FFT work buffer: one extra complex slot per 16.
RR_HD int cslot(int p) { return p + (p >> 4); }
RR_HD int cfft_buf_complex(int m) { return m + (m >> 4) + 1; }

template <class T>
RR_HD void sr_bfly(T &a0r, T &a0i, T &a1r, T &a1i, T &a2r, T &a2i, T &a3r, T &a3i, T wre, T wim, bool zero)
{
  typedef Arith<T> A;
  T t1, t2, t5, t6;
  if (zero) { t1 = a2r; t2 = a2i; t5 = a3r; t6 = a3i; }           // TRANSFORM_ZERO, fft.c:228-235
  else {                                                           // TRANSFORM, fft.c:222-226
    t1 = A::add(A::mul(a2r, wre), A::mul(a2i, wim));
    t2 = A::sub(A::mul(a2i, wre), A::mul(a2r, wim));
    t5 = A::sub(A::mul(a3r, wre), A::mul(a3i, wim));
    t6 = A::add(A::mul(a3r, wim), A::mul(a3i, wre));
  }
  const T t3 = A::sub(t5, t1); t5 = A::add(t5, t1);                // BUTTERFLIES, fft.c:200-207
  a2r = A::sub(a0r, t5); a0r = A::add(a0r, t5);
  a3i = A::sub(a1i, t3); a1i = A::add(a1i, t3);
  const T t4 = A::sub(t2, t6); t6 = A::add(t2, t6);
  a3r = A::sub(a1r, t4); a1r = A::add(a1r, t4);
  a2i = A::sub(a0i, t6); a0i = A::add(a0i, t6);
}

template <class T> RR_HD void leaf_fft4(T *re, T *im)             // fft4, fft.c:274-286
{
  typedef Arith<T> A;
  const T s01r = A::add(re[0], re[1]), d01r = A::sub(re[0], re[1]);
  const T s01i = A::add(im[0], im[1]), d01i = A::sub(im[0], im[1]);
  const T s32r = A::add(re[3], re[2]), d32r = A::sub(re[3], re[2]);
  const T s23i = A::add(im[2], im[3]), d23i = A::sub(im[2], im[3]);
  re[0] = A::add(s01r, s32r); re[2] = A::sub(s01r, s32r);
  im[0] = A::add(s01i, s23i); im[2] = A::sub(s01i, s23i);
  im[1] = A::add(d01i, d32r); im[3] = A::sub(d01i, d32r);
  re[1] = A::add(d01r, d23i); re[3] = A::sub(d01r, d23i);
}

template <class T> RR_HD void leaf_fft2(T *re, T *im)             // the two size-2 children inside fft8
{
  typedef Arith<T> A;
  const T ar = re[0], ai = im[0], br = re[1], bi = im[1];
  re[0] = A::add(ar, br); im[0] = A::add(ai, bi);
  re[1] = A::sub(ar, br); im[1] = A::sub(ai, bi);
}

template <class T> RR_HD void leaf_fft8(T *re, T *im, T sqrthalf) // fft8, fft.c:288-301
{
  leaf_fft4(re, im);
  leaf_fft2(re + 4, im + 4);
  leaf_fft2(re + 6, im + 6);
  sr_bfly(re[0], im[0], re[2], im[2], re[4], im[4], re[6], im[6], sqrthalf, sqrthalf, true);
  sr_bfly(re[1], im[1], re[3], im[3], re[5], im[5], re[7], im[7], sqrthalf, sqrthalf, false);
}

template <class T> RR_HD void leaf_fft16(T *re, T *im, T sqrthalf, T c1, T c3)   // fft16, fft.c:304-318
{
  leaf_fft8(re, im, sqrthalf);
  leaf_fft4(re + 8, im + 8);
  leaf_fft4(re + 12, im + 12);
  sr_bfly(re[0], im[0], re[4], im[4], re[8], im[8], re[12], im[12], sqrthalf, sqrthalf, true);
  sr_bfly(re[2], im[2], re[6], im[6], re[10], im[10], re[14], im[14], sqrthalf, sqrthalf, false);
  sr_bfly(re[1], im[1], re[5], im[5], re[9], im[9], re[13], im[13], c1, c3, false);
  sr_bfly(re[3], im[3], re[7], im[7], re[11], im[11], re[15], im[15], c3, c1, false);
}

// Leaf phase for `lanes` independent transforms: gathers from linear natural-order tiles
// (src + lane*src_stride, complex j at floats 2j,2j+1) and writes the padded work buffers
// (dst + lane*dst_stride, complex p at floats 2*cslot(p)).
template <class T>
RR_PROG void cfft_leaf_task(const CfftSched &s, bool inverse_unused, int task, const T *src, T *dst, T sqrthalf,
                          T c16_1, T c16_3)
{
  (void)inverse_unused;
  if (task < s.n16) {
    T re[16], im[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int g = s.gather16[e * s.n16 + task];
      re[e] = src[2 * g]; im[e] = src[2 * g + 1];
    }
    leaf_fft16(re, im, sqrthalf, c16_1, c16_3);
    const int base = cslot(s.leaf16_off[task]);                  // offsets are multiples of 16: slots stay contiguous
#pragma unroll
    for (int e = 0; e < 16; ++e) { dst[2 * (base + e)] = re[e]; dst[2 * (base + e) + 1] = im[e]; }
  } else {
    const int t8 = task - s.n16;
    T re[8], im[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = s.gather8[e * s.n8 + t8];
      re[e] = src[2 * g]; im[e] = src[2 * g + 1];
    }
    leaf_fft8(re, im, sqrthalf);
    const int base = cslot(s.leaf8_off[t8]);                     // multiples of 8: never straddle a pad slot
#pragma unroll
    for (int e = 0; e < 8; ++e) { dst[2 * (base + e)] = re[e]; dst[2 * (base + e) + 1] = im[e]; }
  }
}

// One butterfly of the combining pass of size S = 1 << lg (pass(), fft.c:237-256).
template <class T>
RR_PROG void cfft_pass_item(const CfftSched &s, int lg, int item, T *buf, const T *pyramid)
{
  const int qbits = lg - 2, q = 1 << qbits;
  const int node = item >> qbits, k = item & (q - 1);
  const int p0 = s.node_off[s.level_begin[lg] + node] + k;
  const T *tw = pyramid + s.pyr_off[lg];
  T *a0 = buf + 2 * cslot(p0), *a1 = buf + 2 * cslot(p0 + q), *a2 = buf + 2 * cslot(p0 + 2 * q),
    *a3 = buf + 2 * cslot(p0 + 3 * q);
  T a0r = a0[0], a0i = a0[1], a1r = a1[0], a1i = a1[1], a2r = a2[0], a2i = a2[1], a3r = a3[0], a3i = a3[1];
  sr_bfly(a0r, a0i, a1r, a1i, a2r, a2i, a3r, a3i, tw[k], tw[q - k], k == 0);
  a0[0] = a0r; a0[1] = a0i; a1[0] = a1r; a1[1] = a1i; a2[0] = a2r; a2[1] = a2i; a3[0] = a3r; a3[1] = a3i;
}

// Whole complex FFT for `lanes` transforms living in one CTA.
template <class T>
RR_PROG void cfft_run(const CfftSched &s, int lanes, const T *src, int src_stride, T *dst, int dst_stride,
                    const T *pyramid, T sqrthalf, T c16_1, T c16_3)
{
  const int nleaf = s.n16 + s.n8;
  cta_for(lanes * nleaf, [&](int w) {
    const int lane = w / nleaf, task = w - lane * nleaf;
    cfft_leaf_task<T>(s, false, task, src + (long long)lane * src_stride, dst + (long long)lane * dst_stride, sqrthalf,
                      c16_1, c16_3);
  });
  for (int lg = 5; lg <= s.bits; ++lg) {
    const int per_lane = s.level_cnt[lg] << (lg - 2);
    cta_for(lanes * per_lane, [&](int w) {
      const int lane = w / per_lane, item = w - lane * per_lane;
      cfft_pass_item<T>(s, lg, item, dst + (long long)lane * dst_stride, pyramid);
    });
  }
}

// ---------------------------------------------------------------------------------------------------
// Overlap-save DFT FIR stage (dft_filter.h:60-190)
// ---------------------------------------------------------------------------------------------------
enum DftInMode { DFT_IN_FREQ_UP = 0, DFT_IN_COPY = 1, DFT_IN_ZERO_STUFF = 2 };

template <class T> struct DftParams {
  // geometry
  int N, overlap, L, step;       // step: 1, M > 1 (time-domain decimation) or -m (F-domain decimation by 2^m)
  int in_mode;                   // DftInMode
  int Pf, Ni;                    // forward / inverse real transform sizes
  int remL0;
  int q;                         // inputs consumed per block (modes FREQ_UP / COPY)
  int kept;                      // outputs per block when step <= 1
  // schedules and tables (device memory)
  CfftSched fwd, inv;
  const T *pyr_f, *pyr_i;        // twiddle pyramids
  const T *tcos_f, *tcos_i;      // cos(2 pi i / Pf), i <= Pf/4 ; cos(2 pi i / Ni), i <= Ni/4
  const T *coef;                 // N packed spectrum values of the filter
  T sqrthalf, c16_1, c16_3;
  // data
  LaneView in, out;
  long long out_preload;         // coordinate of this stage's output 0 in the next FIFO
  long long block0;              // first block of this launch
  int nblocks, nlanes, lanes_per_cta;
  int xstride, ystride;          // per-lane smem strides (in T)
};

template <class T> RR_HD int dft_smem_elems(int N, int lanes_per_cta, int *xstride, int *ystride)
{
  const int xs = 2 * cfft_buf_complex(N / 2), ys = N;
  if (xstride) *xstride = xs;
  if (ystride) *ystride = ys;
  return lanes_per_cta * (xs + ys);
}

// Value of the reference's `output[idx]` after the frequency-domain up-sampling block
// (dft_filter.h:86-104), read from the packed Pf-point spectrum held in the padded buffer X.
template <class T> RR_HD T dft_spec_freq_up(const T *X, int Pf, int idx)
{
  const int twoP = Pf << 1, r = idx & (twoP - 1);
  if (idx >= twoP && r == 1) return (T)0;
  if (r < Pf) { const int f = (r == 1) ? 0 : r; return X[2 * cslot(f >> 1) + (f & 1)]; }
  if (r == Pf) return X[2 * cslot(0) + 1];
  if (r == Pf + 1) return (T)0;
  if (!(r & 1)) { const int f = twoP - r; return X[2 * cslot(f >> 1) + (f & 1)]; }
  { const int f = twoP - r + 2; return -X[2 * cslot(f >> 1) + (f & 1)]; }
}

template <class T, class InT, class OutT>
RR_PROG void dft_stage_program(const DftParams<T> &p, long long work, T *smem)
{
  typedef Arith<T> A;
  const int LPC = p.lanes_per_cta;
  const int groups = (p.nlanes + LPC - 1) / LPC;
  const long long b = p.block0 + work / groups;
  const int lane0 = (int)(work % groups) * LPC;
  const int lanes = (p.nlanes - lane0) < LPC ? (p.nlanes - lane0) : LPC;
  T *X = smem, *Y = smem + (long long)LPC * p.xstride;
  const int N = p.N, L = p.L, V = N - p.overlap;

  // ---- block geometry in absolute coordinates ----
  long long Rb; int remLb = p.remL0;
  if (p.in_mode == DFT_IN_ZERO_STUFF) {
    const long long Pb = b * (long long)V;
    Rb = Pb <= p.remL0 ? 0 : (Pb - p.remL0 + L - 1) / L;
    remLb = (int)(p.remL0 + (long long)L * Rb - Pb);
  } else Rb = b * (long long)p.q;
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : N;
  long long in_off[2], out_off[2];                       // lanes_per_cta <= 2
  for (int l = 0; l < 2; ++l) {
    const int lane = lane0 + (l < lanes ? l : 0);
    in_off[l] = lane_offset(p.in, lane);
    out_off[l] = lane_offset(p.out, lane);
  }

  // ---- phase 0: stage the input tile in natural order in Y ----
  cta_for(lanes * span, [&](int w) {
    const int l = w % lanes, j = w / lanes;
    const long long loff = in_off[l];
    T v;
    if (p.in_mode == DFT_IN_ZERO_STUFF) {
      const int d = j - remLb;
      v = (d >= 0 && d % L == 0) ? view_read<InT, T>(p.in, loff, Rb + d / L) : (T)0;
    } else v = view_read<InT, T>(p.in, loff, Rb + j);
    Y[(long long)l * p.ystride + j] = v;
  });

  // ---- phases 1-2: forward complex FFT of Pf/2 points, Y -> X ----
  cfft_run<T>(p.fwd, lanes, Y, p.ystride, X, p.xstride, p.pyr_f, p.sqrthalf, p.c16_1, p.c16_3);

  // ---- phase 3: real-FFT post-processing in place (ff_rdft_calc_c forward, rdft.c:46-77) ----
  {
    const int Pf = p.Pf, per = (Pf >> 2) + 1;
    cta_for(lanes * per, [&](int w) {
      const int l = w / per, i = w - l * per;
      T *d = X + (long long)l * p.xstride;
      if (i == 0) {
        T *z = d + 2 * cslot(0);
        const T e = z[0];
        z[0] = A::add(e, z[1]); z[1] = A::sub(e, z[1]);
      } else if (i == (Pf >> 2)) {
        T *z = d + 2 * cslot(Pf >> 2);
        z[1] = -z[1];
      } else {
        T *za = d + 2 * cslot(i), *zb = d + 2 * cslot((Pf >> 1) - i);
        const T c = p.tcos_f[i], s = p.tcos_f[(Pf >> 2) - i];
        const T evr = A::mul((T)0.5, A::add(za[0], zb[0]));
        const T odi = A::mul((T)0.5, A::sub(zb[0], za[0]));
        const T evi = A::mul((T)0.5, A::sub(za[1], zb[1]));
        const T odr = A::mul((T)0.5, A::add(za[1], zb[1]));
        const T sr = A::add(A::mul(odr, c), A::mul(odi, s));
        const T si = A::sub(A::mul(odi, c), A::mul(odr, s));
        za[0] = A::add(evr, sr); za[1] = A::add(evi, si);
        zb[0] = A::sub(evr, sr); zb[1] = A::sub(si, evi);
      }
    });
  }

  if (p.step == 0) {   // spectrum-only mode (filter bank preparation, rate_base.h:184): emit the packed spectrum
    cta_for(lanes * p.Pf, [&](int w) {
      const int l = w % lanes, t = w / lanes;
      view_write<OutT, T>(p.out, out_off[l], (long long)t, X[(long long)l * p.xstride + 2 * cslot(t >> 1) + (t & 1)]);
    });
    return;
  }

  // ---- phase 4: spectrum assembly, filter multiply, inverse pre-processing, X -> Y (natural order) ----
  {
    const int Ni = p.Ni, per = (Ni >> 2) + 1, Pf = p.Pf;
    const bool freq_up = p.in_mode == DFT_IN_FREQ_UP;
    cta_for(lanes * per, [&](int w) {
      const int l = w / per, i = w - l * per;
      const T *Xs = X + (long long)l * p.xstride;
      T *d = Y + (long long)l * p.ystride;
      auto spec = [&](int idx) -> T {
        return freq_up ? dft_spec_freq_up<T>(Xs, Pf, idx) : Xs[2 * cslot(idx >> 1) + (idx & 1)];
      };
      auto cmul = [&](int idx, T &re, T &im) {           // dft_filter.h:140-145
        const T t = spec(idx), o1 = spec(idx + 1), c0 = p.coef[idx], c1 = p.coef[idx + 1];
        re = A::sub(A::mul(c0, t), A::mul(c1, o1));
        im = A::add(A::mul(c1, t), A::mul(c0, o1));
      };
      if (i == 0) {
        const T d0 = A::mul(spec(0), p.coef[0]);
        T d1;
        if (p.step > 0) d1 = A::mul(spec(1), p.coef[1]);
        else d1 = A::sub(A::mul(p.coef[Ni], spec(Ni)), A::mul(p.coef[Ni + 1], spec(Ni + 1)));   // dft_filter.h:185
        const T e0 = A::add(d0, d1), e1 = A::sub(d0, d1);       // rdft.c:44-46
        d[0] = A::mul(e0, (T)0.5); d[1] = A::mul(e1, (T)0.5);   // rdft.c:79-80
      } else if (i == (Ni >> 2)) {
        T re, im;
        cmul(Ni >> 1, re, im);
        d[Ni >> 1] = re; d[(Ni >> 1) + 1] = -im;                // rdft.c:77
      } else {
        const int i1 = 2 * i, i2 = Ni - i1;
        T ar, ai, br, bi;
        cmul(i1, ar, ai);
        cmul(i2, br, bi);
        const T c = p.tcos_i[i], s = p.tcos_i[(Ni >> 2) - i];
        const T evr = A::mul((T)0.5, A::add(ar, br));           // RDFT_UNMANGLE(-,+), k2 = -0.5
        const T odi = A::mul((T)-0.5, A::sub(br, ar));
        const T evi = A::mul((T)0.5, A::sub(ai, bi));
        const T odr = A::mul((T)-0.5, A::add(ai, bi));
        const T sr = A::sub(A::mul(odr, c), A::mul(odi, s));
        const T si = A::add(A::mul(odi, c), A::mul(odr, s));
        d[i1] = A::add(evr, sr); d[i1 + 1] = A::add(evi, si);
        d[i2] = A::sub(evr, sr); d[i2 + 1] = A::sub(si, evi);
      }
    });
  }

  // ---- phases 5-6: inverse complex FFT of Ni/2 points, Y -> X ----
  cfft_run<T>(p.inv, lanes, Y, p.ystride, X, p.xstride, p.pyr_i, p.sqrthalf, p.c16_1, p.c16_3);

  // ---- phase 7: emit the valid samples ----
  {
    int first = 0, stride = 1, count; long long k0;
    if (p.step == 1) { count = V; k0 = b * (long long)V; }
    else if (p.step > 1) {
      const long long v0 = b * (long long)V;
      const int M = p.step;
      first = (int)((M - v0 % M) % M); stride = M;
      k0 = (v0 + M - 1) / M;
      count = first < V ? (V - first + M - 1) / M : 0;
    } else { count = p.kept; k0 = b * (long long)p.kept; }
    cta_for(lanes * count, [&](int w) {
      const int l = w % lanes, j = w / lanes;
      const int t = first + j * stride;
      const T v = X[(long long)l * p.xstride + 2 * cslot(t >> 1) + (t & 1)];
      view_write<OutT, T>(p.out, out_off[l], p.out_preload + k0 + j, v);
    });
  }
}

// ---------------------------------------------------------------------------------------------------
// Polyphase FIR stages
// ---------------------------------------------------------------------------------------------------
template <class T> struct PolyParams {
  int n, L, order, phase_bits;
  long long at0, step;           // vpoly0: integer units of 1/L input samples; vpolyN: 32.32 fixed point (.all)
  int pre;                       // stage_t.pre (0 for these stages)
  const T *coefs;                // vpoly0: [L][n]; vpolyN: [2^phase_bits][n][order+1]
  LaneView in, out;
  long long out_preload;
  long long out0;                // first output index of this launch
  long long nout;                // outputs per lane in this launch
  int nlanes, tile;              // outputs per CTA tile
  int win_cap;                   // smem window capacity (samples)
};

// 128-bit product helper: low and high 64 bits of a*b (a, b >= 0).
RR_HD void mul_u64_wide(unsigned long long a, unsigned long long b, unsigned long long &lo, unsigned long long &hi)
{
#if defined(__CUDA_ARCH__)
  lo = a * b; hi = __umul64hi(a, b);
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  lo = (unsigned long long)p; hi = (unsigned long long)(p >> 64);
#endif
}

// vpoly0: y_i = sum_j c[r][j] * x[q + j], (q, r) = divmod(at0 + i*step, L), summed in tap order.
template <class T, class InT, class OutT>
RR_PROG void poly0_program(const PolyParams<T> &p, long long work, T *smem)
{
  typedef Arith<T> A;
  const long long tiles = (p.nout + p.tile - 1) / p.tile;
  const int lane = (int)(work / tiles);
  const long long tix = work - lane * tiles;
  const long long i0 = p.out0 + tix * p.tile;
  const long long rest = p.out0 + p.nout - i0;
  const int cnt = rest < p.tile ? (int)rest : p.tile;
  const long long at_first = p.at0 + i0 * p.step;
  const long long q0 = at_first / p.L;
  const int r0 = (int)(at_first % p.L);
  const long long at_span = (long long)r0 + (long long)(cnt - 1) * p.step;
  const int win = (int)(at_span / p.L) + p.n;
  const long long in_off = lane_offset(p.in, lane), out_off = lane_offset(p.out, lane);

  cta_for(win, [&](int j) { smem[j] = view_read<InT, T>(p.in, in_off, q0 + p.pre + j); });
  const bool narrow = at_span < 0x7fffffffll;              // tile-relative positions fit 32 bits
  cta_for(cnt, [&](int j) {
    int q, r;
    if (narrow) {
      const unsigned at = (unsigned)r0 + (unsigned)j * (unsigned)p.step;
      q = (int)(at / (unsigned)p.L); r = (int)(at - (unsigned)q * (unsigned)p.L);
    } else {
      const long long at = (long long)r0 + (long long)j * p.step;
      q = (int)(at / p.L); r = (int)(at - (long long)q * p.L);
    }
    const T *c = p.coefs + (long long)r * p.n, *x = smem + q;
    T sum = (T)0;
    for (int k = 0; k < p.n; ++k) sum = A::add(sum, A::mul(c[k], x[k]));
    view_write<OutT, T>(p.out, out_off, p.out_preload + i0 + j, sum);
  });
}

// vpoly1..3: 32.32 fixed-point position, Horner-interpolated coefficients.
template <class T, class InT, class OutT>
RR_PROG void polyN_program(const PolyParams<T> &p, long long work, T *smem)
{
  typedef Arith<T> A;
  const long long tiles = (p.nout + p.tile - 1) / p.tile;
  const int lane = (int)(work / tiles);
  const long long tix = work - lane * tiles;
  const long long i0 = p.out0 + tix * p.tile;
  const long long rest = p.out0 + p.nout - i0;
  const int cnt = rest < p.tile ? (int)rest : p.tile;
  // at_first = at0 + i0*step as a 128-bit value: integer part may exceed 32 bits on long streams
  unsigned long long lo, hi;
  mul_u64_wide((unsigned long long)i0, (unsigned long long)p.step, lo, hi);
  const unsigned long long lo2 = lo + (unsigned long long)p.at0;
  hi += lo2 < lo ? 1 : 0;
  const long long q0 = (long long)((hi << 32) | (lo2 >> 32));
  const unsigned long long f0 = lo2 & 0xffffffffull;
  const unsigned long long span = f0 + (unsigned long long)(cnt - 1) * (unsigned long long)p.step;
  const int win = (int)(span >> 32) + p.n;
  const long long in_off = lane_offset(p.in, lane), out_off = lane_offset(p.out, lane);
  const int order = p.order;

  cta_for(win, [&](int j) { smem[j] = view_read<InT, T>(p.in, in_off, q0 + p.pre + j); });
  cta_for(cnt, [&](int j) {
    const unsigned long long at = f0 + (unsigned long long)j * (unsigned long long)p.step;
    const int q = (int)(at >> 32);
    const uint32_t fraction = (uint32_t)at;
    const int phase = (int)(fraction >> (32 - p.phase_bits));
    const T t = A::mul((T)(uint32_t)(fraction << p.phase_bits), (T)(1.0 / 4294967296.0));
    const T *c = p.coefs + (long long)phase * p.n * (order + 1), *x = smem + q;
    T sum = (T)0;
    if (order == 1)
      for (int k = 0; k < p.n; ++k)
        sum = A::add(sum, A::mul(A::add(A::mul(c[2 * k], t), c[2 * k + 1]), x[k]));
    else if (order == 2)
      for (int k = 0; k < p.n; ++k)
        sum = A::add(sum, A::mul(A::add(A::mul(A::add(A::mul(c[3 * k], t), c[3 * k + 1]), t), c[3 * k + 2]), x[k]));
    else
      for (int k = 0; k < p.n; ++k)
        sum = A::add(sum, A::mul(A::add(A::mul(A::add(A::mul(A::add(A::mul(c[4 * k], t), c[4 * k + 1]), t),
                                                       c[4 * k + 2]), t), c[4 * k + 3]), x[k]));
    view_write<OutT, T>(p.out, out_off, p.out_preload + i0 + j, sum);
  });
}

// ---------------------------------------------------------------------------------------------------
// Half-band 2:1 decimator (h8..h13)
// ---------------------------------------------------------------------------------------------------
template <class T> struct HalfbandParams {
  int ncoef, pre;                // one-sided coefficient count; stage_t.pre (= 2*ncoef)
  T coef[13];
  LaneView in, out;
  long long out_preload, out0, nout;
  int nlanes, tile;
};

template <class T, class InT, class OutT>
RR_PROG void halfband_program(const HalfbandParams<T> &p, long long work, T *smem)
{
  typedef Arith<T> A;
  const long long tiles = (p.nout + p.tile - 1) / p.tile;
  const int lane = (int)(work / tiles);
  const long long tix = work - lane * tiles;
  const long long k0 = p.out0 + tix * p.tile;
  const long long rest = p.out0 + p.nout - k0;
  const int cnt = rest < p.tile ? (int)rest : p.tile;
  const int reach = 2 * p.ncoef - 1;
  const long long x0 = 2 * k0 + p.pre - reach;           // first coordinate needed
  const int win = 2 * (cnt - 1) + 2 * reach + 1;
  const long long in_off = lane_offset(p.in, lane), out_off = lane_offset(p.out, lane);

  cta_for(win, [&](int j) { smem[j] = view_read<InT, T>(p.in, in_off, x0 + j); });
  cta_for(cnt, [&](int j) {
    const T *x = smem + 2 * j + reach;
    T sum = A::mul(x[0], (T)0.5);
    for (int k = 0; k < p.ncoef; ++k)
      sum = A::add(sum, A::mul(A::add(x[-(2 * k + 1)], x[2 * k + 1]), p.coef[k]));
    view_write<OutT, T>(p.out, out_off, p.out_preload + k0 + j, sum);
  });
}

}  // namespace b200rate
