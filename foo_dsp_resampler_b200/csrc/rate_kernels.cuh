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
#include <string.h>

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
  // add / sub with a product as operand: only the packed type needs a different spelling (rate_kernels_pk.cuh)
  static RR_HD float addp(float a, float b) { return add(a, b); }
  static RR_HD float subp(float a, float b) { return sub(a, b); }
};
template <> struct Arith<double> {
  static RR_HD double mul(double a, double b) { return a * b; }
  static RR_HD double add(double a, double b) { return a + b; }
  static RR_HD double sub(double a, double b) { return a - b; }
  static RR_HD double addp(double a, double b) { return a + b; }
  static RR_HD double subp(double a, double b) { return a - b; }
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
RR_PROG bool cta_leader() { return threadIdx.x == 0; }
RR_PROG int cta_threads() { return blockDim.x; }
#else
template <class F> inline void cta_for(int count, F f)
{
  for (int i = 0; i < count; ++i) f(i);
}
inline void cta_sync() {}
inline bool cta_leader() { return true; }
inline int cta_threads() { return 256; }
#endif

// 64-bit / 32-bit split that avoids the slow 64-bit divide for the common case of a small dividend.
RR_HD void divmod_ll(long long a, int b, long long &q, int &r)
{
  if (a >= 0 && a < 0x7fffffffll) { const unsigned ua = (unsigned)a, uq = ua / (unsigned)b; q = uq; r = (int)(ua - uq * (unsigned)b); }
  else { q = a / b; r = (int)(a - q * b); }
}

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

// Address of a coordinate for asynchronous copies; *valid = false where view_read would return zero
// (the returned pointer is then the buffer base, never dereferenced beyond a zero-byte request).
template <class E> RR_HD const E *view_addr(const LaneView &v, long long lane_off, long long coord, bool *valid)
{
  const E *base = static_cast<const E *>(v.base);
  if (coord < v.lo || coord >= v.hi) { *valid = false; return base; }
  const unsigned long long phys = (unsigned long long)(coord - v.origin) & v.mask;
  *valid = true;
  return base + lane_off + (long long)phys * v.elem_stride;
}

// Fast path: when every coordinate of [c0, c1) is stored contiguously (no zero region, no ring wrap) the
// per-element bounds / mask logic collapses to one pointer per lane; element j is then ptr[j * elem_stride].
RR_HD bool view_range_direct(const LaneView &v, long long c0, long long c1)
{
  return c0 >= v.lo && c1 <= v.hi && c0 >= v.origin &&
         (v.mask == ~0ull || (unsigned long long)(c1 - 1 - v.origin) <= v.mask);
}
template <class E> RR_HD E *view_ptr(const LaneView &v, long long lane_off, long long coord)
{
  return static_cast<E *>(v.base) + lane_off + (coord - v.origin) * v.elem_stride;
}

// LDGSTS: global -> shared without staging registers; `valid == false` zero-fills the destination.
#if defined(__CUDA_ARCH__)
template <class E> RR_PROG void async_copy_elem(E *smem_dst, const E *gsrc, bool valid)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? (int)sizeof(E) : 0;
  if (sizeof(E) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
RR_PROG void async_copy_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> RR_PROG void async_copy_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#else
template <class E> inline void async_copy_elem(E *dst, const E *src, bool valid) { *dst = valid ? *src : (E)0; }
inline void async_copy_commit() {}
template <int N> inline void async_copy_wait() {}
#endif

// TMA bulk copies (cp.async.bulk, SASS UBLKCP): one thread moves a contiguous, 16-byte aligned range global -> shared
// without touching registers or the LSU issue slots of the other threads; completion is signalled on an mbarrier by
// transaction bytes. Used where a kernel's input window is contiguous in memory (polyphase windows, tables).
#if defined(__CUDA_ARCH__)
RR_PROG void tma_bar_init(unsigned long long *bar, int arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
RR_PROG void tma_bar_expect(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
RR_PROG void tma_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
RR_PROG void tma_bar_wait(unsigned long long *bar, unsigned parity)
{
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
#else
inline void tma_bar_init(unsigned long long *, int) {}
inline void tma_bar_expect(unsigned long long *, unsigned) {}
inline void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
inline void tma_bar_wait(unsigned long long *, unsigned) {}
#endif

// ---------------------------------------------------------------------------------------------------
// Complex FFT of size M = 1 << bits over FFmpeg's split-radix DAG (fft.c:186-346)
// ---------------------------------------------------------------------------------------------------
// Complex values live in shared memory as 2-element vectors so that every access is one 8/16-byte LDS/STS.
template <class T> struct alignas(2 * sizeof(T)) C2 { T x, y; };

// Device-side schedule. Node lists are per power-of-two size; the gather tables are the out-of-place
// permutation of ff_fft_permute_c (fft.c:169-177) folded into the leaf loads and stored transposed
// ([element][leaf]) so consecutive threads read consecutive entries.
struct CfftSched {
  int bits;
  int n16, n8;                   // leaves: nodes of size 16, and size-8 quarter-children of size-32 nodes
  const uint16_t *leaf16_off;    // [n16] offset of the leaf in the permuted array
  const uint16_t *leaf8_off;     // [n8]
  const uint16_t *gather16;      // [16][n16] natural index feeding permuted element off+e (stored at cslot() of it)
  const uint16_t *gather8;       // [8][n8]
  const uint16_t *node_off;      // concatenated node offsets for sizes 32 .. M
  int level_begin[17], level_cnt[17];
  int qchild_begin[17], qchild_cnt[17];   // per size: the nodes that are quarter children (fused passes)
  int pyr_off[17];               // start of the twiddle row of each size inside the pyramid
  int pyr_len;
};

// Bank-conflict padding of the FFT work buffer: one extra complex slot per 16.
RR_HD int cslot(int p) { return p + (p >> 4); }
// Per-lane stride of a padded buffer holding m complex values: congruent to `unit` modulo 2*unit complex
// (unit = 16 words), so two lanes never meet in a bank when accessed lane-interleaved.
RR_HD int lane_stride_complex(int min_complex, int unit) { return ((min_complex + 2 * unit - 1) / (2 * unit)) * 2 * unit + unit; }

// Read-only global loads (LDG.CONSTANT on the device).
#if defined(__CUDA_ARCH__)
template <class E> RR_HD E ldg(const E *p)
{
  if constexpr (sizeof(E) == 16 && alignof(E) == 16) {                  // any 16-byte POD: one LDG.128
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    E r;
    memcpy(&r, &v, 16);
    return r;
  } else if constexpr (sizeof(E) <= 8 && (sizeof(E) & (sizeof(E) - 1)) == 0 && !__is_class(E)) return __ldg(p);
  else return *p;
}
RR_HD C2<float> ldg(const C2<float> *p) { const float2 v = __ldg(reinterpret_cast<const float2 *>(p)); return C2<float>{v.x, v.y}; }
RR_HD C2<double> ldg(const C2<double> *p) { const double2 v = __ldg(reinterpret_cast<const double2 *>(p)); return C2<double>{v.x, v.y}; }
#else
template <class E> RR_HD E ldg(const E *p) { return *p; }
#endif

template <class T>
RR_HD void sr_bfly(T &a0r, T &a0i, T &a1r, T &a1i, T &a2r, T &a2i, T &a3r, T &a3i, T wre, T wim, bool zero)
{
  typedef Arith<T> A;
  T t1, t2, t5, t6;
  if (zero) { t1 = a2r; t2 = a2i; t5 = a3r; t6 = a3i; }           // TRANSFORM_ZERO, fft.c:228-235
  else {                                                           // TRANSFORM, fft.c:222-226
    t1 = A::addp(A::mul(a2r, wre), A::mul(a2i, wim));
    t2 = A::subp(A::mul(a2i, wre), A::mul(a2r, wim));
    t5 = A::subp(A::mul(a3r, wre), A::mul(a3i, wim));
    t6 = A::addp(A::mul(a3r, wim), A::mul(a3i, wre));
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

// Leaf task `task` for LPC transforms at once: gathers from the natural-order tiles
// (src + lane*src_stride complex, value j stored at cslot(j)) and writes the padded work buffers
// (dst + lane*dst_stride complex, permuted position p at cslot(p)).
// The gather indices are loaded once and shared by the lanes.
template <class T, int LPC>
RR_PROG void cfft_leaf_task(const CfftSched &s, int task, int lanes, const C2<T> *src, int src_stride, C2<T> *dst,
                            int dst_stride, T sqrthalf, T c16_1, T c16_3)
{
  if (task < s.n16) {
    int g[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) g[e] = cslot(ldg(s.gather16 + e * s.n16 + task));
    const int base = cslot(ldg(s.leaf16_off + task));             // multiples of 16: slots stay contiguous
#pragma unroll
    for (int l = 0; l < LPC; ++l) {
      if (l < lanes) {
        T re[16], im[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { const C2<T> v = src[l * src_stride + g[e]]; re[e] = v.x; im[e] = v.y; }
        leaf_fft16(re, im, sqrthalf, c16_1, c16_3);
#pragma unroll
        for (int e = 0; e < 16; ++e) dst[l * dst_stride + base + e] = C2<T>{re[e], im[e]};
      }
    }
  } else {
    const int t8 = task - s.n16;
    int g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = cslot(ldg(s.gather8 + e * s.n8 + t8));
    const int base = cslot(ldg(s.leaf8_off + t8));                // multiples of 8: never straddle a pad slot
#pragma unroll
    for (int l = 0; l < LPC; ++l) {
      if (l < lanes) {
        T re[8], im[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { const C2<T> v = src[l * src_stride + g[e]]; re[e] = v.x; im[e] = v.y; }
        leaf_fft8(re, im, sqrthalf);
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[l * dst_stride + base + e] = C2<T>{re[e], im[e]};
      }
    }
  }
}

// One butterfly of the combining pass of size S = 1 << lg (pass(), fft.c:237-256), for every lane.
template <class T, int LPC>
RR_PROG void cfft_pass_item(const CfftSched &s, int lg, int item, int lanes, C2<T> *buf, int stride, const T *pyramid)
{
  const int qbits = lg - 2, q = 1 << qbits;
  const int node = item >> qbits, k = item & (q - 1);
  const int p0 = ldg(s.node_off + s.level_begin[lg] + node) + k;
  const T *tw = pyramid + s.pyr_off[lg];
  const T wre = tw[k], wim = tw[q - k];
  const int i0 = cslot(p0), i1 = cslot(p0 + q), i2 = cslot(p0 + 2 * q), i3 = cslot(p0 + 3 * q);
#pragma unroll
  for (int l = 0; l < LPC; ++l) {
    if (l < lanes) {
      C2<T> *b = buf + l * stride;
      C2<T> a0 = b[i0], a1 = b[i1], a2 = b[i2], a3 = b[i3];
      sr_bfly(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y, wre, wim, k == 0);
      b[i0] = a0; b[i1] = a1; b[i2] = a2; b[i3] = a3;
    }
  }
}

// Whole complex FFT for the CTA's lanes: natural-order tiles `src` -> padded buffers `dst`, in two calls so
// the caller can start refilling `src` as soon as the leaves have consumed it.
// When a phase has fewer work items than half the CTA, the two lanes of an item go to two threads instead
// of one (the index / twiddle loads are then duplicated, but twice as many threads have work).
template <class T, int LPC>
RR_PROG void cfft_leaves(const CfftSched &s, int lanes, const C2<T> *src, int src_stride, C2<T> *dst, int dst_stride,
                         T sqrthalf, T c16_1, T c16_3)
{
  const int count = s.n16 + s.n8;
  if (LPC == 2 && lanes == 2 && 2 * count <= cta_threads()) {
    cta_for(2 * count, [&](int w) {
      const int l = w & 1;
      cfft_leaf_task<T, 1>(s, w >> 1, 1, src + l * src_stride, src_stride, dst + l * dst_stride, dst_stride, sqrthalf,
                           c16_1, c16_3);
    });
    return;
  }
  cta_for(count, [&](int task) {
    cfft_leaf_task<T, LPC>(s, task, lanes, src, src_stride, dst, dst_stride, sqrthalf, c16_1, c16_3);
  });
}

// Two consecutive combining passes (sizes S = 1 << lg and 2S) in one shared-memory round trip. A node of size
// 2S consists of a first half of size S (whose own pass is still due) and two finished quarters of size S/2.
// Task (node, k), k < S/4, keeps 8 values in registers: butterfly k of the S-pass on the first half, then
// butterflies k and k + S/4 of the 2S-pass, which consume exactly those four results plus two values of each
// quarter. Same butterflies on the same operands as the level-by-level order (fft.c:265-272), hence the same
// bits; one barrier and a third of the shared-memory traffic less per pair of levels.
// Where the very last pass may put its results instead of shared memory: complex element c (= samples 2c,
// 2c+1 of the block) of lane l goes to out[l][c] when c < half. Used when the block's valid samples form an
// aligned, contiguous, same-type range, which saves a shared-memory round trip and a barrier per block.
template <class T> struct PassSink { C2<T> *out[2]; int half; };

template <class T, int LPC, bool SINK>
RR_PROG void cfft_fused_item(const CfftSched &s, int lg, int item, int lanes, C2<T> *buf, int stride, const T *pyramid,
                             const PassSink<T> *sink, int lane_base)
{
  const int qbits = lg - 2, q = 1 << qbits;
  const int node = item >> qbits, k = item & (q - 1);
  const int o = ldg(s.node_off + s.level_begin[lg + 1] + node) + k;
  const T *twa = pyramid + s.pyr_off[lg], *twb = pyramid + s.pyr_off[lg + 1];
  const T ar = twa[k], ai = twa[q - k];
  const T b0r = twb[k], b0i = twb[2 * q - k], b1r = twb[k + q], b1i = twb[q - k];
  const int i0 = cslot(o), i1 = cslot(o + q), i2 = cslot(o + 2 * q), i3 = cslot(o + 3 * q);
  const int j0 = cslot(o + 4 * q), j1 = cslot(o + 5 * q), j2 = cslot(o + 6 * q), j3 = cslot(o + 7 * q);
#pragma unroll
  for (int l = 0; l < LPC; ++l) {
    if (l < lanes) {
      C2<T> *b = buf + l * stride;
      C2<T> a0 = b[i0], a1 = b[i1], a2 = b[i2], a3 = b[i3];
      C2<T> c0 = b[j0], c1 = b[j1], d0 = b[j2], d1 = b[j3];
      sr_bfly(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y, ar, ai, k == 0);          // size S, index k
      sr_bfly(a0.x, a0.y, a2.x, a2.y, c0.x, c0.y, d0.x, d0.y, b0r, b0i, k == 0);        // size 2S, index k
      sr_bfly(a1.x, a1.y, a3.x, a3.y, c1.x, c1.y, d1.x, d1.y, b1r, b1i, false);         // size 2S, index k + S/4
      if (SINK) {
        C2<T> *g = sink->out[lane_base + l];
        const int h = sink->half;
        if (o < h) g[o] = a0;
        if (o + q < h) g[o + q] = a1;
        if (o + 2 * q < h) g[o + 2 * q] = a2;
        if (o + 3 * q < h) g[o + 3 * q] = a3;
        if (o + 4 * q < h) g[o + 4 * q] = c0;
        if (o + 5 * q < h) g[o + 5 * q] = c1;
        if (o + 6 * q < h) g[o + 6 * q] = d0;
        if (o + 7 * q < h) g[o + 7 * q] = d1;
      } else {
        b[i0] = a0; b[i1] = a1; b[i2] = a2; b[i3] = a3;
        b[j0] = c0; b[j1] = c1; b[j2] = d0; b[j3] = d1;
      }
    }
  }
}

// One butterfly of the size-S pass on a quarter-child node (not covered by a fused task).
template <class T, int LPC>
RR_PROG void cfft_qchild_item(const CfftSched &s, int lg, int item, int lanes, C2<T> *buf, int stride, const T *pyramid)
{
  const int qbits = lg - 2, q = 1 << qbits;
  const int node = item >> qbits, k = item & (q - 1);
  const int p0 = ldg(s.node_off + s.qchild_begin[lg] + node) + k;
  const T *tw = pyramid + s.pyr_off[lg];
  const T wre = tw[k], wim = tw[q - k];
  const int i0 = cslot(p0), i1 = cslot(p0 + q), i2 = cslot(p0 + 2 * q), i3 = cslot(p0 + 3 * q);
#pragma unroll
  for (int l = 0; l < LPC; ++l) {
    if (l < lanes) {
      C2<T> *b = buf + l * stride;
      C2<T> a0 = b[i0], a1 = b[i1], a2 = b[i2], a3 = b[i3];
      sr_bfly(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y, wre, wim, k == 0);
      b[i0] = a0; b[i1] = a1; b[i2] = a2; b[i3] = a3;
    }
  }
}

// All combining passes. `sink` != nullptr: the final (fused, whole-array) pass stores straight to global memory.
template <class T, int LPC>
RR_PROG void cfft_passes(const CfftSched &s, int lanes, C2<T> *dst, int dst_stride, const T *pyramid,
                         const PassSink<T> *sink = nullptr)
{
  const bool two = LPC == 2 && lanes == 2;
  const int nt = cta_threads();
  int lg = 5;
  if ((s.bits - 4) & 1) {                                  // odd number of levels: the smallest one goes alone
    const int count = s.level_cnt[lg] << (lg - 2);
    if (two && 2 * count <= nt)
      cta_for(2 * count, [&](int w) {
        cfft_pass_item<T, 1>(s, lg, w >> 1, 1, dst + (w & 1) * dst_stride, dst_stride, pyramid);
      });
    else
      cta_for(count, [&](int item) { cfft_pass_item<T, LPC>(s, lg, item, lanes, dst, dst_stride, pyramid); });
    ++lg;
  }
  for (; lg < s.bits; lg += 2) {
    const int nfused = s.level_cnt[lg + 1] << (lg - 2);    // one task per (2S-node, k)
    const int nplain = s.qchild_cnt[lg] << (lg - 2);       // single butterflies, bundled in threes (same cost)
    const int nbundles = (nplain + 2) / 3;
    const int count = nfused + nbundles;
    if (sink && lg + 1 == s.bits) {                        // top pair: every element is produced by a fused task
      if (two && 2 * count <= nt)
        cta_for(2 * count, [&](int w) {
          cfft_fused_item<T, 1, true>(s, lg, w >> 1, 1, dst + (w & 1) * dst_stride, dst_stride, pyramid, sink, w & 1);
        });
      else
        cta_for(count, [&](int t) { cfft_fused_item<T, LPC, true>(s, lg, t, lanes, dst, dst_stride, pyramid, sink, 0); });
      return;
    }
    if (two && 2 * count <= nt)
      cta_for(2 * count, [&](int w) {
        const int t = w >> 1;
        C2<T> *b = dst + (w & 1) * dst_stride;
        if (t < nfused) cfft_fused_item<T, 1, false>(s, lg, t, 1, b, dst_stride, pyramid, nullptr, 0);
        else {
          const int first = (t - nfused) * 3;
          for (int j = first; j < first + 3 && j < nplain; ++j) cfft_qchild_item<T, 1>(s, lg, j, 1, b, dst_stride, pyramid);
        }
      });
    else
      cta_for(count, [&](int t) {
        if (t < nfused) cfft_fused_item<T, LPC, false>(s, lg, t, lanes, dst, dst_stride, pyramid, nullptr, 0);
        else {
          const int first = (t - nfused) * 3;
          for (int j = first; j < first + 3 && j < nplain; ++j)
            cfft_qchild_item<T, LPC>(s, lg, j, lanes, dst, dst_stride, pyramid);
        }
      });
  }
}

// ---------------------------------------------------------------------------------------------------
// Overlap-save DFT FIR stage (dft_filter.h:60-190)
// ---------------------------------------------------------------------------------------------------
enum DftInMode { DFT_IN_FREQ_UP = 0, DFT_IN_COPY = 1, DFT_IN_ZERO_STUFF = 2 };

template <class T> struct DftParams {
  // geometry
  int N, overlap, L, step;       // step: 1, M > 1 (time-domain decimation), -m (F-domain decimation by 2^m),
                                 // 0 = emit the forward spectrum only (filter bank preparation)
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
  int nblocks, nlanes;
  int xstride, ystride, zstride; // per-lane smem strides, in complex elements (zstride == 0: no prefetch buffer)
  int lean_tables;               // only the twiddle pyramids in shared memory (once if both transforms have one size), cosines from global
};

// Tables the kernel stages in shared memory once per CTA (the emulation passes the global pointers).
template <class T> struct DftTables { const T *pyr_f, *pyr_i, *tcos_f, *tcos_i; };

template <class T> RR_HD int dft_table_elems(const DftParams<T> &p)
{
  if (p.lean_tables) return p.fwd.pyr_len + (p.fwd.bits == p.inv.bits ? 0 : p.inv.pyr_len) + 8;
  return p.fwd.pyr_len + p.inv.pyr_len + (p.Pf >> 2) + 1 + (p.Ni >> 2) + 1 + 8;
}

// Per-lane shared-memory strides (complex elements). The natural-order tile stride is an odd multiple of
// 8 complex (= 16 words) so that the lane-interleaved phases 0 and 7 stay bank-conflict free.
// Buffers per lane: X (FFT work buffer), Y (natural-order input of the inverse transform) and, when the
// next block's input is prefetched while this one is processed, Z (natural-order input of the forward
// transform, Pf reals). Without prefetch the forward input shares Y. All use the padded cslot layout.
template <class T> RR_HD size_t dft_smem_bytes(int Pf, int Ni, int lanes_per_cta, bool prefetch, int *xstride, int *ystride,
                                               int *zstride)
{
  const int unit = 32 / (int)sizeof(T);                  // 16 words, in complex elements
  const int mf = Pf / 2, mi = Ni / 2, mx = mf > mi ? mf : mi;
  const int xs = lane_stride_complex(mx + (mx >> 4) + 1, unit);
  const int ys = prefetch ? lane_stride_complex(mi + (mi >> 4) + 1, unit) : xs;   // without Z, Y also takes the forward input
  const int zs = prefetch ? lane_stride_complex(mf + (mf >> 4) + 1, unit) : 0;
  if (xstride) *xstride = xs;
  if (ystride) *ystride = ys;
  if (zstride) *zstride = zs;
  return sizeof(C2<T>) * (size_t)lanes_per_cta * (size_t)(xs + ys + zs);
}

// (output[idx], output[idx+1]) for even idx of the reference's buffer after the frequency-domain
// up-sampling block (dft_filter.h:86-104), from the packed Pf-point spectrum in the padded buffer X.
template <class T> RR_HD C2<T> dft_spec_freq_up(const C2<T> *X, int Pf, int idx)
{
  const int twoP = Pf << 1, r = idx & (twoP - 1);
  if (r == 0) { const T a0 = X[0].x; return C2<T>{a0, idx == 0 ? a0 : (T)0}; }
  if (r < Pf) return X[cslot(r >> 1)];
  if (r == Pf) return C2<T>{X[0].y, (T)0};
  const C2<T> v = X[cslot((twoP - r) >> 1)];
  return C2<T>{v.x, -v.y};
}

// Per-thread cache of the filter spectrum values its phase-4 items need (they do not depend on the block,
// so a persistent CTA loads them once): ca[k] / cb[k] belong to item threadIdx.x + k*blockDim.x.
template <class T, int MAXI> struct CoefCache { C2<T> ca[MAXI > 0 ? MAXI : 1], cb[MAXI > 0 ? MAXI : 1]; };

template <class T, int MAXI>
RR_PROG void dft_load_coef_cache(const DftParams<T> &p, CoefCache<T, MAXI> &cc)
{
#if defined(__CUDACC__)
  if (MAXI > 0 && p.step != 0) {
    const C2<T> *coef = reinterpret_cast<const C2<T> *>(p.coef);
#pragma unroll
    for (int k = 0; k < (MAXI > 0 ? MAXI : 1); ++k) {
      const int i = threadIdx.x + k * blockDim.x;
      if (i >= 1 && i < (p.Ni >> 2)) { cc.ca[k] = ldg(coef + i); cc.cb[k] = ldg(coef + (p.Ni >> 1) - i); }
    }
  }
#else
  (void)p; (void)cc;
#endif
}

// cta_for variant that also hands the per-thread iteration number to the body (device) or -1 (emulation).
#if defined(__CUDACC__)
template <int MAXI, class F> RR_PROG void cta_for_cached(int count, F f)
{
  if (MAXI > 0) {
#pragma unroll
    for (int k = 0; k < (MAXI > 0 ? MAXI : 1); ++k) {
      const int i = threadIdx.x + k * blockDim.x;
      if (i < count) f(i, k);
    }
    for (int i = threadIdx.x + MAXI * blockDim.x; i < count; i += blockDim.x) f(i, -1);
  } else
    for (int i = threadIdx.x; i < count; i += blockDim.x) f(i, -1);
  __syncthreads();
}
#else
template <int MAXI, class F> inline void cta_for_cached(int count, F f)
{
  for (int i = 0; i < count; ++i) f(i, -1);
}
#endif

// Geometry of one work item (block, lane group) in absolute coordinates.
template <class T> struct DftItem {
  long long b, Rb, in_off0, in_off1, out_off0, out_off1;
  int lanes, remLb;
};

template <class T, int LPC>
RR_PROG DftItem<T> dft_item(const DftParams<T> &p, long long work)
{
  DftItem<T> it;
  const int groups = (p.nlanes + LPC - 1) / LPC;
  long long bq; int gr;
  divmod_ll(work, groups, bq, gr);
  it.b = p.block0 + bq;
  const int lane0 = gr * LPC;
  it.lanes = (p.nlanes - lane0) < LPC ? (p.nlanes - lane0) : LPC;
  it.remLb = p.remL0;
  if (p.in_mode == DFT_IN_ZERO_STUFF) {
    const long long Pb = it.b * (long long)(p.N - p.overlap);
    it.Rb = Pb <= p.remL0 ? 0 : (Pb - p.remL0 + p.L - 1) / p.L;
    it.remLb = (int)(p.remL0 + (long long)p.L * it.Rb - Pb);
  } else it.Rb = it.b * (long long)p.q;
  it.in_off0 = lane_offset(p.in, lane0);
  it.out_off0 = lane_offset(p.out, lane0);
  const int lane1 = lane0 + (it.lanes > 1 ? 1 : 0);
  it.in_off1 = lane_offset(p.in, lane1);
  it.out_off1 = lane_offset(p.out, lane1);
  return it;
}

// Phase 0: bring the input tile of a work item into a natural-order (cslot-padded) buffer, lane fastest so
// interleaved input is read fully coalesced. ASYNC: issue LDGSTS copies and return without waiting.
template <class T, class InT, int LPC, bool ASYNC>
RR_PROG void dft_stage_tile(const DftParams<T> &p, const DftItem<T> &it, C2<T> *tile, int tile_stride)
{
  T *tr = reinterpret_cast<T *>(tile);
  const int span = p.in_mode == DFT_IN_FREQ_UP ? p.Pf : p.N, L = p.L;
  const bool direct = p.in_mode != DFT_IN_ZERO_STUFF && view_range_direct(p.in, it.Rb, it.Rb + span);
  const InT *s0 = view_ptr<const InT>(p.in, it.in_off0, it.Rb), *s1 = view_ptr<const InT>(p.in, it.in_off1, it.Rb);
  const int es = p.in.elem_stride;
  auto body = [&](int w) {
    int l, j;
    if (LPC == 1 || it.lanes == 1) { l = 0; j = w; } else { l = w & 1; j = w >> 1; }
    if (direct) {                                         // interior block: plain pointer arithmetic
      T *dst = tr + 2 * (l * tile_stride + cslot(j >> 1)) + (j & 1);
      const InT *src = (l ? s1 : s0) + j * es;
      if (ASYNC) async_copy_elem<T>(dst, reinterpret_cast<const T *>(src), true);
      else *dst = (T)*src;
      return;
    }
    long long coord = it.Rb + j;
    bool on_grid = true;
    if (p.in_mode == DFT_IN_ZERO_STUFF) {
      const int d = j - it.remLb;
      on_grid = d >= 0 && d % L == 0;
      coord = it.Rb + (on_grid ? d / L : 0);
    }
    T *dst = tr + 2 * (l * tile_stride + cslot(j >> 1)) + (j & 1);
    if (ASYNC) {
      bool valid;
      const T *src = reinterpret_cast<const T *>(view_addr<InT>(p.in, l ? it.in_off1 : it.in_off0, coord, &valid));
      async_copy_elem<T>(dst, src, valid && on_grid);
    } else *dst = on_grid ? view_read<InT, T>(p.in, l ? it.in_off1 : it.in_off0, coord) : (T)0;
  };
#if defined(__CUDACC__)
  if (ASYNC) {
    for (int w = threadIdx.x; w < it.lanes * span; w += blockDim.x) body(w);
    async_copy_commit();
    return;
  }
#endif
  cta_for(it.lanes * span, body);
}

// `items` is a two-entry array in shared memory: items[slot] describes `work` (written by the CTA leader
// before the previous barrier), items[slot ^ 1] receives the description of `work_next` (< 0: none) so
// that the 64-bit coordinate arithmetic is done by one thread per item instead of by all of them. With a
// prefetch buffer (p.zstride > 0) the tile of `work` must already be in flight into Z (the kernel primes it).
template <class T, class InT, class OutT, int LPC, int MAXI>
RR_PROG void dft_stage_program(const DftParams<T> &p, const DftTables<T> &tab, const CoefCache<T, MAXI> &cc,
                               DftItem<T> *items, int slot, long long work_next, C2<T> *smem)
{
  typedef Arith<T> A;
  const DftItem<T> it = items[slot];
  if (work_next >= 0 && cta_leader()) items[slot ^ 1] = dft_item<T, LPC>(p, work_next);
  const long long b = it.b;
  const int lanes = it.lanes;
  C2<T> *X = smem, *Y = smem + LPC * p.xstride, *Z = Y + LPC * p.ystride;
  const int N = p.N, V = N - p.overlap;
  const bool prefetch = p.zstride > 0;
  const long long out_off0 = it.out_off0, out_off1 = it.out_off1;

  // ---- phase 0: the forward transform's input tile, natural order ----
  if (prefetch) { async_copy_wait<0>(); cta_sync(); }
  else dft_stage_tile<T, InT, LPC, false>(p, it, Y, p.ystride);
  const C2<T> *fwd_in = prefetch ? Z : Y;
  const int fwd_stride = prefetch ? p.zstride : p.ystride;

  // ---- phases 1-2: forward complex FFT of Pf/2 points -> X; refill Z for the next item meanwhile ----
  cfft_leaves<T, LPC>(p.fwd, lanes, fwd_in, fwd_stride, X, p.xstride, p.sqrthalf, p.c16_1, p.c16_3);
  if (prefetch && work_next >= 0) {
    const DftItem<T> nx = items[slot ^ 1];              // published before the barrier that ended the leaves
    dft_stage_tile<T, T, LPC, true>(p, nx, Z, p.zstride);
  }
  cfft_passes<T, LPC>(p.fwd, lanes, X, p.xstride, tab.pyr_f);

  // ---- phase 3: real-FFT post-processing in place (ff_rdft_calc_c forward, rdft.c:46-77) ----
  {
    const int Pf = p.Pf;
    cta_for((Pf >> 2) + 1, [&](int i) {
      T c = (T)0, s = (T)0;
      if (i > 0 && i < (Pf >> 2)) { c = tab.tcos_f[i]; s = tab.tcos_f[(Pf >> 2) - i]; }
#pragma unroll
      for (int l = 0; l < LPC; ++l) {
        if (l < lanes) {
          C2<T> *d = X + l * p.xstride;
          if (i == 0) {
            const C2<T> z = d[0];
            d[0] = C2<T>{A::add(z.x, z.y), A::sub(z.x, z.y)};
          } else if (i == (Pf >> 2)) {
            C2<T> *z = d + cslot(Pf >> 2);
            z->y = -z->y;
          } else {
            const int ia = cslot(i), ib = cslot((Pf >> 1) - i);
            const C2<T> za = d[ia], zb = d[ib];
            const T evr = A::mul((T)0.5, A::add(za.x, zb.x));
            const T odi = A::mul((T)0.5, A::sub(zb.x, za.x));
            const T evi = A::mul((T)0.5, A::sub(za.y, zb.y));
            const T odr = A::mul((T)0.5, A::add(za.y, zb.y));
            const T sr = A::add(A::mul(odr, c), A::mul(odi, s));
            const T si = A::sub(A::mul(odi, c), A::mul(odr, s));
            d[ia] = C2<T>{A::add(evr, sr), A::add(evi, si)};
            d[ib] = C2<T>{A::sub(evr, sr), A::sub(si, evi)};
          }
        }
      }
    });
  }

  if (p.step == 0) {   // spectrum-only mode (filter bank preparation, rate_base.h:184): emit the packed spectrum
    cta_for(p.Pf, [&](int t) {
      const C2<T> v = X[cslot(t >> 1)];
      view_write<OutT, T>(p.out, out_off0, (long long)t, (t & 1) ? v.y : v.x);
    });
    return;
  }

  // ---- phase 4: spectrum assembly, filter multiply, inverse pre-processing, X -> Y (natural order) ----
  {
    const int Ni = p.Ni, Pf = p.Pf;
    const bool freq_up = p.in_mode == DFT_IN_FREQ_UP;
    const C2<T> *coef = reinterpret_cast<const C2<T> *>(p.coef);
    cta_for_cached<MAXI>((Ni >> 2) + 1, [&](int i, int k) {
      const bool pair = i > 0 && i < (Ni >> 2);
      C2<T> ca, cb;
      T c = (T)0, s = (T)0;
      if (pair) {
        if (k >= 0) { ca = cc.ca[k]; cb = cc.cb[k]; } else { ca = ldg(coef + i); cb = ldg(coef + (Ni >> 1) - i); }
        c = tab.tcos_i[i]; s = tab.tcos_i[(Ni >> 2) - i];
      } else if (i == 0) {
        ca = ldg(coef);
        cb = p.step > 0 ? C2<T>{(T)0, (T)0} : ldg(coef + (Ni >> 1));
      } else { ca = ldg(coef + (Ni >> 2)); cb = ca; }
#pragma unroll
      for (int l = 0; l < LPC; ++l) {
        if (l < lanes) {
          const C2<T> *Xs = X + l * p.xstride;
          C2<T> *d = Y + l * p.ystride;
          auto spec = [&](int bin) -> C2<T> {              // (output[2 bin], output[2 bin + 1])
            return freq_up ? dft_spec_freq_up<T>(Xs, Pf, 2 * bin) : Xs[cslot(bin)];
          };
          auto cmul = [&](const C2<T> &cf, const C2<T> &v) -> C2<T> {   // dft_filter.h:140-145
            return C2<T>{A::sub(A::mul(cf.x, v.x), A::mul(cf.y, v.y)), A::add(A::mul(cf.y, v.x), A::mul(cf.x, v.y))};
          };
          if (pair) {
            const C2<T> a = cmul(ca, spec(i)), bb = cmul(cb, spec((Ni >> 1) - i));
            const T evr = A::mul((T)0.5, A::add(a.x, bb.x));           // RDFT_UNMANGLE(-,+), k2 = -0.5
            const T odi = A::mul((T)-0.5, A::sub(bb.x, a.x));
            const T evi = A::mul((T)0.5, A::sub(a.y, bb.y));
            const T odr = A::mul((T)-0.5, A::add(a.y, bb.y));
            const T sr = A::sub(A::mul(odr, c), A::mul(odi, s));
            const T si = A::add(A::mul(odi, c), A::mul(odr, s));
            d[cslot(i)] = C2<T>{A::add(evr, sr), A::add(evi, si)};
            d[cslot((Ni >> 1) - i)] = C2<T>{A::sub(evr, sr), A::sub(si, evi)};
          } else if (i == 0) {
            const C2<T> v0 = spec(0);
            const T d0 = A::mul(v0.x, ca.x);
            T d1;
            if (p.step > 0) d1 = A::mul(v0.y, ca.y);
            else { const C2<T> vn = spec(Ni >> 1); d1 = A::sub(A::mul(cb.x, vn.x), A::mul(cb.y, vn.y)); }  // dft_filter.h:185
            const T e0 = A::add(d0, d1), e1 = A::sub(d0, d1);            // rdft.c:44-46
            d[0] = C2<T>{A::mul(e0, (T)0.5), A::mul(e1, (T)0.5)};        // rdft.c:79-80
          } else {
            const C2<T> m = cmul(ca, spec(Ni >> 2));
            d[cslot(Ni >> 2)] = C2<T>{m.x, -m.y};                        // rdft.c:77
          }
        }
      }
    });
  }

  // ---- output geometry of this block ----
  int first = 0, stride = 1, count; long long k0;
  if (p.step == 1) { count = V; k0 = b * (long long)V; }
  else if (p.step > 1) {
    const long long v0 = b * (long long)V;
    const int M = p.step;
    first = (int)((M - v0 % M) % M); stride = M;
    k0 = (v0 + M - 1) / M;
    count = first < V ? (V - first + M - 1) / M : 0;
  } else { count = p.kept; k0 = b * (long long)p.kept; }
  const long long c0 = p.out_preload + k0;
  const bool direct = view_range_direct(p.out, c0, c0 + count);
  OutT *d0 = view_ptr<OutT>(p.out, out_off0, c0), *d1 = view_ptr<OutT>(p.out, out_off1, c0);
  const int es = p.out.elem_stride;
  // planar same-type output of an interior block: whole complex slots (two samples) can be stored at once
  const bool pairs = direct && stride == 1 && es == 1 && !(count & 1) && sizeof(OutT) == sizeof(T) &&
                     !(((size_t)d0 | (size_t)d1) & (2 * sizeof(T) - 1));
  // ... and then the top pass of the inverse transform writes them itself (needs a fused top pair)
  const bool sink_ok = pairs && p.inv.bits >= 6;

  // ---- phases 5-6: inverse complex FFT of Ni/2 points, Y -> X ----
  cfft_leaves<T, LPC>(p.inv, lanes, Y, p.ystride, X, p.xstride, p.sqrthalf, p.c16_1, p.c16_3);
  if (sink_ok) {
    PassSink<T> sink;
    sink.out[0] = reinterpret_cast<C2<T> *>(d0); sink.out[1] = reinterpret_cast<C2<T> *>(d1); sink.half = count >> 1;
    cfft_passes<T, LPC>(p.inv, lanes, X, p.xstride, tab.pyr_i, &sink);
    return;
  }
  cfft_passes<T, LPC>(p.inv, lanes, X, p.xstride, tab.pyr_i);

  // ---- phase 7: emit the valid samples ----
  {
    const T *Xr = reinterpret_cast<const T *>(X);
    if (pairs) {
      const int half = count >> 1;
      cta_for(lanes * half, [&](int w) {
        int l, c;
        if (LPC == 1 || lanes == 1) { l = 0; c = w; } else { l = w & 1; c = w >> 1; }
        reinterpret_cast<C2<T> *>(l ? d1 : d0)[c] = X[l * p.xstride + cslot(c)];
      });
    } else
    cta_for(lanes * count, [&](int w) {
      int l, j;
      if (LPC == 1 || lanes == 1) { l = 0; j = w; } else { l = w & 1; j = w >> 1; }
      const int t = first + j * stride;
      const T v = Xr[2 * (l * p.xstride + cslot(t >> 1)) + (t & 1)];
      if (direct) (l ? d1 : d0)[j * es] = (OutT)v;
      else view_write<OutT, T>(p.out, l ? out_off1 : out_off0, c0 + j, v);
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
  int row_pitch;                 // vpolyN: elements between the coefficient rows a warp stages in shared memory (0: rows are read from global)
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

// vpoly0, phase-stationary variant for rational ratios with enough phases: thread (slot, channel) owns the
// outputs i = out0 + m*L + slot, m = 0, 1, ... -- they all use the same coefficient row r, which therefore
// lives in registers for the whole tile, and consecutive m advance the input position by exactly `step`
// (exact integer phase arithmetic: at(i + L) = at(i) + L*step, so q += step and r is unchanged).
// The input windows of all CH channels of a stream are staged in shared memory; writes of (frame, channel)
// pairs are contiguous for interleaved output. Taps are summed in order, un-fused (Arith<T>).
template <class T> struct Poly0FastParams {
  PolyParams<T> base;
  int F, ncols, MM, CH;          // slots per CTA column, columns per period, periods per tile, lanes per CTA
  int win;                       // shared-memory window per lane (samples; odd multiple of 16 words)
  long long mtiles;              // tiles along m
  int double_buffer;             // two window sets: the next tile is prefetched while this one is computed
};

// Tile geometry of one work item of the phase-stationary kernel.
struct Poly0Tile {
  long long i_first, q_first;
  int lane0, nslots, mcount, r_first, win;
};

template <class T> RR_HD Poly0Tile poly0_tile(const Poly0FastParams<T> &fp, long long work)
{
  const PolyParams<T> &p = fp.base;
  Poly0Tile t;
  const int L = p.L;
  long long mt, rest; int col;
  if (work < 0x7fffffffll && fp.mtiles < 0x7fffffffll) {
    const unsigned uw = (unsigned)work, um = (unsigned)fp.mtiles, ur = uw / um;
    mt = uw - ur * um; rest = ur;
  } else { mt = work % fp.mtiles; rest = work / fp.mtiles; }
  { long long g; divmod_ll(rest, fp.ncols, g, col); t.lane0 = (int)g * fp.CH; }
  const int slot0 = col * fp.F;
  t.nslots = (L - slot0) < fp.F ? (L - slot0) : fp.F;
  const long long periods = (p.nout + L - 1) / L;
  const long long m0 = mt * fp.MM;
  t.mcount = (periods - m0) < fp.MM ? (int)(periods - m0) : fp.MM;
  // first input position touched by this tile: output out0 + m0*L + slot0
  t.i_first = p.out0 + m0 * L + slot0;
  const long long at_first = p.at0 + t.i_first * p.step;
  divmod_ll(at_first, L, t.q_first, t.r_first);
  const long long at_last = (long long)t.r_first + (long long)(t.nslots - 1) * p.step;   // relative to q_first*L
  t.win = (int)(at_last / L) + (t.mcount - 1) * (int)p.step + p.n;
  return t;
}

// Stage the input windows of the tile's CH lanes. ASYNC: LDGSTS, returns after committing the group.
template <class T, class InT, bool ASYNC>
RR_PROG void poly0_fast_load(const Poly0FastParams<T> &fp, const Poly0Tile &t, T *buf)
{
  const PolyParams<T> &p = fp.base;
  const bool direct = view_range_direct(p.in, t.q_first + p.pre, t.q_first + p.pre + t.win);
  const int es = p.in.elem_stride;
  for (int l = 0; l < fp.CH; ++l) {
    const long long off = lane_offset(p.in, t.lane0 + l);
    T *dst = buf + l * fp.win;
    const InT *sp = view_ptr<const InT>(p.in, off, t.q_first + p.pre);
    auto body = [&](int j) {
      if (direct) {
        if (ASYNC) async_copy_elem<T>(dst + j, reinterpret_cast<const T *>(sp + j * es), true);
        else dst[j] = (T)sp[j * es];
      } else if (ASYNC) {
        bool valid;
        const T *src = reinterpret_cast<const T *>(view_addr<InT>(p.in, off, t.q_first + p.pre + j, &valid));
        async_copy_elem<T>(dst + j, src, valid);
      } else dst[j] = view_read<InT, T>(p.in, off, t.q_first + p.pre + j);
    };
#if defined(__CUDACC__)
    for (int j = threadIdx.x; j < t.win; j += blockDim.x) body(j);
#else
    for (int j = 0; j < t.win; ++j) body(j);
#endif
  }
  if (ASYNC) async_copy_commit();
}

template <class T, class OutT, int NT>
RR_PROG void poly0_fast_compute(const Poly0FastParams<T> &fp, const Poly0Tile &t, const T *buf)
{
  typedef Arith<T> A;
  const PolyParams<T> &p = fp.base;
  const int CH = fp.CH, L = p.L;
  cta_for(t.nslots * CH, [&](int tid) {
    const int fs = tid / CH, ch = tid - fs * CH;
    const unsigned at_rel = (unsigned)t.r_first + (unsigned)fs * (unsigned)p.step;      // < L + 512*step < 2^30
    const int q = (int)(at_rel / (unsigned)L), r = (int)(at_rel - (unsigned)q * (unsigned)L);
    T c[NT];
    const T *row = p.coefs + (long long)r * NT;
#pragma unroll
    for (int k = 0; k < NT; ++k) c[k] = ldg(row + k);
    const T *x = buf + ch * fp.win + q;
    const long long out_off = lane_offset(p.out, t.lane0 + ch);
    long long i = t.i_first + fs;
    const long long i_end = p.out0 + p.nout;
    // interior tile: every output of the tile exists and is stored contiguously -> direct stores
    const long long i_tile_end = t.i_first + (long long)t.mcount * L;
    const bool direct = i_tile_end <= i_end && view_range_direct(p.out, p.out_preload + t.i_first, p.out_preload + i_tile_end);
    OutT *dp = view_ptr<OutT>(p.out, out_off, p.out_preload + i);
    const long long dstep = (long long)L * p.out.elem_stride;
    int m = 0;
    for (; m + 1 < t.mcount; m += 2, x += 2 * p.step, i += 2 * L, dp += 2 * dstep) {   // two independent accumulators
      const T *x1 = x + p.step;
      T s0 = (T)0, s1 = (T)0;
#pragma unroll
      for (int k = 0; k < NT; ++k) { s0 = A::add(s0, A::mul(c[k], x[k])); s1 = A::add(s1, A::mul(c[k], x1[k])); }
      if (direct) { dp[0] = (OutT)s0; dp[dstep] = (OutT)s1; }
      else {
        if (i < i_end) view_write<OutT, T>(p.out, out_off, p.out_preload + i, s0);
        if (i + L < i_end) view_write<OutT, T>(p.out, out_off, p.out_preload + i + L, s1);
      }
    }
    if (m < t.mcount) {
      T s0 = (T)0;
#pragma unroll
      for (int k = 0; k < NT; ++k) s0 = A::add(s0, A::mul(c[k], x[k]));
      if (direct) dp[0] = (OutT)s0;
      else if (i < i_end) view_write<OutT, T>(p.out, out_off, p.out_preload + i, s0);
    }
  });
}

// ---------------------------------------------------------------------------------------------------
// vpoly0, two adjacent slots per thread, for the fp64 engine (what RR_open selects for Best quality).
// The scheme of the lane-pair kernel poly0_pair2 (rate_kernels_pk.cuh) for scalar lanes: a thread owns the slot pair
// (s, s + 1) of every period (output index mod L) of one lane, its two coefficient rows stay in registers for the
// whole launch, the windows of its two outputs overlap almost entirely (they start DLO or DLO + 1 samples apart,
// DLO = floor(step / L)), so one pass over NT + DLO + 1 window samples feeds both: 1.8x fewer shared-memory loads per
// output than the one-slot kernel, which is bound by exactly those loads. The slot pairs are dealt to the threads on
// the host so that the sixteen lanes of a half-warp start in sixteen different 8-byte banks -- a pair of an overfull
// bank is moved one bank down by starting its pass ONE SAMPLE EARLY (that position carries no tap; see
// Poly0PairParams::shift for the lane-pair kernel), so every pass covers NT + DLO + 2 window positions and each thread
// skips the end positions without a tap of its outputs; the lane's window of a tile of periods is staged with one TMA
// bulk copy. DFMA in tap order (the fp64 contract is 1e-12, not bit equality).
// ---------------------------------------------------------------------------------------------------
constexpr int kDualPad = 2;      // elements in front of the first lane's window (a shifted pass may start at -1)
template <class T> struct Poly0DualParams {
  PolyParams<T> base;
  const T *coef;                 // [2 n + 3][TS] transposed per-thread rows by window position: n + 1 of the first slot, n + 2 of the second
  const uint16_t *slot, *qs;     // [TS] first slot of the thread's pair (0xffff: idle), 1 + the first window sample of its pass within the period
  const uint8_t *flags;          // [TS] bit 0: the second slot's window starts DLO samples later (else DLO + 1); bit 1: it exists;
                                 //      bit 2: the pass starts one sample early
  int TS, NL;                    // threads per lane, lanes per CTA
  int MM, win;                   // periods per tile, window capacity per lane (samples, even)
  long long m_begin, mtiles;     // first period of the launch (output index / L), tiles along the periods
  int groups, group_fastest;     // lane groups; work order: consecutive work items = the lane groups of one tile of periods
                                 // (the channels of an interleaved output frame are then written at about the same time
                                 // and merge in L2 instead of reaching DRAM as partial sectors)
};

struct Poly0DualTile {
  long long c0;                  // first window coordinate (input FIFO) of the tile
  int lane0, nl, mcount, len;    // lanes [lane0, lane0 + nl), periods, window samples needed
  int tma, head;                 // window staged by bulk copies; window sample j then sits at row index head + j
  int i_lo, i_hi;                // outputs of the tile, relative to its first period's slot 0
  long long i_base;              // output index of the first period's slot 0
};

template <class T> RR_PROG Poly0DualTile poly0_dual_tile(const Poly0DualParams<T> &dp, long long work)
{
  const PolyParams<T> &p = dp.base;
  Poly0DualTile t;
  long long grp, mt;
  if (dp.group_fastest) { int g; divmod_ll(work, dp.groups, mt, g); grp = g; }
  else { int m; divmod_ll(work, (int)dp.mtiles, grp, m); mt = m; }
  t.lane0 = (int)grp * dp.NL;
  t.nl = p.nlanes - t.lane0 < dp.NL ? p.nlanes - t.lane0 : dp.NL;
  const long long m0 = dp.m_begin + mt * dp.MM, m_end = (p.out0 + p.nout + p.L - 1) / p.L;
  t.mcount = m_end - m0 < dp.MM ? (int)(m_end - m0) : dp.MM;
  t.i_base = m0 * p.L;
  const long long lo = p.out0 > t.i_base ? p.out0 : t.i_base, hi_all = p.out0 + p.nout, hi_t = t.i_base + (long long)t.mcount * p.L;
  t.i_lo = (int)(lo - t.i_base);
  t.i_hi = (int)((hi_all < hi_t ? hi_all : hi_t) - t.i_base);
  // window: from the first sample of (period m0, slot 0) to the last sample of the tile's last slot
  const long long at_first = p.at0 + t.i_base * p.step;
  const long long q_first = at_first / p.L;
  const long long q_last = (p.at0 + (hi_t - 1) * p.step) / p.L;
  t.c0 = q_first + p.pre;
  t.len = (int)(q_last - q_first) + p.n + 2;
  t.tma = 0; t.head = 0;
  if (p.in.elem_stride == 1 && t.len + 2 <= dp.win && view_range_direct(p.in, t.c0 - 1, t.c0 + t.len + 1)) {
    const T *s0 = view_ptr<const T>(p.in, lane_offset(p.in, t.lane0), t.c0);
    t.head = sizeof(T) == 8 ? (int)(((size_t)s0 >> 3) & 1) : 0;
    t.tma = sizeof(T) == 8;
    for (int l = 0; l < t.nl; ++l) {
      const T *q0 = view_ptr<const T>(p.in, lane_offset(p.in, t.lane0 + l), t.c0) - t.head;
      if ((size_t)q0 & 15) t.tma = 0;
    }
    if (!t.tma) t.head = 0;
  }
  return t;
}

template <class T> RR_PROG void poly0_dual_load(const Poly0DualParams<T> &dp, const Poly0DualTile &t, T *buf, unsigned long long *bar,
                                               int tid, int nthreads)
{
  const PolyParams<T> &p = dp.base;
  if (t.tma) {
    if (tid != 0) return;
    const unsigned bytes = (unsigned)(((t.len + t.head + 1) & ~1) * (int)sizeof(T));
    tma_bar_expect(bar, bytes * (unsigned)t.nl);
    for (int l = 0; l < t.nl; ++l)
      tma_load_1d(buf + kDualPad + l * dp.win, view_ptr<const T>(p.in, lane_offset(p.in, t.lane0 + l), t.c0) - t.head, bytes, bar);
    return;
  }
  for (int l = 0; l < t.nl; ++l) {
    const long long off = lane_offset(p.in, t.lane0 + l);
    for (int j = tid; j < t.len; j += nthreads) buf[kDualPad + l * dp.win + j] = view_read<T, T>(p.in, off, t.c0 + j);
  }
}

template <class T, class OutT, int NT, int DLO> struct Poly0DualThread {
  int lane, s0, q;               // lane within the CTA, first slot (< 0: idle), first position of the pass within the period
  int f1;                        // the second output's taps start at position DLO + f1 (0 / 1 / 2)
  bool sh, two;                  // the first output's taps start at position 1 (else 0); the second slot exists
  T c0[NT + 1], c1[NT + 2];      // by window position
};

template <class T, class OutT, int NT, int DLO>
RR_PROG Poly0DualThread<T, OutT, NT, DLO> poly0_dual_setup(const Poly0DualParams<T> &dp, int tid)
{
  Poly0DualThread<T, OutT, NT, DLO> st;
  st.lane = tid / dp.TS;
  const int ts = tid - st.lane * dp.TS;
  const int s = st.lane < dp.NL ? (int)ldg(dp.slot + ts) : 0xffff;
  st.s0 = s == 0xffff ? -1 : s;
  st.q = (int)ldg(dp.qs + ts) - 1;
  const unsigned fl = ldg(dp.flags + ts);
  st.two = (fl & 2) != 0; st.sh = (fl & 4) != 0;
  st.f1 = ((fl & 1) ? 0 : 1) + (st.sh ? 1 : 0);
#pragma unroll
  for (int k = 0; k <= NT; ++k) st.c0[k] = ldg(dp.coef + k * dp.TS + ts);
#pragma unroll
  for (int k = 0; k <= NT + 1; ++k) st.c1[k] = ldg(dp.coef + (NT + 1 + k) * dp.TS + ts);
  return st;
}

template <class T, class OutT, int NT, int DLO>
RR_PROG void poly0_dual_compute(const Poly0DualParams<T> &dp, const Poly0DualTile &t, const T *buf,
                                const Poly0DualThread<T, OutT, NT, DLO> &st)
{
  const PolyParams<T> &p = dp.base;
  if (st.s0 < 0 || st.lane >= t.nl) return;
  const int L = p.L, step = (int)p.step, es = p.out.elem_stride;
  const long long out_off = lane_offset(p.out, t.lane0 + st.lane);
  const bool direct = view_range_direct(p.out, p.out_preload + t.i_base + t.i_lo, p.out_preload + t.i_base + t.i_hi);
  OutT *d = view_ptr<OutT>(p.out, out_off, p.out_preload + t.i_base + st.s0);
  const T *xw = buf + kDualPad + st.lane * dp.win + t.head + st.q;
  const bool sh = st.sh;
  const int f1 = st.f1;
  auto tap = [&](int j, T xv, T &a0, T &a1) {              // position j of the pass: which taps of the two outputs it carries
    if (j <= NT) {
      if (j == 0 ? !sh : (j == NT ? sh : true)) a0 += st.c0[j] * xv;
    }
    if (j >= DLO) {
      const int jj = j - DLO;                                // 0 .. NT + 1
      if (jj == 0 ? f1 == 0 : (jj == 1 ? f1 <= 1 : (jj == NT ? f1 >= 1 : (jj == NT + 1 ? f1 == 2 : true)))) a1 += st.c1[jj] * xv;
    }
  };
  auto emit = [&](int m, T a0, T a1) {
    const int i = m * L + st.s0;
    if (i >= t.i_lo && i < t.i_hi) {
      if (direct) d[(long long)m * L * es] = (OutT)a0;
      else view_write<OutT, T>(p.out, out_off, p.out_preload + t.i_base + i, a0);
    }
    if (st.two && i + 1 >= t.i_lo && i + 1 < t.i_hi) {
      if (direct) d[((long long)m * L + 1) * es] = (OutT)a1;
      else view_write<OutT, T>(p.out, out_off, p.out_preload + t.i_base + i + 1, a1);
    }
  };
  constexpr int NW = NT + DLO + 2;
  int m = 0;
  for (; m + 1 < t.mcount; m += 2, xw += 2 * step) {        // two periods at a time: four independent DFMA chains
    const T *xb = xw + step;
    T a0 = (T)0, a1 = (T)0, b0 = (T)0, b1 = (T)0;
#pragma unroll
    for (int j = 0; j < NW; ++j) { tap(j, xw[j], a0, a1); tap(j, xb[j], b0, b1); }
    emit(m, a0, a1); emit(m + 1, b0, b1);
  }
  if (m < t.mcount) {
    T a0 = (T)0, a1 = (T)0;
#pragma unroll
    for (int j = 0; j < NW; ++j) tap(j, xw[j], a0, a1);
    emit(m, a0, a1);
  }
}

// One vpolyN output from a coefficient row in shared memory, taps in groups of one 16-byte chunk per coefficient order
// (conflict-free vector loads for rows an odd number of chunks apart); same products, same order as the scalar form.
template <class T, int ORDER> RR_PROG T polyN_eval_row(const T *row, const T *x, T t, int n)
{
  typedef Arith<T> A;
  constexpr int kPer = 16 / (int)sizeof(T);
  struct alignas(16) Q { T v[kPer]; };
  T sum = (T)0;
  for (int g = 0; g < n / kPer; ++g) {
    T c[(ORDER + 1) * kPer];
#pragma unroll
    for (int i = 0; i <= ORDER; ++i) {
      const Q qv = reinterpret_cast<const Q *>(row)[g * (ORDER + 1) + i];
#pragma unroll
      for (int k = 0; k < kPer; ++k) c[i * kPer + k] = qv.v[k];
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      T v = c[k * (ORDER + 1)];
#pragma unroll
      for (int o = 1; o <= ORDER; ++o) v = A::add(A::mul(v, t), c[k * (ORDER + 1) + o]);
      sum = A::add(sum, A::mul(v, x[g * kPer + k]));
    }
  }
  return sum;
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
  // one output: position and phase from the 32.32 accumulator, Horner-interpolated taps summed in tap order
  auto position = [&](int j, int &q, int &phase, T &t) {
    const unsigned long long at = f0 + (unsigned long long)j * (unsigned long long)p.step;
    q = (int)(at >> 32);
    const uint32_t fraction = (uint32_t)at;
    phase = (int)(fraction >> (32 - p.phase_bits));
    t = A::mul((T)(uint32_t)(fraction << p.phase_bits), (T)(1.0 / 4294967296.0));
  };
  auto evaluate = [&](const T *c, const T *x, T t) -> T {
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
    return sum;
  };
  const int row_elems = p.n * (order + 1);
#if defined(__CUDA_ARCH__)
  if (p.row_pitch > 0) {
    // The phases of a warp's 32 outputs are scattered over the bank (up to 288 KB, L2-resident): a thread reading its
    // own row touches 32 different lines per load instruction. Instead the warp copies its 32 rows with coalesced
    // 16-byte loads into shared memory (row pitch an odd number of 16-byte units: conflict-free 16-byte reads), then
    // every thread evaluates its output from its row.
    constexpr int kPer = 16 / (int)sizeof(T);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int chunks = row_elems / kPer, pitch = p.row_pitch;
    T *rows = smem + ((p.win_cap + kPer - 1) / kPer) * kPer + (size_t)warp * 32 * pitch;
    for (int j0 = warp * 32; j0 < cnt; j0 += nwarps * 32) {
      const int j = j0 + lane < cnt ? j0 + lane : cnt - 1;
      int q, phase; T t;
      position(j, q, phase, t);
      const unsigned long long mine = (unsigned long long)(p.coefs + (long long)phase * row_elems);
      for (int ch = lane; ch < 32 * chunks; ch += 32) {
        const int r = ch / chunks, col = ch - r * chunks;
        const float4 *src = reinterpret_cast<const float4 *>(__shfl_sync(0xffffffffu, mine, r));
        *reinterpret_cast<float4 *>(rows + r * pitch + col * kPer) = __ldg(src + col);
      }
      __syncwarp();
      if (j0 + lane < cnt) {
        const T *row = rows + lane * pitch, *x = smem + q;
        const T y = order == 1 ? polyN_eval_row<T, 1>(row, x, t, p.n) : order == 2 ? polyN_eval_row<T, 2>(row, x, t, p.n)
                                                                                   : polyN_eval_row<T, 3>(row, x, t, p.n);
        view_write<OutT, T>(p.out, out_off, p.out_preload + i0 + j, y);
      }
      __syncwarp();
    }
    __syncthreads();
    return;
  }
#endif
  cta_for(cnt, [&](int j) {
    int q, phase; T t;
    position(j, q, phase, t);
    view_write<OutT, T>(p.out, out_off, p.out_preload + i0 + j, evaluate(p.coefs + (long long)phase * row_elems, smem + q, t));
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
  int nlanes, tile;              // tile: outputs per lane per CTA (multiple of 4)
  int CH;                        // lanes per CTA: all channels of a stream when the input is interleaved (1, 2, 4, 8)
  int half;                      // per-parity window capacity per lane (4 mod 32 words: no bank sharing between lanes)
  int qbits;                     // log2(tile / 4)
};

// y[k] = 0.5 x[2k+pre] + sum_t c[t] (x[2k+pre-(2t+1)] + x[2k+pre+(2t+1)]), summed in that order.
// The window is split by sample parity in shared memory (the centre tap reads one parity, all other taps the
// other one), so consecutive outputs read consecutive words; each thread produces four consecutive outputs
// from 16-byte shared loads (6.75 loaded values per output instead of 25).
// OPT = 8 (fp64 engine): eight consecutive outputs per thread, 5 loaded values per output, and rows padded by 16 bytes
// after every 64 so that the eight threads of a quarter-warp (80 bytes apart) hit eight different 16-byte bank groups.
template <int OPT> RR_HD int hb_pad(int j) { return OPT == 8 ? j + 2 * (j >> 3) : j; }

// Geometry of one tile (divisions, 64-bit coordinates), computed once per tile by the CTA leader.
template <class InT, class OutT> struct HbTile {
  long long k0, x0, in_off0, out_off0;
  const InT *src0; OutT *dst0;
  int cnt, win, in_direct, out_direct;
};
template <class T, class InT, class OutT, int NC>
RR_PROG HbTile<InT, OutT> hb_make_tile(const HalfbandParams<T> &p, long long work)
{
  constexpr int c = NC;
  const int CH = p.CH;                                   // 1, 2, 4 or 8; > 1 only for interleaved input
  HbTile<InT, OutT> t;
  const long long tiles = (p.nout + p.tile - 1) / p.tile;
  long long group; int tix_i;
  divmod_ll(work, (int)tiles, group, tix_i);
  const int lane0 = (int)group * CH;
  t.k0 = p.out0 + (long long)tix_i * p.tile;
  const long long rest = p.out0 + p.nout - t.k0;
  t.cnt = rest < p.tile ? (int)rest : p.tile;
  t.x0 = 2 * t.k0 + p.pre - (2 * c - 1);                 // first coordinate needed (window index u = coord - x0)
  t.win = 2 * (t.cnt - 1) + 2 * (2 * c - 1) + 1;
  t.in_off0 = lane_offset(p.in, lane0); t.out_off0 = lane_offset(p.out, lane0);
  t.in_direct = view_range_direct(p.in, t.x0, t.x0 + t.win);
  t.out_direct = view_range_direct(p.out, p.out_preload + t.k0, p.out_preload + t.k0 + t.cnt);
  t.src0 = view_ptr<const InT>(p.in, t.in_off0, t.x0);
  t.dst0 = view_ptr<OutT>(p.out, t.out_off0, p.out_preload + t.k0);
  return t;
}

// The window of a tile travels global -> registers -> shared memory in two steps, so that a persistent CTA can have
// the loads of its next tile in flight while it computes the current one. Thread `tid` of `nthreads` takes elements
// w = tid + k * nthreads, k < kHbRaw (vector path: 16-byte groups of four channels, k < kHbRaw / 4).
constexpr int kHbRaw = 40;
template <class InT> struct HbRegs { InT v[kHbRaw]; };
template <class T, class InT, class OutT> RR_HD bool hb_vec_path(const HalfbandParams<T> &p, const HbTile<InT, OutT> &tl)
{
  // interleaved frames of 4 or 8 channels of floats: one 16-byte load brings four channels of a frame
  return tl.in_direct && p.CH >= 4 && sizeof(InT) == 4 && p.in.ch_stride == 1 && p.in.elem_stride == p.CH && !((size_t)tl.src0 & 15);
}
template <class T, class InT, class OutT>
RR_PROG void hb_fetch(const HalfbandParams<T> &p, const HbTile<InT, OutT> &tl, HbRegs<InT> &r, int tid, int nthreads)
{
  const int CH = p.CH, chbits = CH == 8 ? 3 : CH == 4 ? 2 : CH == 2 ? 1 : 0;
  if (hb_vec_path(p, tl)) {
    struct alignas(16) V4 { InT a, b, c, d; };
    const int nvec = tl.win << (chbits - 2);
#pragma unroll
    for (int k = 0; k < kHbRaw / 4; ++k) {
      const int w = tid + k * nthreads;
      if (w < nvec) {
        const V4 v = ldg(reinterpret_cast<const V4 *>(tl.src0) + w);
        r.v[4 * k] = v.a; r.v[4 * k + 1] = v.b; r.v[4 * k + 2] = v.c; r.v[4 * k + 3] = v.d;
      } else { r.v[4 * k] = r.v[4 * k + 1] = r.v[4 * k + 2] = r.v[4 * k + 3] = (InT)0; }
    }
    return;
  }
  const int total = tl.win << chbits;
#pragma unroll
  for (int k = 0; k < kHbRaw; ++k) {
    const int w = tid + k * nthreads;
    InT v = (InT)0;
    if (w < total) {
      const int l = w & (CH - 1), u = w >> chbits;        // channel fastest: coalesced for interleaved input
      if (tl.in_direct) v = ldg(tl.src0 + (long long)l * p.in.ch_stride + (long long)u * p.in.elem_stride);
      else v = view_read<InT, InT>(p.in, tl.in_off0 + (long long)l * p.in.ch_stride, tl.x0 + u);
    }
    r.v[k] = v;
  }
}
// window index u: even u -> P0[u/2], odd u -> P1[(u+1)/2 + shift] with shift chosen so that the centre tap of
// output j (u = 2j + reach, odd) sits at P1[j + 4]: rows stay 16-byte aligned for the vector loads of the compute phase
template <class T, class InT, class OutT, int NC, int OPT>
RR_PROG void hb_put(const HalfbandParams<T> &p, const HbTile<InT, OutT> &tl, const HbRegs<InT> &r, T *smem, int tid, int nthreads)
{
  const int CH = p.CH, chbits = CH == 8 ? 3 : CH == 4 ? 2 : CH == 2 ? 1 : 0;
  T *P0 = smem, *P1 = smem + (long long)CH * p.half;     // even-u / odd-u samples, [lane][half]
  const int shift = 4 - NC;
  const bool vec = hb_vec_path(p, tl);
  // A thread's elements are nthreads apart: the same lane(s), frames u0 + k * du with du even (nthreads is a multiple
  // of 16 frames), so the parity array, the row and the index step are per-thread constants; the index step is a
  // multiple of 8, which the row padding maps to a constant step as well.
  const int ebits = vec ? chbits - 2 : chbits;           // log2 (elements or 16-byte vectors per frame)
  const int l0 = vec ? 4 * (tid & ((1 << ebits) - 1)) : (tid & (CH - 1));
  const int u0 = tid >> ebits, didx = (nthreads >> ebits) >> 1;
  const bool odd = u0 & 1;
  const int idx0 = odd ? ((u0 + 1) >> 1) + shift : (u0 >> 1);
  T *row = (odd ? P1 : P0) + l0 * p.half + hb_pad<OPT>(idx0);
  const int dph = hb_pad<OPT>(didx), n = (tl.win << chbits) >> (vec ? 2 : 0), half = p.half;
#pragma unroll
  for (int k = 0; k < (kHbRaw / 4); ++k) {
    if (!vec) break;
    const int w = tid + k * nthreads;
    if (w < n && (!odd || idx0 + k * didx >= 4)) {       // odd samples below the first centre tap are never read
      T *d = row + k * dph;
      d[0] = (T)r.v[4 * k]; d[half] = (T)r.v[4 * k + 1]; d[2 * half] = (T)r.v[4 * k + 2]; d[3 * half] = (T)r.v[4 * k + 3];
    }
  }
#pragma unroll
  for (int k = 0; k < kHbRaw; ++k) {
    if (vec) break;
    const int w = tid + k * nthreads;
    if (w < n && (!odd || idx0 + k * didx >= 4)) row[k * dph] = (T)r.v[k];
  }
}

// Outputs of work item `w` (< CH * tile / OPT) of the tile: OPT consecutive outputs of one lane.
template <class T, class InT, class OutT, int NC, int OPT>
RR_PROG void hb_compute_item(const HalfbandParams<T> &p, const HbTile<InT, OutT> &tl, const T *smem, int w)
{
  typedef Arith<T> A;
  constexpr int c = NC;
  const int CH = p.CH, cnt = tl.cnt;
  const T *P0 = smem, *P1 = smem + (long long)CH * p.half;
  const int qbits = p.qbits - (OPT == 8 ? 1 : 0);        // log2(tile / OPT); p.qbits = log2(tile / 4)
  const int l = w >> qbits, j = OPT * (w & ((1 << qbits) - 1));
  if (j >= cnt) return;
  const T *e = P0 + l * p.half + hb_pad<OPT>(j), *o = P1 + l * p.half + hb_pad<OPT>(j);
  // outputs j..j+OPT-1 use P0[j .. j+2c+OPT-2] and the centres P1[j+4 .. j+OPT+3]; j is a multiple of OPT, so the
  // padding of a logical offset from j is a compile-time constant
  constexpr int kPer = 16 / (int)sizeof(T);              // samples per 16-byte shared load
  constexpr int kVecs = (2 * c + OPT - 1 + kPer - 1) / kPer;
  struct alignas(16) Q { T v[kPer]; };
  T x[kVecs * kPer], ctr[OPT];
#pragma unroll
  for (int i = 0; i < kVecs; ++i) {
    const Q qv = *reinterpret_cast<const Q *>(e + hb_pad<OPT>(i * kPer));
#pragma unroll
    for (int k = 0; k < kPer; ++k) x[i * kPer + k] = qv.v[k];
  }
#pragma unroll
  for (int i = 0; i < OPT / kPer; ++i) {
    const Q qv = *reinterpret_cast<const Q *>(o + hb_pad<OPT>(4 + i * kPer));
#pragma unroll
    for (int k = 0; k < kPer; ++k) ctr[i * kPer + k] = qv.v[k];
  }
  T y[OPT];
#pragma unroll
  for (int r = 0; r < OPT; ++r) {
    T sum = A::mul(ctr[r], (T)0.5);
#pragma unroll
    for (int t = 0; t < c; ++t) sum = A::add(sum, A::mul(A::add(x[r + c - 1 - t], x[r + c + t]), p.coef[t]));
    y[r] = sum;
  }
  const long long off = tl.out_off0 + (long long)l * p.out.ch_stride;
  const long long cbase = p.out_preload + tl.k0 + j;
  if (j + OPT <= cnt && tl.out_direct) {
    const int es = p.out.elem_stride;
    OutT *d = tl.dst0 + (long long)l * p.out.ch_stride + j * es;
    if (es == 1 && !((size_t)d & (4 * sizeof(OutT) - 1))) {            // planar, aligned: vector stores
      struct alignas(4 * sizeof(OutT)) O4 { OutT a, b, c, d; };
#pragma unroll
      for (int r = 0; r < OPT; r += 4)
        reinterpret_cast<O4 *>(d)[r >> 2] = O4{(OutT)y[r], (OutT)y[r + 1], (OutT)y[r + 2], (OutT)y[r + 3]};
    } else {
#pragma unroll
      for (int r = 0; r < OPT; ++r) d[r * es] = (OutT)y[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < OPT; ++r)
      if (j + r < cnt) view_write<OutT, T>(p.out, off, cbase + r, y[r]);
  }
}

// One tile, start to end (the generic kernel; the emulation): y[k] = 0.5 x[2k+pre] + sum_t c[t] (x[2k+pre-(2t+1)] +
// x[2k+pre+(2t+1)]), summed in that order. The window is split by sample parity in shared memory (the centre tap
// reads one parity, all other taps the other one), so consecutive outputs read consecutive words; each thread
// produces OPT consecutive outputs from 16-byte shared loads.
template <class T, class InT, class OutT, int NC, int OPT = 4>
RR_PROG void halfband_program(const HalfbandParams<T> &p, long long work, T *smem)
{
#if defined(__CUDA_ARCH__)
  __shared__ HbTile<InT, OutT> tile_s;
  if (threadIdx.x == 0) tile_s = hb_make_tile<T, InT, OutT, NC>(p, work);
  __syncthreads();
  const HbTile<InT, OutT> tl = tile_s;                   // the program's last phase ends with a barrier: safe to reuse
  const int nthreads = blockDim.x;
  // the window in rounds of kHbRaw elements per thread
  const int total = hb_vec_path(p, tl) ? (tl.win * p.CH) >> 2 : tl.win * p.CH;
  (void)total;
  HbRegs<InT> r;
  hb_fetch(p, tl, r, (int)threadIdx.x, nthreads);
  hb_put<T, InT, OutT, NC, OPT>(p, tl, r, smem, (int)threadIdx.x, nthreads);
  __syncthreads();
  const int items = (p.CH * p.tile) / OPT;
  for (int w = threadIdx.x; w < items; w += nthreads) hb_compute_item<T, InT, OutT, NC, OPT>(p, tl, smem, w);
  __syncthreads();
#else
  const HbTile<InT, OutT> tl = hb_make_tile<T, InT, OutT, NC>(p, work);
  const int nthreads = 256;
  for (int tid = 0; tid < nthreads; ++tid) {
    HbRegs<InT> r;
    hb_fetch(p, tl, r, tid, nthreads);
    hb_put<T, InT, OutT, NC, OPT>(p, tl, r, smem, tid, nthreads);
  }
  const int items = (p.CH * p.tile) / OPT;
  for (int w = 0; w < items; ++w) hb_compute_item<T, InT, OutT, NC, OPT>(p, tl, smem, w);
#endif
}

}  // namespace b200rate
