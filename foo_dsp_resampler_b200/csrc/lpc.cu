// lpc.cu -- track-edge extrapolation by linear prediction, batched over (stream, channel) lanes on the device.
//
// What it replaces: lpc_extrapolate2 (lpc/lpc.cpp:25-71) as the plugin calls it at the start and the end of a track
// (foo_dsp_rate.cpp:165, :244-245, :288): from the first / last `prime` frames of a lane predict the frames before
// the beginning and after the end, so that the resampler's filters see a continuation instead of a step.
// Results are bit-identical to the reference: every operation below is the reference's operation in the reference's
// order (IEEE round-to-nearest, no contraction) -- only the SCHEDULE is different:
//
//   lpc_analyse_kernel   one warp per (lane, job).  Welch window in float (lpc.cpp:84-91), then the 33 lags
//                        (lpc.cpp:101-108): lane l of the warp owns lag l+1, every lane adds up lag 0 as well; the
//                        windowed samples pass through shared memory in tiles as doubles, each lag is ONE sequential
//                        double sum over ascending i exactly like the reference's inner loop (a product of two floats
//                        is exact in double, so DFMA == DMUL + DADD here). Levinson-Durbin + damping
//                        (lpc.cpp:111-166) on lane 0, un-fused. Output: 32 float coefficients per (lane, job),
//                        oldest-sample-first, zeros above the usable order.
//   lpc_extend_kernel    one THREAD per (lane, job, direction) chain, 32 chains per warp. The reference evaluates
//                        s = ((0 - x0*c0) - x1*c1) - ... - x31*c31 per new sample (lpc.cpp:168-191), a 32-deep
//                        dependent chain per sample. Here the 32 partial sums of the next 32 outputs live in
//                        registers (transposed-form FIR): when a sample becomes known its 32 products are
//                        subtracted from 32 different accumulators -- independent instructions -- and each
//                        accumulator still receives its terms oldest sample first, so the rounding sequence of every
//                        output is the reference's. The dependent path per sample is one FMUL + one FADD + the clamp.
//
// A "job" is one base segment with its own analysis: lpc_extrapolate2 is one job with both directions;
// the two edges of a track (backward from the first `prime` frames, forward from the last) are two jobs in one launch.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>

#include "b200_ratelib.h"
#include "engine.hpp"

namespace b200rate {
namespace {

constexpr int kOrd = 32;          // LPC_ORDER (lpc/lpc.h:24); the register layouts below are built for it
constexpr int kTile = 1024;       // windowed samples per shared-memory tile
constexpr int kHist = kOrd;       // samples before a tile that its lags reach back to

struct LpcJob {
  long long base;                 // first frame of the base segment, relative to a stream's frame 0 (may be negative)
  long long len;                  // frames in the base segment
  long long extra_bkwd, extra_fwd;
};

struct LpcParams {
  float *data;                    // frame 0 of stream 0
  long long stream_stride;        // floats between frame 0 of consecutive streams
  int nch;
  int order;                      // requested lpc_order, 1..32
  long long nlanes;               // nstreams * nch
  int njobs;
  LpcJob job[2];
  float *coefs;                   // [njobs][nlanes][32]
  double *debug;                  // optional [njobs][nlanes][33 + 32 + 1]: lags, lpc, usable order
};

__device__ __forceinline__ float *lane_ptr(const LpcParams &p, long long lane)
{
  long long s = lane / p.nch;
  int c = static_cast<int>(lane - s * p.nch);
  return p.data + s * p.stream_stride + c;
}

// lpc.cpp:111-166 on one thread; r[0..order], a[0..31] in shared memory, element i at [i * S] (S = 32: one column
// per thread of a warp, conflict-free). Returns the usable order.
template <int S>
__device__ int levinson_damped(const double *r, int order, double *a)
{
  double err = __dmul_rn(r[0], 1. + 1e-10);
  const double floor_ = __dadd_rn(__dmul_rn(1e-9, r[0]), 1e-10);
  int used = order;
  for (int i = 0; i < order; ++i) {
    if (err < floor_) {
      for (int j = i; j < order; ++j) a[j * S] = 0;
      used = i;
      break;
    }
    double k = -r[(i + 1) * S];
    for (int j = 0; j < i; ++j) k = __dsub_rn(k, __dmul_rn(a[j * S], r[(i - j) * S]));
    k = __ddiv_rn(k, err);
    a[i * S] = k;
    int j = 0;
    for (; j < i / 2; ++j) {
      double lo = a[j * S], hi = a[(i - 1 - j) * S];
      a[j * S] = __dadd_rn(lo, __dmul_rn(k, hi));
      a[(i - 1 - j) * S] = __dadd_rn(hi, __dmul_rn(k, lo));
    }
    if (i & 1) a[j * S] = __dadd_rn(a[j * S], __dmul_rn(a[j * S], k));
    err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(k, k)));
  }
  double damp = 0.999;
  for (int j = 0; j < used; ++j) {
    a[j * S] = __dmul_rn(a[j * S], damp);
    damp = __dmul_rn(damp, 0.999);
  }
  if (used == 0) {
    used = 1;
    a[0] = -1;
  }
  for (int j = order; j < kOrd; ++j) a[j * S] = 0;
  return used;
}

__global__ void __launch_bounds__(32) lpc_analyse_kernel(const LpcParams p)
{
  __shared__ double xs[kHist + kTile];
  __shared__ double r_s[kOrd + 1];
  __shared__ double a_s[kOrd];
  const int l = threadIdx.x;
  const long long lane = blockIdx.x;
  const int jb = blockIdx.y;
  const LpcJob job = p.job[jb];
  const float *src = lane_ptr(p, lane) + job.base * p.nch;
  const long long len = job.len;
  const float half = __fdiv_rn(static_cast<float>(static_cast<unsigned long long>(len + 1)), 2.0f);

  // windowed sample i of the base segment as the reference stores it in tdata[] (float), 0 outside
  auto windowed = [&](long long i) -> float {
    if (i >= len) return 0.f;
    float x = __ldg(src + i * p.nch);
    float k = __fdiv_rn(__fsub_rn(static_cast<float>(static_cast<int>(i) + 1), half), half);
    return __fmul_rn(x, __fsub_rn(1.0f, __fmul_rn(k, k)));
  };

  xs[l] = 0;                                    // x[i - j] for i < j: a zero product leaves a sum unchanged
  double acc = 0, acc0 = 0;
  constexpr int kPer = kTile / 32;
  float pre[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) pre[u] = windowed(l + 32 * u);
  for (long long t0 = 0; t0 < len; t0 += kTile) {
#pragma unroll
    for (int u = 0; u < kPer; ++u) xs[kHist + l + 32 * u] = static_cast<double>(pre[u]);
    __syncwarp();
    if (t0 + kTile < len) {
#pragma unroll
      for (int u = 0; u < kPer; ++u) pre[u] = windowed(t0 + kTile + l + 32 * u);
    }
    const int n = static_cast<int>(len - t0 < kTile ? len - t0 : kTile);
    const double *own = xs + kHist - 1 - l;     // x[i - (l+1)]
    const double *cur = xs + kHist;             // x[i]
    int t = 0;
    for (; t + 4 <= n; t += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double xi = cur[t + u];
        acc0 = __fma_rn(xi, xi, acc0);
        acc = __fma_rn(xi, own[t + u], acc);
      }
    }
    for (; t < n; ++t) {
      double xi = cur[t];
      acc0 = __fma_rn(xi, xi, acc0);
      acc = __fma_rn(xi, own[t], acc);
    }
    __syncwarp();
    double tail = xs[kTile + l];                // the last kHist samples become the next tile's history
    __syncwarp();
    xs[l] = tail;
  }
  r_s[l + 1] = acc;
  if (l == 0) r_s[0] = acc0;
  __syncwarp();
  int used = 0;
  if (l == 0) used = levinson_damped<1>(r_s, p.order, a_s);
  __syncwarp();
  // history slot j (0 = oldest of the 32 samples behind an output) meets lpc[31 - j]
  const long long slot = static_cast<long long>(jb) * p.nlanes + lane;
  p.coefs[slot * kOrd + l] = static_cast<float>(a_s[kOrd - 1 - l]);
  if (p.debug) {
    double *d = p.debug + slot * (kOrd + 1 + kOrd + 1);
    d[l] = r_s[l];
    if (l == 0) d[kOrd] = r_s[kOrd];
    d[kOrd + 1 + l] = a_s[l];
    if (l == 0) d[kOrd + 1 + kOrd] = used;
  }
}

// The same analysis for LARGE batches: one THREAD per (lane, job), 32 of them per warp. A thread keeps the last 32
// windowed samples of its lane in registers (a delay line) and all 33 lag sums, so a term costs one shared-memory read
// and 33 DFMAs instead of 33 reads -- the narrow kernel above is bound by the shared-memory pipe (4 wavefronts per
// term), this one by the FP64 pipe. Each lag is still one sequential sum over ascending i. All jobs of a launch have
// the same length (both entry points guarantee it). Raw samples of the next tile arrive by LDGSTS during the sums.
constexpr int kWTile = 32;        // terms per tile
constexpr int kWStep = 8;         // terms per unrolled trip; the delay line moves down by kWStep after each

__global__ void __launch_bounds__(32) lpc_analyse_wide_kernel(const LpcParams p)
{
  __shared__ float stage[2][32][kWTile];            // raw samples [buffer][job of the warp][term]
  __shared__ double cols[(2 * kOrd + 2) * 33];      // tile as doubles [term][job] (pitch 33), later r / a columns
  const int l = threadIdx.x;
  const long long slots = p.nlanes * p.njobs;
  const long long g = static_cast<long long>(blockIdx.x) * 32 + l;
  const bool live = g < slots;
  const long long gl = live ? g : slots - 1;
  const int jb = static_cast<int>(gl / p.nlanes);
  const long long lane = gl - jb * p.nlanes;
  const long long len = p.job[0].len;
  const float *src = lane_ptr(p, lane) + p.job[jb].base * p.nch;
  const float half = __fdiv_rn(static_cast<float>(static_cast<unsigned long long>(len + 1)), 2.0f);

  // thread l fetches term l of every job of the warp: the 32 source pointers travel by shuffle
  auto fetch = [&](int buf, long long t0) {
    const long long i = t0 + l;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float *sk = reinterpret_cast<const float *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(src), k));
      if (i < len) {
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&stage[buf][k][l]));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(sk + i * p.nch));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double acc[kOrd + 1], dl[kOrd + kWStep];          // dl[j]: windowed sample kOrd - j behind the trip's first term
#pragma unroll
  for (int j = 0; j <= kOrd; ++j) acc[j] = 0;
#pragma unroll
  for (int j = 0; j < kOrd + kWStep; ++j) dl[j] = 0;
  fetch(0, 0);
  int buf = 0;
  for (long long t0 = 0; t0 < len; t0 += kWTile, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    {   // window (lpc.cpp:84-91) and widen term l of every job
      const long long i = t0 + l;
      const float k = __fdiv_rn(__fsub_rn(static_cast<float>(static_cast<int>(i) + 1), half), half);
      const float w = __fsub_rn(1.0f, __fmul_rn(k, k));
#pragma unroll 8
      for (int j = 0; j < 32; ++j) cols[l * 33 + j] = i < len ? static_cast<double>(__fmul_rn(stage[buf][j][l], w)) : 0.0;
    }
    __syncwarp();
    if (t0 + kWTile < len) fetch(buf ^ 1, t0 + kWTile);
    for (int t = 0; t < kWTile; t += kWStep) {
#pragma unroll
      for (int u = 0; u < kWStep; ++u) {
        const double x = cols[(t + u) * 33 + l];
        dl[kOrd + u] = x;
        acc[0] = __fma_rn(x, x, acc[0]);
#pragma unroll
        for (int j = 1; j <= kOrd; ++j) acc[j] = __fma_rn(x, dl[kOrd + u - j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < kOrd; ++j) dl[j] = dl[j + kWStep];
    }
    __syncwarp();
  }
  // r and a columns of this thread: element i at cols[i * 32 + l]
  double *r_c = cols + l, *a_c = cols + (kOrd + 1) * 32 + l;
#pragma unroll
  for (int j = 0; j <= kOrd; ++j) r_c[j * 32] = acc[j];
  const int used = levinson_damped<32>(r_c, p.order, a_c);
  if (!live) return;
#pragma unroll 8
  for (int j = 0; j < kOrd; ++j) p.coefs[g * kOrd + j] = static_cast<float>(a_c[(kOrd - 1 - j) * 32]);
  if (p.debug) {
    double *d = p.debug + g * (kOrd + 1 + kOrd + 1);
    for (int j = 0; j <= kOrd; ++j) d[j] = r_c[j * 32];
    for (int j = 0; j < kOrd; ++j) d[kOrd + 1 + j] = a_c[j * 32];
    d[kOrd + 1 + kOrd] = used;
  }
}

__device__ __forceinline__ float clamp10(float s)
{
  return s > 10.f ? 10.f : (s < -10.f ? -10.f : s);
}

// U consecutive outputs per loop trip with compile-time register indices, then the accumulators move down by U
// (U register moves per U samples). The fully unrolled form (U = 32, no moves) is a 39 KB loop body that runs out of
// the instruction cache -- measured: 1.05 no-instruction stalls per issued instruction (profiles/README.md).
template <int U>
__global__ void __launch_bounds__(32) lpc_extend_kernel(const LpcParams p, const int nslots, const int4 slots)
{
  // slots.{x,y,z,w}: (job * 2 + direction) of the chain groups present in this launch; direction 1 = backward
  const long long id = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  if (id >= p.nlanes * nslots) return;
  const int q = static_cast<int>(id / p.nlanes);
  const long long lane = id - q * p.nlanes;
  const int code = q == 0 ? slots.x : q == 1 ? slots.y : q == 2 ? slots.z : slots.w;
  const int jb = code >> 1;
  const bool back = code & 1;
  const LpcJob job = p.job[jb];
  const long long extra = back ? job.extra_bkwd : job.extra_fwd;
  float *base = lane_ptr(p, lane) + job.base * p.nch;          // frame 0 of the base segment, this channel
  // time runs forward along +step from `origin`: sample m of the chain's own axis (m < 0: base, m >= 0: new)
  const long long step = back ? -static_cast<long long>(p.nch) : p.nch;
  float *origin = back ? base - p.nch : base + job.len * p.nch;

  float c[kOrd], A[kOrd + U];                                   // A[k]: partial sum of output n0 + k
  const float *cf = p.coefs + (static_cast<long long>(jb) * p.nlanes + lane) * kOrd;
#pragma unroll
  for (int j = 0; j < kOrd; ++j) c[j] = __ldg(cf + j);         // c[j] multiplies the sample 32 - j behind an output
#pragma unroll
  for (int j = 0; j < kOrd + U; ++j) A[j] = 0.f;
  // The loop starts 32 samples before the first new one: there a step takes its sample from the base segment instead
  // of from its accumulator, which is how the 32 partial sums of outputs 0..31 get their base-segment terms (oldest
  // first, from zero) with the code of the main loop.
  float *out = origin - kOrd * step;
  for (long long n0 = -kOrd; n0 < extra; n0 += U, out += U * step) {
#pragma unroll
    for (int s = 0; s < U; ++s) {
      float y;
      if (n0 < 0) {
        y = -(n0 + s) <= job.len ? out[s * step] : 0.f;         // a shorter base: nothing there (its coefficient is 0)
      } else {
        y = clamp10(A[s]);
        if (n0 + s < extra) out[s * step] = y;
      }
      // output n0 + s + d sits in A[s + d] and takes this sample with c[32 - d]; d = 32 opens a new sum
#pragma unroll
      for (int d = 1; d <= kOrd; ++d) A[s + d] = __fsub_rn(A[s + d], __fmul_rn(y, c[kOrd - d]));
    }
#pragma unroll
    for (int k = 0; k < kOrd; ++k) A[k] = A[k + U];
#pragma unroll
    for (int k = kOrd; k < kOrd + U; ++k) A[k] = 0.f;
  }
}

int cuda_fail(cudaError_t e, const char *what)
{
  set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? RR_ENOMEM : RR_INTERNAL;
}

// Stream-ordered scratch for the coefficient rows comes from a pool of the library's own, one per device, that keeps
// its memory between calls (the default pool hands it back to the driver at every synchronisation, which put a
// driver allocation of ~0.5 ms in front of every call).
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, stream);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      e = cudaMemPoolCreate(&pools[dev], &props);
      if (e != cudaSuccess) { pools[dev] = nullptr; return e; }
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = pools[dev];
  }
  return cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
}

// Thread-per-job analysis once there are enough jobs, else warp-per-job. Measured on a B200 (profiles/README.md): the
// warp-per-job kernel costs ~18 ps per (job, term) once the SMs are full, the thread-per-job kernel ~104 ns per term
// whatever the job count up to one warp per SM sub-partition (18 944 jobs) -- they cross near 5 800 jobs.
void launch_analysis(const LpcParams &p, cudaStream_t stream)
{
  const long long slots = p.nlanes * p.njobs;
  static const char *force = getenv("B200RATE_LPC_ANALYSIS");     // "wide" / "narrow": probes only
  bool wide = slots >= 6144 && (p.njobs == 1 || p.job[0].len == p.job[1].len);
  if (force && force[0] == 'w' && (p.njobs == 1 || p.job[0].len == p.job[1].len)) wide = true;
  if (force && force[0] == 'n') wide = false;
  if (wide)
    lpc_analyse_wide_kernel<<<static_cast<unsigned>((slots + 31) / 32), 32, 0, stream>>>(p);
  else
    lpc_analyse_kernel<<<dim3(static_cast<unsigned>(p.nlanes), p.njobs), 32, 0, stream>>>(p);
}

// Both kernels for up to two jobs on `stream`; the coefficient scratch is stream-ordered.
int run_jobs(float *d_data, size_t nstreams, long long stream_stride, int nch, int order, const LpcJob *jobs, int njobs,
             cudaStream_t stream, double *d_debug)
{
  if (!d_data || nch < 1 || nstreams < 1 || order < 1 || order > kOrd) {
    set_last_error("lpc: bad arguments (lpc_order must be 1..32)");
    return RR_INVPARAM;
  }
  LpcParams p{};
  p.data = d_data;
  p.stream_stride = stream_stride;
  p.nch = nch;
  p.order = order;
  p.nlanes = static_cast<long long>(nstreams) * nch;
  p.njobs = njobs;
  p.debug = d_debug;
  int codes[4] = {0, 0, 0, 0}, nslots = 0;
  for (int j = 0; j < njobs; ++j) {
    p.job[j] = jobs[j];
    if (jobs[j].len < 0 || jobs[j].len > 0x7ffffff0ll || jobs[j].extra_bkwd < 0 || jobs[j].extra_fwd < 0) {
      set_last_error("lpc: bad segment length");
      return RR_INVPARAM;
    }
    if (jobs[j].extra_fwd) codes[nslots++] = j * 2;
    if (jobs[j].extra_bkwd) codes[nslots++] = j * 2 + 1;
  }
  if (nslots == 0) return RR_OK;
  if (p.nlanes > 0x7fffffffll) { set_last_error("lpc: too many lanes for one launch"); return RR_INVPARAM; }
  cudaError_t e = scratch_alloc(reinterpret_cast<void **>(&p.coefs), sizeof(float) * kOrd * p.nlanes * njobs, stream);
  if (e != cudaSuccess) return cuda_fail(e, "lpc: cudaMallocAsync");
  launch_analysis(p, stream);
  const long long chains = p.nlanes * nslots;
  const int4 sl = make_int4(codes[0], codes[1], codes[2], codes[3]);
  const unsigned grid = static_cast<unsigned>((chains + 31) / 32);
  static const int variant = getenv("B200RATE_LPC_UNROLL") ? atoi(getenv("B200RATE_LPC_UNROLL")) : 8;
  switch (variant) {
    case 4: lpc_extend_kernel<4><<<grid, 32, 0, stream>>>(p, nslots, sl); break;
    case 16: lpc_extend_kernel<16><<<grid, 32, 0, stream>>>(p, nslots, sl); break;
    case 32: lpc_extend_kernel<32><<<grid, 32, 0, stream>>>(p, nslots, sl); break;
    default: lpc_extend_kernel<8><<<grid, 32, 0, stream>>>(p, nslots, sl); break;
  }
  e = cudaGetLastError();
  cudaError_t e2 = cudaFreeAsync(p.coefs, stream);
  if (e != cudaSuccess) return cuda_fail(e, "lpc: kernel launch");
  if (e2 != cudaSuccess) return cuda_fail(e2, "lpc: cudaFreeAsync");
  return RR_OK;
}

}  // namespace
}  // namespace b200rate

using namespace b200rate;

extern "C" {

int RRX_track_edge_lengths(unsigned in_rate, unsigned out_rate, unsigned *add, unsigned *drop, unsigned *prime_len,
                           unsigned *inbuf_frames)
{
  if (!in_rate || !out_rate) return RR_INVPARAM;
  // util.h:38-49 with N = 20, M = 8192: the same duration at both rates, at most 1/20 s and 8192 frames
  unsigned a = in_rate, b = out_rate;
  while (b) { unsigned t = a % b; a = b; b = t; }
  const unsigned g = a;
  unsigned per_in = in_rate / g, per_out = out_rate / g;
  unsigned reps = (g + 19) / 20;
  const unsigned longer = per_in > per_out ? per_in : per_out;
  if (static_cast<unsigned long long>(longer) * reps > 8192u) reps = 8192u / longer;
  if (reps < 1) reps = 1;
  if (add) *add = per_in * reps;
  if (drop) *drop = per_out * reps;
  // foo_dsp_rate.cpp:100-101
  unsigned block = in_rate / 10;
  block = block < 2048u ? 2048u : block > 65536u ? 65536u : block;
  unsigned prime = in_rate / 20;
  prime = prime < 1024u ? 1024u : prime > 16384u ? 16384u : prime;
  if (prime < 2u * RRX_LPC_ORDER + 1) prime = 2u * RRX_LPC_ORDER + 1;
  if (prime_len) *prime_len = prime;
  if (inbuf_frames) *inbuf_frames = block;
  return RR_OK;
}

int RRX_lpc_extrapolate_batch(float *d_data, size_t nstreams, size_t stream_stride_frames, size_t data_len, int nchannels,
                              int lpc_order, size_t extra_bkwd, size_t extra_fwd, void *stream)
{
  LpcJob j{0, static_cast<long long>(data_len), static_cast<long long>(extra_bkwd), static_cast<long long>(extra_fwd)};
  return run_jobs(d_data, nstreams, static_cast<long long>(stream_stride_frames) * nchannels, nchannels, lpc_order, &j, 1,
                  static_cast<cudaStream_t>(stream), nullptr);
}

int RRX_lpc_extend_tracks(float *d_padded, size_t nstreams, size_t track_frames, size_t prime_len, int nchannels,
                          int lpc_order, size_t extra, void *stream)
{
  if (prime_len > track_frames) prime_len = track_frames;     // foo_dsp_rate.cpp:243
  LpcJob j[2] = {{0, static_cast<long long>(prime_len), static_cast<long long>(extra), 0},
                 {static_cast<long long>(track_frames - prime_len), static_cast<long long>(prime_len), 0,
                  static_cast<long long>(extra)}};
  if (!d_padded) { set_last_error("lpc: NULL buffer"); return RR_INVPARAM; }
  return run_jobs(d_padded + extra * nchannels, nstreams, static_cast<long long>(track_frames + 2 * extra) * nchannels,
                  nchannels, lpc_order, j, 2, static_cast<cudaStream_t>(stream), nullptr);
}

int RRX_lpc_analysis_dump(const float *d_data, size_t nstreams, size_t stream_stride_frames, size_t data_len, int nchannels,
                          int lpc_order, double *d_out, void *stream)
{
  if (!d_data || !d_out || nchannels < 1 || nstreams < 1 || lpc_order < 1 || lpc_order > kOrd) {
    set_last_error("lpc: bad arguments (lpc_order must be 1..32)");
    return RR_INVPARAM;
  }
  LpcParams p{};
  p.data = const_cast<float *>(d_data);
  p.stream_stride = static_cast<long long>(stream_stride_frames) * nchannels;
  p.nch = nchannels;
  p.order = lpc_order;
  p.nlanes = static_cast<long long>(nstreams) * nchannels;
  p.njobs = 1;
  p.job[0] = LpcJob{0, static_cast<long long>(data_len), 0, 0};
  p.debug = d_out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = scratch_alloc(reinterpret_cast<void **>(&p.coefs), sizeof(float) * kOrd * p.nlanes, st);
  if (e != cudaSuccess) return cuda_fail(e, "lpc: cudaMallocAsync");
  launch_analysis(p, st);
  e = cudaGetLastError();
  cudaFreeAsync(p.coefs, st);
  return e == cudaSuccess ? RR_OK : cuda_fail(e, "lpc: kernel launch");
}

int RRX_lpc_extrapolate2(float *data, size_t data_len, int nchannels, int lpc_order, size_t extra_bkwd, size_t extra_fwd)
{
  if (!data || nchannels < 1 || lpc_order < 1 || lpc_order > kOrd) {
    set_last_error("lpc: bad arguments (lpc_order must be 1..32)");
    return RR_INVPARAM;
  }
  if (!extra_bkwd && !extra_fwd) return RR_OK;
  const size_t total = extra_bkwd + data_len + extra_fwd;
  float *d = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "lpc: cudaStreamCreate");
  int rc = RR_OK;
  e = cudaMalloc(reinterpret_cast<void **>(&d), sizeof(float) * total * nchannels);
  if (e != cudaSuccess) { rc = cuda_fail(e, "lpc: cudaMalloc"); cudaStreamDestroy(st); return rc; }
  float *d0 = d + extra_bkwd * nchannels;
  e = cudaMemcpyAsync(d0, data, sizeof(float) * data_len * nchannels, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    LpcJob j{0, static_cast<long long>(data_len), static_cast<long long>(extra_bkwd), static_cast<long long>(extra_fwd)};
    rc = run_jobs(d0, 1, static_cast<long long>(total) * nchannels, nchannels, lpc_order, &j, 1, st, nullptr);
    if (rc == RR_OK && extra_bkwd)
      e = cudaMemcpyAsync(data - extra_bkwd * nchannels, d, sizeof(float) * extra_bkwd * nchannels, cudaMemcpyDeviceToHost, st);
    if (rc == RR_OK && e == cudaSuccess && extra_fwd)
      e = cudaMemcpyAsync(data + data_len * nchannels, d0 + data_len * nchannels, sizeof(float) * extra_fwd * nchannels,
                          cudaMemcpyDeviceToHost, st);
  }
  cudaError_t es = cudaStreamSynchronize(st);
  if (rc == RR_OK && (e != cudaSuccess || es != cudaSuccess)) rc = cuda_fail(e != cudaSuccess ? e : es, "lpc: transfer");
  cudaFree(d);
  cudaStreamDestroy(st);
  return rc;
}

int RRX_lpc_extrapolate_bkwd(float *data, size_t data_len, size_t prime_len, int nchannels, int lpc_order, size_t extra_bkwd)
{
  (void)data_len;                                              // lpc/lpc.h:28-32
  return RRX_lpc_extrapolate2(data, prime_len, nchannels, lpc_order, extra_bkwd, 0);
}

int RRX_lpc_extrapolate_fwd(float *data, size_t data_len, size_t prime_len, int nchannels, int lpc_order, size_t extra_fwd)
{
  if (!data || prime_len > data_len) { set_last_error("lpc: prime_len exceeds data_len"); return RR_INVPARAM; }
  return RRX_lpc_extrapolate2(data + (data_len - prime_len) * nchannels, prime_len, nchannels, lpc_order, 0, extra_fwd);
}

}  // extern "C"
