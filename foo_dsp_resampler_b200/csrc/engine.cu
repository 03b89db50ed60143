// engine.cu -- device engine: table upload, stage launches, and the two front-ends (batch / streaming).
//
// Built by nvcc for sm_100a this is the product. With -DB200RATE_EMU (tests/emu only) the "device" is
// host memory and a launch is a serial loop over the same CTA programs, which lets the host-side
// orchestration and all index arithmetic be checked against the oracle where no GPU exists. The EMU
// build is test infrastructure and is never part of libb200rate.so.
//
// Coordinates: FIFO i feeds stage i and starts with `preload` zeros (rate/rate_base.h:417-421); stage i
// output k lands at coordinate preload[i+1] + k of FIFO i+1. All block / phase positions are closed-form
// functions of absolute indices, so any range of any stage can be (re)computed independently -- the
// property the reference only has implicitly through its FIFO state (rate/rate_base.h:96-128).
#include "engine.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <type_traits>
#include <vector>

#include "fft_tables.hpp"
#include "rate_kernels.cuh"
#include "rate_kernels_pk.cuh"
#include "rate_kernels_f64.cuh"
#include "rate_kernels_fused.cuh"

#ifndef B200RATE_EMU
#include <cuda_runtime.h>
#endif

namespace b200rate {

// ===================================================================================================
// error plumbing
// ===================================================================================================
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
const char *last_error() { return g_last_error.c_str(); }

// ===================================================================================================
// backend: CUDA, or host emulation for tests
// ===================================================================================================
#ifdef B200RATE_EMU
#include "emu_backend.inc"   // tests/emu: host memory as the "device" (test infrastructure, not in this tree)
#else
constexpr int kEmulated = 0;
typedef cudaStream_t stream_t;
static int cuda_fail(cudaError_t e, const char *what)
{
  set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? RR_ENOMEM : RR_INTERNAL;
}
#define CUDA_TRY(expr)                                           \
  do {                                                           \
    cudaError_t e_ = (expr);                                     \
    if (e_ != cudaSuccess) return cuda_fail(e_, #expr);          \
  } while (0)
static int be_set_device(int dev)
{
  if (dev >= 0) CUDA_TRY(cudaSetDevice(dev));
  int cur = -1;
  CUDA_TRY(cudaGetDevice(&cur));
  CUDA_TRY(cudaFree(0));
  return RR_OK;
}
static int be_malloc(void **p, size_t n) { CUDA_TRY(cudaMalloc(p, n ? n : 1)); return RR_OK; }
static void be_free(void *p) { if (p) cudaFree(p); }
static int be_h2d(void *d, const void *h, size_t n, stream_t s)
{
  CUDA_TRY(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
  return RR_OK;
}
static int be_d2h(void *h, const void *d, size_t n, stream_t s)
{
  CUDA_TRY(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s));
  return RR_OK;
}
static int be_sync(stream_t s) { CUDA_TRY(cudaStreamSynchronize(s)); return RR_OK; }
static int be_stream_create(stream_t *s) { CUDA_TRY(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)); return RR_OK; }
static void be_stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
static int be_num_sms()
{
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
static size_t be_max_smem()
{
  int dev = 0, n = 48 * 1024;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return static_cast<size_t>(n);
}
static int be_current_device()
{
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : 0;
}
// A handle / batch lives on the device it was opened on; every public entry point selects that device for the
// duration of the call (the caller may drive it from any host thread, whose current device is arbitrary) and
// restores the caller's device on return.
struct DeviceScope {
  int prev_ = -1, want_;
  bool ok_ = true;
  explicit DeviceScope(int device) : want_(device)
  {
    if (cudaGetDevice(&prev_) != cudaSuccess) { ok_ = false; return; }
    if (prev_ != want_ && cudaSetDevice(want_) != cudaSuccess) ok_ = false;
  }
  ~DeviceScope() { if (ok_ && prev_ != want_) cudaSetDevice(prev_); }
  bool ok() const { return ok_; }
};
#endif
#ifndef B200RATE_EMU
typedef cudaEvent_t event_t;
// page-locked host memory for the staging buffers of the host-facing entry points
static int be_host_alloc(void **p, size_t n) { CUDA_TRY(cudaHostAlloc(p, n ? n : 1, cudaHostAllocDefault)); return RR_OK; }
static void be_host_free(void *p) { if (p) cudaFreeHost(p); }
// true when `p` is page-locked host memory CUDA knows about (cudaHostAlloc / cudaHostRegister): an asynchronous
// copy from / to it really is asynchronous, so it can be used directly instead of the library's staging buffer
static bool be_host_is_pinned(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
static int be_memset(void *d, int v, size_t n, stream_t s) { CUDA_TRY(cudaMemsetAsync(d, v, n, s)); return RR_OK; }
static int be_event_create(event_t *e) { CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming)); return RR_OK; }
static void be_event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
static int be_event_record(event_t e, stream_t s) { CUDA_TRY(cudaEventRecord(e, s)); return RR_OK; }
static int be_event_sync(event_t e) { CUDA_TRY(cudaEventSynchronize(e)); return RR_OK; }
#endif
#define RR_DEVICE_SCOPE(dev)                                                               \
  DeviceScope device_scope_(dev);                                                          \
  if (!device_scope_.ok()) { set_last_error("cannot select the handle's device"); return RR_INTERNAL; }

// ===================================================================================================
// kernels
// ===================================================================================================
constexpr int kDftThreads = 256;
constexpr int kTileThreads = 256;
constexpr int kPolyTile = 2048;      // outputs per CTA tile
constexpr int kHalfTile = 2048;
constexpr int kCopyTile = 4096;      // elements per CTA tile (lane-fastest)
// filter-spectrum values cached in registers per thread (phase-4 items threadIdx.x + k*blockDim.x, k < depth)
template <class T> struct DftCacheDepth { static constexpr int value = sizeof(T) == 4 ? 8 : 4; };

struct CopyParams {                  // out[coord c0 + j] = in[coord c0 + j + in_shift]
  LaneView in, out;
  long long c0, n, in_shift;
  int nlanes;
};

template <class InT, class OutT>
RR_PROG void copy_program(const CopyParams &p, long long work)
{
  const long long total = p.n * p.nlanes;
  const long long e0 = work * kCopyTile;
  const int cnt = (total - e0) < kCopyTile ? (int)(total - e0) : kCopyTile;
  const long long j0 = e0 / p.nlanes;
  const unsigned l0 = (unsigned)(e0 - j0 * p.nlanes);
  cta_for(cnt, [&](int t) {
    const unsigned u = l0 + (unsigned)t, dj = u / (unsigned)p.nlanes;
    const int l = (int)(u - dj * (unsigned)p.nlanes);
    const long long j = j0 + dj;
    const OutT v = (OutT)view_read<InT, InT>(p.in, lane_offset(p.in, l), p.c0 + j + p.in_shift);
    view_write<OutT, OutT>(p.out, lane_offset(p.out, l), p.c0 + j, v);
  });
}

#ifndef B200RATE_EMU
extern __shared__ __align__(16) unsigned char rr_smem_raw[];

// Persistent CTA: stage the twiddle / cosine tables in shared memory and the filter spectrum values this
// thread will need in registers, once; then loop over (block, lane group) work items, prefetching the next
// item's input tile with LDGSTS while the current one is transformed.
template <class T, class InT, class OutT, int LPC>
__global__ void __launch_bounds__(2 * kDftThreads) dft_kernel(const __grid_constant__ DftParams<T> p, long long nwork, int data_bytes)
{
  C2<T> *data = reinterpret_cast<C2<T> *>(rr_smem_raw);
  T *t = reinterpret_cast<T *>(rr_smem_raw + data_bytes);
  const int nf = p.fwd.pyr_len, ni = p.inv.pyr_len, tf = (p.Pf >> 2) + 1, ti = (p.Ni >> 2) + 1;
  DftTables<T> tab;
  if (p.lean_tables) {                                   // pyramids only (shared when both sizes agree), cosines stay in global memory
    const bool same = p.fwd.bits == p.inv.bits;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) t[i] = p.pyr_f[i];
    if (!same) for (int i = threadIdx.x; i < ni; i += blockDim.x) t[nf + i] = p.pyr_i[i];
    tab = DftTables<T>{t, same ? t : t + nf, p.tcos_f, p.tcos_i};
  } else {
    for (int i = threadIdx.x; i < nf; i += blockDim.x) t[i] = p.pyr_f[i];
    for (int i = threadIdx.x; i < ni; i += blockDim.x) t[nf + i] = p.pyr_i[i];
    for (int i = threadIdx.x; i < tf; i += blockDim.x) t[nf + ni + i] = p.tcos_f[i];
    for (int i = threadIdx.x; i < ti; i += blockDim.x) t[nf + ni + tf + i] = p.tcos_i[i];
    tab = DftTables<T>{t, t + nf, t + nf + ni, t + nf + ni + tf};
  }
  CoefCache<T, DftCacheDepth<T>::value> cc;
  dft_load_coef_cache(p, cc);
  __shared__ DftItem<T> items[2];
  if (threadIdx.x == 0 && (long long)blockIdx.x < nwork) items[0] = dft_item<T, LPC>(p, blockIdx.x);
  __syncthreads();
  if (p.zstride > 0 && (long long)blockIdx.x < nwork)             // prime the prefetch pipeline
    dft_stage_tile<T, T, LPC, true>(p, items[0], data + LPC * (p.xstride + p.ystride), p.zstride);
  int slot = 0;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x, slot ^= 1) {
    const long long next = w + gridDim.x < nwork ? w + gridDim.x : -1;
    dft_stage_program<T, InT, OutT, LPC, DftCacheDepth<T>::value>(p, tab, cc, items, slot, next, data);
  }
}
// DFT blocks that do not fit shared memory (N > 16384 fp32 / 8192 fp64: transition bands below ~1 % of Nyquist):
// the same CTA program with its work buffers in a per-CTA slice of global scratch (L2-resident) and the tables
// read from global memory. One lane per CTA, no prefetch buffer. Slow path, there for coverage of RR_config.
template <class T, class InT, class OutT>
__global__ void __launch_bounds__(2 * kDftThreads) dft_big_kernel(const __grid_constant__ DftParams<T> p, long long nwork,
                                                                   C2<T> *scratch, unsigned long long per_cta)
{
  C2<T> *data = scratch + (size_t)blockIdx.x * per_cta;
  const DftTables<T> tab{p.pyr_f, p.pyr_i, p.tcos_f, p.tcos_i};
  const CoefCache<T, 0> cc{};
  __shared__ DftItem<T> items[2];
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    if (threadIdx.x == 0) items[0] = dft_item<T, 1>(p, w);
    __syncthreads();
    dft_stage_program<T, InT, OutT, 1, 0>(p, tab, cc, items, 0, -1, data);
    __syncthreads();
  }
}

// Lane-pair DFT stage (rate_kernels_pk.cuh): a CTA is `groups` independent groups of `gthreads` threads, each
// a persistent worker with its own forward and inverse buffer behind the shared tables (twiddle pyramids,
// task tables, forward permutation).
// FB / IB > 0: specialised on the transform sizes (everything inlined and static); 0: any size.
// Tables of the lane-pair DFT kernels, staged once per CTA: twiddle pyramids, task tables, forward permutation.
__device__ __forceinline__ PkTables pk_stage_tables(const DftPkParams &pp)
{
  float *pf = reinterpret_cast<float *>(rr_smem_raw + pp.lay_pyr_f), *pi = reinterpret_cast<float *>(rr_smem_raw + pp.lay_pyr_i);
  uint16_t *tf = reinterpret_cast<uint16_t *>(rr_smem_raw + pp.lay_ltab_f), *ti = reinterpret_cast<uint16_t *>(rr_smem_raw + pp.lay_ltab_i);
  uint16_t *pm = reinterpret_cast<uint16_t *>(rr_smem_raw + pp.lay_perm_f);
  const int nf = pp.n_pyr_f, ni = pp.n_pyr_i, ef = pp.n_ltab_f, ei = pp.n_ltab_i, mf = 1 << pp.fb;
  for (int i = threadIdx.x; i < nf; i += blockDim.x) pf[i] = pp.base.pyr_f[i];
  for (int i = threadIdx.x; i < ni; i += blockDim.x) pi[i] = pp.base.pyr_i[i];
  for (int i = threadIdx.x; i < ef; i += blockDim.x) tf[i] = pp.ltab_f[i];
  for (int i = threadIdx.x; i < ei; i += blockDim.x) ti[i] = pp.ltab_i[i];
  for (int i = threadIdx.x; i < mf; i += blockDim.x) pm[i] = pp.perm_f[i];
  return PkTables{pf, pi, tf, ti, pm};
}

// GROUPS == kPkMaxGroups: every group has a forward and an inverse buffer. GROUPS == kPkInplaceGroups: one buffer per
// group (the spectrum phase goes through registers, pk_spectrum_inplace), which lets a fifth group -- four more
// warps to hide shared-memory latency with -- share the SM at 102 registers per thread.
template <int MODE, int FB, int IB, bool STEREO, int GROUPS = kPkMaxGroups>
__global__ void __launch_bounds__(kPkGroupThreads * GROUPS) dftp_kernel(const __grid_constant__ DftPkParams pp, long long nwork)
{
  constexpr bool kInplace = GROUPS != kPkMaxGroups;
  const PkTables tb = pk_stage_tables(pp);
  const int gi = threadIdx.x / kPkGroupThreads;          // pp.gthreads == kPkGroupThreads: a literal keeps it out of registers
  const Grp g{(int)threadIdx.x % kPkGroupThreads, kPkGroupThreads, 1 + gi};
  CPk *F = reinterpret_cast<CPk *>(rr_smem_raw + pp.lay_data) + (size_t)gi * (pp.fslots + pp.bslots), *B = F + pp.fslots;   // in place: fslots == 0
  __shared__ PkItem items[GROUPS][2];
  __syncthreads();
  // 32-bit work counters (the host splits launches of more than 2^30 items)
  int w = (int)blockIdx.x * pp.groups + gi;
  const int stride = (int)gridDim.x * pp.groups, nw = (int)nwork;
  if (w < nw) {
    if (g.tid == 0) items[gi][0] = pk_make_item(pp, w);
  }
  for (int n = 0; w < nw; w += stride, n ^= 1) {
    const int next = w + stride < nw ? w + stride : -1;
    dftp_program<MODE, FB, IB, STEREO, kInplace>(pp, g, tb, items[gi], n, next, F, B);
  }
}

// DFT stage + vpoly0 in one kernel (rate_kernels_fused.cuh): a work item is a run of consecutive blocks of a lane pair.
template <int MODE, int FB, int IB, int NT, int DLO>
__global__ void __launch_bounds__(kPkGroupThreads * kPkMaxGroups) dft_poly_kernel(const __grid_constant__ DftPolyParams fp, long long nwork)
{
  const DftPkParams &pp = fp.dft;
  const PkTables tb = pk_stage_tables(pp);
  const int gi = threadIdx.x / kPkGroupThreads;
  const Grp g{(int)threadIdx.x % kPkGroupThreads, kPkGroupThreads, 1 + gi};
  CPk *F = reinterpret_cast<CPk *>(rr_smem_raw + pp.lay_data) + (size_t)gi * (pp.fslots + pp.halo_slots + pp.bslots);
  CPk *B = F + pp.fslots + pp.halo_slots;
  __shared__ PkFusedBlock blk[kPkMaxGroups][2];
  __syncthreads();
  const int stride = (int)gridDim.x * pp.groups, nw = (int)nwork;
  for (int w = (int)blockIdx.x * pp.groups + gi; w < nw; w += stride) {
    long long b_first; int count, pair, halo_first;
    if (!pk_fused_run(fp, w, &b_first, &count, &pair, &halo_first)) continue;
    grp_sync(g);                                          // the previous run is done with blk[] and with F
    if (g.tid == 0) blk[gi][0] = pk_fused_block(fp, pair, b_first, halo_first != 0);
    grp_sync(g);
    pk_tile_now<FB, true>(pp, g, blk[gi][0].it, F, tb.perm_f);
    for (int k = 0, slot = 0; k < count; ++k, slot ^= 1)
      dft_poly_program<MODE, FB, IB, NT, DLO>(fp, g, tb, blk[gi], slot, pair, k + 1 < count ? b_first + k + 1 : -1, k == 0, F, B);
  }
}
template <class T, class InT, class OutT>
__global__ void __launch_bounds__(kTileThreads) poly0_kernel(const __grid_constant__ PolyParams<T> p, long long nwork)
{
  T *smem = reinterpret_cast<T *>(rr_smem_raw);
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) poly0_program<T, InT, OutT>(p, w, smem);
}
// Persistent CTA with a two-deep LDGSTS pipeline: the input windows of tile k+1 are in flight while
// tile k is computed (when the input needs a type conversion the copy is synchronous instead).
template <class T, class InT, class OutT, int NT>
__global__ void __launch_bounds__(512) poly0_fast_kernel(const __grid_constant__ Poly0FastParams<T> p, long long nwork)
{
  T *smem = reinterpret_cast<T *>(rr_smem_raw);
  constexpr bool kAsync = std::is_same<InT, T>::value;
  const int set = p.win * p.CH;
  long long w = blockIdx.x;
  __shared__ Poly0Tile tiles[3];                     // tile geometry is computed by one thread per tile
  if (kAsync && p.double_buffer) {
    if (threadIdx.x == 0) {
      if (w < nwork) tiles[0] = poly0_tile(p, w);
      if (w + gridDim.x < nwork) tiles[1] = poly0_tile(p, w + gridDim.x);
    }
    __syncthreads();
    if (w < nwork) poly0_fast_load<T, InT, true>(p, tiles[0], smem);
    for (int it = 0; w < nwork; w += gridDim.x, ++it) {
      const int cur = it & 1, ts = it % 3, tn = (it + 1) % 3, tnn = (it + 2) % 3;
      const long long next = w + gridDim.x;
      if (next < nwork) { poly0_fast_load<T, InT, true>(p, tiles[tn], smem + (cur ^ 1) * set); async_copy_wait<1>(); }
      else async_copy_wait<0>();
      // slot tnn was the current tile of the previous iteration (all reads done before its last barrier)
      if (threadIdx.x == 0 && next + gridDim.x < nwork) tiles[tnn] = poly0_tile(p, next + gridDim.x);
      __syncthreads();
      const Poly0Tile t = tiles[ts];
      poly0_fast_compute<T, OutT, NT>(p, t, smem + cur * set);                 // ends with a barrier
    }
  } else {
    for (; w < nwork; w += gridDim.x) {
      const Poly0Tile t = poly0_tile(p, w);
      poly0_fast_load<T, InT, false>(p, t, smem);
      __syncthreads();
      poly0_fast_compute<T, OutT, NT>(p, t, smem);
    }
  }
}
// vpoly0 for lane pairs (rate_kernels_pk.cuh): persistent CTA with one window buffer -- with several CTAs per SM the
// others cover this one's load; the slots of a period are dealt to the threads once per CTA. The window of a tile is
// staged by one TMA bulk copy per lane pair (completion on an mbarrier) when it is a contiguous 16-byte aligned range
// of a pair-interleaved FIFO, else with LDGSTS copies by all threads.
#define RR_POLY0_PAIR_KERNEL_BODY(SETUP, TILE)                                                                        \
  const Poly0FastParams<float> &p = pp.fast;                                                                          \
  Pk *smem = reinterpret_cast<Pk *>(rr_smem_raw) + 2;  /* 16 bytes of slack in front: a shifted window starts at -1 */  \
  const int set = p.win * pp.P;                        /* the windows, then the slot table */                          \
  uint16_t *slot_of = reinterpret_cast<uint16_t *>(smem + set);                                                       \
  __shared__ Poly0PairTile tiles[2];                                                                                  \
  __shared__ int cnt[17];                                                                                             \
  __shared__ uint16_t ovf[kPolyDealOverflow];                                                                         \
  __shared__ __align__(8) unsigned long long bar;                                                                     \
  const int tid = threadIdx.x, nt = blockDim.x;                                                                       \
  long long w = blockIdx.x;                                                                                           \
  if (w >= nwork) return;                                                                                             \
  for (int i = tid; i < pp.tslots; i += nt) slot_of[i] = 0xffff;                                                      \
  if (tid < 17) cnt[tid] = 0;                                                                                         \
  if (tid == 0) { tma_bar_init(&bar, 1); tiles[0] = poly0_pair_make_tile(pp, w); }                                    \
  __syncthreads();                                                                                                    \
  if (pp.spread) poly0_pair_deal(pp, tiles[0].t, slot_of, cnt, ovf, tid, nt); /* one column: holds for every tile */  \
  poly0_pair_load(pp, tiles[0].t, tiles[0].tma, tiles[0].head, smem, &bar, tid, nt);                                  \
  __syncthreads();                                                                                                    \
  if (pp.spread) { poly0_pair_deal_overflow(pp, tiles[0].t, slot_of, cnt, ovf, tid); __syncthreads(); }               \
  const auto st = SETUP;                                                                                              \
  unsigned phase = 0;                                                                                                 \
  for (int it = 0; w < nwork; w += gridDim.x, ++it) {                                                                 \
    const int ts = it & 1;                                                                                            \
    const long long next = w + gridDim.x;                                                                             \
    if (tid == 0 && next < nwork) tiles[ts ^ 1] = poly0_pair_make_tile(pp, next);                                     \
    if (tiles[ts].tma) { tma_bar_wait(&bar, phase); phase ^= 1; } else async_copy_wait<0>();                          \
    __syncthreads();                                                                                                  \
    TILE;                                                                                                             \
    __syncthreads();                                                                                                  \
    if (next < nwork) poly0_pair_load(pp, tiles[ts ^ 1].t, tiles[ts ^ 1].tma, tiles[ts ^ 1].head, smem, &bar, tid, nt); \
  }

template <int NT>
__global__ void __launch_bounds__(512, 2) poly0_pair_kernel(const __grid_constant__ Poly0PairParams pp, long long nwork)
{
  RR_POLY0_PAIR_KERNEL_BODY((poly0_pair_setup<NT>(pp, tiles[0].t, slot_of, tid)), (poly0_pair_tile<NT>(pp, tiles[ts], smem, st)))
}
// Two adjacent slots per thread (poly0_pair2_*): same skeleton.
template <int NT, int DLO>
__global__ void __launch_bounds__(256, 3) poly0_pair2_kernel(const __grid_constant__ Poly0PairParams pp, long long nwork)
{
  RR_POLY0_PAIR_KERNEL_BODY((poly0_pair2_setup<NT, DLO>(pp, tiles[0].t, slot_of, tid)), (poly0_pair2_tile<NT, DLO>(pp, tiles[ts], smem, st)))
}
#undef RR_POLY0_PAIR_KERNEL_BODY
template <class T, class InT, class OutT>
__global__ void __launch_bounds__(kTileThreads) polyN_kernel(const __grid_constant__ PolyParams<T> p, long long nwork)
{
  T *smem = reinterpret_cast<T *>(rr_smem_raw);
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) polyN_program<T, InT, OutT>(p, w, smem);
}
// vpoly0 with two slots per thread for scalar (fp64) lanes: persistent CTA, one window buffer staged by TMA bulk copies
// (the other CTAs of the SM cover the load), coefficient rows in registers for the whole launch.
template <class T, class OutT, int NT, int DLO>
__global__ void __launch_bounds__(256, 2) poly0_dual_kernel(const __grid_constant__ Poly0DualParams<T> dp, long long nwork)
{
  T *smem = reinterpret_cast<T *>(rr_smem_raw);
  __shared__ Poly0DualTile tiles[2];
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, nt = blockDim.x;
  long long w = blockIdx.x;
  if (w >= nwork) return;
  if (tid == 0) { tma_bar_init(&bar, 1); tiles[0] = poly0_dual_tile(dp, w); }
  __syncthreads();
  poly0_dual_load(dp, tiles[0], smem, &bar, tid, nt);
  const Poly0DualThread<T, OutT, NT, DLO> st = poly0_dual_setup<T, OutT, NT, DLO>(dp, tid);
  unsigned phase = 0;
  for (int it = 0; w < nwork; w += gridDim.x, ++it) {
    const int ts = it & 1;
    const long long next = w + gridDim.x;
    if (tid == 0 && next < nwork) tiles[ts ^ 1] = poly0_dual_tile(dp, next);
    if (tiles[ts].tma) { tma_bar_wait(&bar, phase); phase ^= 1; }
    __syncthreads();
    poly0_dual_compute<T, OutT, NT, DLO>(dp, tiles[ts], smem, st);
    __syncthreads();
    if (next < nwork) poly0_dual_load(dp, tiles[ts ^ 1], smem, &bar, tid, nt);
  }
}
// fp64 engine: eight outputs per thread from padded rows (kHalfOpt, rate_kernels.cuh)
template <class T> struct HalfOpt { static constexpr int value = sizeof(T) == 8 ? 8 : 4; };
template <class T, class InT, class OutT, int NC>
__global__ void __launch_bounds__(kTileThreads, 2) halfband_kernel(const __grid_constant__ HalfbandParams<T> p, long long nwork)
{
  T *smem = reinterpret_cast<T *>(rr_smem_raw);
  if constexpr (sizeof(T) == 8) {
    // fp64 engine: the global loads of the CTA's next tile are in flight (in registers) while the current one is computed
    constexpr int OPT = HalfOpt<T>::value;
    __shared__ HbTile<InT, OutT> tiles[2];
    const int tid = threadIdx.x, nthreads = blockDim.x, items = (p.CH * p.tile) / OPT;
    long long w = blockIdx.x;
    if (tid == 0 && w < nwork) tiles[0] = hb_make_tile<T, InT, OutT, NC>(p, w);
    __syncthreads();
    HbRegs<InT> r;
    if (w < nwork) hb_fetch(p, tiles[0], r, tid, nthreads);
    for (int it = 0; w < nwork; w += gridDim.x, it ^= 1) {
      const long long next = w + gridDim.x;
      if (tid == 0 && next < nwork) tiles[it ^ 1] = hb_make_tile<T, InT, OutT, NC>(p, next);
      hb_put<T, InT, OutT, NC, OPT>(p, tiles[it], r, smem, tid, nthreads);
      __syncthreads();
      if (next < nwork) hb_fetch(p, tiles[it ^ 1], r, tid, nthreads);
      for (int i = tid; i < items; i += nthreads) hb_compute_item<T, InT, OutT, NC, OPT>(p, tiles[it], smem, i);
      __syncthreads();
    }
  } else {
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) halfband_program<T, InT, OutT, NC, HalfOpt<T>::value>(p, w, smem);
  }
}
template <int NC>
__global__ void __launch_bounds__(kTileThreads) halfband_pair_kernel(const __grid_constant__ HalfbandPairParams p, long long nwork)
{
  Pk *smem = reinterpret_cast<Pk *>(rr_smem_raw);
  float cf[NC];                                          // coefficients in registers for the whole launch
#pragma unroll
  for (int t = 0; t < NC; ++t) cf[t] = p.base.coef[t];
  // two window buffers: the next tile's LDGSTS copies are in flight while this one is computed; tile geometry
  // comes from one thread per tile
  const int set = 2 * p.base.half * p.G + 8;
  __shared__ HalfbandPairTile tiles[3];
  long long w = blockIdx.x;
  if (w >= nwork) return;
  if (threadIdx.x == 0) {
    tiles[0] = halfband_pair_tile<NC>(p, w);
    if (w + gridDim.x < nwork) tiles[1] = halfband_pair_tile<NC>(p, w + gridDim.x);
  }
  __syncthreads();
  halfband_pair_load<NC>(p, tiles[0], smem, threadIdx.x, blockDim.x);
  for (int it = 0; w < nwork; w += gridDim.x, ++it) {
    const int ts = it % 3, tn = (it + 1) % 3, tnn = (it + 2) % 3;
    const long long next = w + gridDim.x;
    if (next < nwork) { halfband_pair_load<NC>(p, tiles[tn], smem + ((it + 1) & 1) * set, threadIdx.x, blockDim.x); async_copy_wait<1>(); }
    else async_copy_wait<0>();
    if (threadIdx.x == 0 && next + gridDim.x < nwork) tiles[tnn] = halfband_pair_tile<NC>(p, next + gridDim.x);
    __syncthreads();
    halfband_pair_compute<NC>(p, cf, tiles[ts], smem + (it & 1) * set, threadIdx.x, blockDim.x);
    __syncthreads();
  }
}
template <class InT, class OutT>
__global__ void __launch_bounds__(kTileThreads) copy_kernel(const __grid_constant__ CopyParams p, long long nwork)
{
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) copy_program<InT, OutT>(p, w);
}

// Launch bookkeeping shared by every handle of the process: distinct handles may be driven from distinct host
// threads (the reference's threading contract, rate/rate_uni.c:210), so the cache is guarded by a mutex. The
// opt-in dynamic shared-memory limit of a kernel is raised ONCE per (device, kernel) to the device maximum and
// never lowered, so launches of one kernel with different shared-memory sizes cannot disturb each other.
struct LaunchKey {
  int device; const void *fn; int threads; size_t smem;
  bool operator<(const LaunchKey &o) const
  {
    if (device != o.device) return device < o.device;
    if (fn != o.fn) return fn < o.fn;
    if (threads != o.threads) return threads < o.threads;
    return smem < o.smem;
  }
};
struct LaunchCache {
  std::mutex mu;
  std::map<std::pair<int, const void *>, int> armed;   // (device, kernel) -> largest dynamic shared memory it may use
  std::map<LaunchKey, int> blocks_per_sm;
  std::map<int, int> sms;                               // device -> SM count
};
static LaunchCache &launch_cache()
{
  static LaunchCache c;
  return c;
}

// Resident CTAs of `kernel` on the current device for (threads, smem), and the device's SM count.
template <class Kernel>
static int launch_geometry(Kernel kernel, int threads, size_t smem, long long *resident)
{
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  const void *fn = reinterpret_cast<const void *>(kernel);
  LaunchCache &c = launch_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto sm = c.sms.find(dev);
  if (sm == c.sms.end()) {
    int n = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sm = c.sms.emplace(dev, n).first;
  }
  auto arm = c.armed.find(std::make_pair(dev, fn));
  if (arm == c.armed.end()) {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
    int optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int max_dyn = optin - static_cast<int>(fa.sharedSizeBytes);
    if (max_dyn > 48 * 1024 - static_cast<int>(fa.sharedSizeBytes))
      CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    arm = c.armed.emplace(std::make_pair(dev, fn), max_dyn).first;
  }
  if (static_cast<long long>(smem) > arm->second) { set_last_error("kernel needs more shared memory than an SM has"); return RR_INTERNAL; }
  const LaunchKey key{dev, fn, threads, smem};
  auto it = c.blocks_per_sm.find(key);
  if (it == c.blocks_per_sm.end()) {
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) { set_last_error("kernel does not fit on an SM"); return RR_INTERNAL; }
    it = c.blocks_per_sm.emplace(key, occ).first;
  }
  *resident = static_cast<long long>(it->second) * sm->second;
  return RR_OK;
}

template <class Kernel, class Params, class... Extra>
static int launch_persistent(Kernel kernel, const Params &p, long long nwork, int threads, size_t smem, stream_t s,
                             Extra... extra)
{
  if (nwork <= 0) return RR_OK;
  long long resident = 0;
  const int rc = launch_geometry(kernel, threads, smem, &resident);
  if (rc != RR_OK) return rc;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(nwork, resident));
  kernel<<<grid, threads, smem, s>>>(p, nwork, extra...);
  CUDA_TRY(cudaGetLastError());
  return RR_OK;
}
#endif  // !B200RATE_EMU

static int launch_dftp(const DftPkParams &pp, long long nwork, stream_t s)
{
  if (nwork <= 0) return RR_OK;
#ifdef B200RATE_EMU
  (void)s;
  std::vector<CPk> mem(static_cast<size_t>(pp.fslots + pp.bslots) + 1);
  CPk *F = mem.data(), *B = F + pp.fslots;
  const Grp g{0, 1, 0};
  const PkTables tb{pp.base.pyr_f, pp.base.pyr_i, pp.ltab_f, pp.ltab_i, pp.perm_f};
  for (long long w = 0; w < nwork; ++w) {
    PkItem items[2];
    items[0] = pk_make_item(pp, w);
    if (pp.fslots == 0 && pp.stereo) dftp_program<PK_SPEC_UP2, 10, 11, true, true>(pp, g, tb, items, 0, -1, F, B);   // one buffer per group
    else if (pp.fslots == 0) dftp_program<PK_SPEC_UP2, 10, 11, false, true>(pp, g, tb, items, 0, -1, F, B);
    else if (pp.spec_mode == PK_SPEC_UP2 && pp.stereo) dftp_program<PK_SPEC_UP2, 0, 0, true>(pp, g, tb, items, 0, -1, F, B);
    else if (pp.spec_mode == PK_SPEC_UP2) dftp_program<PK_SPEC_UP2, 0, 0, false>(pp, g, tb, items, 0, -1, F, B);
    else if (pp.spec_mode == PK_SPEC_SAME && pp.stereo) dftp_program<PK_SPEC_SAME, 0, 0, true>(pp, g, tb, items, 0, -1, F, B);
    else if (pp.spec_mode == PK_SPEC_SAME) dftp_program<PK_SPEC_SAME, 0, 0, false>(pp, g, tb, items, 0, -1, F, B);
    else if (pp.stereo) dftp_program<PK_SPEC_GEN, 0, 0, true>(pp, g, tb, items, 0, -1, F, B);
    else dftp_program<PK_SPEC_GEN, 0, 0, false>(pp, g, tb, items, 0, -1, F, B);
  }
  return RR_OK;
#else
  const size_t smem = pk_smem_layout(pp).total;
  const int threads = pp.groups * pp.gthreads;
  auto go = [&](auto kernel) -> int {
    long long resident = 0;
    const int rc = launch_geometry(kernel, threads, smem, &resident);
    if (rc != RR_OK) return rc;
    const long long ctas = (nwork + pp.groups - 1) / pp.groups;
    kernel<<<static_cast<unsigned>(std::min(ctas, resident)), threads, smem, s>>>(pp, nwork);
    CUDA_TRY(cudaGetLastError());
    return RR_OK;
  };
  // the common Best-quality sizes get kernels specialised on (forward, inverse) transform size; every kernel exists for
  // adjacent stereo frames as input (the caller's buffer of a stereo stream, a pair-interleaved FIFO) and for any layout
#define RR_DFTP(MODE, FBV, IBV) (pp.stereo ? go(dftp_kernel<MODE, FBV, IBV, true>) : go(dftp_kernel<MODE, FBV, IBV, false>))
#define RR_DFTP_G(MODE, FBV, IBV, G) (pp.stereo ? go(dftp_kernel<MODE, FBV, IBV, true, G>) : go(dftp_kernel<MODE, FBV, IBV, false, G>))
  if (pp.spec_mode == PK_SPEC_UP2) {
    if (pp.fb == 10 && pp.ib == 11 && pp.fslots == 0 && pp.groups == 5) return RR_DFTP_G(PK_SPEC_UP2, 10, 11, 5);   // one buffer per group
    if (pp.fb == 10 && pp.ib == 11 && pp.fslots == 0 && pp.groups == 6) return RR_DFTP_G(PK_SPEC_UP2, 10, 11, 6);
    if (pp.fb == 10 && pp.ib == 11) return RR_DFTP(PK_SPEC_UP2, 10, 11);      // N = 4096, x2 (44.1 <-> 48 family)
    return RR_DFTP(PK_SPEC_UP2, 0, 0);
  }
  if (pp.spec_mode == PK_SPEC_SAME) {
    if (pp.fb == 11 && pp.ib == 11) return RR_DFTP(PK_SPEC_SAME, 11, 11);     // N = 4096, 1:1 pre-filter
    return RR_DFTP(PK_SPEC_SAME, 0, 0);
  }
  if (pp.fb == 11 && pp.ib == 10) return RR_DFTP(PK_SPEC_GEN, 11, 10);        // N = 4096, F-domain / 2
  return RR_DFTP(PK_SPEC_GEN, 0, 0);
#undef RR_DFTP
#undef RR_DFTP_G
#endif
}

// fp64 DFT stage: kernel and launcher live in dft64.cu (rate_kernels_f64.cuh)
int launch_dft64(const Dft64Params &dp, long long nwork, void *stream);

// nwork = lane pairs x runs per pair
static int launch_dft_poly(const DftPolyParams &fp, long long nwork, stream_t s)
{
  if (nwork <= 0) return RR_OK;
  const DftPkParams &pp = fp.dft;
  const int dlo = static_cast<int>(fp.step / fp.L);
#ifdef B200RATE_EMU
  (void)s;
  std::vector<CPk> mem(static_cast<size_t>(pp.fslots + pp.halo_slots + pp.bslots) + 1);
  CPk *F = mem.data(), *B = F + pp.fslots + pp.halo_slots;
  const Grp g{0, 1, 0};
  const PkTables tb{pp.base.pyr_f, pp.base.pyr_i, pp.ltab_f, pp.ltab_i, pp.perm_f};
  for (long long w = 0; w < nwork; ++w) {
    long long b_first; int count, pair, halo_first;
    if (!pk_fused_run(fp, w, &b_first, &count, &pair, &halo_first)) continue;
    PkFusedBlock blk[2];
    blk[0] = pk_fused_block(fp, pair, b_first, halo_first != 0);
    pk_tile_now<10, true>(pp, g, blk[0].it, F, tb.perm_f);
    for (int k = 0, slot = 0; k < count; ++k, slot ^= 1) {
      const long long nb = k + 1 < count ? b_first + k + 1 : -1;
#define RR_FP(D) dft_poly_program<PK_SPEC_UP2, 10, 11, 24, D>(fp, g, tb, blk, slot, pair, nb, k == 0, F, B)
      if (dlo == 0) RR_FP(0); else if (dlo == 1) RR_FP(1); else RR_FP(2);
#undef RR_FP
    }
  }
  return RR_OK;
#else
  const size_t smem = pk_smem_layout(pp).total;
  const int threads = pp.groups * pp.gthreads;
  auto go = [&](auto kernel) -> int {
    long long resident = 0;
    const int rc = launch_geometry(kernel, threads, smem, &resident);
    if (rc != RR_OK) return rc;
    const long long ctas = (nwork + pp.groups - 1) / pp.groups;
    kernel<<<static_cast<unsigned>(std::min(ctas, resident)), threads, smem, s>>>(fp, nwork);
    CUDA_TRY(cudaGetLastError());
    return RR_OK;
  };
  if (dlo == 0) return go(dft_poly_kernel<PK_SPEC_UP2, 10, 11, 24, 0>);
  if (dlo == 1) return go(dft_poly_kernel<PK_SPEC_UP2, 10, 11, 24, 1>);
  return go(dft_poly_kernel<PK_SPEC_UP2, 10, 11, 24, 2>);
#endif
}

static int launch_halfband_pair(const HalfbandPairParams &hp, long long nwork, size_t smem, stream_t s)
{
  if (nwork <= 0) return RR_OK;
#ifdef B200RATE_EMU
  (void)s;
  std::vector<Pk> mem(smem / sizeof(Pk) + 2);
#define RR_HBP(NC) { float cf[NC]; for (int t = 0; t < NC; ++t) cf[t] = hp.base.coef[t]; \
                     for (long long w = 0; w < nwork; ++w) { const HalfbandPairTile t = halfband_pair_tile<NC>(hp, w); \
                                                             halfband_pair_load<NC>(hp, t, mem.data(), 0, 1); \
                                                             halfband_pair_compute<NC>(hp, cf, t, mem.data(), 0, 1); } return RR_OK; }
#else
#define RR_HBP(NC) return launch_persistent(halfband_pair_kernel<NC>, hp, nwork, kTileThreads, smem, s)
#endif
  switch (hp.base.ncoef) {
    case 8: RR_HBP(8);
    case 9: RR_HBP(9);
    case 10: RR_HBP(10);
    case 11: RR_HBP(11);
    case 12: RR_HBP(12);
    case 13: RR_HBP(13);
    default: return RR_INTERNAL;
  }
#undef RR_HBP
}

static size_t poly0_pair_smem(const Poly0PairParams &pp)
{
  return sizeof(Pk) * (static_cast<size_t>(pp.fast.win) * pp.P + 2) + 2 * static_cast<size_t>(pp.tslots) + 16;
}
static int launch_poly0_pair(const Poly0PairParams &pp, int threads, long long nwork, stream_t s)
{
  if (nwork <= 0) return RR_OK;
#ifdef B200RATE_EMU
  (void)s; (void)threads;
  if (getenv("B200RATE_TRACE")) fprintf(stderr, "poly0pair L %d n %d step %lld P %d PG %d CL %d tslots %d spread %d shift %d MM %d win %d nwork %lld\n", pp.fast.base.L, pp.fast.base.n, pp.fast.base.step, pp.P, pp.PG, pp.CL, pp.tslots, pp.spread, pp.shift, pp.fast.MM, pp.fast.win, nwork);
  std::vector<Pk> bufv(static_cast<size_t>(pp.fast.win) * pp.P + 8);
  struct { Pk *p; Pk *data() const { return p; } } buf = {bufv.data() + 2};   // slack either side, like the kernel's buffer
  std::vector<uint16_t> slot_of(static_cast<size_t>(pp.tslots) + 1);
  for (long long w = 0; w < nwork; ++w) {
    const Poly0PairTile pt = poly0_pair_make_tile(pp, w);
    const Poly0Tile &t = pt.t;
    int cnt[17] = {0};
    uint16_t ovf[kPolyDealOverflow];
    std::fill(slot_of.begin(), slot_of.end(), static_cast<uint16_t>(0xffff));
    if (pp.spread) { poly0_pair_deal(pp, t, slot_of.data(), cnt, ovf, 0, 1); poly0_pair_deal_overflow(pp, t, slot_of.data(), cnt, ovf, 0); }
    poly0_pair_load(pp, t, pt.tma, pt.head, buf.data(), nullptr, 0, 1);
    if (pp.CL == 2) {
      const int dlo = static_cast<int>(pp.fast.base.step / pp.fast.base.L);
      for (int th = 0; th < pp.tslots * pp.P; ++th) {
#define RR_P2(NT, D) if (pp.fast.base.n == NT && dlo == D) poly0_pair2_tile<NT, D>(pp, pt, buf.data(), poly0_pair2_setup<NT, D>(pp, t, slot_of.data(), th))
        RR_P2(16, 0); RR_P2(16, 1); RR_P2(16, 2); RR_P2(24, 0); RR_P2(24, 1); RR_P2(24, 2); RR_P2(28, 0); RR_P2(28, 1); RR_P2(28, 2);
        RR_P2(32, 0); RR_P2(32, 1); RR_P2(32, 2);
#undef RR_P2
      }
      continue;
    }
    for (int th = 0; th < pp.tslots * pp.P * pp.PG; ++th) {
      if (pp.fast.base.n == 16) poly0_pair_tile<16>(pp, pt, buf.data(), poly0_pair_setup<16>(pp, t, slot_of.data(), th));
      else if (pp.fast.base.n == 24) poly0_pair_tile<24>(pp, pt, buf.data(), poly0_pair_setup<24>(pp, t, slot_of.data(), th));
      else if (pp.fast.base.n == 28) poly0_pair_tile<28>(pp, pt, buf.data(), poly0_pair_setup<28>(pp, t, slot_of.data(), th));
      else poly0_pair_tile<32>(pp, pt, buf.data(), poly0_pair_setup<32>(pp, t, slot_of.data(), th));
    }
  }
  return RR_OK;
#else
  const size_t smem = poly0_pair_smem(pp);
  if (pp.CL == 2) {
    const int dlo = static_cast<int>(pp.fast.base.step / pp.fast.base.L);
#define RR_P2(NT, D) if (pp.fast.base.n == NT && dlo == D) return launch_persistent(poly0_pair2_kernel<NT, D>, pp, nwork, threads, smem, s)
    RR_P2(16, 0); RR_P2(16, 1); RR_P2(16, 2); RR_P2(24, 0); RR_P2(24, 1); RR_P2(24, 2); RR_P2(28, 0); RR_P2(28, 1); RR_P2(28, 2);
        RR_P2(32, 0); RR_P2(32, 1); RR_P2(32, 2);
#undef RR_P2
    return RR_INTERNAL;
  }
  if (pp.fast.base.n == 16) return launch_persistent(poly0_pair_kernel<16>, pp, nwork, threads, smem, s);
  if (pp.fast.base.n == 24) return launch_persistent(poly0_pair_kernel<24>, pp, nwork, threads, smem, s);
  if (pp.fast.base.n == 28) return launch_persistent(poly0_pair_kernel<28>, pp, nwork, threads, smem, s);   // Best quality, bandwidth >= 97 %
  return launch_persistent(poly0_pair_kernel<32>, pp, nwork, threads, smem, s);
#endif
}

// Typed dispatch. For the fp32 engine every buffer is float; the fp64 engine reads float at the caller
// boundary and double in between.
template <class T> struct Launch {
  static constexpr bool kIsF32 = std::is_same<T, float>::value;

#ifdef B200RATE_EMU
  template <class F> static int serial(long long nwork, size_t smem_bytes, F body)
  {
    std::vector<unsigned char> smem(smem_bytes + 64);
    for (long long w = 0; w < nwork; ++w) body(w, reinterpret_cast<T *>(smem.data()));
    return RR_OK;
  }
#endif

#define RR_DISPATCH_IO(CALL)                                                       \
  do {                                                                             \
    if (kIsF32 || (!in_f32 && !out_f32)) return CALL(T, T);                        \
    if (in_f32 && out_f32) return CALL(float, float);                              \
    if (in_f32) return CALL(float, T);                                             \
    return CALL(T, float);                                                         \
  } while (0)

  // lpc: lanes per CTA (1 or 2); data_bytes: shared memory for the sample buffers, the tables follow
  // big != nullptr: the block does not fit shared memory; *big is a global scratch area of big_ctas slices
  static int dft(const DftParams<T> &p, int lpc, bool in_f32, bool out_f32, long long nwork, size_t data_bytes, stream_t s,
                 void *big = nullptr, int big_ctas = 0)
  {
    const size_t smem = data_bytes + sizeof(T) * static_cast<size_t>(dft_table_elems(p));
#ifndef B200RATE_EMU
    if (big) {
      if (nwork <= 0) return RR_OK;
      const unsigned long long per_cta = data_bytes / sizeof(C2<T>);
      const unsigned grid = static_cast<unsigned>(std::min<long long>(nwork, big_ctas));
#define RR_CALLB(I, O) (dft_big_kernel<T, I, O><<<grid, 2 * kDftThreads, 0, s>>>(p, nwork, static_cast<C2<T> *>(big), per_cta), \
                        cudaGetLastError() == cudaSuccess ? RR_OK : RR_INTERNAL)
      RR_DISPATCH_IO(RR_CALLB);
#undef RR_CALLB
    }
#else
    (void)big; (void)big_ctas;
#endif
#ifdef B200RATE_EMU
    const DftTables<T> tab{p.pyr_f, p.pyr_i, p.tcos_f, p.tcos_i};
    const CoefCache<T, 0> cc{};
    // emulation: the "asynchronous" prefetch of each item's tile is performed right before the item
#define RR_CALLE(I, O, LPC)                                                                                   \
  serial(nwork, smem, [&](long long w, T *sm) {                                                               \
    C2<T> *data = reinterpret_cast<C2<T> *>(sm);                                                              \
    DftItem<T> items[2];                                                                                      \
    items[0] = dft_item<T, LPC>(p, w);                                                                        \
    if (p.zstride > 0) dft_stage_tile<T, T, LPC, true>(p, items[0], data + LPC * (p.xstride + p.ystride), p.zstride); \
    dft_stage_program<T, I, O, LPC, 0>(p, tab, cc, items, 0, -1, data);                                       \
  })
#define RR_CALL1(I, O) RR_CALLE(I, O, 1)
#define RR_CALL2(I, O) RR_CALLE(I, O, 2)
#else
    // a block that leaves room for only one CTA per SM gets twice the threads (same register budget per SM)
    const int threads = smem > 113 * 1024 ? 2 * kDftThreads : kDftThreads;
#define RR_CALL1(I, O) launch_persistent(dft_kernel<T, I, O, 1>, p, nwork, threads, smem, s, static_cast<int>(data_bytes))
#define RR_CALL2(I, O) launch_persistent(dft_kernel<T, I, O, 2>, p, nwork, threads, smem, s, static_cast<int>(data_bytes))
#endif
    (void)s;
    if (lpc == 2) RR_DISPATCH_IO(RR_CALL2);
    RR_DISPATCH_IO(RR_CALL1);
#undef RR_CALL1
#undef RR_CALL2
#undef RR_CALLE
  }
  static int poly0(const PolyParams<T> &p, bool in_f32, bool out_f32, long long nwork, size_t smem, stream_t s)
  {
#ifdef B200RATE_EMU
#define RR_CALL(I, O) serial(nwork, smem, [&](long long w, T *sm) { poly0_program<T, I, O>(p, w, sm); })
#else
#define RR_CALL(I, O) launch_persistent(poly0_kernel<T, I, O>, p, nwork, kTileThreads, smem, s)
#endif
    (void)s;
    RR_DISPATCH_IO(RR_CALL);
#undef RR_CALL
  }
  static int poly0_fast(const Poly0FastParams<T> &p, int threads, bool in_f32, bool out_f32, long long nwork, size_t smem,
                        stream_t s)
  {
    (void)s; (void)threads;
#ifdef B200RATE_EMU
#define RR_CALLN(I, O, NT)                                                   \
  serial(nwork, smem, [&](long long w, T *sm) {                              \
    const Poly0Tile t = poly0_tile(p, w);                                    \
    poly0_fast_load<T, I, false>(p, t, sm);                                  \
    poly0_fast_compute<T, O, NT>(p, t, sm);                                  \
  })
#else
#define RR_CALLN(I, O, NT) launch_persistent(poly0_fast_kernel<T, I, O, NT>, p, nwork, threads, smem, s)
#endif
#define RR_CALL16(I, O) RR_CALLN(I, O, 16)
#define RR_CALL24(I, O) RR_CALLN(I, O, 24)
#define RR_CALL28(I, O) RR_CALLN(I, O, 28)
#define RR_CALL32(I, O) RR_CALLN(I, O, 32)
    if (p.base.n == 16) RR_DISPATCH_IO(RR_CALL16);
    if (p.base.n == 24) RR_DISPATCH_IO(RR_CALL24);
    if (p.base.n == 28) RR_DISPATCH_IO(RR_CALL28);
    if (p.base.n == 32) RR_DISPATCH_IO(RR_CALL32);
    return RR_INTERNAL;
#undef RR_CALL16
#undef RR_CALL24
#undef RR_CALL28
#undef RR_CALL32
#undef RR_CALLN
  }
  static int poly0_dual(const Poly0DualParams<T> &dp, bool out_f32, long long nwork, size_t smem, stream_t s)
  {
    (void)s;
    const int dlo = static_cast<int>(dp.base.step / dp.base.L), threads = dp.TS * dp.NL;
#ifdef B200RATE_EMU
#define RR_CALLD(O, NT, D)                                                                        \
  serial(nwork, smem, [&](long long w, T *sm) {                                                   \
    const Poly0DualTile t = poly0_dual_tile(dp, w);                                               \
    poly0_dual_load(dp, t, sm, nullptr, 0, 1);                                                    \
    for (int th = 0; th < threads; ++th)                                                          \
      poly0_dual_compute<T, O, NT, D>(dp, t, sm, poly0_dual_setup<T, O, NT, D>(dp, th));          \
  })
#else
#define RR_CALLD(O, NT, D) launch_persistent(poly0_dual_kernel<T, O, NT, D>, dp, nwork, threads, smem, s)
#endif
#define RR_CALLDN(O)                                                                              \
  do {                                                                                            \
    if (dp.base.n == 16) { if (dlo == 0) return RR_CALLD(O, 16, 0); if (dlo == 1) return RR_CALLD(O, 16, 1); return RR_CALLD(O, 16, 2); } \
    if (dlo == 0) return RR_CALLD(O, 24, 0);                                                      \
    if (dlo == 1) return RR_CALLD(O, 24, 1);                                                      \
    return RR_CALLD(O, 24, 2);                                                                    \
  } while (0)
    if (out_f32 && !kIsF32) RR_CALLDN(float);
    RR_CALLDN(T);
#undef RR_CALLDN
#undef RR_CALLD
  }
  static int polyN(const PolyParams<T> &p, bool in_f32, bool out_f32, long long nwork, size_t smem, stream_t s)
  {
#ifdef B200RATE_EMU
#define RR_CALL(I, O) serial(nwork, smem, [&](long long w, T *sm) { polyN_program<T, I, O>(p, w, sm); })
#else
#define RR_CALL(I, O) launch_persistent(polyN_kernel<T, I, O>, p, nwork, kTileThreads, smem, s)
#endif
    (void)s;
    RR_DISPATCH_IO(RR_CALL);
#undef RR_CALL
  }
  static int halfband(const HalfbandParams<T> &p, bool in_f32, bool out_f32, long long nwork, size_t smem, stream_t s)
  {
    (void)s;
#ifdef B200RATE_EMU
#define RR_CALLN(I, O, NC) serial(nwork, smem, [&](long long w, T *sm) { halfband_program<T, I, O, NC, sizeof(T) == 8 ? 8 : 4>(p, w, sm); })
#else
#define RR_CALLN(I, O, NC) launch_persistent(halfband_kernel<T, I, O, NC>, p, nwork, kTileThreads, smem, s)
#endif
#define RR_CALL8(I, O) RR_CALLN(I, O, 8)
#define RR_CALL9(I, O) RR_CALLN(I, O, 9)
#define RR_CALL10(I, O) RR_CALLN(I, O, 10)
#define RR_CALL11(I, O) RR_CALLN(I, O, 11)
#define RR_CALL12(I, O) RR_CALLN(I, O, 12)
#define RR_CALL13(I, O) RR_CALLN(I, O, 13)
    switch (p.ncoef) {
      case 8: RR_DISPATCH_IO(RR_CALL8);
      case 9: RR_DISPATCH_IO(RR_CALL9);
      case 10: RR_DISPATCH_IO(RR_CALL10);
      case 11: RR_DISPATCH_IO(RR_CALL11);
      case 12: RR_DISPATCH_IO(RR_CALL12);
      case 13: RR_DISPATCH_IO(RR_CALL13);
      default: return RR_INTERNAL;
    }
#undef RR_CALL8
#undef RR_CALL9
#undef RR_CALL10
#undef RR_CALL11
#undef RR_CALL12
#undef RR_CALL13
#undef RR_CALLN
  }
  static int copy(const CopyParams &p, bool in_f32, bool out_f32, stream_t s)
  {
    const long long nwork = (p.n * p.nlanes + kCopyTile - 1) / kCopyTile;
#ifdef B200RATE_EMU
#define RR_CALL(I, O) serial(nwork, 0, [&](long long w, T *) { copy_program<I, O>(p, w); })
#else
#define RR_CALL(I, O) launch_persistent(copy_kernel<I, O>, p, nwork, kTileThreads, 0, s)
#endif
    (void)s;
    RR_DISPATCH_IO(RR_CALL);
#undef RR_CALL
  }
#undef RR_DISPATCH_IO
};

// ===================================================================================================
// stage geometry (host) -- closed forms of the reference's per-call bookkeeping
// ===================================================================================================
struct StageGeom {
  int kind = 0, preload = 0;
  // half-band
  int hb_c = 0;
  // dft
  int N = 0, ov = 0, V = 0, L = 1, step = 1, in_mode = 0, Pf = 0, Ni = 0, remL0 = 0, q = 0, kept = 0, filter = 0;
  // poly
  int n = 0, Lp = 1, order = 0, phase_bits = 0, pre_post = 0;
  long long at0 = 0, pstep = 1;    // vpoly0: units of 1/L input samples; vpoly1..3: 32.32 fixed point
};

typedef __int128 i128;

static inline long long ceil_div(long long a, long long b) { return a <= 0 ? 0 : (a + b - 1) / b; }

// ---- DFT block geometry ----
static long long dft_Rb(const StageGeom &g, long long b)
{
  if (g.in_mode != DFT_IN_ZERO_STUFF) return b * static_cast<long long>(g.q);
  const long long Pb = b * static_cast<long long>(g.V);
  return Pb <= g.remL0 ? 0 : (Pb - g.remL0 + g.L - 1) / g.L;
}
static int dft_remLb(const StageGeom &g, long long b)
{
  if (g.in_mode != DFT_IN_ZERO_STUFF) return g.remL0;
  return static_cast<int>(g.remL0 + static_cast<long long>(g.L) * dft_Rb(g, b) - b * static_cast<long long>(g.V));
}
static long long dft_span(const StageGeom &g, long long b)   // inputs a block touches
{
  if (g.in_mode == DFT_IN_FREQ_UP) return g.Pf;
  if (g.in_mode == DFT_IN_COPY) return g.N;
  return (g.N - dft_remLb(g, b) + g.L - 1) / g.L;
}
static long long dft_kfirst(const StageGeom &g, long long b)  // first output index of block b
{
  if (g.step == 1) return b * static_cast<long long>(g.V);
  if (g.step > 1) return (b * static_cast<long long>(g.V) + g.step - 1) / g.step;
  return b * static_cast<long long>(g.kept);
}
static long long dft_block_of(const StageGeom &g, long long k)
{
  if (g.step == 1) return k / g.V;
  if (g.step > 1) return (k * g.step) / g.V;
  return k / g.kept;
}
// blocks runnable once `W` samples are in the input FIFO (dft_filter.h:78), starting the search at `from`
static long long dft_blocks_ready(const StageGeom &g, long long W, long long from)
{
  long long b = from;
  while (dft_remLb(g, b) + static_cast<long long>(g.L) * (W - dft_Rb(g, b)) >= g.N) ++b;
  return b;
}

// ---- polyphase geometry ----
static long long poly_q(const StageGeom &g, long long i)      // first input coordinate (minus pre) of output i
{
  if (g.order == 0) return (g.at0 + i * g.pstep) / g.Lp;
  return static_cast<long long>((static_cast<i128>(g.at0) + static_cast<i128>(i) * g.pstep) >> 32);
}
static long long poly_ready(const StageGeom &g, long long W)  // outputs available once W samples are in
{
  const long long avail = W - g.pre_post;
  if (avail <= 0) return 0;
  if (g.order == 0) return ceil_div(avail * g.Lp - g.at0, g.pstep);
  const i128 lim = (static_cast<i128>(avail) << 32) - g.at0;
  return lim <= 0 ? 0 : static_cast<long long>((lim + g.pstep - 1) / g.pstep);
}
static long long half_ready(const StageGeom &g, long long W)  // rate_filters_generic.h:83 in closed form
{
  const long long x = W - 4 * g.hb_c + 1;
  return x <= 0 ? 0 : x / 2;
}

struct StageRange {
  long long w0 = 0, wn = 0;             // dft: first block / blocks; others: first output / outputs
  long long prod_lo = 0, prod_hi = 0;   // output indices produced
  long long need_lo = 0, need_hi = 0;   // input FIFO coordinates read
};

// Work a stage must do so that outputs [klo, khi) exist, and the input coordinates that work reads.
static StageRange stage_range_for_outputs(const StageGeom &g, long long klo, long long khi)
{
  StageRange r;
  if (khi <= klo) return r;
  if (g.kind == RR_STAGE_HALFBAND) {
    r.w0 = klo; r.wn = khi - klo; r.prod_lo = klo; r.prod_hi = khi;
    r.need_lo = 2 * klo + 1; r.need_hi = 2 * khi + 4 * g.hb_c - 2;
  } else if (g.kind == RR_STAGE_DFT) {
    const long long b0 = dft_block_of(g, klo), b1 = dft_block_of(g, khi - 1) + 1;
    r.w0 = b0; r.wn = b1 - b0;
    r.prod_lo = dft_kfirst(g, b0); r.prod_hi = dft_kfirst(g, b1);
    r.need_lo = dft_Rb(g, b0); r.need_hi = dft_Rb(g, b1 - 1) + dft_span(g, b1 - 1);
  } else {
    r.w0 = klo; r.wn = khi - klo; r.prod_lo = klo; r.prod_hi = khi;
    r.need_lo = poly_q(g, klo); r.need_hi = poly_q(g, khi - 1) + g.n;
  }
  return r;
}

// ===================================================================================================
// Engine<T>: plan + device tables + stage launch
// ===================================================================================================
template <class T> class Engine {
 public:
  Design design;
  int ns = 0;
  StageGeom geom[RR_MAX_STAGES];
  int launches = 0;
  int device_id = 0;                                      // the device every buffer, stream and launch of this engine lives on
  const char *kernel_name[RR_MAX_STAGES] = {nullptr};   // what run_stage launched last for each stage

  ~Engine() { for (void *p : allocs_) be_free(p); }

  int init(const RR_config &cfg, int device)
  {
    int rc = build_design(cfg, static_cast<int>(sizeof(T)), design);
    if (rc != RR_OK) { set_last_error("invalid resampler configuration"); return rc; }
    if ((rc = be_set_device(device)) != RR_OK) return rc;
    device_id = be_current_device();
    ns = design.plan.num_stages;
    num_sms_ = be_num_sms();
    max_smem_ = be_max_smem();
    leaf_constants<T>(sqrthalf_, c16_1_, c16_3_);
    for (int i = 0; i < ns; ++i) {
      const rr_stage_plan &sp = design.plan.st[i];
      StageGeom &g = geom[i];
      g.kind = sp.kind; g.preload = sp.preload;
      if (sp.kind == RR_STAGE_HALFBAND) g.hb_c = sp.hb_coefs;
      else if (sp.kind == RR_STAGE_DFT) {
        g.N = sp.dft_length; g.ov = sp.num_taps - 1; g.V = g.N - g.ov; g.L = sp.L; g.step = sp.step_int;
        g.filter = sp.dft_filter_num;
        const bool pow2 = sp.L >= 2 && (sp.L & (sp.L - 1)) == 0;
        g.in_mode = pow2 ? DFT_IN_FREQ_UP : sp.L == 1 ? DFT_IN_COPY : DFT_IN_ZERO_STUFF;
        g.Pf = pow2 ? g.N / g.L : g.N;
        g.Ni = g.step < 0 ? g.N >> (-g.step) : g.N;
        g.remL0 = sp.remL;
        g.q = (g.N - g.ov - g.remL0 + g.L - 1) / g.L;
        g.kept = g.step < 0 ? g.N - ((((1 << (-g.step)) - 1) * g.N + g.ov) >> (-g.step)) : g.V;
        if (g.Pf < 64 || g.Ni < 64) { set_last_error("DFT stage too short"); return RR_INTERNAL; }
      } else {
        g.n = sp.n; g.Lp = sp.L; g.order = sp.interp_order; g.phase_bits = sp.phase_bits; g.pre_post = sp.pre_post;
        if (g.order == 0) { g.at0 = sp.at >> 32; g.pstep = sp.step >> 32; }
        else { g.at0 = sp.at; g.pstep = sp.step; }
      }
    }
    return upload_tables();
  }

  // Launch stage i: dft -> blocks [w0, w0+wn); others -> outputs [w0, w0+wn), for every lane of the views.
  int run_stage(int i, const LaneView &in, bool in_f32, const LaneView &out, bool out_f32, long long out_preload,
                long long w0, long long wn, int nlanes, stream_t s)
  {
    if (wn <= 0 || nlanes <= 0) return RR_OK;
    const StageGeom &g = geom[i];
    ++launches;
    if (g.kind == RR_STAGE_DFT) {
      DftParams<T> p = dft_params_[i];
      p.in = in; p.out = out; p.out_preload = out_preload;
      p.block0 = w0; p.nblocks = static_cast<int>(wn); p.nlanes = nlanes;
      if constexpr (std::is_same<T, float>::value) {
        // any two lanes form a pair (channels of a stream, or different streams of a mono / odd-channel batch)
        if (use_pair_kernel_ && use_pair_dft_ && !(nlanes & 1) && wn * (nlanes / 2) < (1ll << 30)) {
          DftPkParams pp;
          if (make_pair_params(i, p, pp, 0, wn * (nlanes / 2))) {
            last_dft_kernel_ = 1;
            const bool spec = (pp.spec_mode == PK_SPEC_UP2 && pp.fb == 10 && pp.ib == 11) ||
                              (pp.spec_mode == PK_SPEC_SAME && pp.fb == 11 && pp.ib == 11) ||
                              (pp.spec_mode == PK_SPEC_GEN && pp.fb == 11 && pp.ib == 10);
            kernel_name[i] = spec ? "dftp_kernel (lane pairs, size-specialised)" : "dftp_kernel (lane pairs)";
            return launch_dftp(pp, wn * (nlanes / 2), s);
          }
        }
      }
      if constexpr (sizeof(T) == 8) {
        // fp64 engine: in-place radix-16 kernel (rate_kernels_f64.cuh) for the block sizes that fit shared memory
        if (use_dft64_ && wn * nlanes < (1ll << 30)) {
          int rc;
          if (!d64_tab_[i].built && (rc = build_d64_tab(i))) return rc;
          const D64Tab &dt = d64_tab_[i];
          if (dt.ok) {
            Dft64Params dp = dt.params;
            dp.base = p;
            dp.in_f32 = in_f32 ? 1 : 0; dp.out_f32 = out_f32 ? 1 : 0;
            last_dft_kernel_ = 2;
            kernel_name[i] = "dft64_kernel (fp64, in-place radix-16)";
            return launch_dft64(dp, wn * nlanes, reinterpret_cast<void *>(s));
          }
        }
      }
      last_dft_kernel_ = 0;
      kernel_name[i] = "dft_kernel";
      const int lpc = dft_lanes_per_cta(g, nlanes);
      // prefetch the next tile with LDGSTS when the input needs no conversion and the staging buffer still
      // leaves room for two CTAs per SM
      const bool same_type = Launch<T>::kIsF32 || !in_f32;
      bool prefetch = same_type;
      size_t data_bytes = dft_smem_bytes<T>(g.Pf, g.Ni, lpc, prefetch, &p.xstride, &p.ystride, &p.zstride);
      p.lean_tables = 0;
      if (prefetch && data_bytes + sizeof(T) * dft_table_elems(p) > 112 * 1024) {
        // no room for two CTAs per SM with a prefetch buffer. One CTA per SM exposes the tile load entirely, so
        // rather keep the buffer and shrink the tables (pyramids only, cosines from global memory) if that fits
        p.lean_tables = 1;
        if (lpc == 1 && data_bytes + sizeof(T) * dft_table_elems(p) + 1024 <= max_smem_) {
          // keep prefetch
        } else {
          p.lean_tables = 0;
          prefetch = false;
          data_bytes = dft_smem_bytes<T>(g.Pf, g.Ni, lpc, false, &p.xstride, &p.ystride, &p.zstride);
        }
      }
      if (dft_big_[i]) {
        data_bytes = dft_smem_bytes<T>(g.Pf, g.Ni, 1, false, &p.xstride, &p.ystride, &p.zstride);
        int rc = ensure_big_scratch(data_bytes);
        if (rc != RR_OK) return rc;
        kernel_name[i] = "dft_big_kernel (global-memory work buffers)";
        return Launch<T>::dft(p, 1, in_f32, out_f32, wn * nlanes, data_bytes, s, big_scratch_, kBigCtas);
      }
      const long long groups = (nlanes + lpc - 1) / lpc;
      return Launch<T>::dft(p, lpc, in_f32, out_f32, wn * groups, data_bytes, s);
    }
    if (g.kind == RR_STAGE_HALFBAND) {
      HalfbandParams<T> p = half_params_[i];
      p.in = in; p.out = out; p.out_preload = out_preload; p.out0 = w0; p.nout = wn; p.nlanes = nlanes;
      if constexpr (std::is_same<T, float>::value) {
        if (use_pair_kernel_ && use_pair_half_ && !(nlanes & 1)) {
          // lane-pair kernel: all pairs of a stream per CTA when its frames are interleaved (coalesced window load)
          HalfbandPairParams hp;
          hp.base = p;
          const int nchan = in.nch;
          const bool group_ok = in.ch_stride == 1 && out.nch == nchan && nlanes % nchan == 0 && (nchan == 2 || nchan == 4 || nchan == 8) &&
                                !(out.nch & 1);
          hp.G = group_ok ? nchan / 2 : 1;
          hp.base.CH = 2 * hp.G;
          hp.base.tile = half_pair_tile_ / (2 * hp.G);       // outputs per pair per CTA, power of two
          hp.base.qbits = 0;
          while ((4 << hp.base.qbits) < hp.base.tile) ++hp.base.qbits;
          // rows of 8-byte values: 16-byte aligned (even) and 4 mod 16, so the pairs of a frame land 8 banks apart
          hp.base.half = ((hp.base.tile + 2 * p.ncoef + 8 + 15) / 16) * 16 + 4;
          const long long tiles = (wn + hp.base.tile - 1) / hp.base.tile;
          const size_t smem = 2 * sizeof(Pk) * (2 * static_cast<size_t>(hp.base.half) * hp.G + 8);   // two window buffers
          kernel_name[i] = "halfband_pair_kernel";
          return launch_halfband_pair(hp, tiles * (nlanes / (2 * hp.G)), smem, s);
        }
      }
      // all channels of a stream per CTA when the input is interleaved (coalesced window load)
      const int nchan = in.nch;
      const bool group_ok = in.ch_stride == 1 && out.nch == nchan && nlanes % nchan == 0 &&
                            (nchan == 2 || nchan == 4 || nchan == 8);
      p.CH = group_ok ? nchan : 1;
      p.tile = (sizeof(T) == 8 && !getenv("B200RATE_HALF_TILE_G2048") ? 2 * kHalfTile : kHalfTile) / p.CH;   // power of two (the window must fit kHbRaw registers per thread)
      p.qbits = 0;
      while ((4 << p.qbits) < p.tile) ++p.qbits;
      p.half = ((p.tile + 2 * p.ncoef + 8 + 31) / 32) * 32 + 4;
      // padded rows (hb_pad); 2 mod 4 doubles: rows four lanes apart (one staging store instruction) sit 64 bytes apart
      if (sizeof(T) == 8) p.half = ((hb_pad<8>(p.tile + 2 * p.ncoef + 8) + 2 + 31) / 32) * 32 + 6;
      const long long tiles = (wn + p.tile - 1) / p.tile;
      const size_t smem = sizeof(T) * 2 * static_cast<size_t>(p.half) * p.CH;
      kernel_name[i] = "halfband_kernel";
      return Launch<T>::halfband(p, in_f32, out_f32, tiles * ((nlanes + p.CH - 1) / p.CH), smem, s);
    }
    PolyParams<T> p = poly_params_[i];
    p.in = in; p.out = out; p.out_preload = out_preload; p.out0 = w0; p.nout = wn; p.nlanes = nlanes;
    p.tile = kPolyTile;
    if constexpr (std::is_same<T, float>::value) {
      if (use_pair_kernel_ && use_pair_poly_ && g.order == 0 && (g.n == 16 || g.n == 24 || g.n == 28 || g.n == 32) && !(nlanes & 1) &&
          g.Lp >= 48 && g.Lp <= 512 && g.pstep < (1 << 16)) {
        // lane-pair kernel: one column per period (L <= 512 slots), P pairs of a stream and PG period groups per CTA
        Poly0PairParams pp{};
        pp.fast.base = p;
        const int L = g.Lp, step = static_cast<int>(g.pstep);
        const int r_first = static_cast<int>((((g.at0 + static_cast<i128>(w0) * g.pstep) % L) + L) % L);
        // two adjacent slots per thread when their windows overlap almost entirely (step / L < 3) and a period
        // still has enough slot pairs for a CTA
        pp.CL = (use_pair_poly2_ && step / L <= 2 && L >= 96) ? 2 : 1;
        int bucket[16] = {0}, maxb = 0;
        for (int fs = 0; fs < L; fs += pp.CL) maxb = std::max(maxb, ++bucket[((r_first + static_cast<long long>(fs) * step) / L) & 15]);
        const int ncl = (L + pp.CL - 1) / pp.CL;
        int rows = (ncl + 15) / 16;
        auto spill_of = [&](int r) { int sp = 0; for (int b2 = 0; b2 < 16; ++b2) sp += std::max(0, bucket[b2] - r); return sp; };
        const int max_threads = pp.CL == 2 ? 256 : 512;      // launch bounds of poly0_pair2_kernel / poly0_pair_kernel
        // two-slot kernel: move clusters of overfull banks one bank down by starting their window a sample early
        // (Poly0PairParams::shift): m[b] clusters leave bank b for bank b - 1 until no bank holds more than `rows`
        pp.shift = 0;
        if (pp.CL == 2 && use_pair_shift_) {
          const int rows0 = (ncl + 15) / 16;
          int mv[16] = {0};
          bool ok = false;
          for (int pass = 0; pass < 64 && !ok; ++pass) {
            ok = true;
            for (int b2 = 0; b2 < 16; ++b2) {
              const int load = bucket[b2] - mv[b2] + mv[(b2 + 1) & 15];
              if (load > rows0) { mv[b2] += load - rows0; ok = false; }
            }
          }
          for (int b2 = 0; b2 < 16; ++b2) ok = ok && mv[b2] >= 0 && mv[b2] <= bucket[b2];
          if (ok) {
            pp.shift = 1; rows = rows0;
            for (int b2 = 0; b2 < 16; ++b2) pp.keep[b2] = static_cast<unsigned char>(bucket[b2] - mv[b2]);
          }
        }
        if (!pp.shift && spill_of(rows) > ncl / 10 && 16 * (rows + 1) <= max_threads) ++rows;   // a few idle lanes cost less than two-way conflicts
        pp.spread = 1; pp.tslots = 16 * rows;               // overfull banks spill into the holes (poly0_pair_deal_overflow)
        if (!pp.shift && (maxb > 2 * rows || spill_of(rows) > kPolyDealOverflow)) pp.spread = 0;   // few distinct banks (steep up-sampling): keep order
        pp.P = 1;
        while (!(in.nch & 1) && !(out.nch & 1) && in.nch == out.nch && 2 * pp.P < in.nch && in.nch % (4 * pp.P) == 0 &&
               pp.tslots * 2 * pp.P <= 256)
          pp.P *= 2;                                          // several pairs per CTA as channels of one stream ...
        // ... or the stereo pairs of two consecutive streams: 74 slot pairs of a period fill 5 full warps instead of
        // 2.3 of 3 (the two-slot kernel is bound by issue and latency, not by a pipe)
        if (pp.P == 1 && in.nch == 2 && out.nch == 2 && nlanes % 4 == 0 && pp.tslots * 2 <= 256 &&
            pp.tslots % 32 != 0 && !getenv("B200RATE_PAIR2_P1"))
          pp.P = 2;
        // one period group; as many CTAs per SM as 64 registers per thread allow (1024 threads), each with one
        // window buffer of an even number of periods (the other CTAs cover its load)
        pp.PG = 1;
        const int threads = pp.tslots * pp.P;
        // window budget per CTA from the CTAs an SM will hold: 1024 threads' worth, but at 80 registers a 160-thread CTA
        // fits four times (measured with 6 / 5 / 4 / 3 assumed: cfg4 0.825 / 0.804 / 0.792 / 0.860 ms per 256 streams)
        int ctas = std::max(1, std::min(16, 1024 / threads));
        if (threads >= 160) ctas = std::min(ctas, 4);
        if (const char *e = getenv("B200RATE_PAIR_CTAS")) ctas = std::max(1, atoi(e));          // probes
        const size_t budget = std::min<size_t>(64 * 1024, (max_smem_ - 2048) / ctas - 1024);
        auto window_of = [&](int mm) {
          const long long wd = ((L - 1) + static_cast<long long>(L - 1) * step) / L + static_cast<long long>(mm - 1) * step + g.n + 1 + 4;
          return ((wd + 15) / 16) * 16 + 8;
        };
        int MM = 32;
        while (MM > 2 && window_of(MM) * pp.P * sizeof(Pk) > budget) MM -= 2;
        if (window_of(MM) * pp.P * sizeof(Pk) <= budget && nlanes % (2 * pp.P) == 0 && threads <= max_threads) {
          const long long periods = (wn + L - 1) / L;
          pp.fast.F = L; pp.fast.ncols = 1; pp.fast.MM = MM; pp.fast.CH = 2 * pp.P;
          pp.fast.win = static_cast<int>(window_of(MM));
          pp.fast.mtiles = (periods + MM - 1) / MM;
          pp.fast.double_buffer = 0;
          const long long nwork = static_cast<long long>(nlanes / (2 * pp.P)) * pp.fast.mtiles;
          kernel_name[i] = pp.CL == 2 ? "poly0_pair2_kernel" : "poly0_pair_kernel";
          return launch_poly0_pair(pp, pp.tslots * pp.P * pp.PG, nwork, s);
        }
      }
    }
    if constexpr (sizeof(T) == 8) {
      // fp64 engine: two slots per thread, host-dealt slot pairs, TMA-staged windows (poly0_dual_kernel)
      if (use_dual_poly_ && g.order == 0 && (g.n == 16 || g.n == 24) && !in_f32 && in.elem_stride == 1 && g.Lp >= 32 &&
          (g.Lp + 1) / 2 <= 128 && g.pstep / g.Lp <= 2 && g.pstep < (1 << 15) && g.at0 >= 0 && g.at0 < g.Lp && p.pre == 0) {
        int rc;
        if (!dual_tab_[i].built && (rc = build_dual_tab(i))) return rc;
        const DualTab &dt = dual_tab_[i];
        Poly0DualParams<T> dp;
        dp.base = p;
        dp.coef = dt.coef; dp.slot = dt.slot; dp.qs = dt.qs; dp.flags = dt.flags; dp.TS = dt.TS;
        // lanes per CTA: two (160 threads for 5 rows of slot pairs: three CTAs per SM at 128 registers, three tiles in
        // flight) measured better than three (two CTAs) or one (five CTAs): cfg3 4.11 / 4.31 / 4.95 ms
        dp.NL = std::max(1, std::min(std::min(nlanes, 256 / dt.TS), 2));
        if (const char *e = getenv("B200RATE_DUAL_NL")) dp.NL = std::max(1, std::min(std::min(nlanes, 256 / dt.TS), atoi(e)));   // probes
        // periods per tile: about ten (measured on cfg3, step 320: 4 / 7 / 9 / 12 periods 1.342 / 1.131 / 1.106 / 1.127 ms;
        // 44.1 -> 48 kHz, step 147: 10 / 15 / 20 periods 1.962 / 2.036 / 2.149 ms), within 4096 doubles per lane window
        const int budget = getenv("B200RATE_DUAL_BUDGET") ? std::max(1024, std::min(6144, atoi(getenv("B200RATE_DUAL_BUDGET")))) : 4096;
        dp.MM = std::max(2, std::min(getenv("B200RATE_DUAL_BUDGET") ? 64 : 10, static_cast<int>((budget - g.n - 8) / g.pstep)));
        dp.win = static_cast<int>(((static_cast<long long>(dp.MM) * g.pstep + g.n + 8 + 1) & ~1ll) + 2);
        dp.m_begin = w0 / g.Lp;
        const long long m_end = (w0 + wn + g.Lp - 1) / g.Lp;
        dp.mtiles = (m_end - dp.m_begin + dp.MM - 1) / dp.MM;
        const long long groups = (nlanes + dp.NL - 1) / dp.NL;
        dp.groups = static_cast<int>(groups);
        dp.group_fastest = (out.nch > 1 && groups < 0x7fffffffll && !getenv("B200RATE_DUAL_TILE_FASTEST")) ? 1 : 0;
        const size_t smem = sizeof(T) * (static_cast<size_t>(dp.win) * dp.NL + kDualPad + 2);
        if (dp.mtiles < 0x7fffffffll && smem <= 100 * 1024) {
          kernel_name[i] = "poly0_dual_kernel (two slots per thread, TMA windows)";
          return Launch<T>::poly0_dual(dp, out_f32, groups * dp.mtiles, smem, s);
        }
      }
    }
    if (g.order == 0 && (g.n == 16 || g.n == 24 || g.n == 28 || g.n == 32)) {
      // phase-stationary kernel: needs enough phases to fill a CTA and a window that fits shared memory
      const int nch = in.nch;
      // lanes per CTA: as many channels of a stream as still leave >= 8 periods per tile in a 24 KB window
      // (the coefficient row load is amortised over the periods); writes stay contiguous per frame
      int CH = (nch <= 8 && nlanes % nch == 0) ? nch : 1;
      int ncols = 1, F = g.Lp, threads = 32, MM = 16;
      auto window_of = [&](int f, int mm) {
        long long w = ((g.Lp - 1) + static_cast<long long>(f - 1) * g.pstep) / g.Lp + static_cast<long long>(mm - 1) * g.pstep + g.n + 1;
        const long long unit = 64 / sizeof(T);
        return ((w + 2 * unit - 1) / (2 * unit)) * 2 * unit + unit;
      };
      for (;; CH >>= 1) {
        ncols = (g.Lp * CH + 511) / 512;
        F = (g.Lp + ncols - 1) / ncols;
        threads = ((F * CH + 31) / 32) * 32;
        MM = 16;
        while (MM > 2 && window_of(F, MM) * CH * sizeof(T) > 24 * 1024) MM >>= 1;
        if (MM >= 8 || CH == 1 || (CH & 1)) break;
      }
      const long long periods = (wn + g.Lp - 1) / g.Lp;
      auto window = [&](int mm) { return window_of(F, mm); };
      if (F * CH >= 64 && g.pstep < (1 << 20) && window(MM) * CH * sizeof(T) <= 48 * 1024) {
        Poly0FastParams<T> fp;
        fp.base = p; fp.F = F; fp.ncols = ncols; fp.MM = MM; fp.CH = CH;
        fp.win = static_cast<int>(window(MM));
        fp.mtiles = (periods + MM - 1) / MM;
        fp.double_buffer = 1;
        const long long nwork = static_cast<long long>(nlanes / CH) * ncols * fp.mtiles;
        kernel_name[i] = "poly0_fast_kernel";
        return Launch<T>::poly0_fast(fp, threads, in_f32, out_f32, nwork, 2 * sizeof(T) * static_cast<size_t>(fp.win) * CH, s);
      }
    }
    // window of one tile: worst-case start phase (L-1) plus (tile-1) steps, plus the taps
    long long win;
    if (g.order == 0) win = ((g.Lp - 1) + static_cast<long long>(p.tile - 1) * g.pstep) / g.Lp + g.n + 1;
    else win = static_cast<long long>(((static_cast<i128>(0xffffffffll) + static_cast<i128>(p.tile - 1) * g.pstep) >> 32)) + g.n + 1;
    while (win * sizeof(T) > max_smem_ / 2 && p.tile > 64) {   // very steep decimation: shrink the tile
      p.tile >>= 1;
      if (g.order == 0) win = ((g.Lp - 1) + static_cast<long long>(p.tile - 1) * g.pstep) / g.Lp + g.n + 1;
      else win = static_cast<long long>(((static_cast<i128>(0xffffffffll) + static_cast<i128>(p.tile - 1) * g.pstep) >> 32)) + g.n + 1;
    }
    p.win_cap = static_cast<int>(win);
    const long long tiles = (wn + p.tile - 1) / p.tile;
    size_t smem = sizeof(T) * static_cast<size_t>(win);
    p.row_pitch = 0;
    if (g.order > 0) {
      // interpolated stages: every warp stages the coefficient rows of its 32 outputs in shared memory (polyN_program)
      const int per = 16 / static_cast<int>(sizeof(T)), row_elems = g.n * (g.order + 1);
      const int pitch = ((row_elems / per) | 1) * per;
      const size_t rows = sizeof(T) * static_cast<size_t>(kTileThreads) * pitch;
      const size_t base = sizeof(T) * static_cast<size_t>((win + per - 1) / per * per);
      if (g.n % per == 0 && base + rows + 1024 <= max_smem_ / 2) { p.row_pitch = pitch; smem = base + rows; }
    }
    kernel_name[i] = g.order == 0 ? "poly0_kernel" : "polyN_kernel";
    if (g.order == 0) return Launch<T>::poly0(p, in_f32, out_f32, tiles * nlanes, smem, s);
    return Launch<T>::polyN(p, in_f32, out_f32, tiles * nlanes, smem, s);
  }

  // ---- DFT stage i and the vpoly0 stage i + 1 as one kernel (rate_kernels_fused.cuh) ----
  // Static part of the decision: plan shapes the fused kernel is instantiated for (x2 F-domain up-sampling with
  // N = 4096, i.e. the 44.1 <-> 48 kHz family at Best quality, followed by a 24-tap rational polyphase stage).
  bool fused_config_ok(int i) const
  {
    if (!std::is_same<T, float>::value || !use_pair_kernel_ || !use_pair_dft_ || !use_pair_poly_ || !use_fused_ || i + 1 >= ns) return false;
    const StageGeom &d = geom[i], &q = geom[i + 1];
    if (d.kind != RR_STAGE_DFT || q.kind != RR_STAGE_POLY || q.order != 0) return false;
    if (pk_spec_mode(d) != PK_SPEC_UP2 || ilog2(d.Pf) - 1 != 10 || ilog2(d.Ni) - 1 != 11) return false;
    if (d.step != 1 || (d.V & 1) || d.V < 8 * kFusedHalo) return false;
    if (q.n != 24 || design.plan.st[i + 1].pre != 0 || q.at0 < 0 || q.at0 >= q.Lp) return false;
    // at least sixteen threads of a group must be free of polyphase work: they fetch the next tile meanwhile
    if (q.pstep / q.Lp > 2 || q.pstep >= (1 << 15) || (q.Lp + 1) / 2 > kFusedPolyThreads - 32) return false;
    return true;
  }

  // in: view of the DFT stage's input; out: view the polyphase stage writes (final output or next FIFO); blocks
  // [b0, b0 + nb) of the DFT stage, outputs [k0, k0 + nk) of the polyphase stage. Returns RR_RATEERROR (nothing
  // launched) when this call's buffers do not allow the fused kernel (input not adjacent stereo frames).
  int run_fused(int i, const LaneView &in, const LaneView &out, long long out_preload, long long b0, long long nb, long long k0,
                long long nk, int nlanes, stream_t s)
  {
    if (nb <= 0 || nk <= 0 || nlanes <= 0) return RR_OK;
    if constexpr (!std::is_same<T, float>::value) return RR_RATEERROR;
    else {
      if ((nlanes & 1) || out.mask != ~0ull) return RR_RATEERROR;
      const StageGeom &d = geom[i], &q = geom[i + 1];
      int rc;
      if (!fused_tab_[i].built && (rc = build_fused_tab(i))) return rc;
      DftPolyParams fp;
      memset(&fp, 0, sizeof(fp));
      DftParams<float> p = dft_params_[i];
      p.in = in; p.out = out; p.out_preload = 0;          // the DFT stage's own output view is not used
      p.block0 = b0; p.nblocks = static_cast<int>(nb); p.nlanes = nlanes;
      if (!make_pair_params(i, p, fp.dft, kFusedHalo / 2) || !fp.dft.stereo || fp.dft.groups < 1) return RR_RATEERROR;
      fp.L = q.Lp; fp.n = q.n; fp.at0 = q.at0; fp.step = q.pstep;
      fp.poly_preload = q.preload;
      fp.out0 = k0; fp.nout = nk; fp.out = out; fp.out_preload = out_preload;
      fp.coef = fused_tab_[i].coef; fp.slot = fused_tab_[i].slot; fp.qs = fused_tab_[i].qs; fp.flags = fused_tab_[i].flags;
      fp.tile_t0 = fused_tab_[i].tile_t0;
      fp.block0 = b0; fp.nblocks = static_cast<int>(nb);
      // runs: whole lanes when there are enough of them, otherwise pieces of at least 16 blocks (each piece but the
      // first recomputes one block for its history) so that every group of every SM gets a dozen items
      const long long pairs = nlanes / 2, target = 12ll * num_sms_ * fp.dft.groups;
      long long runs = std::max(1ll, std::min((target + pairs - 1) / pairs, std::max(1ll, nb / 16)));
      const long long len = (nb + runs - 1) / runs;
      runs = (nb + len - 1) / len;
      fp.run_len = static_cast<int>(len); fp.runs_per_pair = static_cast<int>(runs);
      if (pairs * runs >= (1ll << 30)) return RR_RATEERROR;
      (void)d;
      ++launches;
      kernel_name[i] = "dft_poly_kernel (DFT stage + vpoly0 fused, lane pairs)";
      kernel_name[i + 1] = "(fused into the DFT stage's kernel)";
      return launch_dft_poly(fp, pairs * runs, s);
    }
  }

  int copy(const LaneView &in, bool in_f32, const LaneView &out, bool out_f32, long long c0, long long n,
           long long in_shift, int nlanes, stream_t s)
  {
    if (n <= 0) return RR_OK;
    ++launches;
    CopyParams p{in, out, c0, n, in_shift, nlanes};
    return Launch<T>::copy(p, in_f32, out_f32, s);
  }

  int dft_spectrum_host(int instance, void *out, int max_n) const
  {
    if (!dft_coef_dev_[instance]) return 0;
    const int N = design.dft[instance].dft_length;
    const int n = std::min(N, max_n);
    if (n > 0) {
      if (be_d2h(out, dft_coef_dev_[instance], sizeof(T) * static_cast<size_t>(n), 0) != RR_OK) return -1;
      if (be_sync(0) != RR_OK) return -1;
    }
    return N;
  }

  // Algorithmic flop count of one stage's work range, per lane (SURVEY.md 8d accounting: split-radix
  // 4M log2 M - 6M + 8 per M-point complex FFT, 18 per real-FFT pair, 6 per multiplied bin, 1 + 3c per
  // half-band output, 2n (x (1 + order)) per polyphase output).
  double stage_flops(int i, const StageRange &r) const
  {
    auto sr = [](double M) { return 4 * M * std::log2(M) - 6 * M + 8; };
    const StageGeom &g = geom[i];
    if (g.kind == RR_STAGE_DFT)
      return static_cast<double>(r.wn) * (sr(g.Pf / 2.) + 18. * g.Pf / 4 + 6. * g.Ni / 2 + 18. * g.Ni / 4 + sr(g.Ni / 2.));
    if (g.kind == RR_STAGE_HALFBAND) return static_cast<double>(r.wn) * (1 + 3. * g.hb_c);
    return static_cast<double>(r.wn) * g.n * (2. + 2. * g.order);
  }
  double flops_of(const StageRange *r) const
  {
    double total = 0;
    for (int i = 0; i < ns; ++i) total += stage_flops(i, r[i]);
    return total;
  }

 private:
  std::vector<void *> allocs_;
  int num_sms_ = 148;
  size_t max_smem_ = 48 * 1024;
  T sqrthalf_, c16_1_, c16_3_;
  DftParams<T> dft_params_[RR_MAX_STAGES];
  PolyParams<T> poly_params_[RR_MAX_STAGES];
  HalfbandParams<T> half_params_[RR_MAX_STAGES];
  T *dft_coef_dev_[2] = {nullptr, nullptr};
  bool dft_big_[RR_MAX_STAGES] = {false};
  void *big_scratch_ = nullptr;              // global work buffers of dft_big_kernel, kBigCtas slices of big_slice_ bytes
  size_t big_slice_ = 0;
  static constexpr int kBigCtas = 296;

  // ---- per-thread tables of poly0_dual_kernel (fp64 vpoly0, two slots per thread): see build_fused_tab ----
  struct DualTab { const T *coef = nullptr; const uint16_t *slot = nullptr, *qs = nullptr; const uint8_t *flags = nullptr; int TS = 0; bool built = false; };
  DualTab dual_tab_[RR_MAX_STAGES];
  bool use_dual_poly_ = getenv("B200RATE_NO_DUAL_POLY") == nullptr;

  // ---- tables of dft64_kernel (fp64 DFT stage, rate_kernels_f64.cuh) ----
  struct D64Tab { bool built = false, ok = false; Dft64Params params; };
  D64Tab d64_tab_[RR_MAX_STAGES];
  bool use_dft64_ = getenv("B200RATE_NO_DFT64") == nullptr;

  int build_d64_tab(int i)
  {
    D64Tab &dt = d64_tab_[i];
    dt.built = true;
    dt.ok = false;
    if constexpr (sizeof(T) == 8) {
      const StageGeom &g = geom[i];
      Dft64Params &dp = dt.params;
      memset(&dp, 0, sizeof(dp));
      if (g.in_mode == DFT_IN_FREQ_UP && (g.L == 2 || g.L == 4 || g.L == 8) && g.step >= 1) { dp.mode = D64_UP2; dp.up_bits = g.L == 2 ? 1 : g.L == 4 ? 2 : 3; }
      else if (g.Ni == g.Pf && g.step >= 1 && g.in_mode != DFT_IN_FREQ_UP) dp.mode = D64_SAME;
      else if (g.step < 0 && g.in_mode != DFT_IN_FREQ_UP && g.Pf == g.N) dp.mode = D64_DECIM;
      else return RR_OK;                                   // other shapes stay on the generic kernel
      dp.fb = ilog2(g.Pf) - 1;
      dp.ib = dp.mode == D64_DECIM ? ilog2(g.Ni) - 1 : dp.fb;
      if (dp.fb < 5 || dp.fb > 13 || dp.ib < 5 || dp.ib > 13) return RR_OK;
      const D64Plan pf = d64_plan(dp.fb), pi = d64_plan(dp.ib);
      dp.npf = pf.n; dp.npi = pi.n;
      // pass twiddle rows: forward pass ps works on sub-blocks of 2^lgS points with stride 2^(lgS - lr)
      std::vector<double> tw;
      auto add_row = [&](int lgS, int lr, int sign) -> int {
        const int s = 1 << (lgS - lr);
        if (s == 1) return -1;                             // the stride-1 pass has no twiddles
        const int off = static_cast<int>(tw.size() / 2);
        const std::vector<double> row = unit_circle(1 << lgS, s, sign);
        tw.insert(tw.end(), row.begin(), row.end());
        return off;
      };
      int lgS = dp.fb;
      for (int ps = 0; ps < pf.n; ++ps) { dp.lr_f[ps] = pf.lr[ps]; dp.tw_f[ps] = add_row(lgS, pf.lr[ps], -1); lgS -= pf.lr[ps]; }
      lgS = 0;
      for (int ps = 0; ps < pi.n; ++ps) {                  // the transposed network runs the plan backwards
        const int lr = pi.lr[pi.n - 1 - ps];
        lgS += lr;
        dp.lr_i[ps] = lr; dp.tw_i[ps] = add_row(lgS, lr, +1);
      }
      dp.ntw = static_cast<int>(tw.size() / 2);
      dp.fslots = d64_buf_slots(dp.fb); dp.bslots = d64_buf_slots(dp.ib);
      // 8 / L mod 8: the L transforms' slots of one position sit in different bank groups (stored by neighbouring threads)
      dp.hstride = dp.bslots + (dp.mode == D64_UP2 ? 8 >> dp.up_bits : 4);
      dp.group_slots = dp.mode == D64_UP2 ? (dp.hstride << dp.up_bits) : dp.mode == D64_SAME ? dp.fslots : dp.fslots + dp.bslots;
      // groups: 64 threads for blocks up to 1024 complex points, else 128; as many as shared memory and the launch bound allow
      dp.gthreads = dp.fb + (dp.mode == D64_UP2 ? dp.up_bits - 1 : 0) <= 10 ? 64 : 128;
      if (const char *e = getenv("B200RATE_D64_GT")) dp.gthreads = atoi(e) == 64 ? 64 : atoi(e) == 256 ? 256 : 128;
      const size_t per_group = sizeof(CD) * static_cast<size_t>(dp.group_slots), fixed = sizeof(CD) * static_cast<size_t>(dp.ntw) + 2048;
      if (fixed + per_group > max_smem_) return RR_OK;
      dp.groups = static_cast<int>(std::min<size_t>((max_smem_ - fixed) / per_group, static_cast<size_t>(kD64MaxThreads / dp.gthreads)));
      dp.groups = std::min(dp.groups, kD64MaxGroups);
      dp.lane_major = getenv("B200RATE_D64_LANE_MAJOR") ? 1 : 0;
      if (const char *e = getenv("B200RATE_D64_GROUPS")) dp.groups = std::max(1, std::min(dp.groups, atoi(e)));
      const DftFilterDesign &f = design.dft[g.filter];
      const std::vector<double> H = real_spectrum(f.coefs_time, 0.25);
      const std::vector<double> ta = unit_circle(g.Pf, g.Pf / 4 + 1, -1), tb = unit_circle(g.Ni, g.Ni / 4 + 1, +1);
      const double *dH = nullptr, *dta = nullptr, *dtb = nullptr, *dtw = nullptr;
      int rc;
      if ((rc = upload(H, &dH)) || (rc = upload(ta, &dta)) || (rc = upload(tb, &dtb)) || (rc = upload(tw, &dtw))) return rc;
      if ((rc = be_sync(0))) return rc;
      dp.H = reinterpret_cast<const CD *>(dH); dp.ta = reinterpret_cast<const CD *>(dta);
      dp.tb = reinterpret_cast<const CD *>(dtb); dp.tw = reinterpret_cast<const CD *>(dtw);
      dt.ok = true;
    }
    (void)i;
    return RR_OK;
  }

  int build_dual_tab(int i)
  {
    const StageGeom &q = geom[i];
    const int L = q.Lp, n = q.n, nsp = (L + 1) / 2, dlo = static_cast<int>(q.pstep / q.Lp);
    const int step = static_cast<int>(q.pstep), at0 = static_cast<int>(q.at0);
    // deal the slot pairs to half-warps by the 8-byte bank of their first window sample. Banks hold unequal numbers of
    // pairs; mv[b] pairs of bank b move to bank b - 1 by starting their pass one sample early, until no bank holds more
    // than `rows` (16-element relaxation; feasible for every (L, step, phase) tried). If it is not, what does not fit a
    // bank's rows fills the holes (a two-way conflict each).
    std::vector<std::vector<int>> bucket(16);
    for (int sp = 0; sp < nsp; ++sp) bucket[static_cast<size_t>(((at0 + 2 * sp * step) / L) & 15)].push_back(sp);
    int rows = (nsp + 15) / 16;
    int mv[16] = {0};
    bool shifted = use_pair_shift_;
    for (int pass = 0; pass < 64 && shifted; ++pass) {
      bool changed = false;
      for (int b = 0; b < 16; ++b) {
        const int load = static_cast<int>(bucket[static_cast<size_t>(b)].size()) - mv[b] + mv[(b + 1) & 15];
        if (load > rows) { mv[b] += load - rows; changed = true; }
      }
      if (!changed) break;
      if (pass == 63) shifted = false;
    }
    for (int b = 0; b < 16 && shifted; ++b) shifted = mv[b] <= static_cast<int>(bucket[static_cast<size_t>(b)].size());
    auto spill_of = [&](int r) { int sp = 0; for (const auto &b : bucket) sp += std::max(0, static_cast<int>(b.size()) - r); return sp; };
    if (!shifted) while (rows < 8 && spill_of(rows) > nsp / 8) ++rows;
    const int TS = 16 * rows;
    std::vector<int> owner(static_cast<size_t>(TS), -1), early(static_cast<size_t>(TS), 0), rest;
    for (int b = 0; b < 16; ++b) {
      const std::vector<int> &bk = bucket[static_cast<size_t>(b)];
      const int keep = shifted ? static_cast<int>(bk.size()) - mv[b] : static_cast<int>(bk.size());
      const int b2 = (b + 15) & 15, keep2 = static_cast<int>(bucket[static_cast<size_t>(b2)].size()) - mv[b2];
      for (int k = 0; k < static_cast<int>(bk.size()); ++k) {
        if (k < keep) {
          if (k < rows) owner[static_cast<size_t>(k) * 16 + b] = bk[static_cast<size_t>(k)];
          else rest.push_back(bk[static_cast<size_t>(k)]);
        } else {
          const size_t t = static_cast<size_t>(keep2 + k - keep) * 16 + b2;
          owner[t] = bk[static_cast<size_t>(k)]; early[t] = 1;
        }
      }
    }
    for (int t = 0; t < TS && !rest.empty(); ++t)
      if (owner[static_cast<size_t>(t)] < 0) { owner[static_cast<size_t>(t)] = rest.back(); rest.pop_back(); }
    if (!rest.empty()) return RR_INTERNAL;
    std::vector<uint16_t> slot(static_cast<size_t>(TS), 0xffff), qs(static_cast<size_t>(TS), 1);
    std::vector<uint8_t> flags(static_cast<size_t>(TS), 0);
    std::vector<T> coef(static_cast<size_t>(2 * n + 3) * TS, static_cast<T>(0));
    for (int t = 0; t < TS; ++t) {
      const int sp = owner[static_cast<size_t>(t)];
      if (sp < 0) continue;
      const int sh = early[static_cast<size_t>(t)];
      const int s0 = 2 * sp, a0 = at0 + s0 * step, q0 = a0 / L, r0 = a0 % L, a1 = a0 + step, q1 = a1 / L, r1 = a1 % L;
      const bool two = s0 + 1 < L, d_lo = q1 - q0 == dlo;
      const int f1 = (d_lo ? 0 : 1) + sh;
      slot[static_cast<size_t>(t)] = static_cast<uint16_t>(s0);
      qs[static_cast<size_t>(t)] = static_cast<uint16_t>(q0 - sh + 1);
      flags[static_cast<size_t>(t)] = static_cast<uint8_t>((d_lo ? 1 : 0) | (two ? 2 : 0) | (sh ? 4 : 0));
      for (int j = 0; j <= n; ++j) {                       // position j of the pass carries tap j - sh of the first output
        const int k = j - sh;
        if (k >= 0 && k < n) coef[static_cast<size_t>(j) * TS + t] = static_cast<T>(design.poly_bank[static_cast<size_t>(r0) * n + k]);
      }
      for (int j = 0; j <= n + 1; ++j) {                   // position dlo + j carries tap j - f1 of the second output
        const int k = j - f1;
        if (two && k >= 0 && k < n) coef[static_cast<size_t>(n + 1 + j) * TS + t] = static_cast<T>(design.poly_bank[static_cast<size_t>(r1) * n + k]);
      }
    }
    DualTab &dt = dual_tab_[i];
    int rc;
    if ((rc = upload(coef, &dt.coef)) || (rc = upload(slot, &dt.slot)) || (rc = upload(qs, &dt.qs)) || (rc = upload(flags, &dt.flags))) return rc;
    if ((rc = be_sync(0))) return rc;
    dt.TS = TS; dt.built = true;
    return RR_OK;
  }

  struct FusedTab { const float *coef = nullptr; const uint16_t *slot = nullptr, *qs = nullptr; const uint8_t *flags = nullptr; int tile_t0 = 0; bool built = false; };
  FusedTab fused_tab_[RR_MAX_STAGES];
  // The fused DFT + vpoly0 kernel is an option, not the default: it removes the intermediate FIFO (a third of the HBM
  // traffic and half the device memory of a 48 -> 44.1 kHz batch) but measures 30 % slower than the two kernels on a
  // B200 -- the stages are issue / latency bound, not HBM bound, and the polyphase phase occupies only the 74 threads
  // of a group that own a slot pair (profiles/README.md, round 2). B200RATE_FUSED=1 selects it.
  // groups per CTA of the single-buffer DFT kernel: 5 (102 registers per thread) or 6 (85); 4 = the two-buffer kernel
  int pk_inplace_groups_ = getenv("B200RATE_DFT_GROUPS") ? atoi(getenv("B200RATE_DFT_GROUPS")) : kPkInplaceGroups;
  bool use_fused_ = getenv("B200RATE_FUSED") != nullptr && atoi(getenv("B200RATE_FUSED")) != 0;

  // Per-thread tables of the fused kernel's polyphase phase: thread t owns the slot pair (s, s + 1) of every period
  // (output index mod L); the slot pairs are dealt to the threads so that the sixteen lanes of a half-warp start
  // their windows in sixteen different 8-byte banks.
  int build_fused_tab(int i)
  {
    const StageGeom &q = geom[i + 1];
    const int L = q.Lp, n = q.n, nsp = (L + 1) / 2, dlo = static_cast<int>(q.pstep / q.Lp);
    const int step = static_cast<int>(q.pstep), at0 = static_cast<int>(q.at0);
    std::vector<uint16_t> slot(kFusedPolyThreads, 0xffff), qs(kFusedPolyThreads, 0);
    std::vector<uint8_t> flags(kFusedPolyThreads, 0);
    std::vector<float> coef(static_cast<size_t>(2 * n + 1) * kFusedPolyThreads, 0.f);
    // deal: bucket the slot pairs by the bank of their first window sample; entry k of every bucket goes to half-warp
    // k, what does not fit fills the holes
    std::vector<std::vector<int>> bucket(16);
    for (int sp = 0; sp < nsp; ++sp) bucket[static_cast<size_t>(((at0 + 2 * sp * step) / L) & 15)].push_back(sp);
    const int rows = kFusedPolyThreads / 16;
    std::vector<int> owner(kFusedPolyThreads, -1), rest;
    for (int b = 0; b < 16; ++b)
      for (size_t k = 0; k < bucket[static_cast<size_t>(b)].size(); ++k) {
        if (static_cast<int>(k) < rows) owner[k * 16 + static_cast<size_t>(b)] = bucket[static_cast<size_t>(b)][k];
        else rest.push_back(bucket[static_cast<size_t>(b)][k]);
      }
    // compact: threads beyond the last used half-warp stay idle; leftovers go to the first holes
    for (int t = 0; t < kFusedPolyThreads && !rest.empty(); ++t)
      if (owner[static_cast<size_t>(t)] < 0) { owner[static_cast<size_t>(t)] = rest.back(); rest.pop_back(); }
    if (!rest.empty()) return RR_INTERNAL;
    int used = 0;
    for (int t = 0; t < kFusedPolyThreads; ++t) if (owner[static_cast<size_t>(t)] >= 0) used = t + 1;
    fused_tab_[i].tile_t0 = (used + 15) / 16 * 16;
    if (fused_tab_[i].tile_t0 > kFusedPolyThreads - 16) return RR_INTERNAL;
    for (int t = 0; t < kFusedPolyThreads; ++t) {
      const int sp = owner[static_cast<size_t>(t)];
      if (sp < 0) continue;
      const int s0 = 2 * sp, a0 = at0 + s0 * step, q0 = a0 / L, r0 = a0 % L, a1 = a0 + step, q1 = a1 / L, r1 = a1 % L;
      const bool two = s0 + 1 < L, d_lo = q1 - q0 == dlo;
      slot[static_cast<size_t>(t)] = static_cast<uint16_t>(s0);
      qs[static_cast<size_t>(t)] = static_cast<uint16_t>(q0);
      flags[static_cast<size_t>(t)] = static_cast<uint8_t>((d_lo ? 1 : 0) | (two ? 2 : 0));
      for (int k = 0; k < n; ++k)
        coef[static_cast<size_t>(k) * kFusedPolyThreads + t] = static_cast<float>(design.poly_bank[static_cast<size_t>(r0) * n + k]);
      // window position dlo + j (j = 0 .. n) carries tap j of the second output when d == dlo, tap j - 1 when d == dlo + 1
      for (int j = 0; j <= n; ++j) {
        const int k = d_lo ? j : j - 1;
        if (two && k >= 0 && k < n)
          coef[static_cast<size_t>(n + j) * kFusedPolyThreads + t] = static_cast<float>(design.poly_bank[static_cast<size_t>(r1) * n + k]);
      }
    }
    FusedTab &ft = fused_tab_[i];
    int rc;
    if ((rc = upload(coef, &ft.coef)) || (rc = upload(slot, &ft.slot)) || (rc = upload(qs, &ft.qs)) || (rc = upload(flags, &ft.flags))) return rc;
    if ((rc = be_sync(0))) return rc;
    ft.built = true;
    return RR_OK;
  }

  int ensure_big_scratch(size_t slice_bytes)
  {
    if (big_scratch_ && big_slice_ >= slice_bytes) return RR_OK;
    void *p = nullptr;
    int rc = be_malloc(&p, slice_bytes * kBigCtas + 256);
    if (rc != RR_OK) return rc;
    allocs_.push_back(p);                    // an outgrown area stays allocated until the engine goes (launches may be in flight)
    big_scratch_ = p; big_slice_ = slice_bytes;
    return RR_OK;
  }

  struct DevSched { CfftSched fwd, inv; const T *pyramid; const uint16_t *pk_ltab = nullptr, *pk_perm[2] = {nullptr, nullptr}; int pk_ltab_len = 0; };
  std::map<int, std::vector<uint16_t>> pk_perm_inv_host_;   // by complex bits
  const PkSpecConst *pk_spec_dev_[RR_MAX_STAGES] = {nullptr};
  std::map<int, DevSched> sched_;          // by complex bits
  std::map<int, const T *> tcos_;          // by real bits
  // debugging switches: generic kernels only / per stage kind
  bool use_pair_kernel_ = getenv("B200RATE_NO_PAIR_KERNEL") == nullptr;
  bool use_pair_dft_ = getenv("B200RATE_NO_PAIR_DFT") == nullptr, use_pair_poly_ = getenv("B200RATE_NO_PAIR_POLY") == nullptr;
  // outputs per CTA tile of halfband_pair_kernel: two work units per thread and tile halve the barriers per output
  // (cfg5 stages 0 / 1: 1.095 + 0.660 / 0.832 + 0.496 / 0.805 + 0.441 / 0.913 + 0.518 ms for 1024 / 2048 / 4096 / 8192).
  // The generic kernel's fp64 instance likewise (1.210 -> 1.089 ms on cfg3 stage 0); its window travels through kHbRaw
  // registers per thread, which bounds the tile.
  int half_pair_tile_ = getenv("B200RATE_HALF_TILE") ? std::max(512, std::min(8192, atoi(getenv("B200RATE_HALF_TILE")))) : 2 * kHalfTile;
  bool use_pair_shift_ = getenv("B200RATE_NO_PAIR_SHIFT") == nullptr;   // probes: the old deal with overflow into holes
  bool use_pair_half_ = getenv("B200RATE_NO_PAIR_HALF") == nullptr, use_pair_poly2_ = getenv("B200RATE_NO_PAIR_POLY2") == nullptr;
  int last_dft_kernel_ = 0;

  // Parameters of the lane-pair kernel for DFT stage i, or false when it does not apply (transform too
  // large for its shared-memory layout).
  bool make_pair_params(int i, const DftParams<float> &p, DftPkParams &pp, int halo_slots = 0, long long nwork = 0)
  {
    const StageGeom &g = geom[i];
    const int fb = ilog2(g.Pf) - 1, ib = ilog2(g.Ni) - 1;
    if (fb < 6 || ib < 6 || fb > 13 || ib > 13) return false;
    auto sf = sched_.find(fb), si = sched_.find(ib);
    if (sf == sched_.end() || si == sched_.end() || !sf->second.pk_ltab || !si->second.pk_ltab) return false;
    pp.base = p;
    pp.fb = fb; pp.ib = ib;
    pp.ltab_f = sf->second.pk_ltab; pp.ltab_i = si->second.pk_ltab;
    pp.perm_f = sf->second.pk_perm[0]; pp.perm_i = si->second.pk_perm[1];
    pp.fslots = pk_buf_slots(g.Pf >> 1); pp.bslots = pk_buf_slots(g.Ni >> 1); pp.halo_slots = halo_slots;
    pp.gthreads = kPkGroupThreads;
    pp.spec_mode = pk_spec_mode(g);
    pp.stereo = p.in.ch_stride == 1 && p.in.elem_stride == 2 && !(reinterpret_cast<size_t>(p.in.base) & 7) &&
                !(p.in.stream_stride & 1) && g.in_mode != DFT_IN_ZERO_STUFF;
    pp.spec = pk_spec_dev_[i];
    if (pp.spec_mode != PK_SPEC_GEN && !pp.spec) return false;
    pp.n_pyr_f = pk_pyr_len(fb); pp.n_pyr_i = pk_pyr_len(ib); pp.n_ltab_f = sf->second.pk_ltab_len; pp.n_ltab_i = si->second.pk_ltab_len;
    // one buffer per group and more groups per SM where the kernel exists (x2 up-sampling with N = 4096, no fused
    // polyphase stage): the forward transform works in the first slots of the inverse buffer
    if (halo_slots == 0 && pk_inplace_groups_ > kPkMaxGroups && pp.spec_mode == PK_SPEC_UP2 && fb == 10 && ib == 11) {
      pp.fslots = 0;
      pp.bslots = std::max(pk_buf_slots(g.Pf >> 1), pk_buf_slots(g.Ni >> 1));
      pp.groups = pk_inplace_groups_;
      // small launches (one stream): six groups per SM when the whole rounds of items that saves outweigh the slower
      // rounds (six groups sharing an SM take 1.25x as long per round as five: 1.875 vs 1.80 ms in steady state)
      if (pp.groups == kPkInplaceGroups && nwork > 0 && !getenv("B200RATE_DFT_GROUPS")) {
        const long long r5 = (nwork + 5ll * num_sms_ - 1) / (5ll * num_sms_), r6 = (nwork + 6ll * num_sms_ - 1) / (6ll * num_sms_);
        if (r6 * 5 < r5 * 4) pp.groups = 6;
      }
      if (pk_smem_layout(pp).total + 2048 <= max_smem_) {
        const PkSmemLayout lay = pk_smem_layout(pp);
        pp.lay_pyr_f = lay.pyr_f; pp.lay_pyr_i = lay.pyr_i; pp.lay_ltab_f = lay.ltab_f; pp.lay_ltab_i = lay.ltab_i;
        pp.lay_perm_f = lay.perm_f; pp.lay_data = lay.data;
        return true;
      }
      pp.fslots = pk_buf_slots(g.Pf >> 1); pp.bslots = pk_buf_slots(g.Ni >> 1);
    }
    for (pp.groups = kPkMaxGroups; pp.groups >= 1; --pp.groups)
      if (pk_smem_layout(pp).total + 1024 <= max_smem_) break;
    if (pp.groups < 1) return false;
    const PkSmemLayout lay = pk_smem_layout(pp);
    pp.lay_pyr_f = lay.pyr_f; pp.lay_pyr_i = lay.pyr_i; pp.lay_ltab_f = lay.ltab_f; pp.lay_ltab_i = lay.ltab_i;
    pp.lay_perm_f = lay.perm_f; pp.lay_data = lay.data;
    return true;
  }
  static int pk_spec_mode(const StageGeom &g)
  {
    if (g.in_mode == DFT_IN_FREQ_UP && g.L == 2 && g.step == 1) return PK_SPEC_UP2;
    if (g.Ni == g.Pf && g.step >= 1 && g.in_mode != DFT_IN_FREQ_UP) return PK_SPEC_SAME;
    return PK_SPEC_GEN;
  }

  // Per-index constants of the fused spectrum phase of stage i (modes UP2 / SAME): record i (i = 0 stands
  // for M/2 and also carries the two real bins). Needs the filter spectrum (make_spectrum).
  int build_pk_spec(int i)
  {
    const StageGeom &g = geom[i];
    const int mode = pk_spec_mode(g);
    const int fb = ilog2(g.Pf) - 1, ib = ilog2(g.Ni) - 1;
    if (mode == PK_SPEC_GEN || fb < 6 || ib < 6 || fb > 13 || ib > 13) return RR_OK;
    const int N = g.N, M = g.Pf >> 1, Mi = g.Ni >> 1, n = M >> 1;
    std::vector<float> spec(static_cast<size_t>(N));
    if (dft_spectrum_host(g.filter, spec.data(), N) != N) return RR_INTERNAL;
    const C2<float> *coef = reinterpret_cast<const C2<float> *>(spec.data());
    const std::vector<float> tf = cos_quarter_table<float>(ilog2(g.Pf)), ti = cos_quarter_table<float>(ilog2(g.Ni));
    const std::vector<uint16_t> &perm = pk_perm_inv_host_[ib];
    std::vector<PkSpecConst> rec(static_cast<size_t>(n) + 1);
    memset(rec.data(), 0, rec.size() * sizeof(PkSpecConst));
    for (int k = 0; k < n; ++k) {
      const int ii = k ? k : n;
      PkSpecConst &r = rec[static_cast<size_t>(k)];
      r.tfc = tf[static_cast<size_t>(ii)]; r.tfs = tf[static_cast<size_t>(n - ii)];
      if (mode == PK_SPEC_UP2) {
        r.c0 = coef[ii]; r.c1 = coef[Mi - ii]; r.c2 = coef[M - ii]; r.c3 = coef[M + ii];
        r.tic = ti[static_cast<size_t>(ii)]; r.tis = ti[static_cast<size_t>(M - ii)];
        r.s01 = perm[static_cast<size_t>(ii)] | (static_cast<unsigned>(perm[static_cast<size_t>(Mi - ii)]) << 16);
        r.s23 = perm[static_cast<size_t>(M - ii)] | (static_cast<unsigned>(perm[static_cast<size_t>(M + ii)]) << 16);
      } else {
        r.c0 = coef[ii]; r.c1 = coef[M - ii]; r.c2 = r.c0; r.c3 = r.c0;
        r.tic = r.tfc; r.tis = r.tfs;
        r.s01 = perm[static_cast<size_t>(ii)] | (static_cast<unsigned>(perm[static_cast<size_t>(M - ii)]) << 16);
        r.s23 = r.s01;
      }
    }
    PkSpecConst &sp = rec[0];                      // index 0 has no use for c2, c3, s23: the two real bins go there
    sp.c2 = coef[0];
    sp.s23 = perm[0];
    sp.c3 = C2<float>{0.f, 0.f};
    if (mode == PK_SPEC_UP2) { sp.c3 = coef[M]; sp.s23 |= static_cast<unsigned>(perm[static_cast<size_t>(M)]) << 16; }
    // where the record's two inputs live in the forward buffer, and whether it is the special index 0
    for (int k = 0; k < n; ++k) {
      PkSpecConst &r = rec[static_cast<size_t>(k)];
      r.fab = k ? static_cast<unsigned>(pk_slot(k)) | (static_cast<unsigned>(pk_slot(M - k)) << 16)
                : static_cast<unsigned>(pk_slot(0)) | (static_cast<unsigned>(pk_slot(n)) << 16);
      r.first = k == 0 ? 1u : 0u;
    }
    rec.resize(static_cast<size_t>(n));
    // Record e is handled by lane e % 32 of a warp: order the records so that the eight lanes of every quarter-warp
    // read and write eight different 16-byte bank groups (slot mod 8) in each of their two loads and two / four stores
    // -- as far as a greedy choice gets. In natural order the scattered stores cost 1.75 wavefronts where 1 would do.
    {
      auto slots_of = [&](const PkSpecConst &r, int *out) {
        int c = 0;
        out[c++] = r.fab & 0xffff; out[c++] = r.fab >> 16; out[c++] = r.s01 & 0xffff; out[c++] = r.s01 >> 16;
        if (mode == PK_SPEC_UP2) { out[c++] = r.s23 & 0xffff; out[c++] = r.s23 >> 16; }
        return c;
      };
      std::vector<PkSpecConst> pool(rec), ordered;
      std::vector<char> used(pool.size(), 0);
      for (size_t base = 0; base < pool.size(); base += 8) {
        int cnt[6][8];
        memset(cnt, 0, sizeof(cnt));
        for (size_t l = 0; l < 8 && base + l < pool.size(); ++l) {
          int best = -1, best_cost = 1 << 30;
          for (size_t c = 0; c < pool.size(); ++c) {
            if (used[c]) continue;
            int sl[6];
            const int ns = slots_of(pool[c], sl);
            int cost = 0;
            for (int a = 0; a < ns; ++a) cost += cnt[a][sl[a] & 7];
            if (cost < best_cost) { best_cost = cost; best = static_cast<int>(c); if (!cost) break; }
          }
          int sl[6];
          const int ns = slots_of(pool[static_cast<size_t>(best)], sl);
          for (int a = 0; a < ns; ++a) ++cnt[a][sl[a] & 7];
          used[static_cast<size_t>(best)] = 1;
          ordered.push_back(pool[static_cast<size_t>(best)]);
        }
      }
      rec.swap(ordered);
    }
    return upload(rec, &pk_spec_dev_[i]);
  }

  int dft_lanes_per_cta(const StageGeom &g, int nlanes) const
  {
    if (nlanes < 2) return 1;
    const size_t two = dft_smem_bytes<T>(g.Pf, g.Ni, 2, false, nullptr, nullptr, nullptr);
    return two <= 100 * 1024 ? 2 : 1;      // two lanes share every table / index load; keep 2 CTAs per SM
  }

  template <class E> int upload(const std::vector<E> &v, const E **dev)
  {
    void *p = nullptr;
    int rc = be_malloc(&p, v.size() * sizeof(E) + 16);
    if (rc != RR_OK) return rc;
    allocs_.push_back(p);
    if (!v.empty() && (rc = be_h2d(p, v.data(), v.size() * sizeof(E), 0)) != RR_OK) return rc;
    *dev = static_cast<const E *>(p);
    return RR_OK;
  }

  int get_sched(int bits, const DevSched **out)
  {
    auto it = sched_.find(bits);
    if (it == sched_.end()) {
      const CfftHostSched h = build_cfft_sched(bits);
      DevSched d{};
      int rc;
      const uint16_t *leaf16 = nullptr, *leaf8 = nullptr, *g16[2] = {nullptr, nullptr}, *g8[2] = {nullptr, nullptr}, *nodes = nullptr;
      if ((rc = upload(h.leaf16_off, &leaf16)) || (rc = upload(h.leaf8_off, &leaf8)) ||
          (rc = upload(h.gather16[0], &g16[0])) || (rc = upload(h.gather16[1], &g16[1])) ||
          (rc = upload(h.gather8[0], &g8[0])) || (rc = upload(h.gather8[1], &g8[1])) ||
          (rc = upload(h.node_off, &nodes)) || (rc = upload(twiddle_pyramid<T>(h), &d.pyramid)))
        return rc;
      for (int inv = 0; inv < 2; ++inv) {
        CfftSched &c = inv ? d.inv : d.fwd;
        c.bits = bits;
        c.n16 = static_cast<int>(h.leaf16_off.size()); c.n8 = static_cast<int>(h.leaf8_off.size());
        c.leaf16_off = leaf16; c.leaf8_off = leaf8; c.gather16 = g16[inv]; c.gather8 = g8[inv]; c.node_off = nodes;
        c.pyr_len = h.pyr_len;
        for (int l = 0; l < 17; ++l) {
          c.level_begin[l] = h.level_begin[l]; c.level_cnt[l] = h.level_cnt[l]; c.pyr_off[l] = h.pyr_off[l];
          c.qchild_begin[l] = h.qchild_begin[l]; c.qchild_cnt[l] = h.qchild_cnt[l];
        }
      }
      if (std::is_same<T, float>::value && bits >= 6 && bits <= 13) {
        const PkHostSched ph = build_pk_sched(h);
        if ((rc = upload(ph.local, &d.pk_ltab)) || (rc = upload(ph.perm[0], &d.pk_perm[0])) || (rc = upload(ph.perm[1], &d.pk_perm[1])))
          return rc;
        d.pk_ltab_len = static_cast<int>(ph.local.size());
        pk_perm_inv_host_[bits] = ph.perm[1];
      }
      it = sched_.emplace(bits, d).first;
    }
    *out = &it->second;
    return RR_OK;
  }

  int get_tcos(int real_bits, const T **out)
  {
    auto it = tcos_.find(real_bits);
    if (it == tcos_.end()) {
      const T *dev = nullptr;
      int rc = upload(cos_quarter_table<T>(real_bits), &dev);
      if (rc) return rc;
      it = tcos_.emplace(real_bits, dev).first;
    }
    *out = it->second;
    return RR_OK;
  }

  static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

  int upload_tables()
  {
    int rc;
    for (int i = 0; i < ns; ++i) {
      const StageGeom &g = geom[i];
      if (g.kind == RR_STAGE_HALFBAND) {
        HalfbandParams<T> &p = half_params_[i];
        memset(&p, 0, sizeof(p));
        p.ncoef = g.hb_c; p.pre = design.plan.st[i].pre;
        const double *c = half_band_coefs(g.hb_c);
        for (int k = 0; k < g.hb_c; ++k) p.coef[k] = static_cast<T>(c[k]);
      } else if (g.kind == RR_STAGE_POLY) {
        PolyParams<T> &p = poly_params_[i];
        memset(&p, 0, sizeof(p));
        p.n = g.n; p.L = g.Lp; p.order = g.order; p.phase_bits = g.phase_bits; p.at0 = g.at0; p.step = g.pstep;
        p.pre = design.plan.st[i].pre;
        std::vector<T> bank(design.poly_bank.size());
        for (size_t k = 0; k < bank.size(); ++k) bank[k] = static_cast<T>(design.poly_bank[k]);
        if ((rc = upload(bank, &p.coefs))) return rc;
      } else {
        // smem budget check: the whole block lives in shared memory
        int xs = 0, ys = 0;
        DftParams<T> probe; memset(&probe, 0, sizeof(probe));
        probe.Pf = g.Pf; probe.Ni = g.Ni; probe.fwd.pyr_len = g.Pf / 4 + 16; probe.inv.pyr_len = g.Ni / 4 + 16;
        const size_t need = dft_smem_bytes<T>(g.Pf, g.Ni, 1, false, &xs, &ys, nullptr) + sizeof(T) * dft_table_elems(probe);
        if (ilog2(g.Pf) - 1 > 16 || ilog2(g.Ni) - 1 > 16) {   // the reference's own table limit (rate/rate_uni.c:134-189): N = 2^17
          set_last_error("DFT length " + std::to_string(g.N) + " exceeds the 131072-point limit of the FFT tables");
          return RR_INTERNAL;
        }
        dft_big_[i] = need > max_smem_;             // work buffers in global scratch instead of shared memory
        DftParams<T> &p = dft_params_[i];
        memset(&p, 0, sizeof(p));
        p.N = g.N; p.overlap = g.ov; p.L = g.L; p.step = g.step; p.in_mode = g.in_mode; p.Pf = g.Pf; p.Ni = g.Ni;
        p.remL0 = g.remL0; p.q = g.q; p.kept = g.kept;
        p.sqrthalf = sqrthalf_; p.c16_1 = c16_1_; p.c16_3 = c16_3_;
        const DevSched *sf = nullptr, *si = nullptr;
        if ((rc = get_sched(ilog2(g.Pf) - 1, &sf)) || (rc = get_sched(ilog2(g.Ni) - 1, &si))) return rc;
        p.fwd = sf->fwd; p.pyr_f = sf->pyramid;
        p.inv = si->inv; p.pyr_i = si->pyramid;
        if ((rc = get_tcos(ilog2(g.Pf), &p.tcos_f)) || (rc = get_tcos(ilog2(g.Ni), &p.tcos_i))) return rc;
        if (!dft_coef_dev_[g.filter] && (rc = make_spectrum(g.filter))) return rc;
        p.coef = dft_coef_dev_[g.filter];
        if constexpr (std::is_same<T, float>::value) { if ((rc = build_pk_spec(i))) return rc; }
      }
    }
    return be_sync(0);
  }

  // Forward real FFT of the wrapped/scaled taps, in the engine's sample type and with the engine's own
  // transform (rate/rate_base.h:172-184 does the same with lsx_safe_rdft): one block, COPY mode, stop
  // after the post-processing phase.
  int make_spectrum(int instance)
  {
    const DftFilterDesign &f = design.dft[instance];
    const int N = f.dft_length;
    std::vector<T> t(static_cast<size_t>(N));
    for (int k = 0; k < N; ++k) t[k] = static_cast<T>(f.coefs_time[k]);
    const T *time_dev = nullptr;
    int rc = upload(t, &time_dev);
    if (rc) return rc;
    void *spec = nullptr;
    if ((rc = be_malloc(&spec, sizeof(T) * static_cast<size_t>(N) + 16))) return rc;
    allocs_.push_back(spec);
    DftParams<T> p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.overlap = 0; p.L = 1; p.step = 0 /* spectrum only */; p.in_mode = DFT_IN_COPY; p.Pf = N; p.Ni = N;
    p.q = N; p.kept = N;
    p.sqrthalf = sqrthalf_; p.c16_1 = c16_1_; p.c16_3 = c16_3_;
    const DevSched *sf = nullptr;
    if ((rc = get_sched(ilog2(N) - 1, &sf))) return rc;
    p.fwd = sf->fwd; p.pyr_f = sf->pyramid; p.inv = sf->inv; p.pyr_i = sf->pyramid;
    if ((rc = get_tcos(ilog2(N), &p.tcos_f))) return rc;
    p.tcos_i = p.tcos_f;
    p.coef = nullptr;
    LaneView v{};
    v.origin = 0; v.mask = ~0ull; v.lo = 0; v.hi = N; v.stream_stride = 0; v.ch_stride = 0; v.elem_stride = 1; v.nch = 1;
    p.in = v; p.in.base = const_cast<T *>(time_dev);
    p.out = v; p.out.base = spec;
    p.block0 = 0; p.nblocks = 1; p.nlanes = 1;
    const size_t data_bytes = dft_smem_bytes<T>(N, N, 1, false, &p.xstride, &p.ystride, &p.zstride);
    void *big = nullptr;
    if (data_bytes + sizeof(T) * dft_table_elems(p) > max_smem_) {
      if ((rc = ensure_big_scratch(data_bytes))) return rc;
      big = big_scratch_;
    }
    if ((rc = Launch<T>::dft(p, 1, false, false, 1, data_bytes, 0, big, kBigCtas))) return rc;
    dft_coef_dev_[instance] = static_cast<T *>(spec);
    return RR_OK;
  }
};

// ===================================================================================================
// Batch front-end
// ===================================================================================================
template <class T> class Batch : public IBatch {
 public:
  Engine<T> eng;
  int nch = 0, nstreams = 0;
  size_t frames_in_max = 0;
  std::vector<long long> cap;        // per FIFO (1..ns-1): samples per lane
  std::vector<T *> buf;
  int launches_ = 0;
  bool timing_ = false;
  int active_streams_ = 0;           // streams processed by the current call (<= nstreams)
#ifndef B200RATE_EMU
  std::vector<cudaEvent_t> ev_;      // 2 per stage when timing is on
  cudaStream_t hs_[3] = {nullptr, nullptr, nullptr};   // host pipeline: H2D, compute, D2H
  cudaEvent_t h2d_done_[2] = {nullptr, nullptr}, comp_done_[2] = {nullptr, nullptr}, d2h_done_[2] = {nullptr, nullptr};
#endif
  float *slot_in_[2] = {nullptr, nullptr}, *slot_out_[2] = {nullptr, nullptr};
  std::vector<char> fused_;          // stage i runs fused with stage i + 1
  // lane pairs a batch needs before its DFT + vpoly0 stages run fused (fewer: not enough runs to fill the device)
  int kFuseMinPairs = getenv("B200RATE_FUSE_MIN_PAIRS") ? atoi(getenv("B200RATE_FUSE_MIN_PAIRS")) : 64;

  int ensure_fifo(int i)             // intermediate FIFO i (input of stage i), allocated on first use
  {
    if (i < 1 || i >= eng.ns || buf[i]) return RR_OK;
    void *p = nullptr;
    const size_t bytes = sizeof(T) * static_cast<size_t>(cap[i]) * nch * nstreams;
    const int rc = be_malloc(&p, bytes);
    if (rc) return rc;
    buf[i] = static_cast<T *>(p);
    return RR_OK;
  }

  ~Batch() override
  {
    DeviceScope scope(eng.device_id);
    for (T *p : buf) be_free(p);
    for (int k = 0; k < 2; ++k) { be_free(slot_in_[k]); be_free(slot_out_[k]); }
#ifndef B200RATE_EMU
    for (cudaEvent_t e : ev_) cudaEventDestroy(e);
    for (int k = 0; k < 3; ++k) if (hs_[k]) cudaStreamDestroy(hs_[k]);
    for (int k = 0; k < 2; ++k) {
      if (h2d_done_[k]) cudaEventDestroy(h2d_done_[k]);
      if (comp_done_[k]) cudaEventDestroy(comp_done_[k]);
      if (d2h_done_[k]) cudaEventDestroy(d2h_done_[k]);
    }
#endif
  }

  int init(const RR_config &cfg, int nchannels, int nstr, size_t fmax, int device)
  {
    const int dev = device >= 0 ? device : be_current_device();
    RR_DEVICE_SCOPE(dev);
    int rc = eng.init(cfg, dev);
    if (rc) return rc;
    nch = nchannels; nstreams = nstr; frames_in_max = fmax; active_streams_ = nstr;
    const size_t nout = frames_out(fmax);
    std::vector<StageRange> r(eng.ns);
    plan_ranges(0, static_cast<long long>(nout), r.data());
    cap.assign(eng.ns + 1, 0); buf.assign(eng.ns + 1, nullptr);
    fused_.assign(eng.ns + 1, 0);
    for (int i = 1; i < eng.ns; ++i) {
      const StageGeom &up = eng.geom[i - 1];
      long long slack = up.kind == RR_STAGE_DFT ? 2ll * up.N : 64;
      cap[i] = ((r[i - 1].prod_hi - r[i - 1].prod_lo) + slack + 3) & ~3ll;   // lanes stay 16-byte aligned
    }
    // A DFT stage and the vpoly0 stage behind it run as one kernel when the batch is large enough to fill the device
    // with runs of blocks and the DFT stage reads adjacent stereo frames; the FIFO between them is then never
    // allocated unless a call cannot use the fused kernel (ensure_fifo).
    for (int i = 0; i + 1 < eng.ns; ++i) {
      const bool stereo_in = i == 0 ? nch == 2 : pair_fifo(i);
      if (eng.fused_config_ok(i) && stereo_in && !(nch & 1) && static_cast<long long>(nch) * nstreams / 2 >= kFuseMinPairs) { fused_[i] = 1; ++i; }
    }
    for (int i = 1; i < eng.ns; ++i)
      if (!fused_[i - 1] && (rc = ensure_fifo(i))) return rc;
    return RR_OK;
  }

  const Design &design() const override { return eng.design; }
  const char *stage_kernel(int stage) const override
  {
    return stage >= 0 && stage < eng.ns && eng.kernel_name[stage] ? eng.kernel_name[stage] : "";
  }

  // FIFO i (input of stage i) is stored pair-interleaved -- [pair of lanes][sample][2], i.e. a view of two-channel
  // "streams" -- when the lane-pair kernels sit on both sides of it: the DFT stage then stores whole 16-byte
  // slots and the polyphase stage stages its window with 16-byte copies. Planar otherwise.
  bool pair_fifo(int i) const
  {
    if (!std::is_same<T, float>::value || (nch & 1) || i < 1 || i >= eng.ns || getenv("B200RATE_NO_PAIR_KERNEL")) return false;
    const StageGeom &prod = eng.geom[i - 1], &cons = eng.geom[i];
    return prod.kind == RR_STAGE_DFT && (cons.kind == RR_STAGE_DFT || (cons.kind == RR_STAGE_POLY && cons.order == 0));
  }

  size_t frames_out(size_t frames_in) const override      // rate_flush target, rate/rate_base.h:457
  {
    return static_cast<size_t>(static_cast<double>(frames_in) / eng.design.plan.factor + .5);
  }

  // Backward dependency pass: per-stage work for final outputs [klo, khi).
  void plan_ranges(long long klo, long long khi, StageRange *r) const
  {
    for (int i = eng.ns - 1; i >= 0; --i) {
      r[i] = stage_range_for_outputs(eng.geom[i], klo, khi);
      const long long pre = eng.geom[i].preload;
      klo = std::max(0ll, r[i].need_lo - pre);
      khi = std::max(0ll, r[i].need_hi - pre);
    }
  }

  void input_window(size_t frames_in, uint64_t out_begin, size_t out_count, uint64_t *first, uint64_t *count) const override
  {
    if (eng.ns == 0) {                       // identity: output frame k is input frame k
      const uint64_t lo = std::min<uint64_t>(out_begin, frames_in), hi = std::min<uint64_t>(out_begin + out_count, frames_in);
      if (first) *first = lo;
      if (count) *count = hi - lo;
      return;
    }
    std::vector<StageRange> r(eng.ns);
    const long long klo = static_cast<long long>(out_begin), khi = klo + static_cast<long long>(out_count);
    plan_ranges(klo, khi, r.data());
    const long long pre0 = eng.geom[0].preload;
    const long long lo = std::max(0ll, r[0].need_lo - pre0);
    const long long hi = std::min<long long>(static_cast<long long>(frames_in), std::max(0ll, r[0].need_hi - pre0));
    if (first) *first = static_cast<uint64_t>(lo);
    if (count) *count = static_cast<uint64_t>(std::max(0ll, hi - lo));
  }

  int process(const float *d_in, uint64_t win_first, size_t win_frames, size_t frames_in, uint64_t out_begin,
              size_t out_count, void *d_out, bool native_out, void *stream) override
  {
    RR_DEVICE_SCOPE(eng.device_id);
    stream_t s = static_cast<stream_t>(stream);
    const int ns = eng.ns, nlanes = nch * active_streams_;
    eng.launches = 0;
#ifndef B200RATE_EMU
    if (timing_ && ev_.size() < static_cast<size_t>(2 * ns)) {
      while (ev_.size() < static_cast<size_t>(2 * ns)) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); ev_.push_back(e); }
    }
#endif
    const long long klo = static_cast<long long>(out_begin), khi = klo + static_cast<long long>(out_count);
    if (ns == 0) {
      // in_rate == out_rate: no stages, FIFO 0 is the output FIFO (rate_base.h:445-447) -- the frames pass through
      if (!d_out || out_count == 0) return RR_OK;
      LaneView in{}, out{};
      in.base = const_cast<float *>(d_in);
      in.origin = static_cast<long long>(win_first); in.mask = ~0ull; in.lo = in.origin;
      in.hi = std::min<long long>(static_cast<long long>(frames_in), static_cast<long long>(win_first + win_frames));
      in.stream_stride = static_cast<long long>(win_frames) * nch; in.ch_stride = 1; in.elem_stride = nch; in.nch = nch;
      out.base = d_out; out.origin = klo; out.mask = ~0ull; out.lo = klo; out.hi = khi; out.nch = nch;
      out.stream_stride = static_cast<long long>(out_count) * nch;
      if (native_out) { out.ch_stride = static_cast<int>(out_count); out.elem_stride = 1; }
      else { out.ch_stride = 1; out.elem_stride = nch; }
      const int rc = eng.copy(in, true, out, !native_out, klo, static_cast<long long>(out_count), 0, nlanes, s);
      launches_ = eng.launches;
      return rc;
    }
    std::vector<StageRange> r(ns);
    plan_ranges(klo, khi, r.data());
    const long long pre0 = eng.geom[0].preload;
    if (!d_out || out_count == 0) return RR_OK;
    {
      uint64_t f = 0, c = 0;
      input_window(frames_in, out_begin, out_count, &f, &c);
      if (c && (f < win_first || f + c > win_first + win_frames)) {
        set_last_error("input window does not cover the samples this output range depends on");
        return RR_INVPARAM;
      }
    }

    // views of the FIFO on either side of stage i for this call
    auto view_in = [&](int i, LaneView &in, bool &in_f32) -> int {
      if (i == 0) {
        in.base = const_cast<float *>(d_in);
        in.origin = pre0 + static_cast<long long>(win_first); in.mask = ~0ull;
        in.lo = in.origin;
        in.hi = pre0 + std::min<long long>(static_cast<long long>(frames_in), static_cast<long long>(win_first + win_frames));
        in.stream_stride = static_cast<long long>(win_frames) * nch; in.ch_stride = 1; in.elem_stride = nch; in.nch = nch;
        in_f32 = true;
        return RR_OK;
      }
      int rc = ensure_fifo(i);
      if (rc) return rc;
      const long long pre = eng.geom[i].preload;
      in.base = buf[i];
      in.origin = pre + r[i - 1].prod_lo; in.mask = ~0ull; in.lo = pre; in.hi = pre + r[i - 1].prod_hi;
      in.stream_stride = cap[i] * nch; in.ch_stride = static_cast<int>(cap[i]); in.elem_stride = 1; in.nch = nch;
      if (pair_fifo(i)) { in.stream_stride = 2 * cap[i]; in.ch_stride = 1; in.elem_stride = 2; in.nch = 2; }
      if (cap[i] > 0x7fffffffll) { set_last_error("intermediate lane too long for one batch"); return RR_INVPARAM; }
      in_f32 = false;
      return RR_OK;
    };
    auto view_out = [&](int i, LaneView &out, bool &out_f32, long long &out_preload) -> int {
      if (i == ns - 1) {
        out.base = d_out;
        out.origin = klo; out.mask = ~0ull; out.lo = klo; out.hi = khi;
        if (native_out) { out.stream_stride = static_cast<long long>(out_count) * nch; out.ch_stride = static_cast<int>(out_count); out.elem_stride = 1; }
        else { out.stream_stride = static_cast<long long>(out_count) * nch; out.ch_stride = 1; out.elem_stride = nch; }
        out.nch = nch;
        out_f32 = !native_out;
        out_preload = 0;
        return RR_OK;
      }
      int rc = ensure_fifo(i + 1);
      if (rc) return rc;
      const long long pre = eng.geom[i + 1].preload;
      if (r[i].prod_hi - r[i].prod_lo > cap[i + 1]) { set_last_error("range exceeds the batch's frames_in_max"); return RR_INVPARAM; }
      out.base = buf[i + 1];
      out.origin = pre + r[i].prod_lo; out.mask = ~0ull; out.lo = out.origin; out.hi = pre + r[i].prod_hi;
      out.stream_stride = cap[i + 1] * nch; out.ch_stride = static_cast<int>(cap[i + 1]); out.elem_stride = 1; out.nch = nch;
      if (pair_fifo(i + 1)) { out.stream_stride = 2 * cap[i + 1]; out.ch_stride = 1; out.elem_stride = 2; out.nch = 2; }
      out_f32 = false;
      out_preload = pre;
      return RR_OK;
    };
    for (int i = 0; i < ns; ++i) {
      LaneView in{}, out{};
      bool in_f32 = false, out_f32 = false;
      long long out_preload = 0;
      int rc = view_in(i, in, in_f32);
      if (rc) return rc;
#ifndef B200RATE_EMU
      if (timing_) CUDA_TRY(cudaEventRecord(ev_[2 * i], s));
#endif
      if (fused_[i]) {
        // stage i (DFT) and stage i + 1 (vpoly0) in one kernel: FIFO i + 1 stays on chip
        if ((rc = view_out(i + 1, out, out_f32, out_preload))) return rc;
        rc = (native_out && i + 1 == ns - 1) ? RR_RATEERROR
                                             : eng.run_fused(i, in, out, out_preload, r[i].w0, r[i].wn, r[i + 1].w0, r[i + 1].wn, nlanes, s);
        if (rc == RR_OK) {
#ifndef B200RATE_EMU
          if (timing_) {
            CUDA_TRY(cudaEventRecord(ev_[2 * i + 1], s));
            CUDA_TRY(cudaEventRecord(ev_[2 * i + 2], s));
            CUDA_TRY(cudaEventRecord(ev_[2 * i + 3], s));
          }
#endif
          ++i;
          continue;
        }
        if (rc != RR_RATEERROR) return rc;              // RR_RATEERROR: not applicable to this call's buffers -> two kernels
        out = LaneView{};
      }
      if ((rc = view_out(i, out, out_f32, out_preload))) return rc;
      rc = eng.run_stage(i, in, in_f32, out, out_f32, out_preload, r[i].w0, r[i].wn, nlanes, s);
      if (rc) return rc;
#ifndef B200RATE_EMU
      if (timing_) CUDA_TRY(cudaEventRecord(ev_[2 * i + 1], s));
#endif
    }
    launches_ = eng.launches;
    return RR_OK;
  }

  int process_streams(const float *d_in, size_t frames_in, void *d_out, int nstreams_now, void *stream) override
  {
    if (nstreams_now < 1 || nstreams_now > nstreams) { set_last_error("bad active stream count"); return RR_INVPARAM; }
    active_streams_ = nstreams_now;
    const int rc = process(d_in, 0, frames_in, frames_in, 0, frames_out(frames_in), d_out, false, stream);
    active_streams_ = nstreams;
    return rc;
  }

  int process_host(const float *h_in, size_t frames_in, float *h_out, size_t total_streams) override
  {
    RR_DEVICE_SCOPE(eng.device_id);
    if (frames_in > frames_in_max) { set_last_error("frames_in exceeds frames_in_max"); return RR_INVPARAM; }
    const size_t nout = frames_out(frames_in);
    const size_t in_elems = frames_in * nch, out_elems = nout * nch;
    int rc;
    for (int k = 0; k < 2; ++k) {              // the four staging slots are allocated independently (a failure leaves the rest usable)
      void *p = nullptr;
      if (!slot_in_[k]) {
        if ((rc = be_malloc(&p, sizeof(float) * frames_in_max * nch * nstreams))) return rc;
        slot_in_[k] = static_cast<float *>(p);
      }
      if (!slot_out_[k]) {
        if ((rc = be_malloc(&p, sizeof(float) * (frames_out(frames_in_max) + 1) * nch * nstreams))) return rc;
        slot_out_[k] = static_cast<float *>(p);
      }
    }
#ifdef B200RATE_EMU
    for (size_t s0 = 0; s0 < total_streams; s0 += nstreams) {
      const int now = static_cast<int>(std::min<size_t>(nstreams, total_streams - s0));
      memcpy(slot_in_[0], h_in + s0 * in_elems, sizeof(float) * in_elems * now);
      if ((rc = process_streams(slot_in_[0], frames_in, slot_out_[0], now, nullptr))) return rc;
      memcpy(h_out + s0 * out_elems, slot_out_[0], sizeof(float) * out_elems * now);
    }
    return RR_OK;
#else
    if (!hs_[0]) {
      for (int k = 0; k < 3; ++k) CUDA_TRY(cudaStreamCreateWithFlags(&hs_[k], cudaStreamNonBlocking));
      for (int k = 0; k < 2; ++k) {
        CUDA_TRY(cudaEventCreateWithFlags(&h2d_done_[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&comp_done_[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d2h_done_[k], cudaEventDisableTiming));
      }
    }
    int total_launches = 0;
    // H2D, kernels and D2H of consecutive sub-batches overlap on three streams; on any failure the streams are
    // drained before returning, so the caller's host buffers are never still in flight
    auto pipeline = [&]() -> int {
      size_t k = 0;
      for (size_t s0 = 0; s0 < total_streams; s0 += nstreams, ++k) {
        const int slot = static_cast<int>(k & 1);
        const int now = static_cast<int>(std::min<size_t>(nstreams, total_streams - s0));
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(hs_[0], comp_done_[slot], 0));       // input slot free again
        CUDA_TRY(cudaMemcpyAsync(slot_in_[slot], h_in + s0 * in_elems, sizeof(float) * in_elems * now,
                                 cudaMemcpyHostToDevice, hs_[0]));
        CUDA_TRY(cudaEventRecord(h2d_done_[slot], hs_[0]));
        CUDA_TRY(cudaStreamWaitEvent(hs_[1], h2d_done_[slot], 0));
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(hs_[1], d2h_done_[slot], 0));        // output slot drained
        const int prc = process_streams(slot_in_[slot], frames_in, slot_out_[slot], now, hs_[1]);
        if (prc) return prc;
        total_launches += launches_;
        CUDA_TRY(cudaEventRecord(comp_done_[slot], hs_[1]));
        CUDA_TRY(cudaStreamWaitEvent(hs_[2], comp_done_[slot], 0));
        CUDA_TRY(cudaMemcpyAsync(h_out + s0 * out_elems, slot_out_[slot], sizeof(float) * out_elems * now,
                                 cudaMemcpyDeviceToHost, hs_[2]));
        CUDA_TRY(cudaEventRecord(d2h_done_[slot], hs_[2]));
      }
      return RR_OK;
    };
    rc = pipeline();
    for (int q = 0; q < 3; ++q) {
      const cudaError_t e = cudaStreamSynchronize(hs_[q]);
      if (e != cudaSuccess && rc == RR_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    launches_ = total_launches;
    return rc;
#endif
  }

  int last_launches() const override { return launches_; }

  void enable_timing(bool on) override { timing_ = on; }

  int stage_times(float *ms, int max_stages) override
  {
    DeviceScope scope(eng.device_id);
    const int n = std::min(eng.ns, max_stages);
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
#ifndef B200RATE_EMU
    if (!timing_ || ev_.size() < static_cast<size_t>(2 * eng.ns)) return 0;
    for (int i = 0; i < n; ++i) {
      if (cudaEventSynchronize(ev_[2 * i + 1]) != cudaSuccess) return 0;
      if (cudaEventElapsedTime(&ms[i], ev_[2 * i], ev_[2 * i + 1]) != cudaSuccess) return 0;
    }
#endif
    return n;
  }

  int stage_work(size_t frames_in, int stage, double *flops_out, double *bytes_out, double *units_out) const override
  {
    if (stage < 0 || stage >= eng.ns) return RR_INVPARAM;
    std::vector<StageRange> r(eng.ns);
    plan_ranges(0, static_cast<long long>(frames_out(frames_in)), r.data());
    const double lanes = static_cast<double>(nch) * nstreams;
    const double in_bytes = stage == 0 ? 4. : sizeof(T), out_bytes = stage == eng.ns - 1 ? 4. : sizeof(T);
    // unique input samples: the coordinates the work range reads, clipped to what really exists upstream
    double in_samples = static_cast<double>(r[stage].need_hi - std::max<long long>(r[stage].need_lo, eng.geom[stage].preload));
    if (stage == 0) in_samples = std::min<double>(in_samples, static_cast<double>(frames_in));
    const double out_samples = stage == eng.ns - 1 ? static_cast<double>(frames_out(frames_in))
                                                   : static_cast<double>(r[stage].prod_hi - r[stage].prod_lo);
    if (flops_out) *flops_out = eng.stage_flops(stage, r[stage]) * lanes;
    if (bytes_out) *bytes_out = (in_samples * in_bytes + out_samples * out_bytes) * lanes;
    if (units_out) *units_out = static_cast<double>(r[stage].wn) * lanes;
    return RR_OK;
  }

  double flops(size_t frames_in) const override
  {
    std::vector<StageRange> r(eng.ns);
    plan_ranges(0, static_cast<long long>(frames_out(frames_in)), r.data());
    return eng.flops_of(r.data()) * nch * nstreams;
  }
};

// ===================================================================================================
// Streaming front-end (RR_push / RR_pull / RR_drain / RR_flow on device ring buffers)
//
// FIFO 0 and the output FIFO are rings of INTERLEAVED FLOAT frames -- the caller's own layout -- so a push is a
// plain host-to-device copy into ring 0 and a pull a plain device-to-host copy out of the last ring: the
// (de)interleave and the float <-> sample type conversion of rate/rate_base.h:559-569 happen inside the first /
// last stage kernel, there is no separate copy pass. Host buffers are staged through page-locked memory owned
// by the handle (two input slots, so the copy-in of push k+1 overlaps the transfer and the kernels of push k);
// caller buffers that are themselves page-locked are used directly. RR_push returns as soon as the caller's
// buffer has been consumed (never while a copy from it is in flight); the kernels run on the handle's own
// non-blocking stream and RR_pull / RR_drain wait for them, so a failure of a kernel launched by RR_push is
// reported by the next RR_pull / RR_drain.
// ===================================================================================================
template <class T> class Stream : public IStream {
 public:
  Engine<T> eng;
  int nch = 0;
  // FIFO state, absolute coordinates (identical for every channel)
  std::vector<long long> W;          // samples written to FIFO i (including preload)
  std::vector<long long> done;       // stage progress: blocks (dft) or outputs (others)
  std::vector<long long> produced;   // outputs produced by stage i
  std::vector<long long> ring_cap;   // per-lane capacity (power of two)
  std::vector<void *> ring;          // 0 and ns: interleaved float; 1 .. ns-1 (and ns with the native tap): planar T
  bool native_tap_ = false;          // last FIFO kept in the engine type (RRX_pull_native of the fp64 engine)
  long long popped = 0;              // read position of the output FIFO
  long long out_shift = 0;           // coordinate of last-stage output 0 in the output FIFO (changes when drain trims)
  uint64_t samples_in = 0, samples_out = 0;   // rate_t counters (rate_base.h:224-231)
  uint64_t in_rate_ = 0, out_rate_ = 0;
  stream_t s_ = nullptr;             // this handle's own (non-blocking) stream: handles on different host threads overlap
  // page-locked staging
  float *pin_in_[2] = {nullptr, nullptr};
  size_t pin_in_cap_[2] = {0, 0};
  event_t pin_in_free_[2] = {};      // recorded after the transfer out of the slot
  bool pin_in_busy_[2] = {false, false};
  unsigned push_seq_ = 0;
  float *pin_out_ = nullptr;
  size_t pin_out_cap_ = 0;
  event_t user_in_done_ = {};        // transfer out of a caller-owned page-locked buffer
  bool events_ok_ = false;
  T *stage_native = nullptr; size_t stage_native_cap = 0;   // device staging of the native tap
  std::vector<void *> retired_;      // outgrown rings

  void release_retired() { for (void *p : retired_) be_free(p); retired_.clear(); }   // only when the stream is idle

  // B200RATE_TRACE_STREAM=1: host-side time of the steps of push / pull, summed per handle and printed at close
  bool trace_ = getenv("B200RATE_TRACE_STREAM") != nullptr;
  double tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tr_calls_[2] = {0, 0};
  static double tick() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

  ~Stream() override
  {
    DeviceScope scope(eng.device_id);
    if (s_) be_sync(s_);
    if (trace_ && tr_calls_[0] && tr_calls_[1])
      fprintf(stderr, "stream trace: %lld pushes: enqueue copy %.1f us, launch stages %.1f us, wait for the copy %.1f us | %lld pulls: "
                      "enqueue copy %.1f us, wait %.1f us\n", tr_calls_[0], tr_[0] / tr_calls_[0], tr_[1] / tr_calls_[0], tr_[2] / tr_calls_[0],
              tr_calls_[1], tr_[3] / tr_calls_[1], tr_[4] / tr_calls_[1]);
    release_retired();
    for (void *p : ring) be_free(p);
    be_free(stage_native);
    for (int k = 0; k < 2; ++k) be_host_free(pin_in_[k]);
    be_host_free(pin_out_);
    if (events_ok_) { be_event_destroy(pin_in_free_[0]); be_event_destroy(pin_in_free_[1]); be_event_destroy(user_in_done_); }
    be_stream_destroy(s_);
  }

  int init(const RR_config &cfg, int nchannels, int device)
  {
    const int dev = device >= 0 ? device : be_current_device();
    RR_DEVICE_SCOPE(dev);
    int rc = eng.init(cfg, dev);
    if (rc) return rc;
    if ((rc = be_stream_create(&s_))) return rc;
    if ((rc = be_event_create(&pin_in_free_[0])) || (rc = be_event_create(&pin_in_free_[1])) || (rc = be_event_create(&user_in_done_)))
      return rc;
    events_ok_ = true;
    nch = nchannels;
    const int ns = eng.ns;
    W.assign(ns + 1, 0); done.assign(ns + 1, 0); produced.assign(ns + 1, 0);
    ring_cap.assign(ns + 1, 0); ring.assign(ns + 1, nullptr);
    for (int i = 0; i < ns; ++i) W[i] = eng.geom[i].preload;
    // the caller-facing rings start large enough for the plugin's chunks (65536 + 2 x 8192 frames, foo_dsp_rate.cpp:165-187)
    for (int i = 0; i <= ns; ++i)
      if ((rc = ensure_ring(i, std::max<long long>(W[i], (i == 0 || i == ns) ? (1 << 17) : (1 << 14))))) return rc;
    return RR_OK;
  }

  const Design &design() const override { return eng.design; }

  // ring i holds interleaved float frames (the caller's layout) rather than planar engine-type lanes
  bool ring_is_frames(int i) const { return i == 0 || (i == eng.ns && !native_tap_); }
  size_t ring_elem_bytes(int i) const { return ring_is_frames(i) ? sizeof(float) : sizeof(T); }

  LaneView ring_view(int i, long long lo, long long hi) const
  {
    LaneView v{};
    v.base = ring[i]; v.origin = 0; v.mask = static_cast<unsigned long long>(ring_cap[i] - 1);
    v.lo = lo; v.hi = hi;
    v.stream_stride = 0; v.nch = nch;
    if (ring_is_frames(i)) { v.ch_stride = 1; v.elem_stride = nch; }
    else { v.ch_stride = static_cast<int>(ring_cap[i]); v.elem_stride = 1; }
    return v;
  }

  // lowest coordinate of FIFO i that may still be read
  long long keep_from(int i) const
  {
    if (i == eng.ns) return popped;
    const StageGeom &g = eng.geom[i];
    if (g.kind == RR_STAGE_HALFBAND) return 2 * done[i] + 1;
    if (g.kind == RR_STAGE_DFT) return dft_Rb(g, done[i]);
    return poly_q(g, done[i]);
  }

  // make ring i able to hold coordinates [keep_from(i), upto)
  int ensure_ring(int i, long long upto)
  {
    const long long keep = ring[i] ? keep_from(i) : 0;
    const long long need = upto - keep;
    if (ring[i] && need <= ring_cap[i]) return RR_OK;
    long long cap = std::max<long long>(ring_cap[i], 1 << 14);
    while (cap < need) cap <<= 1;
    if (cap > 0x40000000ll) { set_last_error("FIFO ring would exceed 2^30 samples per channel"); return RR_ENOMEM; }
    void *p = nullptr;
    const size_t bytes = ring_elem_bytes(i) * static_cast<size_t>(cap) * nch;
    int rc = be_malloc(&p, bytes);
    if (rc) return rc;
    // a fresh ring is all zeros: the preload region (fifo_write0, rate_base.h:420) and everything not yet written
    if ((rc = be_memset(p, 0, bytes, s_))) { be_free(p); return rc; }
    if (ring[i]) {                           // carry the live region over (ring -> bigger ring)
      const bool f32 = ring_is_frames(i);
      LaneView from = ring_view(i, keep, W[i]);
      void *old = ring[i];
      ring[i] = p; ring_cap[i] = cap;
      LaneView to = ring_view(i, keep, W[i]);
      rc = eng.copy(from, f32, to, f32, keep, W[i] - keep, 0, nch, s_);
      retired_.push_back(old);               // still read by the copy just queued: released at the next point the stream is idle
      return rc;
    }
    ring[i] = p; ring_cap[i] = cap;
    return RR_OK;
  }

  void count_input(size_t n)                 // rate_input, rate_base.h:436-441
  {
    samples_in += n;
    while (samples_in > in_rate_ && samples_out > out_rate_) { samples_in -= in_rate_; samples_out -= out_rate_; }
  }

  // new totals for stage i given W[i]; returns outputs produced in total
  void advance_counts(int i, long long *new_done, long long *new_produced) const
  {
    const StageGeom &g = eng.geom[i];
    if (g.kind == RR_STAGE_HALFBAND) { *new_done = std::max(done[i], half_ready(g, W[i])); *new_produced = *new_done; }
    else if (g.kind == RR_STAGE_DFT) { *new_done = dft_blocks_ready(g, W[i], done[i]); *new_produced = dft_kfirst(g, *new_done); }
    else { *new_done = std::max(done[i], poly_ready(g, W[i])); *new_produced = *new_done; }
  }

  // rate_process (rate_base.h:425-432) for all channels: run every stage over what became available
  int process_stages()
  {
    const int ns = eng.ns;
    for (int i = 0; i < ns; ++i) {
      long long nd, np;
      advance_counts(i, &nd, &np);
      if (nd > done[i]) {
        const long long out_pre = i + 1 == ns ? out_shift : eng.geom[i + 1].preload;
        const long long newW = out_pre + np;
        int rc = ensure_ring(i + 1, newW);
        if (rc) return rc;
        LaneView in = ring_view(i, 0, W[i]);
        LaneView out = ring_view(i + 1, W[i + 1], newW);
        rc = eng.run_stage(i, in, ring_is_frames(i), out, ring_is_frames(i + 1), out_pre, done[i], nd - done[i], nch, s_);
        if (rc) return rc;
        done[i] = nd; produced[i] = np; W[i + 1] = newW;
      }
    }
    return RR_OK;
  }

  // frames [c0, c0 + n) of a frame ring <-> a host buffer: at most two transfers (ring wrap)
  int ring_h2d(int i, long long c0, long long n, const float *src)
  {
    float *r = static_cast<float *>(ring[i]);
    const long long cap = ring_cap[i], first = c0 & (cap - 1), n1 = std::min(n, cap - first);
    int rc = be_h2d(r + first * nch, src, sizeof(float) * static_cast<size_t>(n1) * nch, s_);
    if (!rc && n1 < n) rc = be_h2d(r, src + n1 * nch, sizeof(float) * static_cast<size_t>(n - n1) * nch, s_);
    return rc;
  }
  int ring_d2h(int i, long long c0, long long n, float *dst)
  {
    const float *r = static_cast<const float *>(ring[i]);
    const long long cap = ring_cap[i], first = c0 & (cap - 1), n1 = std::min(n, cap - first);
    int rc = be_d2h(dst, r + first * nch, sizeof(float) * static_cast<size_t>(n1) * nch, s_);
    if (!rc && n1 < n) rc = be_d2h(dst + n1 * nch, r, sizeof(float) * static_cast<size_t>(n - n1) * nch, s_);
    return rc;
  }
  int ring_zero(int i, long long c0, long long n)
  {
    float *r = static_cast<float *>(ring[i]);
    const long long cap = ring_cap[i], first = c0 & (cap - 1), n1 = std::min(n, cap - first);
    int rc = be_memset(r + first * nch, 0, sizeof(float) * static_cast<size_t>(n1) * nch, s_);
    if (!rc && n1 < n) rc = be_memset(r, 0, sizeof(float) * static_cast<size_t>(n - n1) * nch, s_);
    return rc;
  }

  int grow_pinned(float **buf, size_t *cap, size_t elems)
  {
    if (elems <= *cap) return RR_OK;
    be_host_free(*buf); *buf = nullptr; *cap = 0;
    size_t want = std::max<size_t>(elems, static_cast<size_t>(1) << 16);
    void *p = nullptr;
    int rc = be_host_alloc(&p, sizeof(float) * want);
    if (rc) return rc;
    *buf = static_cast<float *>(p); *cap = want;
    return RR_OK;
  }

  int push(const float *x, size_t frames) override   // RR_push_x, rate_base.h:616-636
  {
    if (!x || !frames) return RR_OK;
    RR_DEVICE_SCOPE(eng.device_id);
    if (frames > eng.design.plan.isamp_max) frames = static_cast<size_t>(eng.design.plan.isamp_max);
    count_input(frames);
    int rc;
    const size_t elems = frames * static_cast<size_t>(nch);
    if ((rc = ensure_ring(0, W[0] + static_cast<long long>(frames)))) return rc;
    // deinterleave + convert (rate_base.h:565-569) happen in the first stage's loads: ring 0 keeps frames
    if (kEmulated || be_host_is_pinned(x)) {
      const double t0 = trace_ ? tick() : 0;
      if ((rc = ring_h2d(0, W[0], static_cast<long long>(frames), x))) return rc;
      // the caller may refill its buffer as soon as we return: wait for the transfer (not for the kernels)
      if ((rc = be_event_record(user_in_done_, s_))) return rc;
      W[0] += static_cast<long long>(frames);
      const double t1 = trace_ ? tick() : 0;
      if ((rc = process_stages())) return rc;
      const double t2 = trace_ ? tick() : 0;
      rc = be_event_sync(user_in_done_);
      if (trace_) { tr_[0] += t1 - t0; tr_[1] += t2 - t1; tr_[2] += tick() - t2; ++tr_calls_[0]; }
      return rc;
    }
    const int slot = static_cast<int>(push_seq_++ & 1);
    if (pin_in_busy_[slot]) { if ((rc = be_event_sync(pin_in_free_[slot]))) return rc; pin_in_busy_[slot] = false; }
    if ((rc = grow_pinned(&pin_in_[slot], &pin_in_cap_[slot], elems))) return rc;
    memcpy(pin_in_[slot], x, sizeof(float) * elems);
    if ((rc = ring_h2d(0, W[0], static_cast<long long>(frames), pin_in_[slot]))) return rc;
    if ((rc = be_event_record(pin_in_free_[slot], s_))) return rc;
    pin_in_busy_[slot] = true;
    W[0] += static_cast<long long>(frames);
    return process_stages();
  }

  // Takes n frames off the output FIFO and issues their transfer to `dst` (page-locked). Does not wait.
  int pop_frames(float *dst, size_t n)
  {
    int rc = ring_d2h(eng.ns, popped, static_cast<long long>(n), dst);
    popped += static_cast<long long>(n);
    return rc;
  }
  size_t take(size_t max_frames)               // rate_output, rate_base.h:445-450
  {
    const long long avail = W[eng.ns] - popped;
    const size_t n = static_cast<size_t>(std::min<long long>(avail, static_cast<long long>(max_frames)));
    samples_out += n;
    return n;
  }

  int pull(float *y, void *native, size_t max_frames, size_t *got) override   // RR_pull_x, rate_base.h:638-660
  {
    RR_DEVICE_SCOPE(eng.device_id);
    const int last = eng.ns;
    const size_t n = take(max_frames);
    if (got) *got = n;
    if (!n) return RR_OK;
    int rc;
    const size_t elems = n * static_cast<size_t>(nch);
    if (ring_is_frames(last)) {
      const bool direct = y && !native && (kEmulated || be_host_is_pinned(y));
      float *dst = y;
      if (!direct) {
        if ((rc = grow_pinned(&pin_out_, &pin_out_cap_, elems))) return rc;
        dst = pin_out_;
      }
      const double t0 = trace_ ? tick() : 0;
      if ((rc = pop_frames(dst, n))) return rc;
      const double t1 = trace_ ? tick() : 0;
      if ((rc = be_sync(s_))) return rc;
      if (trace_) { tr_[3] += t1 - t0; tr_[4] += tick() - t1; ++tr_calls_[1]; }
      if (!retired_.empty()) release_retired();
      if (y && !direct) memcpy(y, pin_out_, sizeof(float) * elems);
      if (native) {
        // planar view of the same values: the fp32 engine's own type is float, and without stages (in_rate ==
        // out_rate) the fp64 engine's FIFO holds nothing but converted input frames
        if (!std::is_same<T, float>::value && last != 0) {
          set_last_error("RRX_enable_native_tap must precede the first push");
          return RR_INVPARAM;
        }
        T *o = static_cast<T *>(native);
        for (int c = 0; c < nch; ++c)
          for (size_t j = 0; j < n; ++j) o[static_cast<size_t>(c) * max_frames + j] = static_cast<T>(pin_out_[j * nch + c]);
      }
      return RR_OK;
    }
    // native tap: the last FIFO holds planar engine-type lanes
    LaneView src = ring_view(last, popped, popped + static_cast<long long>(n));
    if (elems > stage_native_cap) {
      be_sync(s_);
      be_free(stage_native); stage_native = nullptr; stage_native_cap = 0;
      void *p = nullptr;
      if ((rc = be_malloc(&p, (sizeof(T) + sizeof(float)) * elems))) return rc;
      stage_native = static_cast<T *>(p); stage_native_cap = elems;
    }
    if (y) {
      float *stage_out = reinterpret_cast<float *>(stage_native + elems);
      LaneView dst{};
      dst.base = stage_out; dst.origin = popped; dst.mask = ~0ull; dst.lo = popped; dst.hi = popped + static_cast<long long>(n);
      dst.stream_stride = 0; dst.ch_stride = 1; dst.elem_stride = nch; dst.nch = nch;
      if ((rc = eng.copy(src, false, dst, true, popped, static_cast<long long>(n), 0, nch, s_))) return rc;
      if ((rc = be_d2h(y, stage_out, sizeof(float) * elems, s_))) return rc;
    }
    if (native) {                                // planar, engine type: out[ch * max_frames + i]
      LaneView dst{};
      dst.base = stage_native; dst.origin = popped; dst.mask = ~0ull; dst.lo = popped; dst.hi = popped + static_cast<long long>(n);
      dst.stream_stride = 0; dst.ch_stride = static_cast<int>(n); dst.elem_stride = 1; dst.nch = nch;
      if ((rc = eng.copy(src, false, dst, false, popped, static_cast<long long>(n), 0, nch, s_))) return rc;
      for (int c = 0; c < nch; ++c)
        if ((rc = be_d2h(static_cast<T *>(native) + static_cast<size_t>(c) * max_frames, stage_native + static_cast<size_t>(c) * n,
                         sizeof(T) * n, s_)))
          return rc;
    }
    popped += static_cast<long long>(n);
    return be_sync(s_);
  }

  // RR_flow_x (rate_base.h:571-614): pull what is ready, push, pull what the push produced into the remaining
  // space. Both output transfers and the input transfer are queued behind each other on the handle's stream and
  // waited for once, so the copy-in of this call overlaps the copy-out of the frames that were already there.
  int flow(const float *x, size_t isamp, float *y, size_t osamp, size_t *iused, size_t *ogen) override
  {
    RR_DEVICE_SCOPE(eng.device_id);
    if (!x) isamp = 0;
    if (isamp > eng.design.plan.isamp_max) isamp = static_cast<size_t>(eng.design.plan.isamp_max);
    if (iused) *iused = isamp;
    if (!ring_is_frames(eng.ns)) {               // native tap: the simple sequence
      size_t g1 = 0, g2 = 0;
      int rc = RR_OK;
      if (y && osamp) rc = pull(y, nullptr, osamp, &g1);
      if (!rc && isamp) rc = push(x, isamp);
      if (!rc && y && g1 < osamp) rc = pull(y + g1 * nch, nullptr, osamp - g1, &g2);
      if (ogen) *ogen = g1 + g2;
      return rc;
    }
    int rc;
    const bool want = y && osamp;
    const bool direct = want && (kEmulated || be_host_is_pinned(y));
    float *dst = y;
    if (want && !direct) {
      if ((rc = grow_pinned(&pin_out_, &pin_out_cap_, osamp * static_cast<size_t>(nch)))) return rc;
      dst = pin_out_;
    }
    size_t g1 = 0, g2 = 0;
    if (want) { g1 = take(osamp); if (g1 && (rc = pop_frames(dst, g1))) return rc; }
    if (isamp && (rc = push(x, isamp))) return rc;
    if (want && g1 < osamp) { g2 = take(osamp - g1); if (g2 && (rc = pop_frames(dst + g1 * nch, g2))) return rc; }
    if (ogen) *ogen = g1 + g2;
    if (g1 + g2) {
      if ((rc = be_sync(s_))) return rc;
      if (!direct) memcpy(y, pin_out_, sizeof(float) * (g1 + g2) * nch);
    }
    return RR_OK;
  }

  int drain() override                           // rate_flush, rate_base.h:454-468
  {
    RR_DEVICE_SCOPE(eng.device_id);
    const int last = eng.ns;
    const uint64_t target = static_cast<uint64_t>(static_cast<double>(samples_in) / eng.design.plan.factor + .5);
    if (target <= samples_out) return RR_OK;
    const long long remaining = static_cast<long long>(target - samples_out);
    // Feed 1024-sample blocks of zeros until enough output exists. Only the counters run in the loop;
    // the zeros are written and the stages launched once for the whole extension (same results: every
    // stage is a pure function of absolute positions).
    const std::vector<long long> W_save = W, done_save = done, prod_save = produced;
    long long fed = 0;
    while (W[last] - popped < remaining) {
      fed += 1024; W[0] += 1024;
      count_input(1024);
      for (int i = 0; i < eng.ns; ++i) {
        long long nd, np;
        advance_counts(i, &nd, &np);
        done[i] = nd; produced[i] = np;
        W[i + 1] = (i + 1 == eng.ns ? out_shift : eng.geom[i + 1].preload) + np;
      }
    }
    const long long W0_new = W[0];
    W = W_save; done = done_save; produced = prod_save;
    int rc;
    if (fed > 0) {
      if ((rc = ensure_ring(0, W0_new))) return rc;
      if ((rc = ring_zero(0, W[0], fed))) return rc;
      W[0] = W0_new;
      if ((rc = process_stages())) return rc;
    }
    // fifo_trim_to(remaining): later output continues right after the trimmed end
    W[last] = popped + remaining;
    if (last > 0) out_shift = W[last] - produced[last - 1];
    samples_in = samples_out = 0;
    return be_sync(s_);
  }

  int enable_native_tap() override
  {
    RR_DEVICE_SCOPE(eng.device_id);
    const int last = eng.ns;
    if (native_tap_ || std::is_same<T, float>::value || last == 0) return RR_OK;   // float frames already hold the fp32 engine's values
    if (W[last] != 0 || popped != 0) { set_last_error("RRX_enable_native_tap must precede the first push"); return RR_INVPARAM; }
    int rc = be_sync(s_);
    if (rc) return rc;
    be_free(ring[last]); ring[last] = nullptr; ring_cap[last] = 0;
    native_tap_ = true;
    return ensure_ring(last, 1 << 14);
  }

  int dft_spectrum(int instance, void *out, int max_n) const override
  {
    DeviceScope scope(eng.device_id);
    return eng.dft_spectrum_host(instance, out, max_n);
  }
};

// ===================================================================================================
// factories
// ===================================================================================================
IBatch *create_batch(const RR_config &cfg, int sample_bytes, int nchannels, int nstreams, size_t frames_in_max,
                     int device, int *err)
{
  int rc = RR_INVPARAM;
  IBatch *res = nullptr;
  if (nchannels < 1 || nstreams < 1 || frames_in_max < 1) { set_last_error("bad batch shape"); }
  else if (sample_bytes == 4) {
    auto *b = new Batch<float>();
    rc = b->init(cfg, nchannels, nstreams, frames_in_max, device);
    if (rc) delete b; else res = b;
  } else if (sample_bytes == 8) {
    auto *b = new Batch<double>();
    rc = b->init(cfg, nchannels, nstreams, frames_in_max, device);
    if (rc) delete b; else res = b;
  }
  if (err) *err = rc;
  return res;
}

IStream *create_stream(const RR_config &cfg, int sample_bytes, int nchannels, int device, int *err)
{
  int rc = RR_INVPARAM;
  IStream *res = nullptr;
  if (nchannels < 1) { set_last_error("bad channel count"); }
  else if (sample_bytes == 4) {
    auto *s = new Stream<float>();
    s->in_rate_ = cfg.in_rate; s->out_rate_ = cfg.out_rate;
    rc = s->init(cfg, nchannels, device);
    if (rc) delete s; else res = s;
  } else if (sample_bytes == 8) {
    auto *s = new Stream<double>();
    s->in_rate_ = cfg.in_rate; s->out_rate_ = cfg.out_rate;
    rc = s->init(cfg, nchannels, device);
    if (rc) delete s; else res = s;
  }
  if (err) *err = rc;
  return res;
}

}  // namespace b200rate
