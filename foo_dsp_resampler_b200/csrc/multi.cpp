// multi.cpp -- the multi-GPU layer of the C ABI (RRX_multi_*, include/b200_ratelib.h).
//
// the path shards without a data-path collective --
//   * batched workloads by independent stream (streams share nothing but read-only coefficient banks,
//     rate/rate_base.h:89-92): device k of n converts a contiguous slice of the streams;
//   * one long stream by time chunk: every stage is FIR and every block / phase position is a closed form of the
//     absolute sample index, so the OUTPUT timeline is cut at multiples of the last stage's block and device k
//     converts range k from its own halo'd input window (RRX_batch_input_window / RRX_batch_process_range).
// One host thread per device drives that device's batch object, its CUDA streams and its page-locked staging, so a
// plain C caller (examples/rate_harness.c --gpus N, or foo_dsp_rate.cpp through chain.h:22-43) uses every GPU of
// the box without Python. NCCL is used for exactly one thing, as north_star says: gathering device-resident results
// on one device (RRX_multi_gather); it is loaded with dlopen on first use so that the library has no link-time
// dependency on it and every other entry point works without it.
//
// Built with -DB200RATE_EMU (tests/emu) the "devices" are host memory and the gather is a memcpy: the sharding and
// threading logic is then checked on a machine without GPUs. Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "b200_ratelib.h"
#include "engine.hpp"

#ifndef B200RATE_EMU
#include <cuda_runtime.h>
#include <dlfcn.h>
#endif

using namespace b200rate;

namespace {

#ifdef B200RATE_EMU
int dev_set(int) { return RR_OK; }
int dev_malloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? RR_OK : RR_ENOMEM; }
void dev_free(void *p) { free(p); }
int dev_copy(void *dst, const void *src, size_t n) { memcpy(dst, src, n); return RR_OK; }
int dev_sync() { return RR_OK; }
#else
int cuda_rc(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return RR_OK;
  set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? RR_ENOMEM : RR_INTERNAL;
}
int dev_set(int d) { return cuda_rc(cudaSetDevice(d), "cudaSetDevice"); }
int dev_malloc(void **p, size_t n) { return cuda_rc(cudaMalloc(p, n ? n : 1), "cudaMalloc"); }
void dev_free(void *p) { if (p) cudaFree(p); }
int dev_copy(void *dst, const void *src, size_t n) { return cuda_rc(cudaMemcpy(dst, src, n, cudaMemcpyDefault), "cudaMemcpy"); }
int dev_sync() { return cuda_rc(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); }

// ---- NCCL through dlopen: only the six entry points the gather needs ----
struct Nccl {
  void *lib = nullptr;
  typedef struct ncclComm *comm_t;
  int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load()
  {
    if (lib) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { set_last_error("NCCL is not available (dlopen libnccl.so.2 failed)"); return false; }
#define RR_SYM(N) N = reinterpret_cast<decltype(N)>(dlsym(lib, "nccl" #N))
    RR_SYM(CommInitAll); RR_SYM(CommDestroy); RR_SYM(GroupStart); RR_SYM(GroupEnd); RR_SYM(Send); RR_SYM(Recv); RR_SYM(GetErrorString);
#undef RR_SYM
    if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Send || !Recv) { set_last_error("libnccl lacks an expected symbol"); return false; }
    return true;
  }
};
Nccl g_nccl;
constexpr int kNcclFloat = 7;   // ncclFloat32 (nccl.h: ncclDataType_t)
#endif

struct Shard {
  int device = 0;
  size_t first = 0, count = 0;          // streams [first, first + count) of the batch (stream sharding)
  IBatch *batch = nullptr;
  float *d_out = nullptr;               // device-resident result of the last RRX_multi_process call
  size_t d_out_elems = 0;
};

}  // namespace

extern "C" {

struct RRX_multi_tag {
  RR_config cfg;
  int sample_bytes, nchannels;
  size_t nstreams, frames_in_max;
  std::vector<Shard> shards;
  size_t last_frames_out = 0;
#ifndef B200RATE_EMU
  std::vector<Nccl::comm_t> comms;
  std::vector<cudaStream_t> gather_streams;
#endif
};

}  // extern "C"

namespace {

// Runs f(k) for every shard on its own host thread (the calling thread takes shard 0) and returns the first error.
template <class F> int for_each_shard(RRX_multi *m, F f)
{
  const size_t n = m->shards.size();
  std::vector<int> rc(n, RR_OK);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  auto run = [&](size_t k) {
    try {
      rc[k] = dev_set(m->shards[k].device);
      if (rc[k] == RR_OK) rc[k] = f(k);
    } catch (const std::bad_alloc &) { rc[k] = RR_ENOMEM; }
    catch (...) { rc[k] = RR_INTERNAL; }
    if (rc[k] != RR_OK) msg[k] = last_error();           // the error text is thread-local: carry it to the caller
  };
  for (size_t k = 1; k < n; ++k) th.emplace_back(run, k);
  run(0);
  for (auto &t : th) t.join();
  for (size_t k = 0; k < n; ++k)
    if (rc[k] != RR_OK) { set_last_error("device " + std::to_string(m->shards[k].device) + ": " + msg[k]); return rc[k]; }
  return RR_OK;
}

}  // namespace

extern "C" {

int RRX_multi_open(const RR_config *config, int sample_bytes, int nchannels, size_t nstreams, size_t frames_in_max,
                   const int *devices, int ndevices, RRX_multi **out)
{
  if (!out) return RR_INVPARAM;
  *out = nullptr;
  if (!config || !devices || ndevices < 1 || nchannels < 1 || nstreams < 1 || frames_in_max < 1 || (sample_bytes != 4 && sample_bytes != 8))
    return RR_INVPARAM;
  std::unique_ptr<RRX_multi> m(new RRX_multi());
  m->cfg = *config; m->sample_bytes = sample_bytes; m->nchannels = nchannels; m->nstreams = nstreams; m->frames_in_max = frames_in_max;
  // contiguous slices, the remainder to the low devices; a device without streams still takes part in time-chunked calls
  const size_t base = nstreams / static_cast<size_t>(ndevices), extra = nstreams % static_cast<size_t>(ndevices);
  size_t first = 0;
  for (int k = 0; k < ndevices; ++k) {
    Shard s;
    s.device = devices[k]; s.first = first; s.count = base + (static_cast<size_t>(k) < extra ? 1 : 0);
    first += s.count;
    m->shards.push_back(s);
  }
  RRX_multi *raw = m.get();
  const int rc = for_each_shard(raw, [&](size_t k) {
    Shard &s = raw->shards[k];
    int err = RR_OK;
    // sub-batches of at most 64 streams: the host-buffer pipeline overlaps their transfers with the kernels
    const int sub = static_cast<int>(std::max<size_t>(1, std::min<size_t>(s.count ? s.count : 1, 64)));
    s.batch = create_batch(raw->cfg, raw->sample_bytes, raw->nchannels, sub, raw->frames_in_max, s.device, &err);
    return s.batch ? RR_OK : (err ? err : RR_INTERNAL);
  });
  if (rc != RR_OK) { RRX_multi_close(&raw); m.release(); return rc; }
  *out = m.release();
  return RR_OK;
}

int RRX_multi_devices(const RRX_multi *m) { return m ? static_cast<int>(m->shards.size()) : 0; }

int RRX_multi_shard(const RRX_multi *m, int k, int *device, size_t *first_stream, size_t *stream_count)
{
  if (!m) return RR_NULLHANDLE;
  if (k < 0 || k >= static_cast<int>(m->shards.size())) return RR_INVPARAM;
  if (device) *device = m->shards[static_cast<size_t>(k)].device;
  if (first_stream) *first_stream = m->shards[static_cast<size_t>(k)].first;
  if (stream_count) *stream_count = m->shards[static_cast<size_t>(k)].count;
  return RR_OK;
}

size_t RRX_multi_frames_out(const RRX_multi *m, size_t frames_in)
{
  return m && !m->shards.empty() && m->shards[0].batch ? m->shards[0].batch->frames_out(frames_in) : 0;
}

int RRX_multi_process_host(RRX_multi *m, const float *h_in, size_t frames_in, float *h_out)
{
  if (!m) return RR_NULLHANDLE;
  if (!h_in || !h_out || frames_in > m->frames_in_max) return RR_INVPARAM;
  const size_t nout = RRX_multi_frames_out(m, frames_in);
  const size_t in_elems = frames_in * static_cast<size_t>(m->nchannels), out_elems = nout * static_cast<size_t>(m->nchannels);
  return for_each_shard(m, [&](size_t k) {
    const Shard &s = m->shards[k];
    if (!s.count) return static_cast<int>(RR_OK);
    return s.batch->process_host(h_in + s.first * in_elems, frames_in, h_out + s.first * out_elems, s.count);
  });
}

// One long stream (nstreams == 1): range k of the output timeline on device k, cut at multiples of the last stage's
// block so that no block is computed twice; every device reads only its own halo'd window of the host buffer.
int RRX_multi_process_stream_host(RRX_multi *m, const float *h_in, size_t frames_in_total, float *h_out, size_t *frames_out)
{
  if (!m) return RR_NULLHANDLE;
  if (!h_in || !h_out || m->nstreams != 1) return RR_INVPARAM;
  IBatch *b0 = m->shards[0].batch;
  const size_t nout = b0->frames_out(frames_in_total);
  if (frames_out) *frames_out = nout;
  const rr_plan &plan = b0->design().plan;
  size_t align = 1;
  if (plan.num_stages > 0) {
    const rr_stage_plan &st = plan.st[plan.num_stages - 1];
    if (st.kind == RR_STAGE_DFT) {
      const size_t valid = static_cast<size_t>(st.dft_length - (st.num_taps - 1));
      align = st.step_int == 1 ? valid : st.step_int < 0 ? valid >> (-st.step_int) : 1;
    }
  }
  const size_t n = m->shards.size(), units = (nout + align - 1) / align;
  // pieces of at most frames_in_max input frames per call, consecutive on the device
  return for_each_shard(m, [&](size_t k) {
    const size_t lo = std::min(nout, units * k / n * align), hi = k + 1 == n ? nout : std::min(nout, units * (k + 1) / n * align);
    IBatch *b = m->shards[k].batch;
    const size_t nch = static_cast<size_t>(m->nchannels);
    // output frames per call: what frames_in_max input frames yield, less the halo
    size_t step = b->frames_out(m->frames_in_max);
    step = step > 2 * align + 4096 ? step - align - 4096 : step;
    step = std::max<size_t>(align, step / align * align);
    void *d_in = nullptr, *d_out = nullptr;
    int rc = dev_malloc(&d_in, sizeof(float) * m->frames_in_max * nch);
    if (!rc) rc = dev_malloc(&d_out, sizeof(float) * (step + align) * nch);
    for (size_t ob = lo; !rc && ob < hi; ob += step) {
      const size_t oc = std::min(step, hi - ob);
      uint64_t f = 0, c = 0;
      b->input_window(frames_in_total, ob, oc, &f, &c);
      if (c > m->frames_in_max) { set_last_error("time chunk needs more input frames than frames_in_max"); rc = RR_INVPARAM; break; }
      if ((rc = dev_copy(d_in, h_in + f * nch, sizeof(float) * c * nch))) break;
      if ((rc = b->process(static_cast<const float *>(d_in), f, c, frames_in_total, ob, oc, d_out, false, nullptr))) break;
      if ((rc = dev_sync())) break;
      rc = dev_copy(h_out + ob * nch, d_out, sizeof(float) * oc * nch);
    }
    dev_free(d_in); dev_free(d_out);
    return rc;
  });
}

// Device-resident variant: d_in[k] is device k's slice (float32 [stream_count_k][frames_in][nchannels] on that device).
// The results stay on their devices (RRX_multi_result) until they are gathered.
int RRX_multi_process(RRX_multi *m, const float *const *d_in, size_t frames_in)
{
  if (!m) return RR_NULLHANDLE;
  if (!d_in || frames_in > m->frames_in_max) return RR_INVPARAM;
  const size_t nout = RRX_multi_frames_out(m, frames_in);
  m->last_frames_out = nout;
  const size_t in_elems = frames_in * static_cast<size_t>(m->nchannels), out_elems = nout * static_cast<size_t>(m->nchannels);
  return for_each_shard(m, [&](size_t k) {
    Shard &s = m->shards[k];
    if (!s.count) return static_cast<int>(RR_OK);
    if (!d_in[k]) return static_cast<int>(RR_INVPARAM);
    int rc = RR_OK;
    if (s.d_out_elems < out_elems * s.count) {
      dev_free(s.d_out); s.d_out = nullptr; s.d_out_elems = 0;
      void *p = nullptr;
      if ((rc = dev_malloc(&p, sizeof(float) * out_elems * s.count))) return rc;
      s.d_out = static_cast<float *>(p); s.d_out_elems = out_elems * s.count;
    }
    const size_t sub = 64;                                   // the shard's batch object holds at most 64 streams
    for (size_t s0 = 0; !rc && s0 < s.count; s0 += sub) {
      const int now = static_cast<int>(std::min(sub, s.count - s0));
      rc = s.batch->process_streams(d_in[k] + s0 * in_elems, frames_in, s.d_out + s0 * out_elems, now, nullptr);
    }
    return rc ? rc : dev_sync();
  });
}

int RRX_multi_result(const RRX_multi *m, int k, const float **d_out, size_t *frames_out)
{
  if (!m) return RR_NULLHANDLE;
  if (k < 0 || k >= static_cast<int>(m->shards.size())) return RR_INVPARAM;
  if (d_out) *d_out = m->shards[static_cast<size_t>(k)].d_out;
  if (frames_out) *frames_out = m->last_frames_out;
  return RR_OK;
}

// Gathers the device-resident results of the last RRX_multi_process call on device `root` (an index into the device
// list): d_out_root is float32 [nstreams][frames_out][nchannels] on that device. ncclSend / ncclRecv over NVLink; the
// root's own slice is a device-to-device copy.
int RRX_multi_gather(RRX_multi *m, int root, float *d_out_root)
{
  if (!m) return RR_NULLHANDLE;
  const int n = static_cast<int>(m->shards.size());
  if (root < 0 || root >= n || !d_out_root) return RR_INVPARAM;
  const size_t out_elems = m->last_frames_out * static_cast<size_t>(m->nchannels);
#ifdef B200RATE_EMU
  for (int k = 0; k < n; ++k) {
    const Shard &s = m->shards[static_cast<size_t>(k)];
    if (s.count) memcpy(d_out_root + s.first * out_elems, s.d_out, sizeof(float) * out_elems * s.count);
  }
  return RR_OK;
#else
  bool distinct = true;
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b)
      if (m->shards[static_cast<size_t>(a)].device == m->shards[static_cast<size_t>(b)].device) distinct = false;
  if (n == 1 || !distinct) {                                 // nothing to cross (or one device listed twice): plain copies
    for (int k = 0; k < n; ++k) {
      const Shard &s = m->shards[static_cast<size_t>(k)];
      if (!s.count) continue;
      int rc = dev_set(m->shards[static_cast<size_t>(root)].device);
      if (!rc) rc = dev_copy(d_out_root + s.first * out_elems, s.d_out, sizeof(float) * out_elems * s.count);
      if (rc) return rc;
    }
    return RR_OK;
  }
  if (!g_nccl.load()) return RR_INTERNAL;
  if (m->comms.empty()) {
    std::vector<int> devs;
    for (const Shard &s : m->shards) devs.push_back(s.device);
    m->comms.resize(static_cast<size_t>(n));
    const int e = g_nccl.CommInitAll(m->comms.data(), n, devs.data());
    if (e) { m->comms.clear(); set_last_error(std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?")); return RR_INTERNAL; }
    m->gather_streams.resize(static_cast<size_t>(n));
    for (int k = 0; k < n; ++k) {
      int rc = dev_set(devs[static_cast<size_t>(k)]);
      if (!rc) rc = cuda_rc(cudaStreamCreateWithFlags(&m->gather_streams[static_cast<size_t>(k)], cudaStreamNonBlocking), "cudaStreamCreate");
      if (rc) return rc;
    }
  }
  int e = g_nccl.GroupStart();
  for (int k = 0; k < n && !e; ++k) {
    const Shard &s = m->shards[static_cast<size_t>(k)];
    if (!s.count) continue;
    const size_t cnt = out_elems * s.count;
    if (k == root) {
      dev_set(s.device);
      cudaMemcpyAsync(d_out_root + s.first * out_elems, s.d_out, sizeof(float) * cnt, cudaMemcpyDeviceToDevice, m->gather_streams[static_cast<size_t>(k)]);
      continue;
    }
    e = g_nccl.Send(s.d_out, cnt, kNcclFloat, root, m->comms[static_cast<size_t>(k)], m->gather_streams[static_cast<size_t>(k)]);
    if (!e) e = g_nccl.Recv(d_out_root + s.first * out_elems, cnt, kNcclFloat, k, m->comms[static_cast<size_t>(root)], m->gather_streams[static_cast<size_t>(root)]);
  }
  const int e2 = g_nccl.GroupEnd();
  if (e || e2) { set_last_error(std::string("nccl gather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e ? e : e2) : "?")); return RR_INTERNAL; }
  for (int k = 0; k < n; ++k) {
    int rc = dev_set(m->shards[static_cast<size_t>(k)].device);
    if (!rc) rc = cuda_rc(cudaStreamSynchronize(m->gather_streams[static_cast<size_t>(k)]), "cudaStreamSynchronize");
    if (rc) return rc;
  }
  return RR_OK;
#endif
}

// Page-locked host memory for callers that do not link CUDA themselves (a C host program): buffers from here make
// the host-buffer entry points run at the speed of the host link instead of that of a pageable staging copy.
void *RRX_host_alloc(size_t bytes)
{
#ifdef B200RATE_EMU
  return malloc(bytes ? bytes : 1);
#else
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
#endif
}
void RRX_host_free(void *p)
{
#ifdef B200RATE_EMU
  free(p);
#else
  if (p) cudaFreeHost(p);
#endif
}

void RRX_multi_close(RRX_multi **pm)
{
  if (!pm || !*pm) return;
  RRX_multi *m = *pm;
#ifndef B200RATE_EMU
  for (size_t k = 0; k < m->gather_streams.size(); ++k) {
    if (dev_set(m->shards[k].device) == RR_OK && m->gather_streams[k]) cudaStreamDestroy(m->gather_streams[k]);
  }
  for (auto c : m->comms) if (c) g_nccl.CommDestroy(c);
#endif
  for (Shard &s : m->shards) {
    dev_set(s.device);
    delete s.batch;
    dev_free(s.d_out);
  }
  delete m;
  *pm = nullptr;
}

}  // extern "C"
