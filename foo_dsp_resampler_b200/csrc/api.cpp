// api.cpp -- the C ABI of libb200rate.so (include/b200_ratelib.h).
//
// Part 1 mirrors the reference's facade /root/reference/rate/rate_uni.c:27-231 and the per-engine
// wrappers /root/reference/rate/rate_base.h:517-741: same entry points, same return codes, handle whose
// first member is the vtable. Part 2 are the RRX_* extensions.
#include <cstdlib>
#include <cstring>
#include <new>

#include "b200_ratelib.h"
#include "engine.hpp"
#include "host_design.hpp"

using namespace b200rate;

extern "C" {

// The reference's vtable layout (rate/rate_i.h:25-33); RR_* dispatch through it like rate_uni.c:59-90.
typedef struct RR_vtable_tag {
  int (*init)(RR_handle *h, const RR_config *config, int nchannels);
  int (*flow)(RR_handle *h, const fb_sample_t *ibuf, fb_sample_t *obuf, size_t isamp, size_t osamp, size_t *iused,
              size_t *ogen);
  int (*push)(RR_handle *h, const fb_sample_t *ibuf, size_t isamp);
  int (*pull)(RR_handle *h, fb_sample_t *obuf, size_t osamp, size_t *ogen);
  int (*drain)(RR_handle *h);
  void (*close)(RR_handle *h);
} RR_vtable;

struct RR_handle_tag {
  RR_vtable x;          // must stay first
  IStream *stream;
  int sample_bytes;
  int nchannels;
};

struct RRX_batch_tag {
  IBatch *batch;
  int nchannels, nstreams;
};

}  // extern "C"

namespace {

bool g_initialized = false;
void (*g_oom)(void) = nullptr;

int guarded(int (*body)(void *), void *arg)
{
  try {
    return body(arg);
  } catch (const std::bad_alloc &) {
    if (g_oom) g_oom();                      // the reference's host-OOM contract (rate/xmalloc.c:38-43)
    return RR_ENOMEM;
  } catch (...) {
    set_last_error("unexpected C++ exception");
    return RR_INTERNAL;
  }
}

int h_init(RR_handle *, const RR_config *, int) { return RR_OK; }   // construction happens in the ctor

int h_push(RR_handle *h, const fb_sample_t *ibuf, size_t isamp)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  if (!ibuf || !isamp) return RR_OK;
  struct A { RR_handle *h; const float *x; size_t n; } a{h, ibuf, isamp};
  return guarded([](void *p) { A *a = static_cast<A *>(p); return a->h->stream->push(a->x, a->n); }, &a);
}

int h_pull(RR_handle *h, fb_sample_t *obuf, size_t osamp, size_t *ogen)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  if (!obuf || !osamp) { if (ogen) *ogen = 0; return RR_OK; }
  struct A { RR_handle *h; float *y; size_t n; size_t *g; } a{h, obuf, osamp, ogen};
  return guarded([](void *p) { A *a = static_cast<A *>(p); return a->h->stream->pull(a->y, nullptr, a->n, a->g); }, &a);
}

int h_drain(RR_handle *h)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  return guarded([](void *p) { return static_cast<RR_handle *>(p)->stream->drain(); }, h);
}

// RR_flow_x, rate/rate_base.h:571-614: drain what is ready, feed, then top the output up.
int h_flow(RR_handle *h, const fb_sample_t *ibuf, fb_sample_t *obuf, size_t isamp, size_t osamp, size_t *iused,
           size_t *ogen)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  struct A { RR_handle *h; const float *x; float *y; size_t ni, no; size_t *iu, *og; } a{h, ibuf, obuf, isamp, osamp, iused, ogen};
  return guarded([](void *p) {
    A *a = static_cast<A *>(p);
    return a->h->stream->flow(a->x, a->ni, a->y, a->no, a->iu, a->og);
  }, &a);
}

void h_close(RR_handle *h)
{
  if (!h) return;
  delete h->stream;
  std::free(h);
}

RR_handle *make_handle(const RR_config *config, int nchannels, int sample_bytes, int *err)
{
  int rc = RR_INVPARAM;
  RR_handle *h = nullptr;
  if (config && nchannels > 0) {
    h = static_cast<RR_handle *>(std::calloc(1, sizeof(RR_handle)));
    if (!h) { if (g_oom) g_oom(); rc = RR_ENOMEM; }
    else {
      h->x.init = h_init; h->x.flow = h_flow; h->x.push = h_push; h->x.pull = h_pull; h->x.drain = h_drain; h->x.close = h_close;
      h->sample_bytes = sample_bytes; h->nchannels = nchannels;
      struct A { const RR_config *c; int bytes, nch; IStream *out; int rc; } a{config, sample_bytes, nchannels, nullptr, RR_OK};
      rc = guarded([](void *p) {
        A *a = static_cast<A *>(p);
        a->out = create_stream(*a->c, a->bytes, a->nch, -1, &a->rc);
        return a->rc;
      }, &a);
      h->stream = a.out;
      if (rc != RR_OK) { std::free(h); h = nullptr; }
    }
  }
  if (err) *err = rc;
  return h;
}

}  // namespace

extern "C" {

int init_ratelib(void (*oom)(void))               // rate/rate_uni.c:210-223
{
  g_initialized = false;
  if (!oom) return -1;
  g_oom = oom;
  g_initialized = true;
  return 0;
}

void close_ratelib(void) { g_initialized = false; }   // rate/rate_uni.c:225-231

int RR_open(const RR_config *config, int nchannels, RR_handle **const handle)   // rate/rate_uni.c:27-57
{
  if (!handle) return RR_INVPARAM;
  *handle = nullptr;
  if (!g_initialized) return RR_EXTUNINIT;
  if (!config) return RR_INVPARAM;
  int rc = RR_OK;
  // Best -> the fp64 engine, Normal -> the fp32 engine (the reference picks SSE3/double vs SSE/float)
  RR_handle *h = make_handle(config, nchannels, config->quality == RR_best ? 8 : 4, &rc);
  if (!h) return rc == RR_OK ? RR_ENOMEM : rc;
  *handle = h;
  return RR_OK;
}

int RR_flow(RR_handle *h, const fb_sample_t *ibuf, fb_sample_t *obuf, size_t isamp, size_t osamp, size_t *iused, size_t *ogen)
{
  if (!h) return RR_NULLHANDLE;
  return reinterpret_cast<RR_vtable *>(h)->flow(h, ibuf, obuf, isamp, osamp, iused, ogen);
}
int RR_push(RR_handle *h, const fb_sample_t *ibuf, size_t isamp)
{
  if (!h) return RR_NULLHANDLE;
  return reinterpret_cast<RR_vtable *>(h)->push(h, ibuf, isamp);
}
int RR_pull(RR_handle *h, fb_sample_t *obuf, size_t osamp, size_t *ogen)
{
  if (!h) return RR_NULLHANDLE;
  return reinterpret_cast<RR_vtable *>(h)->pull(h, obuf, osamp, ogen);
}
int RR_drain(RR_handle *h)
{
  if (!h) return RR_NULLHANDLE;
  return reinterpret_cast<RR_vtable *>(h)->drain(h);
}
void RR_close(RR_handle **h)
{
  if (!h || !*h) return;
  reinterpret_cast<RR_vtable *>(*h)->close(*h);
  *h = nullptr;
}

const char *RR_strerror(int error)               // rate/rate_uni.c:92-111
{
  switch (error) {
    case RR_OK: return "OK";
    case RR_ENOMEM: return "Not enough memory";
    case RR_INTERNAL: return "Internal error";
    case RR_NULLHANDLE: return "NULL handle";
    case RR_RATEERROR: return "Error in rate() functions";
    case RR_EXTUNINIT: return "Externals not initialized";
    default: return "Other error";
  }
}

RR_handle *RR_ctor_float(const RR_config *config, int nchannels) { return make_handle(config, nchannels, 4, nullptr); }
RR_handle *RR_ctor_SSE(const RR_config *config, int nchannels) { return make_handle(config, nchannels, 4, nullptr); }
RR_handle *RR_ctor_double(const RR_config *config, int nchannels) { return make_handle(config, nchannels, 8, nullptr); }
RR_handle *RR_ctor_SSE3(const RR_config *config, int nchannels) { return make_handle(config, nchannels, 8, nullptr); }

// ---------------------------------------------------------------------------------------------------
// extensions
// ---------------------------------------------------------------------------------------------------
int RRX_plan(const RR_config *config, int sample_bytes, rr_plan *out)
{
  if (!config || !out || (sample_bytes != 4 && sample_bytes != 8)) return RR_INVPARAM;
  struct A { const RR_config *c; int bytes; rr_plan *out; } a{config, sample_bytes, out};
  return guarded([](void *p) {
    A *a = static_cast<A *>(p);
    Design d;
    const int rc = build_design(*a->c, a->bytes, d);
    if (rc == RR_OK) *a->out = d.plan;
    return rc;
  }, &a);
}

int RRX_design_dump(const RR_config *config, int sample_bytes, int kind, double *out, int max_n)
{
  if (!config || (sample_bytes != 4 && sample_bytes != 8) || kind < 0 || kind > 2) return -1;
  try {
    Design d;
    if (build_design(*config, sample_bytes, d) != RR_OK) return -1;
    const std::vector<double> &v = kind == 2 ? d.poly_bank : d.dft[kind].coefs_time;
    const int n = static_cast<int>(v.size());
    if (out && max_n > 0) std::memcpy(out, v.data(), sizeof(double) * static_cast<size_t>(n < max_n ? n : max_n));
    return n;
  } catch (...) { return -1; }
}

int RRX_plan_dump(const RR_handle *h, rr_plan *out)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream || !out) return RR_INVPARAM;
  *out = h->stream->design().plan;
  return RR_OK;
}

int RRX_pull_native(RR_handle *h, void *out, size_t osamp, size_t *ogen)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  if (!out || !osamp) { if (ogen) *ogen = 0; return RR_OK; }
  struct A { RR_handle *h; void *y; size_t n; size_t *g; } a{h, out, osamp, ogen};
  return guarded([](void *p) { A *a = static_cast<A *>(p); return a->h->stream->pull(nullptr, a->y, a->n, a->g); }, &a);
}

int RRX_enable_native_tap(RR_handle *h)
{
  if (!h) return RR_NULLHANDLE;
  if (!h->stream) return RR_INVPARAM;
  return guarded([](void *p) { return static_cast<RR_handle *>(p)->stream->enable_native_tap(); }, h);
}

int RRX_dft_spectrum(const RR_handle *h, int instance, void *out, int max_n)
{
  if (!h || !h->stream || instance < 0 || instance > 1) return -1;
  return h->stream->dft_spectrum(instance, out, max_n);
}

int RRX_batch_open(const RR_config *config, int sample_bytes, int nchannels, int nstreams, size_t frames_in_max,
                   int device, RRX_batch **out)
{
  if (!out) return RR_INVPARAM;
  *out = nullptr;
  if (!config || (sample_bytes != 4 && sample_bytes != 8)) return RR_INVPARAM;
  struct A { const RR_config *c; int bytes, nch, ns; size_t fmax; int dev; IBatch *res; int rc; }
      a{config, sample_bytes, nchannels, nstreams, frames_in_max, device, nullptr, RR_OK};
  int rc = guarded([](void *p) {
    A *a = static_cast<A *>(p);
    a->res = create_batch(*a->c, a->bytes, a->nch, a->ns, a->fmax, a->dev, &a->rc);
    return a->rc;
  }, &a);
  if (rc != RR_OK) return rc;
  RRX_batch *b = static_cast<RRX_batch *>(std::calloc(1, sizeof(RRX_batch)));
  if (!b) { delete a.res; return RR_ENOMEM; }
  b->batch = a.res; b->nchannels = nchannels; b->nstreams = nstreams;
  *out = b;
  return RR_OK;
}

size_t RRX_batch_frames_out(const RRX_batch *b, size_t frames_in) { return b && b->batch ? b->batch->frames_out(frames_in) : 0; }

int RRX_batch_input_window(const RRX_batch *b, size_t frames_in, uint64_t out_begin, size_t out_count,
                           uint64_t *in_first, uint64_t *in_count)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  b->batch->input_window(frames_in, out_begin, out_count, in_first, in_count);
  return RR_OK;
}

static int batch_run(RRX_batch *b, const float *d_in, uint64_t win_first, size_t win_frames, size_t frames_in,
                     uint64_t out_begin, size_t out_count, void *d_out, bool native, void *stream)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  if (!d_in || !d_out) return RR_INVPARAM;
  struct A { RRX_batch *b; const float *in; uint64_t wf; size_t wn, fin; uint64_t ob; size_t oc; void *out; bool nat; void *s; }
      a{b, d_in, win_first, win_frames, frames_in, out_begin, out_count, d_out, native, stream};
  return guarded([](void *p) {
    A *a = static_cast<A *>(p);
    return a->b->batch->process(a->in, a->wf, a->wn, a->fin, a->ob, a->oc, a->out, a->nat, a->s);
  }, &a);
}

int RRX_batch_process(RRX_batch *b, const float *d_in, size_t frames_in, float *d_out, void *stream)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  return batch_run(b, d_in, 0, frames_in, frames_in, 0, b->batch->frames_out(frames_in), d_out, false, stream);
}

int RRX_batch_process_native(RRX_batch *b, const float *d_in, size_t frames_in, void *d_out_native, void *stream)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  return batch_run(b, d_in, 0, frames_in, frames_in, 0, b->batch->frames_out(frames_in), d_out_native, true, stream);
}

int RRX_batch_process_range(RRX_batch *b, const float *d_in_window, uint64_t window_first, size_t window_frames,
                            size_t frames_in_total, uint64_t out_begin, size_t out_count, float *d_out, void *stream)
{
  return batch_run(b, d_in_window, window_first, window_frames, frames_in_total, out_begin, out_count, d_out, false, stream);
}

int RRX_batch_process_host(RRX_batch *b, const float *h_in, size_t frames_in, float *h_out, size_t total_streams)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  if (!h_in || !h_out || !total_streams) return RR_INVPARAM;
  struct A { RRX_batch *b; const float *in; size_t n; float *out; size_t tot; } a{b, h_in, frames_in, h_out, total_streams};
  return guarded([](void *p) {
    A *a = static_cast<A *>(p);
    return a->b->batch->process_host(a->in, a->n, a->out, a->tot);
  }, &a);
}

int RRX_batch_enable_timing(RRX_batch *b, int on)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  b->batch->enable_timing(on != 0);
  return RR_OK;
}

int RRX_batch_stage_times(RRX_batch *b, float *ms, int max_stages)
{
  if (!b || !b->batch || !ms) return 0;
  return b->batch->stage_times(ms, max_stages);
}

int RRX_batch_stage_work(const RRX_batch *b, size_t frames_in, int stage, double *flops, double *bytes, double *units)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  return b->batch->stage_work(frames_in, stage, flops, bytes, units);
}

const char *RRX_batch_stage_kernel(const RRX_batch *b, int stage)
{
  if (!b || !b->batch) return "";
  return b->batch->stage_kernel(stage);
}

int RRX_batch_plan(const RRX_batch *b, rr_plan *out)
{
  if (!b || !b->batch) return RR_NULLHANDLE;
  if (!out) return RR_INVPARAM;
  *out = b->batch->design().plan;
  return RR_OK;
}

int RRX_batch_last_launches(const RRX_batch *b) { return b && b->batch ? b->batch->last_launches() : 0; }
double RRX_batch_flops(const RRX_batch *b, size_t frames_in) { return b && b->batch ? b->batch->flops(frames_in) : 0.0; }

void RRX_batch_close(RRX_batch **b)
{
  if (!b || !*b) return;
  delete (*b)->batch;
  std::free(*b);
  *b = nullptr;
}

const char *RRX_last_error(void) { return last_error(); }
const char *RRX_version(void) { return "b200rate 0.1 (sm_100a)"; }

}  // extern "C"
