// engine.hpp -- host-side interfaces of the device engine (no CUDA types leak through here).
//
// Two front-ends share one plan/table set:
//   IBatch  -- many equal-length streams processed in one shot, buffers resident in HBM (RRX_batch_*).
//   IStream -- the reference's push/pull/drain protocol on device ring buffers (RR_* entry points),
//              bookkeeping identical to rate/rate_base.h:425-468 but kept in absolute stream coordinates.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "b200_ratelib.h"
#include "host_design.hpp"

namespace b200rate {

// Thread-local description of the last failure (CUDA error string or planner message).
void set_last_error(const std::string &msg);
const char *last_error();

class IBatch {
 public:
  virtual ~IBatch() {}
  virtual const Design &design() const = 0;
  virtual size_t frames_out(size_t frames_in) const = 0;
  // Input frames [first, first+count) that outputs [out_begin, out_begin+out_count) depend on.
  virtual void input_window(size_t frames_in, uint64_t out_begin, size_t out_count, uint64_t *first,
                            uint64_t *count) const = 0;
  // d_in holds input frames [win_first, win_first + win_frames) of each stream (of frames_in in total);
  // produces output frames [out_begin, out_begin + out_count). native_out: planar engine-type output
  // instead of interleaved float.
  virtual int process(const float *d_in, uint64_t win_first, size_t win_frames, size_t frames_in, uint64_t out_begin,
                      size_t out_count, void *d_out, bool native_out, void *stream) = 0;
  // As process() over whole streams, but with only the first `nstreams_now` streams of the batch active.
  virtual int process_streams(const float *d_in, size_t frames_in, void *d_out, int nstreams_now, void *stream) = 0;
  // Host-buffer variant: total_streams streams in (pinned) host memory are moved through the device in
  // sub-batches of the batch's nstreams, H2D / compute / D2H overlapped on three CUDA streams.
  virtual int process_host(const float *h_in, size_t frames_in, float *h_out, size_t total_streams) = 0;
  virtual int last_launches() const = 0;
  virtual double flops(size_t frames_in) const = 0;
  // Per-stage CUDA-event timing of process() calls (events are recorded on the launching stream).
  virtual void enable_timing(bool on) = 0;
  virtual int stage_times(float *ms, int max_stages) = 0;          // of the most recent timed call
  // Algorithmic work of stage i for one whole-stream call on frames_in frames, all lanes:
  // flops (SURVEY.md 8d accounting) and bytes (unique input samples read + output samples written).
  virtual int stage_work(size_t frames_in, int stage, double *flops, double *bytes, double *launch_units) const = 0;
  // Name of the kernel the most recent call launched for stage i ("" before the first call).
  virtual const char *stage_kernel(int stage) const = 0;
};

class IStream {
 public:
  virtual ~IStream() {}
  virtual const Design &design() const = 0;
  virtual int push(const float *host_interleaved, size_t frames) = 0;
  virtual int pull(float *host_interleaved, void *host_native_planar, size_t max_frames, size_t *got) = 0;
  virtual int drain() = 0;
  // RR_flow: pull, push, pull into the remaining space, with one wait for all three transfers.
  virtual int flow(const float *host_in, size_t isamp, float *host_out, size_t osamp, size_t *iused, size_t *ogen) = 0;
  // Keep the last FIFO in the engine's own sample type (un-cast fp64 tap). Before the first push only.
  virtual int enable_native_tap() = 0;
  virtual int dft_spectrum(int instance, void *out, int max_n) const = 0;
};

// sample_bytes: 4 or 8. device < 0: current device. Returns nullptr and sets *err on failure.
IBatch *create_batch(const RR_config &cfg, int sample_bytes, int nchannels, int nstreams, size_t frames_in_max,
                     int device, int *err);
IStream *create_stream(const RR_config &cfg, int sample_bytes, int nchannels, int device, int *err);

}  // namespace b200rate
