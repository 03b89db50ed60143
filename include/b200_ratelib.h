/* b200_ratelib.h -- C ABI of libb200rate.so, the B200-native drop-in for the reference's `rate` library.
 *
 * Part 1 re-declares, with identical names, argument meaning and return codes, the entry points a caller of
 * the reference binds: /root/reference/rate/ratelib.h:25-81 (public API used by chain.h:22-43 and
 * foo_dsp_rate.cpp) and /root/reference/rate/rate_i.h:39-43 (engine constructors, needed to select the
 * fp32 engine with Best quality -- BASELINE config 1). A translation unit written against the reference's
 * ratelib.h links against libb200rate.so unchanged.
 *
 * Part 2 (RRX_*) are extensions that do not exist in the reference: device-resident batch processing so
 * throughput can be measured with inputs already in HBM, plan/coefficient introspection for parity tests,
 * and an un-cast tap of the fp64 engine's output for the 1e-12 contract.
 *
 * Every function here needs a CUDA device except those marked [host-only]; without a usable device
 * RR_open/RR_ctor_* fail (RR_INTERNAL / NULL) -- there is no CPU fallback.
 */
#ifndef B200_RATELIB_H
#define B200_RATELIB_H

#include <stddef.h>
#include <stdint.h>
#include "rr_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ */
/* Part 1: the reference's interface                                                                */
/* ------------------------------------------------------------------------------------------------ */

enum RR_error {          /* rate/ratelib.h:25-34 */
  RR_OK = 0,
  RR_ENOMEM,
  RR_INTERNAL,           /* also: CUDA failure, no device */
  RR_NULLHANDLE,
  RR_RATEERROR,
  RR_EXTUNINIT,          /* RR_open before init_ratelib */
  RR_INVPARAM
};

enum RR_quality { RR_best = 0, RR_norm = 1 };                      /* rate/ratelib.h:36-42 */
enum RR_phase { RR_minimum = 0, RR_linear = 50, RR_maximum = 100 }; /* rate/ratelib.h:44-49 */

typedef float fb_sample_t;                                         /* rate/ratelib.h:51 */

typedef struct RR_config_tag {                                     /* rate/ratelib.h:53-63 */
  size_t in_rate;
  size_t out_rate;
  double phase;          /* 0..100; 50 = linear phase */
  double bandwidth;      /* pass-band end (-3 dB point) in % of Nyquist */
  int allow_aliasing;
  enum RR_quality quality;
} RR_config;

typedef struct RR_handle_tag RR_handle;                            /* opaque; first member is the vtable */

/* rate/rate_uni.c:210-223. Must be called once before RR_open; `oom` is invoked when a HOST allocation
 * fails (the reference's callback contract, rate/xmalloc.c:38-43). Returns 0, or -1 if oom is NULL. */
int init_ratelib(void (*oom)(void));
/* rate/rate_uni.c:225-231 (defined there, not declared in ratelib.h). */
void close_ratelib(void);

/* rate/rate_uni.c:27-57: Best -> fp64 engine, Normal -> fp32 engine. Counts are FRAMES, buffers are
 * caller-owned interleaved float32 HOST memory. */
int RR_open(const RR_config *config, int nchannels, RR_handle **const handle);
/* rate/rate_base.h:571-614: pull, then push, then pull again into the remaining space. */
int RR_flow(RR_handle *h, const fb_sample_t *ibuf, fb_sample_t *obuf, size_t isamp, size_t osamp,
            size_t *iused, size_t *ogen);
/* rate/rate_base.h:616-636: consumes min(isamp, isamp_max) frames (silently, like the reference). Returns when
 * ibuf has been consumed (copied into the handle's page-locked staging, or -- for a page-locked ibuf -- transferred),
 * NOT when the stage kernels have run: they execute asynchronously on the handle's CUDA stream, and a device failure
 * in one of them is reported by the next RR_pull / RR_drain / RR_flow of the handle. */
int RR_push(RR_handle *h, const fb_sample_t *ibuf, size_t isamp);
/* rate/rate_base.h:638-660: writes at most osamp frames; *ogen = frames written (0: nothing ready). */
int RR_pull(RR_handle *h, fb_sample_t *obuf, size_t osamp, size_t *ogen);
/* rate/rate_base.h:662-672: zero-feed until round(frames_in * out_rate / in_rate) frames exist. */
int RR_drain(RR_handle *h);
void RR_close(RR_handle **h);                                      /* rate/rate_uni.c:83-90 */
const char *RR_strerror(int error);                                /* rate/rate_uni.c:92-111 */

/* rate/rate_i.h:39-43. The reference has one constructor per CPU flavour; here _float/_SSE create the
 * fp32 engine and _double/_SSE3 the fp64 engine, with the quality taken from *config. NULL on failure
 * (the reference swallows RR_INVPARAM here, rate/rate_base.h:738; this library returns NULL instead). */
RR_handle *RR_ctor_SSE3(const RR_config *config, int nchannels);
RR_handle *RR_ctor_double(const RR_config *config, int nchannels);
RR_handle *RR_ctor_SSE(const RR_config *config, int nchannels);
RR_handle *RR_ctor_float(const RR_config *config, int nchannels);

/* ------------------------------------------------------------------------------------------------ */
/* Part 2: extensions                                                                               */
/* ------------------------------------------------------------------------------------------------ */

/* [host-only] Stage plan for a configuration: the integers of the parity contract.
 * sample_bytes 4 or 8. Returns RR_OK or RR_INVPARAM. */
int RRX_plan(const RR_config *config, int sample_bytes, rr_plan *out);
/* [host-only] Designed banks, converted to the engine type (what rate_shared_t holds, rate_base.h:89-92),
 * EXCEPT that dft coefficient banks are returned in the time domain (before the forward transform).
 * kind 0: DFT filter 0 time-domain taps placed/scaled as rate_base.h:173-175 (dft_length doubles)
 * kind 1: DFT filter 1 likewise; kind 2: polyphase bank [phase][tap][order..0] (doubles).
 * Returns the number of doubles available; copies min(max_n, that). */
int RRX_design_dump(const RR_config *config, int sample_bytes, int kind, double *out, int max_n);

/* Same as the plan above but from a live handle. */
int RRX_plan_dump(const RR_handle *h, rr_plan *out);
/* Like RR_pull but planar and in the engine's own sample type (float for the fp32 engine, double for the
 * fp64 engine): out[ch * osamp + i]. The fp64 "tap" for the 1e-12 check. The output FIFO of a handle normally
 * holds interleaved float frames (what RR_pull delivers, rate/rate_base.h:559-563); a handle of the fp64 engine
 * keeps it in double only after RRX_enable_native_tap, which must be called before the first RR_push. */
int RRX_enable_native_tap(RR_handle *h);
int RRX_pull_native(RR_handle *h, void *out, size_t osamp, size_t *ogen);
/* Frequency-domain DFT coefficient bank as uploaded to the device, engine type (dft_length values). */
int RRX_dft_spectrum(const RR_handle *h, int instance, void *out, int max_n);

/* ---- device-resident batch converter ----
 * One plan applied to nstreams independent streams of nchannels each, all of the same length, processed
 * in one shot (push everything + drain). Buffers are DEVICE pointers:
 *   d_in  : float32 [nstreams][frames_in ][nchannels]  (interleaved per stream)
 *   d_out : float32 [nstreams][frames_out][nchannels]  with frames_out = RRX_batch_frames_out()
 * Work and intermediate buffers are allocated once in RRX_batch_open for frames_in_max.
 * `stream` is a cudaStream_t (passed as void* so this header needs no CUDA include). */
typedef struct RRX_batch_tag RRX_batch;

int RRX_batch_open(const RR_config *config, int sample_bytes, int nchannels, int nstreams,
                   size_t frames_in_max, int device, RRX_batch **out);
/* round(frames_in * out_rate / in_rate) computed like rate_flush (rate/rate_base.h:457). */
size_t RRX_batch_frames_out(const RRX_batch *b, size_t frames_in);
int RRX_batch_process(RRX_batch *b, const float *d_in, size_t frames_in, float *d_out, void *stream);
/* Time-chunked processing (the primitive for sharding one long stream across GPUs with filter-history
 * halos and an exactly computed start phase). Step 1: ask which input frames
 * [*in_first, *in_first + *in_count) of a stream of frames_in_total frames the output frames
 * [out_begin, out_begin + out_count) depend on. */
int RRX_batch_input_window(const RRX_batch *b, size_t frames_in_total, uint64_t out_begin, size_t out_count,
                           uint64_t *in_first, uint64_t *in_count);
/* Step 2: d_in_window holds input frames [window_first, window_first + window_frames) of every stream
 * (float32 [nstreams][window_frames][nchannels]) and must cover the window reported above; d_out receives
 * output frames [out_begin, out_begin + out_count) as float32 [nstreams][out_count][nchannels]. The samples
 * are bit-identical to the same frames of a whole-stream RRX_batch_process call. */
int RRX_batch_process_range(RRX_batch *b, const float *d_in_window, uint64_t window_first, size_t window_frames,
                            size_t frames_in_total, uint64_t out_begin, size_t out_count, float *d_out,
                            void *stream);
/* Engine-type planar output of the last stage for the same call shape as RRX_batch_process:
 * d_out_native[(s * nchannels + c) * frames_out + i] (float or double). */
int RRX_batch_process_native(RRX_batch *b, const float *d_in, size_t frames_in, void *d_out_native,
                             void *stream);
/* Host-buffer batch: total_streams streams in HOST memory (float32 [total_streams][frames_in][nchannels],
 * ideally pinned) are moved through the device in sub-batches of the batch's nstreams; H2D copy, kernels and
 * D2H copy of consecutive sub-batches overlap on three CUDA streams. Blocking. */
int RRX_batch_process_host(RRX_batch *b, const float *h_in, size_t frames_in, float *h_out, size_t total_streams);
/* Per-stage device timing of RRX_batch_process* calls with CUDA events recorded on the launching stream.
 * RRX_batch_stage_times waits for the most recent timed call and returns the number of stages written. */
int RRX_batch_enable_timing(RRX_batch *b, int on);
int RRX_batch_stage_times(RRX_batch *b, float *ms, int max_stages);
/* Algorithmic work of one stage for a whole-stream call on frames_in frames over all lanes: FLOPs, bytes
 * (unique input samples read + output samples written) and launch units (DFT blocks or output samples). */
int RRX_batch_stage_work(const RRX_batch *b, size_t frames_in, int stage, double *flops, double *bytes, double *units);
/* Name of the kernel the most recent RRX_batch_process* call launched for `stage` (static string, "" if none). */
const char *RRX_batch_stage_kernel(const RRX_batch *b, int stage);
int RRX_batch_plan(const RRX_batch *b, rr_plan *out);
/* Kernel launches issued by the most recent RRX_batch_process* call. */
int RRX_batch_last_launches(const RRX_batch *b);
/* Algorithmic FLOP count of one RRX_batch_process call on frames_in frames (SURVEY.md 8d accounting). */
double RRX_batch_flops(const RRX_batch *b, size_t frames_in);
void RRX_batch_close(RRX_batch **b);

/* ---- several GPUs of one box behind one handle ----
 * The path shards without a data-path collective: a batch by independent stream (device k of n converts a contiguous
 * slice of the streams), one long stream by time chunk (device k converts range k of the output timeline from its own
 * halo'd input window, cut at multiples of the last stage's block, phase and block grid exact). One host thread per
 * device drives it. NCCL (ncclSend / ncclRecv over NVLink, loaded on first use) appears only in RRX_multi_gather.
 * `devices` lists CUDA device ordinals; listing one device several times is allowed (the shards then share it). */
typedef struct RRX_multi_tag RRX_multi;
int RRX_multi_open(const RR_config *config, int sample_bytes, int nchannels, size_t nstreams, size_t frames_in_max,
                   const int *devices, int ndevices, RRX_multi **out);
int RRX_multi_devices(const RRX_multi *m);
/* Slice of the streams device index k owns. */
int RRX_multi_shard(const RRX_multi *m, int k, int *device, size_t *first_stream, size_t *stream_count);
size_t RRX_multi_frames_out(const RRX_multi *m, size_t frames_in);
/* HOST buffers (ideally page-locked), float32 [nstreams][frames][nchannels]: every device moves its slice through
 * itself with the three-stream pipeline of RRX_batch_process_host. Blocking. */
int RRX_multi_process_host(RRX_multi *m, const float *h_in, size_t frames_in, float *h_out);
/* One long stream (the handle was opened with nstreams == 1; frames_in_max bounds the input window of one call):
 * h_in holds all frames_in_total frames, h_out receives all output frames, *frames_out their number. */
int RRX_multi_process_stream_host(RRX_multi *m, const float *h_in, size_t frames_in_total, float *h_out, size_t *frames_out);
/* Device-resident: d_in[k] is device k's slice on that device; results stay on the devices ... */
int RRX_multi_process(RRX_multi *m, const float *const *d_in, size_t frames_in);
int RRX_multi_result(const RRX_multi *m, int k, const float **d_out, size_t *frames_out);
/* ... until gathered on device index `root`: d_out_root is float32 [nstreams][frames_out][nchannels] on that device. */
int RRX_multi_gather(RRX_multi *m, int root, float *d_out_root);
void RRX_multi_close(RRX_multi **m);
/* Page-locked ("pinned") host memory for callers that do not link CUDA themselves: buffers from here let the
 * host-buffer entry points (RR_push / RR_pull, RRX_batch_process_host, RRX_multi_process_*_host) transfer at the
 * speed of the host link, without a staging copy. NULL on failure. */
void *RRX_host_alloc(size_t bytes);
void RRX_host_free(void *p);

/* ------------------------------------------------------------------------------------------------ */
/* Part 3: track-edge extrapolation (the step next to the path in the caller)  */
/* ------------------------------------------------------------------------------------------------ */
/* The plugin does not feed a track's first and last frames to the rate engine as hard edges: it predicts
 * N_samples_to_add_ frames before the beginning and after the end by linear prediction from the first / last
 * PRIME_LEN_ frames, pushes them with the track, and drops N_samples_to_drop_ output frames at each end
 * (foo_dsp_rate.cpp:96-101,165-167,241-313). These entry points run that prediction on the device, bit-identical to
 * lpc/lpc.cpp, so that a device-resident batch conversion reproduces the plugin's track edges. */
#define RRX_LPC_ORDER 32                                         /* lpc/lpc.h:24 */
/* [host-only] foo_dsp_rate.cpp:96-101 + util.h:38-49: frames to add per edge at the input rate, frames to drop per
 * edge at the output rate (the same duration), prime length, and the plugin's input block. NULL outputs are skipped. */
int RRX_track_edge_lengths(unsigned in_rate, unsigned out_rate, unsigned *add, unsigned *drop, unsigned *prime_len,
                           unsigned *inbuf_frames);
/* lpc/lpc.h:26 lpc_extrapolate2, same arguments and memory layout: `data` is HOST memory pointing at frame 0 of
 * data_len interleaved base frames; frames [-extra_bkwd, 0) and [data_len, data_len + extra_fwd) are written.
 * lpc_order 1..32 (the reference's callers use LPC_ORDER = 32), else RR_INVPARAM. Blocking. Runs on the calling thread's
 * current CUDA device (the device-resident entry points below: on the device that owns d_data); no device: RR_INTERNAL
 * with the CUDA error in RRX_last_error -- there is no host fallback. */
int RRX_lpc_extrapolate2(float *data, size_t data_len, int nchannels, int lpc_order, size_t extra_bkwd, size_t extra_fwd);
/* lpc/lpc.h:28-38, the two inline wrappers. */
int RRX_lpc_extrapolate_bkwd(float *data, size_t data_len, size_t prime_len, int nchannels, int lpc_order, size_t extra_bkwd);
int RRX_lpc_extrapolate_fwd(float *data, size_t data_len, size_t prime_len, int nchannels, int lpc_order, size_t extra_fwd);
/* Device-resident batch of the same operation: d_data points at frame 0 of stream 0's base segment, stream s starts
 * stream_stride_frames * s frames later; every (stream, channel) lane is extrapolated independently. Asynchronous on
 * `stream` (a cudaStream_t). */
int RRX_lpc_extrapolate_batch(float *d_data, size_t nstreams, size_t stream_stride_frames, size_t data_len, int nchannels,
                              int lpc_order, size_t extra_bkwd, size_t extra_fwd, void *stream);
/* Both edges of whole tracks in one pair of launches: d_padded is float32 [nstreams][extra + track_frames + extra]
 * [nchannels] with the tracks in the middle; the leading `extra` frames are predicted backward from the first
 * min(prime_len, track_frames) frames, the trailing ones forward from the last (foo_dsp_rate.cpp:243-245). The padded
 * buffer is what RRX_batch_process takes as d_in with frames_in = track_frames + 2 * extra. */
int RRX_lpc_extend_tracks(float *d_padded, size_t nstreams, size_t track_frames, size_t prime_len, int nchannels,
                          int lpc_order, size_t extra, void *stream);
/* Test tap: per (stream, channel) lane 66 doubles -- lags 0..32, the 32 damped predictor coefficients, usable order. */
int RRX_lpc_analysis_dump(const float *d_data, size_t nstreams, size_t stream_stride_frames, size_t data_len, int nchannels,
                          int lpc_order, double *d_out, void *stream);

/* Last CUDA error string seen by this library on the calling thread ("" if none). */
const char *RRX_last_error(void);
/* [host-only] Library version string. */
const char *RRX_version(void);

#ifdef __cplusplus
}
#endif
#endif
