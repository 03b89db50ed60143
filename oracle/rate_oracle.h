/* oracle/rate_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of the reference's `rate` resampling path, written
 * from the algorithm, not copied: every stage is expressed in ABSOLUTE stream coordinates (closed-form
 * block/phase positions) and the fp32 FFT is evaluated level-synchronously -- the same schedule the
 * sm_100a kernels use -- yet it produces the reference's bits because it evaluates the same expression
 * DAG. It is pinned by tests/test_oracle_vs_reference.py against the compiled reference
 * (oracle/_ref/libref_rate.so) and by the committed fixtures in tests/golden/ (generated from that
 * reference by tests/golden/make_golden.py). The reference holds no golden vectors of its own
 * (SURVEY.md section 4).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#ifndef RATE_ORACLE_H
#define RATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "rr_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as RR_config (rate/ratelib.h:53-63). */
typedef struct orc_config {
  size_t in_rate, out_rate;
  double phase;       /* 0..100, 50 = linear */
  double bandwidth;   /* % of Nyquist at -3 dB */
  int allow_aliasing;
  int quality;        /* 0 = best, 1 = normal */
} orc_config;

typedef struct orc_handle orc_handle;

/* sample_bytes: 4 -> follows rate_float.c, 8 -> follows rate_double.c. NULL on invalid ratio. */
orc_handle *orc_open(const orc_config *cfg, int nchannels, int sample_bytes);
void orc_close(orc_handle *h);
/* keep != 0: FIFOs are never compacted, so orc_fifo_read can address any absolute index. */
void orc_keep_history(orc_handle *h, int keep);

/* RR_push_x (rate/rate_base.h:616-636): returns frames accepted (silently clamped to isamp_max). */
size_t orc_push(orc_handle *h, const float *interleaved, size_t frames);
/* RR_pull_x (rate/rate_base.h:638-660). */
size_t orc_pull(orc_handle *h, float *interleaved, size_t max_frames);
/* Like orc_pull without the cast to float: planar out[ch * max_frames + i] in the engine type. */
size_t orc_pull_native(orc_handle *h, void *planar, size_t max_frames);
/* RR_drain_x / rate_flush (rate/rate_base.h:454-468,662-672). */
void orc_drain(orc_handle *h);

int orc_plan_dump(const orc_handle *h, rr_plan *out);
/* Designed banks in the engine's sample type; return the element count available. */
int orc_dft_coefs(const orc_handle *h, int instance, void *out, int max_n);
int orc_poly_coefs(const orc_handle *h, void *out, int max_n);
/* Raw double-precision prototype filters before conversion (designer parity). */
int orc_dft_taps(const orc_handle *h, int instance, double *out, int max_n);

/* FIFO i feeds stage i; FIFO num_stages is the output queue. Absolute counters. */
uint64_t orc_fifo_written(const orc_handle *h, int fifo_index);
uint64_t orc_fifo_consumed(const orc_handle *h, int fifo_index);
/* Copy count samples starting at absolute index start (needs keep_history); returns copied. */
size_t orc_fifo_read(const orc_handle *h, int channel, int fifo_index, uint64_t start, size_t count, void *out);

/* Stand-alone transforms (unit tests; packed [Re0, Re(N/2), Re1, Im1, ...]).
 * orc_rdft_f32 follows ff_rdft_calc_c (rate/fft-float/rdft.c:33-84, fft.c:169-346),
 * orc_rdft_f64 follows lsx_rdft_generic (rate/fft-double/fft4g_dbl.c:26-62). */
void orc_rdft_f32(int n, int inverse, float *data);
void orc_rdft_f64(int n, int inverse, double *data);

/* Designer entry points (rate/effects_i_dsp.c:137-171,181-278); caller frees with orc_free. */
double *orc_design_lpf(double Fp, double Fs, double Fn, double att, int *num_taps, int k, double beta);
void orc_fir_to_phase(double **h, int *len, int *post_len, double phase);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
