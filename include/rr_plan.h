/* rr_plan.h -- plain-C description of a resampler stage plan.
 *
 * One struct that the product (RRX_plan_dump, include/b200_ratelib.h), the oracle restatement
 * (oracle/rate_oracle.h) and the compiled reference tap (oracle/ref_tap.inc) all fill in, so that
 * "every integer the parity contract names"  can be compared with memcmp-like
 * strictness. Field meaning follows the reference's stage_t (rate/rate_base.h:96-128) and
 * dft_filter_t (rate/rate_base.h:83-87).
 */
#ifndef RR_PLAN_H
#define RR_PLAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum rr_stage_kind {
  RR_STAGE_HALFBAND = 0, /* h8..h13, 2:1 decimating half-band FIR  (rate_filters_generic.h:80-249) */
  RR_STAGE_DFT      = 1, /* overlap-save DFT FIR, dft_stage_fn      (dft_filter.h:60-190)           */
  RR_STAGE_POLY     = 2  /* vpoly0..3 polyphase FIR                 (rate_filters_generic.h:272-504) */
};

#define RR_MAX_STAGES 24

typedef struct rr_stage_plan {
  int32_t kind;         /* enum rr_stage_kind */
  int32_t hb_coefs;     /* half-band: number of one-sided coefs (8..13), else 0 */
  int32_t pre;          /* stage_t.pre      */
  int32_t pre_post;     /* stage_t.pre_post */
  int32_t preload;      /* stage_t.preload  */
  int32_t L;            /* stage_t.L (dft: upsampling factor; poly: number of phases when rational) */
  int32_t remL;         /* stage_t.remL at open */
  int32_t remM;         /* stage_t.remM at open */
  int32_t n;            /* poly: taps per phase */
  int32_t phase_bits;   /* poly: stage_t.phase_bits */
  int32_t interp_order; /* poly: 0 = vpoly0 ... 3 = vpoly3 */
  int32_t dft_filter_num;
  int32_t dft_length;   /* dft: N */
  int32_t num_taps;     /* dft: T */
  int32_t post_peak;    /* dft */
  int32_t step_int;     /* dft: stage_t.step.parts.integer (M, or -log2(M) for F-domain decimation) */
  int64_t at;           /* poly: stage_t.at.all at open   */
  int64_t step;         /* poly: stage_t.step.all         */
} rr_stage_plan;

typedef struct rr_plan {
  int32_t num_stages;
  int32_t sample_bytes; /* 4 = float engine, 8 = double engine */
  double factor;        /* in_rate / out_rate */
  uint64_t isamp_max;   /* RR_internal.isamp_max (rate_base.h:531) */
  rr_stage_plan st[RR_MAX_STAGES];
} rr_plan;

#ifdef __cplusplus
}
#endif
#endif
