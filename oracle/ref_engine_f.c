/* oracle/ref_engine_f.c -- TEST INFRASTRUCTURE. Compiles the reference's generic float engine TU
 * (rate/rate_float.c -> rate_base.h) in place and appends the read-only taps. */
#include "rate_float.c"
#define TAP_SUFFIX _f
#include "ref_tap.inc"
