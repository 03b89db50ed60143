/* oracle/ref_engine_d.c -- TEST INFRASTRUCTURE. Compiles the reference's generic double engine TU
 * (rate/rate_double.c -> rate_base.h) in place and appends the read-only taps. */
#include "rate_double.c"
#define TAP_SUFFIX _d
#include "ref_tap.inc"
