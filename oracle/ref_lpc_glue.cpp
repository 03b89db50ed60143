/* oracle/ref_lpc_glue.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * C entry points over the UNMODIFIED reference sources lpc/lpc.cpp (compiled where it lies, see Makefile) and
 * util.h (samples_len is `static`, so it is reached by including the header here; its one SDK dependency,
 * pfc::max_t, comes from shim/pfc/pfc.h). lpc.h uses size_t without including a header that defines it. */
#include <cstddef>
#include <algorithm>
using std::size_t;
#include "lpc/lpc.h"
#include "util.h"

extern "C" void ref_lpc_extrapolate2(float *data, size_t data_len, int nch, int order, size_t extra_bkwd,
                                     size_t extra_fwd)
{
  lpc_extrapolate2(data, data_len, nch, order, extra_bkwd, extra_fwd);
}

extern "C" void ref_lpc_extrapolate_bkwd(float *data, size_t data_len, size_t prime_len, int nch, int order,
                                         size_t extra)
{
  lpc_extrapolate_bkwd(data, data_len, prime_len, nch, order, extra);
}

extern "C" void ref_lpc_extrapolate_fwd(float *data, size_t data_len, size_t prime_len, int nch, int order,
                                        size_t extra)
{
  lpc_extrapolate_fwd(data, data_len, prime_len, nch, order, extra);
}

/* foo_dsp_rate.cpp:96-101, the statements of dsp_rate::reinit that size the edge handling, on top of the
 * reference's own samples_len (util.h:38-49) and LPC_ORDER (lpc/lpc.h:24). */
extern "C" void ref_track_edge_lengths(unsigned in_rate, unsigned out_rate, unsigned *add, unsigned *drop,
                                       unsigned *prime_len, unsigned *inbuf)
{
  unsigned a = in_rate, d = out_rate;
  samples_len(&a, &d, 20, 8192u);
  unsigned ib = std::max(in_rate / 10, 2048u);
  ib = std::min(ib, 65536u);
  unsigned p = std::max(in_rate / 20, 1024u);
  p = std::min(p, 16384u);
  p = std::max<unsigned>(p, 2 * LPC_ORDER + 1);
  *add = a;
  *drop = d;
  *prime_len = p;
  *inbuf = ib;
}

extern "C" int ref_lpc_order(void) { return (int)LPC_ORDER; }
