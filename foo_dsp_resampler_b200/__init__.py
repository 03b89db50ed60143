"""foo_dsp_resampler_b200 -- B200-native SoX-`rate` resampling engine behind the reference's C API.

The product is ``libb200rate.so`` (hand-written CUDA for sm_100a + a C++ host planner); this package is a
thin ctypes mirror of its C ABI (``include/b200_ratelib.h``) for tests, benchmarks and Python callers."""
from ._capi import (RR_BEST, RR_NORM, RR_OK, RR_ENOMEM, RR_INTERNAL, RR_NULLHANDLE, RR_RATEERROR,  # noqa: F401
                    RR_EXTUNINIT, RR_INVPARAM, RRConfig, Plan, RateError, make_config, product)
from .converter import (RateConverter, BatchConverter, TrackBatchConverter, resample, lpc_extrapolate2,  # noqa: F401
                        track_edge_lengths, LPC_ORDER)

__all__ = ["RateConverter", "BatchConverter", "TrackBatchConverter", "resample", "lpc_extrapolate2",
           "track_edge_lengths", "LPC_ORDER", "make_config", "RRConfig", "Plan", "RateError",
           "product", "RR_BEST", "RR_NORM"]
