"""ctypes binding of the C ABI declared in ``include/b200_ratelib.h``.

The product library is ``foo_dsp_resampler_b200/libb200rate.so`` (CUDA, sm_100a). ``bind(path)`` can also
bind another build of the same ABI -- the test suite uses that for the host emulation build under
``tests/emu`` -- but nothing in this package ever does so: the package-level API always loads the CUDA
library and raises if it is missing."""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(PKG_DIR, "libb200rate.so")

RR_OK, RR_ENOMEM, RR_INTERNAL, RR_NULLHANDLE, RR_RATEERROR, RR_EXTUNINIT, RR_INVPARAM = range(7)
RR_BEST, RR_NORM = 0, 1


class RRConfig(C.Structure):
    """RR_config (reference: rate/ratelib.h:53-63)."""
    _fields_ = [("in_rate", C.c_size_t), ("out_rate", C.c_size_t), ("phase", C.c_double),
                ("bandwidth", C.c_double), ("allow_aliasing", C.c_int), ("quality", C.c_int)]


class StagePlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "hb_coefs", "pre", "pre_post", "preload", "L", "remL", "remM", "n", "phase_bits",
        "interp_order", "dft_filter_num", "dft_length", "num_taps", "post_peak", "step_int")] + [
        ("at", C.c_int64), ("step", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Plan(C.Structure):
    _fields_ = [("num_stages", C.c_int32), ("sample_bytes", C.c_int32), ("factor", C.c_double),
                ("isamp_max", C.c_uint64), ("st", StagePlan * 24)]

    def as_dict(self):
        return {"num_stages": self.num_stages, "sample_bytes": self.sample_bytes, "factor": self.factor,
                "isamp_max": self.isamp_max, "stages": [self.st[i].as_dict() for i in range(self.num_stages)]}


OOM_FN = C.CFUNCTYPE(None)

# name -> (restype, argtypes): every symbol include/b200_ratelib.h declares
SYMBOLS = {
    "init_ratelib": (C.c_int, [OOM_FN]),
    "close_ratelib": (None, []),
    "RR_open": (C.c_int, [C.POINTER(RRConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "RR_flow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t),
                          C.POINTER(C.c_size_t)]),
    "RR_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "RR_pull": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "RR_drain": (C.c_int, [C.c_void_p]),
    "RR_close": (None, [C.POINTER(C.c_void_p)]),
    "RR_strerror": (C.c_char_p, [C.c_int]),
    "RR_ctor_SSE3": (C.c_void_p, [C.POINTER(RRConfig), C.c_int]),
    "RR_ctor_double": (C.c_void_p, [C.POINTER(RRConfig), C.c_int]),
    "RR_ctor_SSE": (C.c_void_p, [C.POINTER(RRConfig), C.c_int]),
    "RR_ctor_float": (C.c_void_p, [C.POINTER(RRConfig), C.c_int]),
    "RRX_plan": (C.c_int, [C.POINTER(RRConfig), C.c_int, C.POINTER(Plan)]),
    "RRX_design_dump": (C.c_int, [C.POINTER(RRConfig), C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "RRX_plan_dump": (C.c_int, [C.c_void_p, C.POINTER(Plan)]),
    "RRX_enable_native_tap": (C.c_int, [C.c_void_p]),
    "RRX_pull_native": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "RRX_dft_spectrum": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "RRX_batch_open": (C.c_int, [C.POINTER(RRConfig), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                 C.POINTER(C.c_void_p)]),
    "RRX_batch_frames_out": (C.c_size_t, [C.c_void_p, C.c_size_t]),
    "RRX_batch_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "RRX_batch_input_window": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_size_t, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64)]),
    "RRX_batch_process_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t, C.c_size_t, C.c_uint64,
                                          C.c_size_t, C.c_void_p, C.c_void_p]),
    "RRX_batch_process_native": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "RRX_batch_process_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "RRX_batch_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "RRX_batch_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "RRX_batch_stage_work": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]),
    "RRX_batch_stage_kernel": (C.c_char_p, [C.c_void_p, C.c_int]),
    "RRX_batch_plan": (C.c_int, [C.c_void_p, C.POINTER(Plan)]),
    "RRX_multi_open": (C.c_int, [C.POINTER(RRConfig), C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_int), C.c_int,
                                 C.POINTER(C.c_void_p)]),
    "RRX_multi_devices": (C.c_int, [C.c_void_p]),
    "RRX_multi_shard": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "RRX_multi_frames_out": (C.c_size_t, [C.c_void_p, C.c_size_t]),
    "RRX_multi_process_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "RRX_multi_process_stream_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "RRX_multi_process": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]),
    "RRX_multi_result": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "RRX_multi_gather": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "RRX_multi_close": (None, [C.POINTER(C.c_void_p)]),
    "RRX_host_alloc": (C.c_void_p, [C.c_size_t]),
    "RRX_host_free": (None, [C.c_void_p]),
    "RRX_batch_last_launches": (C.c_int, [C.c_void_p]),
    "RRX_batch_flops": (C.c_double, [C.c_void_p, C.c_size_t]),
    "RRX_batch_close": (None, [C.POINTER(C.c_void_p)]),
    "RRX_track_edge_lengths": (C.c_int, [C.c_uint, C.c_uint] + [C.POINTER(C.c_uint)] * 4),
    "RRX_lpc_extrapolate2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t]),
    "RRX_lpc_extrapolate_bkwd": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t]),
    "RRX_lpc_extrapolate_fwd": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t]),
    "RRX_lpc_extrapolate_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t,
                                            C.c_size_t, C.c_void_p]),
    "RRX_lpc_extend_tracks": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t,
                                        C.c_void_p]),
    "RRX_lpc_analysis_dump": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p]),
    "RRX_last_error": (C.c_char_p, []),
    "RRX_version": (C.c_char_p, []),
}


# entry points that exist only in the CUDA library (csrc/lpc.cu is CUDA-only code)
DEVICE_ONLY = {n for n in SYMBOLS if n.startswith("RRX_lpc_") or n == "RRX_track_edge_lengths"}


@OOM_FN
def _oom():
    raise MemoryError("libb200rate: host allocation failed")


def bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        if path != PRODUCT_SO and name in DEVICE_ONLY and not hasattr(lib, name):
            continue                       # the host emulation build of the test suite has no lpc.cu
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.init_ratelib(_oom) != 0:
        raise RuntimeError("init_ratelib failed")
    return lib


_product = None


def product():
    """The CUDA library. No fallback: a missing or unloadable extension is an error."""
    global _product
    if _product is None:
        if not os.path.exists(PRODUCT_SO):
            raise RuntimeError(
                "libb200rate.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
        _product = bind(PRODUCT_SO)
    return _product


def make_config(in_rate, out_rate, phase=50.0, bandwidth=95.0, allow_aliasing=0, quality=RR_BEST):
    return RRConfig(int(in_rate), int(out_rate), float(phase), float(bandwidth), int(allow_aliasing), int(quality))


class RateError(RuntimeError):
    def __init__(self, lib, code, where):
        self.code = code
        detail = lib.RRX_last_error().decode() if code in (RR_INTERNAL, RR_ENOMEM, RR_INVPARAM) else ""
        super().__init__("%s: %s%s" % (where, lib.RR_strerror(code).decode(), (" (" + detail + ")") if detail else ""))
