"""ctypes driver for the compiled reference checker ``oracle/_ref/libref_rate.so``.

TEST INFRASTRUCTURE. The library is the unmodified reference ``rate/`` sources compiled by
``oracle/Makefile``; here it is driven through the reference's own C entry points
(``rate/ratelib.h:72-81``, ``rate/rate_i.h:39-43``) plus the read-only taps of ``oracle/ref_tap.inc``.
It is only ever used as the checker.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_rate.so")

RR_BEST, RR_NORM = 0, 1


class RRConfig(C.Structure):
    """RR_config, rate/ratelib.h:53-63."""
    _fields_ = [("in_rate", C.c_size_t), ("out_rate", C.c_size_t), ("phase", C.c_double),
                ("bandwidth", C.c_double), ("allow_aliasing", C.c_int), ("quality", C.c_int)]


class StagePlan(C.Structure):
    """rr_stage_plan, include/rr_plan.h."""
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "hb_coefs", "pre", "pre_post", "preload", "L", "remL", "remM", "n", "phase_bits",
        "interp_order", "dft_filter_num", "dft_length", "num_taps", "post_peak", "step_int")] + [
        ("at", C.c_int64), ("step", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Plan(C.Structure):
    """rr_plan, include/rr_plan.h."""
    _fields_ = [("num_stages", C.c_int32), ("sample_bytes", C.c_int32), ("factor", C.c_double),
                ("isamp_max", C.c_uint64), ("st", StagePlan * 24)]

    def as_dict(self):
        return {"num_stages": self.num_stages, "sample_bytes": self.sample_bytes,
                "factor": self.factor, "isamp_max": self.isamp_max,
                "stages": [self.st[i].as_dict() for i in range(self.num_stages)]}


def make_config(in_rate, out_rate, phase=50.0, bandwidth=95.0, allow_aliasing=0, quality=RR_BEST):
    return RRConfig(int(in_rate), int(out_rate), float(phase), float(bandwidth), int(allow_aliasing),
                    int(quality))


_OOM = C.CFUNCTYPE(None)


@_OOM
def _oom_handler():
    raise MemoryError("reference allocator reported OOM")


_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    """Load the reference and build its global FFT tables for the GENERIC engines
    (both SSE and SSE3 hidden from its cpuid; SURVEY.md 8c gotcha 1)."""
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ref_reinit.argtypes = [C.c_int, C.c_int, _OOM]
        for name in ("RR_ctor_float", "RR_ctor_double", "RR_ctor_SSE", "RR_ctor_SSE3"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.POINTER(RRConfig), C.c_int]
        L.RR_open.argtypes = [C.POINTER(RRConfig), C.c_int, C.POINTER(C.c_void_p)]
        L.RR_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.RR_pull.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.RR_flow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                              C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.RR_drain.argtypes = [C.c_void_p]
        L.RR_close.argtypes = [C.POINTER(C.c_void_p)]
        L.RR_strerror.restype = C.c_char_p
        for sfx in ("_f", "_d"):
            getattr(L, "ref_tap_plan_dump" + sfx).argtypes = [C.c_void_p, C.POINTER(Plan)]
            g = getattr(L, "ref_tap_pull_native" + sfx)
            g.restype = C.c_size_t
            g.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            getattr(L, "ref_tap_dft_coefs" + sfx).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            getattr(L, "ref_tap_poly_coefs" + sfx).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, "ref_tap_fifo_occupancy" + sfx).argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_reinit(1, 1, _oom_handler)
        _lib = L
    return _lib


class RefResampler:
    """One reference handle. ``engine``: 'float' -> RR_ctor_float (rate_float.c),
    'double' -> RR_ctor_double (rate_double.c): the two parity oracles named by BASELINE.json."""

    def __init__(self, cfg, nch, engine="float"):
        self.L = lib()
        self.nch = nch
        self.engine = engine
        self.sfx = "_f" if engine == "float" else "_d"
        self.dtype = np.float32 if engine == "float" else np.float64
        ctor = self.L.RR_ctor_float if engine == "float" else self.L.RR_ctor_double
        self.cfg = cfg
        self.h = C.c_void_p(ctor(C.byref(cfg), nch))
        assert self.h.value

    def plan(self):
        p = Plan()
        rc = getattr(self.L, "ref_tap_plan_dump" + self.sfx)(self.h, C.byref(p))
        assert rc == 0
        return p.as_dict()

    def push(self, x):
        """x: float32 [frames, nch] interleaved."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.nch
        rc = self.L.RR_push(self.h, x.ctypes.data, x.shape[0])
        assert rc == 0
        return rc

    def pull(self, osamp):
        out = np.empty((osamp, self.nch), dtype=np.float32)
        ogen = C.c_size_t(0)
        rc = self.L.RR_pull(self.h, out.ctypes.data, osamp, C.byref(ogen))
        assert rc == 0
        return out[:ogen.value]

    def pull_native(self, osamp):
        """Planar native-precision pop: returns [nch, ogen] in the engine's sample type."""
        out = np.empty((self.nch, osamp), dtype=self.dtype)
        n = getattr(self.L, "ref_tap_pull_native" + self.sfx)(self.h, out.ctypes.data, osamp)
        return out[:, :n]

    def drain(self):
        assert self.L.RR_drain(self.h) == 0

    def fifo_occupancy(self, fifo_index, channel=0):
        return getattr(self.L, "ref_tap_fifo_occupancy" + self.sfx)(self.h, channel, fifo_index)

    def dft_coefs(self, instance):
        pl = self.plan()
        n = max([s["dft_length"] for s in pl["stages"] if s["kind"] == 1 and s["dft_filter_num"] == instance] or [0])
        out = np.empty(n, dtype=self.dtype)
        if n:
            getattr(self.L, "ref_tap_dft_coefs" + self.sfx)(self.h, instance, out.ctypes.data, n)
        return out

    def poly_coefs(self, count):
        out = np.empty(count, dtype=self.dtype)
        got = getattr(self.L, "ref_tap_poly_coefs" + self.sfx)(self.h, out.ctypes.data, count)
        return out[:got]

    def close(self):
        if self.h is not None and self.h.value:
            self.L.RR_close(C.byref(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resample(cfg, x, engine="float", chunk=65536, native=False, pull_chunk=None):
    """Push ``x`` ([frames, nch] float32) in ``chunk``-frame pieces, pulling everything after each
    push, then drain. Returns (y, counts) with y [frames_out, nch] (float32, or the native type when
    ``native``) and counts = frames produced after each push/drain (bookkeeping parity)."""
    nch = x.shape[1]
    r = RefResampler(cfg, nch, engine)
    pull_chunk = pull_chunk or max(chunk * 4, 1 << 16)
    outs, counts = [], []

    def pull_all():
        tot = 0
        while True:
            if native:
                y = r.pull_native(pull_chunk).T.copy()
            else:
                y = r.pull(pull_chunk).copy()
            if y.shape[0] == 0:
                break
            outs.append(y)
            tot += y.shape[0]
        counts.append(tot)

    isamp_max = r.plan()["isamp_max"]
    step = min(chunk, isamp_max)
    for s in range(0, x.shape[0], step):
        r.push(x[s:s + step])
        pull_all()
    r.drain()
    pull_all()
    r.close()
    dt = (np.float32 if engine == "float" else np.float64) if native else np.float32
    y = np.concatenate(outs, axis=0) if outs else np.zeros((0, nch), dtype=dt)
    return y, counts
