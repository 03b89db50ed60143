"""Host-side mirror of the reference's resampler interface over the C ABI.

``RateConverter`` is the reference's ``resampler_link`` (chain.h:22-43): open / push / pull / drain / close
on interleaved float32 HOST buffers. ``BatchConverter`` is the device-resident extension (RRX_batch_*).
Both take the C library as a parameter so the test suite can run the identical Python code over the host
emulation build; the package-level constructors always use the CUDA library."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import RR_OK, Plan, RateError


class RateConverter:
    """engine: 'auto' = RR_open (Best -> fp64 engine, Normal -> fp32 engine, rate/rate_uni.c:38-51);
    'float' / 'double' = RR_ctor_float / RR_ctor_double with the config's quality (rate/rate_i.h:39-43)."""

    def __init__(self, cfg, nchannels, engine="auto", lib=None, native_tap=False):
        self.lib = lib or _capi.product()
        self.nch = int(nchannels)
        self.h = C.c_void_p()
        if engine == "auto":
            rc = self.lib.RR_open(C.byref(cfg), self.nch, C.byref(self.h))
            if rc != RR_OK:
                raise RateError(self.lib, rc, "RR_open")
        else:
            ctor = {"float": self.lib.RR_ctor_float, "double": self.lib.RR_ctor_double,
                    "SSE": self.lib.RR_ctor_SSE, "SSE3": self.lib.RR_ctor_SSE3}[engine]
            self.h = C.c_void_p(ctor(C.byref(cfg), self.nch))
            if not self.h.value:
                raise RateError(self.lib, _capi.RR_INTERNAL, "RR_ctor_" + engine)
        self.sample_bytes = self.plan()["sample_bytes"]
        self.dtype = np.float32 if self.sample_bytes == 4 else np.float64
        if native_tap:                      # keep the last FIFO un-cast (fp64 engine): before the first push
            rc = self.lib.RRX_enable_native_tap(self.h)
            if rc != RR_OK:
                raise RateError(self.lib, rc, "RRX_enable_native_tap")

    def plan(self):
        p = Plan()
        rc = self.lib.RRX_plan_dump(self.h, C.byref(p))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_plan_dump")
        return p.as_dict()

    def push(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.nch
        rc = self.lib.RR_push(self.h, x.ctypes.data, x.shape[0])
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RR_push")

    def pull(self, osamp):
        out = np.empty((osamp, self.nch), dtype=np.float32)
        ogen = C.c_size_t(0)
        rc = self.lib.RR_pull(self.h, out.ctypes.data, osamp, C.byref(ogen))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RR_pull")
        return out[:ogen.value]

    def pull_native(self, osamp):
        out = np.empty((self.nch, osamp), dtype=self.dtype)
        ogen = C.c_size_t(0)
        rc = self.lib.RRX_pull_native(self.h, out.ctypes.data, osamp, C.byref(ogen))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_pull_native")
        return out[:, :ogen.value]

    def flow(self, x, osamp):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty((osamp, self.nch), dtype=np.float32)
        iused, ogen = C.c_size_t(0), C.c_size_t(0)
        rc = self.lib.RR_flow(self.h, x.ctypes.data if x.shape[0] else None, out.ctypes.data, x.shape[0], osamp,
                              C.byref(iused), C.byref(ogen))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RR_flow")
        return out[:ogen.value], iused.value

    def drain(self):
        rc = self.lib.RR_drain(self.h)
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RR_drain")

    def dft_spectrum(self, instance):
        n = self.lib.RRX_dft_spectrum(self.h, instance, None, 0)
        out = np.empty(max(n, 0), dtype=self.dtype)
        if n > 0:
            self.lib.RRX_dft_spectrum(self.h, instance, out.ctypes.data, n)
        return out

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.RR_close(C.byref(self.h))
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resample(cfg, x, engine="float", chunk=65536, native=False, pull_chunk=None, lib=None):
    """Push ``x`` ([frames, nch] float32) in ``chunk``-frame pieces, pulling everything after each push,
    then drain: the driving pattern of foo_dsp_rate.cpp:165-187,218-313. Returns (y, counts)."""
    nch = x.shape[1]
    r = RateConverter(cfg, nch, engine, lib=lib, native_tap=native)
    pull_chunk = pull_chunk or max(chunk * 4, 1 << 16)
    outs, counts = [], []

    def pull_all():
        tot = 0
        while True:
            y = (r.pull_native(pull_chunk).T.copy() if native else r.pull(pull_chunk).copy())
            if y.shape[0] == 0:
                break
            outs.append(y)
            tot += y.shape[0]
        counts.append(tot)

    step = min(chunk, r.plan()["isamp_max"])
    for s in range(0, x.shape[0], step):
        r.push(x[s:s + step])
        pull_all()
    r.drain()
    pull_all()
    dt = r.dtype if native else np.float32
    r.close()
    y = np.concatenate(outs, axis=0) if outs else np.zeros((0, nch), dtype=dt)
    return y, counts


class BatchConverter:
    """Device-resident batch: nstreams x nchannels lanes of equal length, one shot. Pointers are raw device
    addresses (ints): pass ``tensor.data_ptr()`` of CUDA tensors, or numpy ``ctypes.data`` under emulation."""

    def __init__(self, cfg, nchannels, nstreams, frames_in_max, engine="float", device=-1, lib=None):
        self.lib = lib or _capi.product()
        self.nch, self.nstreams = int(nchannels), int(nstreams)
        self.sample_bytes = 4 if engine == "float" else 8
        self.b = C.c_void_p()
        rc = self.lib.RRX_batch_open(C.byref(cfg), self.sample_bytes, self.nch, self.nstreams, int(frames_in_max),
                                     int(device), C.byref(self.b))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_open")

    def plan(self):
        p = Plan()
        self.lib.RRX_batch_plan(self.b, C.byref(p))
        return p.as_dict()

    def frames_out(self, frames_in):
        return int(self.lib.RRX_batch_frames_out(self.b, int(frames_in)))

    def process(self, d_in, frames_in, d_out, stream=0):
        rc = self.lib.RRX_batch_process(self.b, d_in, int(frames_in), d_out, stream)
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_process")

    def process_native(self, d_in, frames_in, d_out, stream=0):
        rc = self.lib.RRX_batch_process_native(self.b, d_in, int(frames_in), d_out, stream)
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_process_native")

    def input_window(self, frames_in_total, out_begin, out_count):
        f, c = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.RRX_batch_input_window(self.b, int(frames_in_total), int(out_begin), int(out_count),
                                             C.byref(f), C.byref(c))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_input_window")
        return f.value, c.value

    def process_range(self, d_in_window, window_first, window_frames, frames_in_total, out_begin, out_count,
                      d_out, stream=0):
        rc = self.lib.RRX_batch_process_range(self.b, d_in_window, int(window_first), int(window_frames),
                                              int(frames_in_total), int(out_begin), int(out_count), d_out, stream)
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_process_range")

    def process_host(self, h_in, frames_in, h_out, total_streams):
        """h_in / h_out: host addresses (ints) of float32 [total_streams][frames][nch] buffers."""
        rc = self.lib.RRX_batch_process_host(self.b, h_in, int(frames_in), h_out, int(total_streams))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_process_host")

    def enable_timing(self, on=True):
        self.lib.RRX_batch_enable_timing(self.b, 1 if on else 0)

    def stage_times(self):
        ms = (C.c_float * 24)()
        n = self.lib.RRX_batch_stage_times(self.b, ms, 24)
        return [float(ms[i]) for i in range(n)]

    def stage_kernel(self, stage):
        return (self.lib.RRX_batch_stage_kernel(self.b, int(stage)) or b"").decode()

    def stage_work(self, frames_in, stage):
        f, b, u = C.c_double(0), C.c_double(0), C.c_double(0)
        rc = self.lib.RRX_batch_stage_work(self.b, int(frames_in), int(stage), C.byref(f), C.byref(b), C.byref(u))
        if rc != RR_OK:
            raise RateError(self.lib, rc, "RRX_batch_stage_work")
        return {"flops": f.value, "bytes": b.value, "units": u.value}

    def last_launches(self):
        return int(self.lib.RRX_batch_last_launches(self.b))

    def flops(self, frames_in):
        return float(self.lib.RRX_batch_flops(self.b, int(frames_in)))

    def close(self):
        if self.b is not None and self.b.value:
            self.lib.RRX_batch_close(C.byref(self.b))
        self.b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


LPC_ORDER = 32  # lpc/lpc.h:24


def track_edge_lengths(in_rate, out_rate, lib=None):
    """(add, drop, prime_len, inbuf_frames) as dsp_rate::reinit computes them (foo_dsp_rate.cpp:96-101)."""
    lib = lib or _capi.product()
    v = [C.c_uint(0) for _ in range(4)]
    rc = lib.RRX_track_edge_lengths(int(in_rate), int(out_rate), *[C.byref(x) for x in v])
    if rc != RR_OK:
        raise RateError(lib, rc, "RRX_track_edge_lengths")
    return tuple(int(x.value) for x in v)


def lpc_extrapolate2(buf, first, data_len, extra_bkwd, extra_fwd, lpc_order=LPC_ORDER, lib=None):
    """lpc_extrapolate2 (lpc/lpc.h:26) on a HOST array: ``buf`` is float32 [frames][nch], C-contiguous; the base segment
    is frames [first, first + data_len); frames [first - extra_bkwd, first) and [first + data_len, ... + extra_fwd) are
    written in place."""
    lib = lib or _capi.product()
    assert buf.dtype == np.float32 and buf.flags.c_contiguous and buf.ndim == 2
    assert first >= extra_bkwd and first + data_len + extra_fwd <= buf.shape[0]
    nch = buf.shape[1]
    rc = lib.RRX_lpc_extrapolate2(buf.ctypes.data + first * nch * 4, int(data_len), nch, int(lpc_order),
                                  int(extra_bkwd), int(extra_fwd))
    if rc != RR_OK:
        raise RateError(lib, rc, "RRX_lpc_extrapolate2")


class TrackBatchConverter:
    """Whole tracks, device resident, converted the way the plugin converts a track (foo_dsp_rate.cpp:130-313): `add`
    frames predicted before the beginning and after the end (RRX_lpc_extend_tracks), everything pushed through the
    rate engine (RRX_batch_process), `drop` output frames cut from both ends. Tracks of at most 2 * LPC_ORDER frames
    are converted without extrapolation, like the plugin (foo_dsp_rate.cpp:222-237).

    d_padded: float32 [nstreams][add + track_frames + add][nch] with the tracks in the middle (``padded_frames`` /
    ``track_offset`` say where); d_work: float32 [nstreams][frames_out_padded][nch] scratch the engine writes; the
    track's result is frames [drop, drop + frames_out) of every stream of d_work (``result_slice``)."""

    def __init__(self, cfg, nchannels, nstreams, track_frames, engine="float", device=-1, lib=None):
        self.lib = lib or _capi.product()
        self.nch, self.nstreams, self.track_frames = int(nchannels), int(nstreams), int(track_frames)
        self.add, self.drop, self.prime, _ = track_edge_lengths(cfg.in_rate, cfg.out_rate, self.lib)
        if self.track_frames <= 2 * LPC_ORDER:
            self.add = self.drop = 0
        self.padded_frames = self.track_frames + 2 * self.add
        self.track_offset = self.add
        self.batch = BatchConverter(cfg, nchannels, nstreams, self.padded_frames, engine=engine, device=device, lib=self.lib)
        self.frames_out_padded = self.batch.frames_out(self.padded_frames)
        self.frames_out = self.frames_out_padded - 2 * self.drop
        self.result_slice = slice(self.drop, self.drop + self.frames_out)

    def process(self, d_padded, d_work, stream=0):
        if self.add:
            rc = self.lib.RRX_lpc_extend_tracks(d_padded, self.nstreams, self.track_frames, self.prime, self.nch, LPC_ORDER,
                                                self.add, stream)
            if rc != RR_OK:
                raise RateError(self.lib, rc, "RRX_lpc_extend_tracks")
        self.batch.process(d_padded, self.padded_frames, d_work, stream)

    def close(self):
        self.batch.close()
