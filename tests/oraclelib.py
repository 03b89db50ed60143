"""ctypes driver for the plain-C restatement ``oracle/liboracle_rate.so`` (oracle/rate_oracle.h).

TEST INFRASTRUCTURE: the checker for the CUDA path. Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from reflib import Plan, RRConfig, make_config  # noqa: F401  (same struct layouts)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_rate.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        L = C.CDLL(ORACLE_SO)
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.POINTER(RRConfig), C.c_int, C.c_int]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_keep_history.argtypes = [C.c_void_p, C.c_int]
        L.orc_push.restype = C.c_size_t
        L.orc_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_pull.restype = C.c_size_t
        L.orc_pull.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_pull_native.restype = C.c_size_t
        L.orc_pull_native.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_drain.argtypes = [C.c_void_p]
        L.orc_plan_dump.argtypes = [C.c_void_p, C.POINTER(Plan)]
        L.orc_dft_coefs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_poly_coefs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dft_taps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_fifo_written.restype = C.c_uint64
        L.orc_fifo_written.argtypes = [C.c_void_p, C.c_int]
        L.orc_fifo_consumed.restype = C.c_uint64
        L.orc_fifo_consumed.argtypes = [C.c_void_p, C.c_int]
        L.orc_fifo_read.restype = C.c_size_t
        L.orc_fifo_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_size_t, C.c_void_p]
        L.orc_rdft_f32.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_rdft_f64.argtypes = [C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class OracleResampler:
    def __init__(self, cfg, nch, engine="float", keep_history=False):
        self.L = lib()
        self.nch = nch
        self.dtype = np.float32 if engine == "float" else np.float64
        self.h = C.c_void_p(self.L.orc_open(C.byref(cfg), nch, 4 if engine == "float" else 8))
        if not self.h.value:
            raise ValueError("orc_open rejected the configuration")
        if keep_history:
            self.L.orc_keep_history(self.h, 1)

    def plan(self):
        p = Plan()
        self.L.orc_plan_dump(self.h, C.byref(p))
        return p.as_dict()

    def push(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return self.L.orc_push(self.h, x.ctypes.data, x.shape[0])

    def pull(self, osamp):
        out = np.empty((osamp, self.nch), dtype=np.float32)
        n = self.L.orc_pull(self.h, out.ctypes.data, osamp)
        return out[:n]

    def pull_native(self, osamp):
        out = np.empty((self.nch, osamp), dtype=self.dtype)
        n = self.L.orc_pull_native(self.h, out.ctypes.data, osamp)
        return out[:, :n]

    def drain(self):
        self.L.orc_drain(self.h)

    def dft_coefs(self, instance):
        n = self.L.orc_dft_coefs(self.h, instance, None, 0)
        out = np.empty(n, dtype=self.dtype)
        if n:
            self.L.orc_dft_coefs(self.h, instance, out.ctypes.data, n)
        return out

    def poly_coefs(self):
        n = self.L.orc_poly_coefs(self.h, None, 0)
        out = np.empty(n, dtype=self.dtype)
        if n:
            self.L.orc_poly_coefs(self.h, out.ctypes.data, n)
        return out

    def dft_taps(self, instance):
        n = self.L.orc_dft_taps(self.h, instance, None, 0)
        out = np.empty(n, dtype=np.float64)
        if n:
            self.L.orc_dft_taps(self.h, instance, out.ctypes.data, n)
        return out

    def fifo_written(self, i):
        return self.L.orc_fifo_written(self.h, i)

    def fifo_consumed(self, i):
        return self.L.orc_fifo_consumed(self.h, i)

    def fifo_read(self, ch, i, start, count):
        out = np.empty(count, dtype=self.dtype)
        n = self.L.orc_fifo_read(self.h, ch, i, start, count, out.ctypes.data)
        return out[:n]

    def close(self):
        if self.h is not None and self.h.value:
            self.L.orc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resample(cfg, x, engine="float", chunk=65536, native=False, pull_chunk=None):
    """Same driving pattern and return values as reflib.resample."""
    nch = x.shape[1]
    r = OracleResampler(cfg, nch, engine)
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
    r.close()
    dt = r.dtype if native else np.float32
    y = np.concatenate(outs, axis=0) if outs else np.zeros((0, nch), dtype=dt)
    return y, counts
