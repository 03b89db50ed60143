"""ctypes drivers for the two LPC checkers: the plain-C restatement (oracle/liboracle_rate.so, lpc_oracle.c) and the
compiled reference (oracle/_ref/libref_lpc.so = the unmodified lpc/lpc.cpp + util.h, see oracle/Makefile).

TEST INFRASTRUCTURE. Never imported by the product package."""
import ctypes as C
import os

import numpy as np

import oraclelib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LPC_SO = os.path.join(ROOT, "oracle", "_ref", "libref_lpc.so")
LPC_ORDER = 32

_ref = None
_EX2 = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t]
_EDGE = [C.POINTER(C.c_uint)] * 4


def ref_available():
    return os.path.exists(REF_LPC_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_LPC_SO)
        L.ref_lpc_extrapolate2.argtypes = _EX2
        L.ref_lpc_extrapolate2.restype = None
        for n in ("ref_lpc_extrapolate_bkwd", "ref_lpc_extrapolate_fwd"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t]
            getattr(L, n).restype = None
        L.ref_track_edge_lengths.argtypes = [C.c_uint, C.c_uint] + _EDGE
        L.ref_track_edge_lengths.restype = None
        assert L.ref_lpc_order() == LPC_ORDER
        _ref = L
    return _ref


def oracle():
    L = oraclelib.lib()
    if not getattr(L, "_lpc_bound", False):
        L.orc_lpc_extrapolate2.argtypes = _EX2
        L.orc_lpc_extrapolate2.restype = None
        L.orc_lpc_analyse.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_lpc_analyse.restype = C.c_int
        L.orc_track_edge_lengths.argtypes = [C.c_uint, C.c_uint, C.c_int] + _EDGE
        L.orc_track_edge_lengths.restype = None
        L._lpc_bound = True
    return L


def _run(fn, buf, first, data_len, extra_bkwd, extra_fwd, order):
    assert buf.dtype == np.float32 and buf.flags.c_contiguous
    assert first >= extra_bkwd and first + data_len + extra_fwd <= buf.shape[0]
    fn(buf.ctypes.data + first * buf.shape[1] * 4, data_len, buf.shape[1], order, extra_bkwd, extra_fwd)


def oracle_extrapolate2(buf, first, data_len, extra_bkwd, extra_fwd, order=LPC_ORDER):
    """In place on float32 [frames][nch]: base = frames [first, first + data_len)."""
    _run(oracle().orc_lpc_extrapolate2, buf, first, data_len, extra_bkwd, extra_fwd, order)


def ref_extrapolate2(buf, first, data_len, extra_bkwd, extra_fwd, order=LPC_ORDER):
    _run(ref().ref_lpc_extrapolate2, buf, first, data_len, extra_bkwd, extra_fwd, order)


def oracle_analyse(base, ch, order=LPC_ORDER):
    """(lags[order+1], lpc[order], usable order) of channel ch of float32 [frames][nch]."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    r = np.zeros(order + 1)
    a = np.zeros(order)
    used = oracle().orc_lpc_analyse(base.ctypes.data, base.shape[0], base.shape[1], ch, order, r.ctypes.data, a.ctypes.data)
    return r, a, used


def oracle_edge_lengths(in_rate, out_rate):
    v = [C.c_uint(0) for _ in range(4)]
    oracle().orc_track_edge_lengths(in_rate, out_rate, LPC_ORDER, *[C.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def ref_edge_lengths(in_rate, out_rate):
    v = [C.c_uint(0) for _ in range(4)]
    ref().ref_track_edge_lengths(in_rate, out_rate, *[C.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def signal(kind, frames, nch, seed=1):
    """Deterministic base segments that exercise the branches of lpc/lpc.cpp: 0 tonal + noise, 1 white noise,
    2 silence (error < epsilon at once: order 0 -> the `max_order == 0` fix-up), 3 constant, 4 one pure tone
    (near-singular: early exit of the recursion), 5 random walk (large values: the +-10 clamp can act)."""
    rng = np.random.default_rng(seed * 1000 + kind)
    t = np.arange(frames)[:, None]
    if kind == 0:
        x = 0.5 * np.sin(2 * np.pi * (0.01 + 0.003 * np.arange(nch)) * t) + 0.05 * rng.uniform(-1, 1, (frames, nch))
    elif kind == 1:
        x = rng.uniform(-1, 1, (frames, nch))
    elif kind == 2:
        x = np.zeros((frames, nch))
    elif kind == 3:
        x = np.full((frames, nch), 0.25)
    elif kind == 4:
        x = 0.9 * np.sin(2 * np.pi * 0.05 * t) * np.ones((1, nch))
    elif kind == 5:
        x = np.cumsum(rng.uniform(-1, 1, (frames, nch)), 0) * 0.05
    else:
        x = 40.0 * rng.standard_normal((frames, nch)) * np.exp(np.arange(frames)[:, None] / frames * 3.0)  # clamp acts
    return np.ascontiguousarray(x, dtype=np.float32)


# (frames, nch, lpc_order, extra_bkwd, extra_fwd): plugin sizes (2205 / 2400 / 16384 prime, 8192 extra), short, odd
SHAPES = ((1024, 2, 32, 500, 700), (2205, 1, 32, 2205, 0), (2400, 2, 32, 0, 2400), (16384, 2, 32, 8192, 8192),
          (65, 3, 32, 100, 100), (300, 2, 8, 64, 64), (4410, 6, 32, 0, 2000), (1025, 1, 31, 33, 1), (3000, 2, 1, 40, 40))
KINDS = tuple(range(7))
