"""CPU tier: the PRODUCT library (CUDA build) loads without a GPU, exports every symbol that
include/b200_ratelib.h declares, answers its host-only entry points, and refuses -- loudly, with an error
code, never by falling back to a CPU path -- to create an engine when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

from foo_dsp_resampler_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200_ratelib.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:RR|RRX)_\w+|init_ratelib|close_ratelib)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert set(names) == set(_capi.SYMBOLS), set(names) ^ set(_capi.SYMBOLS)


def test_product_library_exports_every_declared_symbol():
    assert os.path.exists(_capi.PRODUCT_SO), "libb200rate.so missing: run __graft_entry__.build()"
    lib = C.CDLL(_capi.PRODUCT_SO)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_host_only_entry_points():
    lib = _capi.product()
    assert b"sm_100a" in lib.RRX_version()
    assert lib.RR_strerror(_capi.RR_NULLHANDLE) == b"NULL handle"
    cfg = _capi.make_config(44100, 48000)
    p = _capi.Plan()
    assert lib.RRX_plan(C.byref(cfg), 4, C.byref(p)) == _capi.RR_OK
    d = p.as_dict()
    assert [s["kind"] for s in d["stages"]] == [1, 2]
    assert (d["stages"][0]["dft_length"], d["stages"][0]["num_taps"], d["stages"][1]["L"], d["stages"][1]["n"]) == \
        (4096, 553, 80, 24)
    bad = _capi.make_config(1, 48000)
    assert lib.RRX_plan(C.byref(bad), 4, C.byref(p)) == _capi.RR_INVPARAM


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = _capi.product()
    cfg = _capi.make_config(44100, 48000)
    h = C.c_void_p()
    rc = lib.RR_open(C.byref(cfg), 2, C.byref(h))
    assert rc in (_capi.RR_INTERNAL, _capi.RR_ENOMEM) and not h.value
    assert lib.RRX_last_error()                      # the CUDA error string is reported
    assert not lib.RR_ctor_float(C.byref(cfg), 2)
    b = C.c_void_p()
    assert lib.RRX_batch_open(C.byref(cfg), 4, 2, 4, 1000, -1, C.byref(b)) != _capi.RR_OK and not b.value
    # RR_open before init_ratelib (rate/rate_uni.c:35-36)
    lib.close_ratelib()
    assert lib.RR_open(C.byref(cfg), 2, C.byref(h)) == _capi.RR_EXTUNINIT
    assert lib.init_ratelib(_capi._oom) == 0
    assert lib.init_ratelib(C.cast(None, _capi.OOM_FN)) == -1
    assert lib.init_ratelib(_capi._oom) == 0
