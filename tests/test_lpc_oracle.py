"""CPU tier for the track-edge extrapolation: the restatement oracle/lpc_oracle.c against the
compiled reference (lpc/lpc.cpp, util.h) where it is present, and against the golden vectors generated from it
(tests/golden/lpc_reference.npz); the product's host-only edge-length entry point; the plugin track simulation's
bookkeeping. Everything bit for bit."""
import os

import numpy as np
import pytest

import lpclib
import oraclelib
import pluginsim
from foo_dsp_resampler_b200 import _capi, track_edge_lengths

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpc_reference.npz")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _extend(fn, kind, shape):
    n, nch, order, bk, fw = shape
    buf = np.zeros((bk + n + fw, nch), np.float32)
    buf[bk:bk + n] = lpclib.signal(kind, n, nch)
    fn(buf, bk, n, bk, fw, order)
    return buf[:bk], buf[bk + n:]


@pytest.mark.parametrize("kind", lpclib.KINDS)
def test_oracle_equals_golden_reference_vectors(kind):
    g = np.load(GOLD)
    seen = 0
    for si, shape in enumerate(lpclib.SHAPES):
        if "k%d_s%d_bkwd" % (kind, si) not in g:
            continue
        b, f = _extend(lpclib.oracle_extrapolate2, kind, shape)
        assert np.array_equal(_bits(b), _bits(g["k%d_s%d_bkwd" % (kind, si)])), (kind, shape)
        assert np.array_equal(_bits(f), _bits(g["k%d_s%d_fwd" % (kind, si)])), (kind, shape)
        seen += 1
    assert seen >= 7


@pytest.mark.skipif(not lpclib.ref_available(), reason="compiled reference not built")
@pytest.mark.parametrize("kind", lpclib.KINDS)
def test_oracle_equals_compiled_reference(kind):
    for shape in lpclib.SHAPES:
        rb, rf = _extend(lpclib.ref_extrapolate2, kind, shape)
        ob, of = _extend(lpclib.oracle_extrapolate2, kind, shape)
        assert np.array_equal(_bits(rb), _bits(ob)) and np.array_equal(_bits(rf), _bits(of)), (kind, shape)
    # seeds the fixture does not hold
    for seed in range(2, 6):
        n, nch = 1500 + 37 * seed, 1 + seed % 3
        x = lpclib.signal(kind, n, nch, seed)
        a = np.zeros((300 + n + 300, nch), np.float32)
        a[300:300 + n] = x
        b = a.copy()
        lpclib.ref_extrapolate2(a, 300, n, 300, 300)
        lpclib.oracle_extrapolate2(b, 300, n, 300, 300)
        assert np.array_equal(_bits(a), _bits(b))


def test_branches_of_the_reference_are_reached():
    # silence: order 0 -> fixed up to order 1 with lpc[0] = -1 (lpc.cpp:159-163); constant: continues the constant
    _, a, used = lpclib.oracle_analyse(lpclib.signal(2, 500, 1), 0)
    assert used == 1 and a[0] == -1 and not a[1:].any()
    r, a, used = lpclib.oracle_analyse(lpclib.signal(4, 2205, 1), 0)
    assert 1 <= used <= 32 and r[0] > 0
    b, f = _extend(lpclib.oracle_extrapolate2, 6, lpclib.SHAPES[0])
    assert np.abs(np.concatenate([b, f])).max() == 10.0          # the clamp acts on kind 6
    b, f = _extend(lpclib.oracle_extrapolate2, 2, lpclib.SHAPES[0])
    assert not b.any() and not f.any()


def test_edge_lengths_product_oracle_reference():
    g = np.load(GOLD)["edge_lengths"]
    assert len(g) == 17 * 17
    for a, b, add, drop, prime, inbuf in g.tolist():
        assert lpclib.oracle_edge_lengths(a, b) == (add, drop, prime, inbuf)
        assert track_edge_lengths(a, b) == (add, drop, prime, inbuf)          # host-only product entry point
        if lpclib.ref_available():
            assert lpclib.ref_edge_lengths(a, b) == (add, drop, prime, inbuf)
    assert track_edge_lengths(44100, 48000) == (2205, 2400, 2205, 4410)       # the example in util.h:38
    lib = _capi.product()
    assert lib.RRX_track_edge_lengths(0, 48000, None, None, None, None) == _capi.RR_INVPARAM


def test_lpc_entry_points_fail_loudly_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = _capi.product()
    buf = np.zeros((300, 2), np.float32)
    assert lib.RRX_lpc_extrapolate2(buf.ctypes.data + 100 * 8, 100, 2, 32, 100, 100) in (_capi.RR_INTERNAL, _capi.RR_ENOMEM)
    assert lib.RRX_last_error()
    assert lib.RRX_lpc_extrapolate2(buf.ctypes.data + 100 * 8, 100, 2, 33, 100, 100) == _capi.RR_INVPARAM
    assert lib.RRX_lpc_extrapolate2(None, 100, 2, 32, 100, 100) == _capi.RR_INVPARAM


@pytest.mark.parametrize("frames,chunk", [(30000, 1000), (30000, 7919), (4410, 4410), (3000, 512), (50, 50), (4411, 100)])
def test_plugin_track_simulation_is_chunk_invariant_and_whole_track_equivalent(frames, chunk):
    """The plugin's streamed edge handling equals: extend the whole track by `add` predicted frames on both sides,
    convert, cut `drop` frames from both ends -- the form the device batch path (TrackBatchConverter) implements."""
    cfg = oraclelib.make_config(44100, 48000)
    nch = 2
    edge = lpclib.oracle_edge_lengths(44100, 48000)
    add, drop, prime_len, inbuf = edge
    x = lpclib.signal(0, frames, nch, seed=9)
    chunks = [x[i:i + chunk] for i in range(0, frames, chunk)]
    y = pluginsim.convert_track(lambda: oraclelib.OracleResampler(cfg, nch, "float"), lpclib.oracle_extrapolate2, edge,
                                chunks, nch)
    if frames <= 2 * 32:
        whole, _ = oraclelib.resample(cfg, x)
    else:
        prime = min(frames, prime_len)
        buf = np.zeros((add + frames + add, nch), np.float32)
        buf[add:add + frames] = x
        lpclib.oracle_extrapolate2(buf, add, prime, add, 0)
        lpclib.oracle_extrapolate2(buf, add + frames - prime, prime, 0, add)
        whole, _ = oraclelib.resample(cfg, buf)
        whole = whole[drop:whole.shape[0] - drop]
    assert y.shape == whole.shape and np.array_equal(_bits(y), _bits(whole))
    assert y.shape[0] == int(round(frames * 48000 / 44100))
