"""CPU tier: the product's host orchestration and CTA programs, compiled as plain C++ (tests/emu, see
tests/emulib.py), against the oracle. This is what can be checked of the CUDA path without a GPU: stage plans,
the absolute-coordinate bookkeeping of push/pull/drain, every index computation inside the kernels, the batch
front-end and time-chunked ranges. fp32 bit-exact; fp64 within the 1e-12 contract (the engine's fp64 transform
is the split-radix DAG, not Ooura's)."""
import ctypes as C

import numpy as np
import pytest

import emulib
import oraclelib
import signals
from foo_dsp_resampler_b200 import _capi, converter

FP64_TOL = 1e-12

CASES = [
    (44100, 48000, "float", 50, 95, 0, 0, 2), (44100, 96000, "float", 50, 95, 0, 0, 2),
    (192000, 44100, "double", 25, 95, 0, 0, 3), (48000, 44100, "float", 50, 95, 0, 0, 2),
    (384000, 48000, "float", 50, 95, 0, 0, 3), (44100, 48001, "float", 50, 95, 0, 0, 1),
    (44100, 44101, "float", 50, 95, 0, 1, 1), (48000, 47999, "double", 50, 95, 0, 0, 1),
    (8000, 48000, "float", 50, 95, 0, 0, 1), (48000, 32000, "float", 50, 95, 0, 0, 2),
    (96000, 44100, "float", 75, 99, 1, 0, 1), (44100, 48000, "float", 0, 90, 0, 1, 1),
    (22050, 96000, "float", 50, 95, 0, 0, 1), (44100, 11025, "double", 50, 95, 0, 0, 2),
    (44100, 176400, "float", 50, 95, 0, 1, 1), (48000, 8000, "double", 50, 95, 0, 1, 1),
    # even channel counts on the fp32 engine: the lane-pair DFT kernel (rate_kernels_pk.cuh) in all its
    # spectrum modes, input-tile modes (interleaved / planar / zero-stuffed) and output modes
    (96000, 44100, "float", 50, 95, 0, 0, 2), (32000, 48000, "float", 50, 95, 0, 0, 2),
    (384000, 48000, "float", 50, 95, 0, 0, 4), (8000, 48000, "float", 50, 95, 0, 0, 2),
    (44100, 176400, "float", 50, 95, 0, 1, 2), (48000, 8000, "float", 50, 95, 0, 1, 2),
    (44100, 88200, "float", 25, 95, 0, 0, 6), (96000, 48000, "float", 50, 95, 1, 0, 2),
    (44100, 22050, "float", 50, 95, 0, 1, 2), (50000, 40000, "float", 50, 95, 0, 0, 2),
    (48000, 44100, "float", 50, 95, 0, 0, 8), (44100, 48000, "float", 50, 95, 0, 1, 6),
    # narrow transition bands: DFT blocks beyond the shared-memory kernels (work buffers in global scratch)
    (44100, 48000, "float", 50, 99.5, 0, 0, 2), (44100, 48000, "double", 50, 99, 0, 0, 1),
    (192000, 44100, "double", 50, 99, 0, 0, 1),
    # F-domain / 4 (step -2) behind zero-stuffing by 3, the same pair with time-domain / 4, h9 / h8 half-bands,
    # identity (no stages)
    (32000, 24000, "float", 50, 95, 0, 0, 2), (32000, 24000, "float", 50, 95, 1, 0, 2), (32000, 24000, "double", 50, 95, 0, 0, 1),
    (44100, 8000, "float", 50, 95, 0, 1, 2), (32000, 8000, "float", 50, 95, 0, 1, 1), (48000, 48000, "float", 50, 95, 0, 0, 2),
    (48000, 48000, "double", 50, 95, 0, 0, 3),
    # Best quality above 96 % bandwidth: 28-tap polyphase banks (lane-pair, two-slot and generic kernels), N = 8192 blocks
    (44100, 48000, "float", 50, 97, 0, 0, 2), (44100, 96000, "float", 50, 98, 0, 0, 2), (44100, 48000, "double", 50, 97, 0, 0, 2),
    (44100, 48000, "float", 50, 97, 0, 0, 1),
    # fp64 engine, F-domain up-sampling by 4 and 8 (dft64_kernel's radix-L split in the spectrum phase)
    (44100, 176400, "double", 50, 95, 0, 0, 2), (44100, 192000, "double", 25, 95, 0, 0, 1), (8000, 384000, "double", 50, 95, 0, 0, 1),
    # N = 131072, the reference's table limit (rate_uni.c:134-189): bandwidth 99.9 %
    (44100, 48000, "float", 50, 99.9, 0, 0, 2), (96000, 48000, "double", 50, 99.9, 0, 0, 1),
]


def ids(c):
    return "%d-%d-%s-p%d-q%d-%dch" % (c[0], c[1], c[2], c[3], c[6], c[7])


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_stream_front_end(case):
    i, o, eng, ph, bw, al, q, nch = case
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, int(i * (3.3 if bw > 99.8 else 0.3)) + 11)      # N = 131072 blocks consume ~1.2 s each
    yo, co = oraclelib.resample(ocfg, x, engine=eng, chunk=3001, native=True)
    ye, ce = converter.resample(cfg, x, engine=eng, chunk=3001, native=True, lib=L)
    assert ce == co
    assert ye.shape == yo.shape
    if eng == "float":
        assert np.array_equal(ye, yo)
    else:
        assert np.abs(ye - yo).max() <= FP64_TOL


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_plan_and_design(case):
    i, o, eng, ph, bw, al, q, nch = case
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q)
    p = _capi.Plan()
    assert L.RRX_plan(C.byref(cfg), 4 if eng == "float" else 8, C.byref(p)) == 0
    orc = oraclelib.OracleResampler(ocfg, 1, eng)
    assert p.as_dict() == orc.plan()
    # designed banks, before conversion to the engine type: bit-identical doubles
    n = L.RRX_design_dump(C.byref(cfg), 4 if eng == "float" else 8, 2, None, 0)
    if n > 0:
        bank = np.empty(n)
        L.RRX_design_dump(C.byref(cfg), 4 if eng == "float" else 8, 2, bank.ctypes.data, n)
        assert np.array_equal(bank.astype(orc.dtype), orc.poly_coefs())
    if eng == "float":                       # engine-type spectrum of the DFT filters
        r = converter.RateConverter(cfg, 1, eng, lib=L)
        for inst in (0, 1):
            assert np.array_equal(r.dft_spectrum(inst), orc.dft_coefs(inst))
        r.close()
    orc.close()


@pytest.mark.parametrize("case", CASES[:8] + CASES[16:20] + CASES[26:28] + CASES[31:33] + CASES[34:35] + CASES[36:37], ids=ids)
def test_batch_front_end_and_ranges(case):
    i, o, eng, ph, bw, al, q, nch = case
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q)
    nstreams, n = 3, int(i * 0.25) + 7
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(nstreams)])
    b = converter.BatchConverter(cfg, nch, nstreams, n, engine=eng, lib=L)
    nout = b.frames_out(n)
    out = np.zeros((nstreams, nout, nch), np.float32)
    b.process(xs.ctypes.data, n, out.ctypes.data)
    for s in range(nstreams):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine=eng)
        assert ref.shape[0] == nout
        if eng == "float":
            assert np.array_equal(out[s], ref)
        else:
            assert np.abs(out[s] - ref).max() <= 2.0 ** -23
    # time-chunked: ranges computed from their halo'd input windows equal the one-shot result bit for bit
    for ob, oc in ((0, nout // 3), (nout // 3, nout // 2), (nout - 100, 100)):
        f, c = b.input_window(n, ob, oc)
        win = np.ascontiguousarray(xs[:, f:f + c, :])
        part = np.zeros((nstreams, oc, nch), np.float32)
        b.process_range(win.ctypes.data, f, c, n, ob, oc, part.ctypes.data)
        assert np.array_equal(part, out[:, ob:ob + oc, :])
    # host-buffer entry point, more streams than the batch holds
    more = np.concatenate([xs, xs[:2] * 0.5])
    out2 = np.zeros((5, nout, nch), np.float32)
    b.process_host(more.ctypes.data, n, out2.ctypes.data, 5)
    assert np.array_equal(out2[:3], out)
    b.close()


def test_push_pull_protocol_edges():
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(48000, 44100), oraclelib.make_config(48000, 44100)
    x = signals.sweep_noise(48000, 2, 20000)
    ref, _ = oraclelib.resample(ocfg, x, engine="float")
    # RR_flow: pull, push, pull again (rate/rate_base.h:571-614)
    r = converter.RateConverter(cfg, 2, "float", lib=L)
    outs = []
    for s in range(0, x.shape[0], 2500):
        y, used = r.flow(x[s:s + 2500], 3000)
        assert used == min(2500, x.shape[0] - s)
        outs.append(y.copy())
    r.drain()
    while True:
        y = r.pull(4096)
        if not len(y):
            break
        outs.append(y.copy())
    assert np.array_equal(np.concatenate(outs), ref)
    r.drain()                                   # second drain: nothing left
    assert len(r.pull(16)) == 0
    r.close()
    # pushing after a drain continues the stream exactly like the reference does
    orc = oraclelib.OracleResampler(ocfg, 2, "float")
    r = converter.RateConverter(cfg, 2, "float", lib=L)
    got_o, got_e = [], []
    for blk in (x[:7000], x[7000:9000]):
        for obj, got in ((orc, got_o), (r, got_e)):
            obj.push(blk)
            obj.drain()
            got.append(obj.pull(1 << 16).copy())
    assert [len(a) for a in got_o] == [len(a) for a in got_e]
    assert all(np.array_equal(a, b) for a, b in zip(got_o, got_e))
    r.close()
    orc.close()
    # RR_open: Best -> fp64 engine, Normal -> fp32 engine (rate/rate_uni.c:38-51)
    assert converter.RateConverter(cfg, 1, "auto", lib=L).sample_bytes == 8
    assert converter.RateConverter(_capi.make_config(48000, 44100, quality=1), 1, "auto", lib=L).sample_bytes == 4
    # invalid ratio: the reference swallows RR_INVPARAM in its ctor; this library reports it
    h = C.c_void_p()
    bad = _capi.make_config(1, 48000)
    assert L.RR_open(C.byref(bad), 1, C.byref(h)) == _capi.RR_INVPARAM and not h.value
    assert L.RR_push(None, None, 0) == _capi.RR_NULLHANDLE


# Any two lanes can share the lane-pair kernels, also lanes of different streams: mono and odd-channel batches
# with an even number of lanes.
@pytest.mark.parametrize("case", [(48000, 44100, 1, 4, 0), (384000, 48000, 3, 2, 0), (96000, 44100, 1, 2, 0), (44100, 96000, 5, 2, 1),
                                  (44100, 48000, 1, 6, 0), (50000, 44100, 2, 1, 0), (44100, 50000, 2, 2, 0)], ids=lambda c: "%d-%d-%dch-x%d-q%d" % c)
def test_batch_pairs_across_streams(case):
    i, o, nch, nstreams, q = case
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(i, o, 50, 95, 0, q), oraclelib.make_config(i, o, 50, 95, 0, q)
    n = int(i * 0.2) + 5
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(nstreams)])
    b = converter.BatchConverter(cfg, nch, nstreams, n, engine="float", lib=L)
    nout = b.frames_out(n)
    out = np.zeros((nstreams, nout, nch), np.float32)
    b.process(xs.ctypes.data, n, out.ctypes.data)
    for s in range(nstreams):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
        assert ref.shape[0] == nout and np.array_equal(out[s], ref)
    ob, oc = nout // 3, nout // 2
    f, c = b.input_window(n, ob, oc)
    win = np.ascontiguousarray(xs[:, f:f + c, :])
    part = np.zeros((nstreams, oc, nch), np.float32)
    b.process_range(win.ctypes.data, f, c, n, ob, oc, part.ctypes.data)
    assert np.array_equal(part, out[:, ob:ob + oc, :])
    b.close()


FUZZ_RATES = [8000, 11025, 12000, 16000, 22050, 24000, 32000, 37800, 44100, 48000, 50000, 64000, 88200, 96000, 176400, 192000,
              352800, 384000, 47999, 44101]


def fuzz_cases(seed, count):
    """Seeded random configurations with an even channel count (lane-pair kernels): rates, phase, bandwidth,
    aliasing, quality, channels, length, chunk size."""
    import random
    rng = random.Random(seed)
    out = []
    while len(out) < count:
        i, o = rng.choice(FUZZ_RATES), rng.choice(FUZZ_RATES)
        if i == o:
            continue
        out.append((i, o, rng.choice([0, 25, 50, 50, 50, 75, 100]), rng.choice([90, 93, 95, 95, 97, 99]), rng.choice([0, 0, 1]),
                    rng.choice([0, 0, 1]), rng.choice([2, 2, 4, 6, 8]),
                    max(int(i * rng.choice([0.05, 0.11, 0.23])) + rng.randrange(0, 50), 64), rng.choice([977, 4096, 30011])))
    return out


@pytest.mark.parametrize("case", fuzz_cases(7, 24), ids=lambda c: "%d-%d-p%d-b%d-a%d-q%d-%dch" % c[:7])
def test_fuzz_stream_and_batch(case):
    i, o, ph, bw, al, q, nch, n, chunk = case
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, n)
    yo, co = oraclelib.resample(ocfg, x, engine="float", chunk=chunk, native=True)
    ye, ce = converter.resample(cfg, x, engine="float", chunk=chunk, native=True, lib=L)
    assert ce == co and ye.shape == yo.shape and np.array_equal(ye, yo)
    b = converter.BatchConverter(cfg, nch, 2, n, engine="float", lib=L)
    nout = b.frames_out(n)
    xs = np.stack([x, x * 0.5])
    out = np.zeros((2, nout, nch), np.float32)
    b.process(xs.ctypes.data, n, out.ctypes.data)
    ref, _ = oraclelib.resample(ocfg, x, engine="float")
    assert ref.shape[0] == nout and np.array_equal(out[0], ref)
    b.close()


def _golden():
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_cases.json")) as f:
        cases = json.load(f)
    return cases, np.load(os.path.join(here, "golden", "reference_outputs.npz"))


@pytest.mark.parametrize("name", sorted(_golden()[0]))
def test_emulated_path_reproduces_reference_fixture(name):
    """The product's orchestration and CTA programs (host emulation) against the committed outputs of the compiled
    reference itself (tests/golden): plan integers, frame counts after every push, samples -- fp32 bit for bit,
    fp64 within 1e-12 of the reference's Ooura-based engine."""
    cases, outputs = _golden()
    g = cases[name]
    i, o, eng, ph, bw, al, q, nch, frames, chunk = g["case"]
    L = emulib.lib()
    cfg = _capi.make_config(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, frames)
    r = converter.RateConverter(cfg, nch, eng, lib=L)
    assert r.plan() == g["plan"]
    r.close()
    y, counts = converter.resample(cfg, x, engine=eng, chunk=chunk, native=True, lib=L)
    assert counts == g["counts"] and y.shape[0] == g["out_frames"]
    if eng == "float":
        assert np.array_equal(y, outputs[name])
    else:
        assert np.abs(y - outputs[name]).max() <= FP64_TOL


# (in_rate, out_rate, channels, hours, input frames per grid period, output frames per grid period)
# one period of the joint block / phase grid: 384 -> 48 kHz: 8 x 1766 input frames per DFT block; 48 -> 44.1 kHz:
# lcm(1748 input frames per DFT block, 160 input frames per polyphase period) = 69920 input = 64239 output frames
LONG_STREAMS = [(384000, 48000, 8, 10, 14128, 1766), (48000, 44100, 2, 80, 69920, 64239)]


def long_offset_check(make_batch, to_dev, from_dev, case):
    """Time-chunk ranges at the absolute offsets of a many-hour stream (input frame indices beyond 2^32): a window
    at ~90 % of the stream must give the bits of the same window content placed a whole number of grid periods
    earlier (shift invariance of the block grid, dft_filter.h:78-83), and that small-offset placement is checked
    against the oracle run from the start of the stream."""
    i, o, nch, hours, pin, pout = case
    total = i * 3600 * hours
    cfg, ocfg = _capi.make_config(i, o), oraclelib.make_config(i, o)
    oc = 2 * 1766 + 123
    b = make_batch(cfg, nch, 6 * pin // (pin // 14128 if pin > 20000 else 1) + 131072)
    nout_total = b.frames_out(total)
    assert nout_total == total // i * o
    ob_far = int(0.9 * nout_total) // pout * pout + 57
    f_far, c = b.input_window(total, ob_far, oc)
    k = ob_far // pout - 3
    ob_near, f_near = ob_far - k * pout, f_far - k * pin
    assert f_far > (1 << 32) and b.input_window(total, ob_near, oc) == (f_near, c)
    x = signals.sweep_noise(i, nch, c, stream=3)
    dx = to_dev(x)
    res = []
    for ob, f in ((ob_far, f_far), (ob_near, f_near)):
        out = to_dev(np.zeros((1, oc, nch), np.float32))
        b.process_range(dx[1], f, c, total, ob, oc, out[1])
        res.append(from_dev(out))
    assert np.array_equal(res[0], res[1]) and np.abs(res[0]).max() > 0.1
    full = np.concatenate([np.zeros((f_near, nch), np.float32), x, np.zeros((4 * pin // (pin // 14128 if pin > 20000 else 1), nch), np.float32)])
    ref, _ = oraclelib.resample(ocfg, full, engine="float")
    assert np.array_equal(ref[ob_near:ob_near + oc], res[1][0])
    # the very last frames of the stream (drain rule round(n_in / factor), rate_base.h:457)
    tail = 500
    f, c = b.input_window(total, nout_total - tail, tail)
    assert f + c == total
    b.close()


@pytest.mark.parametrize("case", LONG_STREAMS, ids=lambda c: "%d-%d-%dh" % (c[0], c[1], c[3]))
def test_ranges_at_many_hour_offsets(case):
    L = emulib.lib()
    long_offset_check(lambda cfg, nch, fmax: converter.BatchConverter(cfg, nch, 1, fmax, engine="float", lib=L),
                      lambda a: (a, a.ctypes.data), lambda d: d[0].copy(), case)


def test_identity_conversion_passes_frames_through():
    """in_rate == out_rate: the plan has no stages and FIFO 0 is the output FIFO (rate_base.h:445-447)."""
    L = emulib.lib()
    cfg, ocfg = _capi.make_config(48000, 48000), oraclelib.make_config(48000, 48000)
    x = signals.sweep_noise(48000, 2, 5000)
    for eng in ("float", "double"):
        yo, co = oraclelib.resample(ocfg, x, engine=eng, chunk=1777)
        ye, ce = converter.resample(cfg, x, engine=eng, chunk=1777, lib=L)
        assert ce == co and np.array_equal(ye, yo) and np.array_equal(ye, x)
    b = converter.BatchConverter(cfg, 2, 2, 5000, engine="float", lib=L)
    xs = np.stack([x, x * 0.5])
    out = np.zeros((2, 5000, 2), np.float32)
    b.process(xs.ctypes.data, 5000, out.ctypes.data)
    assert np.array_equal(out, xs)
    assert b.input_window(5000, 100, 50) == (100, 50)
    b.close()


FUSED_CASES = [(48000, 44100, 3, 2.0), (44100, 48000, 2, 1.5), (44100, 96000, 1, 1.2), (48000, 44100, 1, 0.05), (44100, 48000, 1, 0.3)]


def fused_check(make_batch, to_dev, from_dev, case):
    """DFT stage + vpoly0 as one kernel (rate_kernels_fused.cuh): whole streams and time-chunk ranges, bit for bit
    against the oracle; the runs of blocks a lane pair is cut into must not show (each run but the first recomputes
    one block for its filter history)."""
    i, o, ns, secs = case
    nch = 2
    cfg, ocfg = _capi.make_config(i, o), oraclelib.make_config(i, o)
    n = (int(i * secs) + 13) & ~1                      # even: every stream starts on a 16-byte boundary
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(ns)])
    b = make_batch(cfg, nch, ns, n)
    nout = b.frames_out(n)
    dx, dy = to_dev(xs), to_dev(np.zeros((ns, nout, nch), np.float32))
    b.process(dx[1], n, dy[1])
    out = from_dev(dy)
    assert b.stage_kernel(0).startswith("dft_poly_kernel") and b.last_launches() == 1
    for s in range(ns):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
        assert ref.shape[0] == nout and np.array_equal(out[s], ref), "stream %d" % s
    for ob, oc in ((0, nout // 3), (nout // 3 + 1, nout // 2), (nout - 100, 100)):
        f, c = b.input_window(n, ob, oc)
        win = to_dev(np.ascontiguousarray(xs[:, f:f + c, :]))
        part = to_dev(np.zeros((ns, oc, nch), np.float32))
        b.process_range(win[1], f, c, n, ob, oc, part[1])
        assert np.array_equal(from_dev(part), out[:, ob:ob + oc, :])
    b.close()


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: "%d-%d-x%d" % c[:3])
def test_fused_dft_poly_kernel(case, monkeypatch):
    monkeypatch.setenv("B200RATE_FUSE_MIN_PAIRS", "1")
    monkeypatch.setenv("B200RATE_FUSED", "1")
    L = emulib.lib()
    fused_check(lambda cfg, nch, ns, n: converter.BatchConverter(cfg, nch, ns, n, engine="float", lib=L),
                lambda a: (a, a.ctypes.data), lambda d: d[0].copy(), case)


def multi_check(L, devices, batch_case, stream_case, to_dev=None):
    """RRX_multi_*: a batch sharded by stream and one long stream cut into time chunks over `devices`, host buffers in and
    out, against the oracle; device-resident slices gathered on one device."""
    ndev = len(devices)
    devs = (C.c_int * ndev)(*devices)
    i, o, nch, ns, n = batch_case
    cfg, ocfg = _capi.make_config(i, o), oraclelib.make_config(i, o)
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(ns)])
    m = C.c_void_p()
    assert L.RRX_multi_open(C.byref(cfg), 4, nch, ns, n, devs, ndev, C.byref(m)) == 0, L.RRX_last_error()
    assert L.RRX_multi_devices(m) == ndev
    nout = L.RRX_multi_frames_out(m, n)
    out = np.zeros((ns, nout, nch), np.float32)
    assert L.RRX_multi_process_host(m, xs.ctypes.data, n, out.ctypes.data) == 0, L.RRX_last_error()
    for s in range(ns):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
        assert ref.shape[0] == nout and np.array_equal(out[s], ref), s
    covered = 0
    for k in range(ndev):
        d, f, c = C.c_int(), C.c_size_t(), C.c_size_t()
        assert L.RRX_multi_shard(m, k, C.byref(d), C.byref(f), C.byref(c)) == 0
        assert d.value == devices[k] and f.value == covered
        covered += c.value
    assert covered == ns
    if to_dev is not None:                               # device-resident slices, results gathered on the last device
        keep, ptrs = [], (C.c_void_p * ndev)()
        for k in range(ndev):
            d, f, c = C.c_int(), C.c_size_t(), C.c_size_t()
            L.RRX_multi_shard(m, k, C.byref(d), C.byref(f), C.byref(c))
            t = to_dev(np.ascontiguousarray(xs[f.value:f.value + c.value]), devices[k])
            keep.append(t)
            ptrs[k] = t[1]
        assert L.RRX_multi_process(m, ptrs, n) == 0, L.RRX_last_error()
        g = to_dev(np.zeros_like(out), devices[-1])
        assert L.RRX_multi_gather(m, ndev - 1, g[1]) == 0, L.RRX_last_error()
        assert np.array_equal(g[2](), out)
    L.RRX_multi_close(C.byref(m))
    assert not m.value
    i, o, nch, n, fmax = stream_case
    cfg, ocfg = _capi.make_config(i, o), oraclelib.make_config(i, o)
    x = signals.sweep_noise(i, nch, n)
    assert L.RRX_multi_open(C.byref(cfg), 4, nch, 1, fmax, devs, ndev, C.byref(m)) == 0, L.RRX_last_error()
    ref, _ = oraclelib.resample(ocfg, x, engine="float")
    out = np.zeros((ref.shape[0] + 8, nch), np.float32)
    fo = C.c_size_t()
    assert L.RRX_multi_process_stream_host(m, x.ctypes.data, n, out.ctypes.data, C.byref(fo)) == 0, L.RRX_last_error()
    assert fo.value == ref.shape[0] and np.array_equal(out[:fo.value], ref)      # seams between chunks and devices included
    L.RRX_multi_close(C.byref(m))


@pytest.mark.parametrize("ndev", [1, 3, 4])
def test_multi_device_layer(ndev):
    multi_check(emulib.lib(), [0] * ndev, (48000, 44100, 2, 7, 9000), (384000, 48000, 4, 384000 // 2, 60000),
                to_dev=lambda a, dev: (a, a.ctypes.data, lambda: a))
