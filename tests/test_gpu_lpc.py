"""GPU parity for the track-edge extrapolation kernels (csrc/lpc.cu), through the C ABI, on
identical bytes: bit for bit against the restatement (oracle/lpc_oracle.c), against the golden vectors generated from
the compiled reference (tests/golden/lpc_reference.npz), and -- where the prebuilt checker travelled -- against the
compiled reference lpc/lpc.cpp itself. The lags and predictor coefficients (doubles) are compared bit for bit too."""
import ctypes as C
import os

import numpy as np
import pytest

import lpclib
import oraclelib
import pluginsim
import foo_dsp_resampler_b200 as pkg
from foo_dsp_resampler_b200 import _capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpc_reference.npz")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def _base(kind, shape, seed=1):
    n, nch, order, bk, fw = shape
    buf = np.zeros((bk + n + fw, nch), np.float32)
    buf[bk:bk + n] = lpclib.signal(kind, n, nch, seed)
    return buf


@pytest.mark.parametrize("kind", lpclib.KINDS)
def test_host_entry_point_equals_oracle_reference_and_golden(kind):
    g = np.load(GOLD)
    for si, shape in enumerate(lpclib.SHAPES):
        n, nch, order, bk, fw = shape
        want = _base(kind, shape)
        got = want.copy()
        lpclib.oracle_extrapolate2(want, bk, n, bk, fw, order)
        pkg.lpc_extrapolate2(got, bk, n, bk, fw, order)
        assert np.array_equal(_bits(got), _bits(want)), (kind, shape, float(np.abs(got - want).max()))
        if "k%d_s%d_bkwd" % (kind, si) in g:
            assert np.array_equal(_bits(got[:bk]), _bits(g["k%d_s%d_bkwd" % (kind, si)]))
            assert np.array_equal(_bits(got[bk + n:]), _bits(g["k%d_s%d_fwd" % (kind, si)]))
        if lpclib.ref_available():
            ref = _base(kind, shape)
            lpclib.ref_extrapolate2(ref, bk, n, bk, fw, order)
            assert np.array_equal(_bits(got), _bits(ref))


def test_inline_wrappers_of_lpc_h():
    lib = _capi.product()
    n, nch, prime, extra = 4410, 2, 2205, 2205
    x = lpclib.signal(0, n, nch, seed=3)
    want = np.zeros((extra + n + extra, nch), np.float32)
    want[extra:extra + n] = x
    got = want.copy()
    lpclib.oracle_extrapolate2(want, extra, prime, extra, 0)                    # lpc_extrapolate_bkwd
    lpclib.oracle_extrapolate2(want, extra + n - prime, prime, 0, extra)        # lpc_extrapolate_fwd
    p0 = got.ctypes.data + extra * nch * 4
    assert lib.RRX_lpc_extrapolate_bkwd(p0, n, prime, nch, 32, extra) == _capi.RR_OK
    assert lib.RRX_lpc_extrapolate_fwd(p0, n, prime, nch, 32, extra) == _capi.RR_OK
    assert np.array_equal(_bits(got), _bits(want))
    assert lib.RRX_lpc_extrapolate_fwd(p0, 100, 200, nch, 32, extra) == _capi.RR_INVPARAM
    assert lib.RRX_lpc_extrapolate2(p0, n, nch, 0, extra, extra) == _capi.RR_INVPARAM
    assert lib.RRX_lpc_extrapolate2(p0, n, nch, 32, 0, 0) == _capi.RR_OK        # nothing to do


def test_lags_and_predictor_bit_for_bit_thread_per_job_kernel():
    """Enough lanes for lpc_analyse_wide_kernel (one thread per lane, delay line in registers)."""
    import torch
    lib = _capi.product()
    nstreams, nch = 6400, 3
    for n, order in ((333, 32), (64, 32), (1000, 7)):
        x = np.stack([lpclib.signal(s % 7, n, nch, seed=40 + s % 13) for s in range(64)])
        x = np.ascontiguousarray(np.tile(x, (nstreams // 64, 1, 1)))
        x[64:] *= np.float32(0.75)
        d = torch.from_numpy(x).cuda()
        out = torch.zeros((nstreams * nch, 66), dtype=torch.float64, device="cuda")
        assert lib.RRX_lpc_analysis_dump(d.data_ptr(), nstreams, n, n, nch, order, out.data_ptr(), None) == _capi.RR_OK
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        for s in list(range(0, 70)) + [nstreams - 1]:
            for c in range(nch):
                r, a, used = lpclib.oracle_analyse(x[s], c, order)
                row = o[s * nch + c]
                assert np.array_equal(_bits(row[:order + 1]), _bits(r)), (n, s, c)
                assert np.array_equal(_bits(row[33:33 + order]), _bits(a)), (n, s, c)
                assert not row[33 + order:65].any() and int(row[65]) == used


def test_lags_and_predictor_bit_for_bit():
    import torch
    lib = _capi.product()
    for kind in lpclib.KINDS:
        for n, nch, order in ((2205, 2, 32), (16384, 1, 32), (1000, 3, 12), (70, 2, 32), (1024, 1, 32), (1056, 2, 32)):
            x = lpclib.signal(kind, n, nch, seed=5)
            d = torch.from_numpy(x).cuda()
            out = torch.zeros((nch, 66), dtype=torch.float64, device="cuda")
            assert lib.RRX_lpc_analysis_dump(d.data_ptr(), 1, n, n, nch, order, out.data_ptr(), None) == _capi.RR_OK
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            for c in range(nch):
                r, a, used = lpclib.oracle_analyse(x, c, order)
                assert np.array_equal(_bits(o[c, :order + 1]), _bits(r)), (kind, n, c)
                assert np.array_equal(_bits(o[c, 33:33 + order]), _bits(a)), (kind, n, c)
                assert not o[c, 33 + order:65].any() and int(o[c, 65]) == used


@pytest.mark.parametrize("nstreams,nch,n,bk,fw", [(5, 2, 2400, 2400, 2400), (67, 1, 1024, 100, 333), (3, 6, 3000, 0, 500),
                                                   (40, 2, 2205, 2205, 0),
                                                   # >= 6144 (lane, job) slots: the thread-per-job analysis kernel
                                                   (9600, 2, 200, 33, 70), (19001, 1, 97, 40, 0)])
def test_device_batch_of_streams(nstreams, nch, n, bk, fw):
    import torch
    lib = _capi.product()
    stride = bk + n + fw + 7                                   # streams need not be packed
    host = np.zeros((nstreams, stride, nch), np.float32)
    for s in range(nstreams):
        host[s, bk:bk + n] = lpclib.signal(s % 7, n, nch, seed=11 + s % 50)
        if s >= 350:
            host[s] *= np.float32(1.0 - 0.4 * s / nstreams)
    want = host.copy()
    for s in range(nstreams):
        lpclib.oracle_extrapolate2(want[s], bk, n, bk, fw)
    d = torch.from_numpy(host).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        rc = lib.RRX_lpc_extrapolate_batch(d.data_ptr() + bk * nch * 4, nstreams, stride, n, nch, 32, bk, fw,
                                           st.cuda_stream)
    assert rc == _capi.RR_OK
    st.synchronize()
    got = d.cpu().numpy()
    assert np.array_equal(_bits(got), _bits(want))
    assert not got[:, bk + n + fw:].any()                      # nothing written past the requested frames


@pytest.mark.parametrize("in_rate,out_rate,nch,frames,nstreams", [(44100, 48000, 2, 30000, 4), (48000, 44100, 2, 4800, 3),
                                                                  (44100, 48000, 1, 1500, 2), (96000, 44100, 2, 20000, 2)])
def test_track_batch_equals_the_plugin_simulation(in_rate, out_rate, nch, frames, nstreams):
    """Device-resident whole-track conversion (RRX_lpc_extend_tracks + RRX_batch_process + cut) == the plugin's
    streamed handling of the same track (tests/pluginsim.py over the oracle), bit for bit."""
    import torch
    cfg = pkg.make_config(in_rate, out_rate)
    ocfg = oraclelib.make_config(in_rate, out_rate)
    edge = lpclib.oracle_edge_lengths(in_rate, out_rate)
    assert pkg.track_edge_lengths(in_rate, out_rate) == edge
    t = pkg.TrackBatchConverter(cfg, nch, nstreams, frames, engine="float", device=0)
    assert abs(t.frames_out - frames * out_rate / in_rate) <= 0.5
    tracks = [lpclib.signal(s % 2, frames, nch, seed=21 + s) for s in range(nstreams)]
    padded = np.zeros((nstreams, t.padded_frames, nch), np.float32)
    for s in range(nstreams):
        padded[s, t.track_offset:t.track_offset + frames] = tracks[s]
    d_in = torch.from_numpy(padded).cuda()
    d_work = torch.zeros((nstreams, t.frames_out_padded, nch), dtype=torch.float32, device="cuda")
    t.process(d_in.data_ptr(), d_work.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_work[:, t.result_slice].cpu().numpy()
    for s in range(nstreams):
        chunks = [tracks[s][i:i + 4096] for i in range(0, frames, 4096)]
        want = pluginsim.convert_track(lambda: oraclelib.OracleResampler(ocfg, nch, "float"), lpclib.oracle_extrapolate2,
                                       edge, chunks, nch)
        assert got[s].shape == want.shape and np.array_equal(_bits(got[s]), _bits(want)), s
    t.close()


def test_plugin_simulation_over_the_product_streaming_api():
    """The same driver over the product: RR_push / RR_pull / RR_drain handles plus RRX_lpc_extrapolate2 on host
    buffers -- the calls a maintainer's dsp_rate would make -- equal the oracle-driven run."""
    cfg, ocfg, nch, frames = pkg.make_config(44100, 48000), oraclelib.make_config(44100, 48000), 2, 20000
    edge = pkg.track_edge_lengths(44100, 48000)
    x = lpclib.signal(0, frames, nch, seed=31)
    chunks = [x[i:i + 3000] for i in range(0, frames, 3000)]
    want = pluginsim.convert_track(lambda: oraclelib.OracleResampler(ocfg, nch, "float"), lpclib.oracle_extrapolate2, edge,
                                   chunks, nch)
    got = pluginsim.convert_track(lambda: pkg.RateConverter(cfg, nch, engine="float"), pkg.lpc_extrapolate2, edge, chunks, nch)
    assert got.shape == want.shape and np.array_equal(_bits(got), _bits(want))
