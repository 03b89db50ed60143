"""GPU parity tests proper: the CUDA path, called through the C ABI (include/b200_ratelib.h), against the
oracle on identical bytes. Bit-exact for the fp32 engine, counts and plan integers; <= 1e-12 for the fp64
engine at the un-cast tap (BASELINE.json north_star)."""
import numpy as np
import pytest

import oraclelib
import signals

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-12          # north_star: double path max |err| <= 1e-12
FP32_TOL = 2.0 ** -22     # north_star: float path max |err| <= 2^-22 full scale (we require 0)

# (in_rate, out_rate, engine, phase, bandwidth, allow_aliasing, quality, nch)
STREAM_CASES = [
    (44100, 48000, "float", 50, 95, 0, 0, 2),     # BASELINE config 1
    (44100, 96000, "float", 50, 95, 0, 0, 2),     # config 2
    (192000, 44100, "double", 25, 95, 0, 0, 8),   # config 3
    (48000, 44100, "float", 50, 95, 0, 0, 2),     # config 4 (one stream)
    (384000, 48000, "float", 50, 95, 0, 0, 8),    # config 5 (short)
    (44100, 48001, "float", 50, 95, 0, 0, 1),     # vpoly2
    (44100, 44101, "float", 50, 95, 0, 1, 1),     # vpoly1
    (48000, 47999, "double", 50, 95, 0, 0, 1),    # vpoly3
    (8000, 48000, "float", 50, 95, 0, 0, 1),      # zero-stuff L=3 + F-domain /2, post stage L=4
    (48000, 32000, "float", 50, 95, 0, 0, 1),     # L=2, time-domain decimation by 3
    (96000, 44100, "float", 75, 99, 1, 0, 1),     # phase 75, aliasing allowed
    (44100, 48000, "float", 0, 90, 0, 1, 1),      # minimum phase, Normal quality
    (22050, 96000, "float", 50, 95, 0, 0, 1),     # pre + arb + post stages
    (44100, 11025, "double", 50, 95, 0, 0, 2),    # half-band + F-domain /2, fp64
    # even channel counts on the fp32 engine run the lane-pair kernels (rate_kernels_pk.cuh): every spectrum
    # mode (x2 up / same size / generic), tile mode (interleaved, planar, zero-stuffed) and output mode
    (96000, 44100, "float", 50, 95, 0, 0, 2),     # half-band -> 1:1 DFT (same-size mode, planar tiles) -> vpoly0
    (32000, 48000, "float", 50, 95, 0, 0, 2),     # zero-stuffed tiles (L = 3)
    (384000, 48000, "float", 50, 95, 0, 0, 4),    # F-domain /2 as last stage, interleaved sink with 4 channels
    (8000, 48000, "float", 50, 95, 0, 0, 2),      # two DFT stages, post stage L = 4 (generic spectrum mode)
    (44100, 176400, "float", 50, 95, 0, 1, 2),    # Normal quality, x4
    (48000, 8000, "float", 50, 95, 0, 1, 2),      # half-bands + time-domain decimation
    (44100, 88200, "float", 25, 95, 0, 0, 6),     # intermediate phase, 6 channels, DFT only
    (96000, 48000, "float", 50, 95, 1, 0, 2),     # aliasing allowed
    (50000, 40000, "float", 50, 95, 0, 0, 2),     # L = 4 / M = 5
    (48000, 44100, "float", 50, 95, 0, 0, 3),     # odd channel count: generic kernels
    # narrow transition bands (bandwidth 99 % is within the plugin's UI range): DFT blocks of 16384 ... 65536
    # points, beyond the shared-memory kernels -> dft_big_kernel (work buffers in global scratch)
    (44100, 48000, "float", 50, 99.5, 0, 0, 2),   # N = 32768, fp32
    (44100, 48000, "double", 50, 99, 0, 0, 2),    # N = 16384, fp64 (what RR_open selects for Best quality)
    (192000, 44100, "double", 50, 99, 0, 0, 1),   # N = 32768 after a half-band stage
    (48000, 44100, "float", 50, 99.7, 0, 0, 1),   # N = 65536
    # stage modes outside the BASELINE plans: F-domain / 4 (step -2) behind zero-stuffing by 3, the same pair with
    # time-domain / 4, h8 / h9 / h11 half-bands, no stages at all
    (32000, 24000, "float", 50, 95, 0, 0, 2),     # L = 3, step -2 (dft_filter.h:157-188), lane-pair kernel
    (32000, 24000, "float", 50, 95, 0, 0, 1),     # ... generic kernel
    (32000, 24000, "double", 50, 95, 0, 0, 2),    # ... fp64
    (32000, 24000, "float", 50, 95, 1, 0, 2),     # L = 3, time-domain / 4
    (44100, 8000, "float", 50, 95, 0, 1, 2),      # Normal quality: h9 + DFT + vpoly0 (n = 16)
    (44100, 8000, "float", 50, 95, 0, 1, 1),
    (32000, 8000, "float", 50, 95, 0, 1, 2),      # h8
    (32000, 8000, "double", 50, 95, 0, 0, 1),     # h11, fp64
    (48000, 48000, "float", 50, 95, 0, 0, 2),     # identity
    (48000, 48000, "double", 50, 95, 0, 0, 3),
    (44100, 48000, "float", 50, 99.9, 0, 0, 2),   # N = 131072: the reference's FFT table limit (rate_uni.c:134-189)
    (96000, 48000, "double", 50, 99.9, 0, 0, 1),
]

# batches (device-resident entry point: pair-interleaved intermediate FIFOs between DFT and polyphase stages)
BATCH_CASES = [
    (44100, 48000, 50, 95, 0, 0, 2, 3), (48000, 44100, 50, 95, 0, 0, 8, 2), (44100, 96000, 50, 95, 0, 0, 4, 2),
    (96000, 44100, 50, 95, 0, 0, 2, 3), (384000, 48000, 50, 95, 0, 0, 8, 1), (22050, 96000, 50, 95, 0, 0, 2, 2),
    (44100, 48000, 50, 95, 0, 1, 6, 1), (48000, 44100, 50, 95, 0, 0, 3, 2),
    # pairs of lanes from different streams: mono and odd-channel batches with an even number of lanes
    (48000, 44100, 50, 95, 0, 0, 1, 4), (384000, 48000, 50, 95, 0, 0, 3, 2), (96000, 44100, 50, 95, 0, 0, 1, 2),
    (44100, 96000, 50, 95, 0, 1, 5, 2), (44100, 48000, 50, 95, 0, 0, 1, 6),
    # polyphase banks with many phases (L = 441, 250): CTA sizes near the kernels' launch bounds
    (50000, 44100, 50, 95, 0, 0, 2, 2), (44100, 50000, 50, 95, 0, 0, 2, 2),
    # Best quality above 96 % bandwidth (inside the plugin's UI range): 28-tap polyphase banks, N = 8192 blocks
    (44100, 48000, 50, 97, 0, 0, 2, 3), (44100, 96000, 50, 98, 0, 0, 2, 2), (44100, 48000, 50, 97, 0, 0, 1, 3),
    # F-domain / 4, time-domain / 4, h9 and h8 through the batch entry point
    (32000, 24000, 50, 95, 0, 0, 2, 2), (32000, 24000, 50, 95, 1, 0, 4, 1), (44100, 8000, 50, 95, 0, 1, 2, 3),
    (32000, 8000, 50, 95, 0, 1, 1, 2),
]


def _cfgs(i, o, ph, bw, al, q):
    import foo_dsp_resampler_b200 as pkg
    return (pkg.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q))


@pytest.mark.parametrize("case", STREAM_CASES, ids=lambda c: "%d-%d-%s-p%d-q%d" % (c[0], c[1], c[2], c[3], c[6]))
def test_stream_matches_oracle(case):
    import foo_dsp_resampler_b200 as pkg
    i, o, eng, ph, bw, al, q, nch = case
    cfg, ocfg = _cfgs(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, int(i * (3.3 if bw > 99.8 else 0.6)) + 17)
    y_ref, c_ref = oraclelib.resample(ocfg, x, engine=eng, chunk=7001, native=True)
    y, c = pkg.resample(cfg, x, engine=eng, chunk=7001, native=True)
    assert c == c_ref                                   # frames available after every push / drain
    assert y.shape == y_ref.shape
    if eng == "float":
        assert np.array_equal(y, y_ref), "max diff %g" % np.abs(y - y_ref).max()
    else:
        assert np.abs(y - y_ref).max() <= FP64_TOL


def test_plan_matches_oracle():
    import foo_dsp_resampler_b200 as pkg
    for i, o, eng, ph, bw, al, q, nch in STREAM_CASES:
        cfg, ocfg = _cfgs(i, o, ph, bw, al, q)
        r = pkg.RateConverter(cfg, 1, eng)
        orc = oraclelib.OracleResampler(ocfg, 1, eng)
        assert r.plan() == orc.plan()
        if eng == "float":
            for inst in (0, 1):
                assert np.array_equal(r.dft_spectrum(inst), orc.dft_coefs(inst))
        r.close()
        orc.close()


def test_chunk_size_invariance_and_float_pull():
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(44100, 48000, 50, 95, 0, 0)
    x = signals.sweep_noise(44100, 2, 50000)
    ref, _ = oraclelib.resample(ocfg, x, engine="float", chunk=65536)
    for chunk in (65536, 4096, 1000, 37 * 13):
        y, _ = pkg.resample(cfg, x, engine="float", chunk=chunk)
        assert np.array_equal(y, ref)
    # RR_open(Best) selects the fp64 engine; through RR_pull both are float-rounded
    refd, cd = oraclelib.resample(ocfg, x, engine="double", chunk=10000)
    y, c = pkg.resample(cfg, x, engine="auto", chunk=10000)
    assert c == cd and np.abs(y - refd).max() <= 2.0 ** -23


def test_flow_and_edge_cases():
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(48000, 44100, 50, 95, 0, 0)
    x = signals.sweep_noise(48000, 2, 30000)
    ref, _ = oraclelib.resample(ocfg, x, engine="float", chunk=65536)
    r = pkg.RateConverter(cfg, 2, "float")
    outs = []
    for s in range(0, x.shape[0], 3000):
        y, used = r.flow(x[s:s + 3000], 4000)
        assert used == min(3000, x.shape[0] - s)
        outs.append(y.copy())
    r.drain()
    while True:
        y = r.pull(5000)
        if not len(y):
            break
        outs.append(y.copy())
    assert np.array_equal(np.concatenate(outs), ref)
    # empty pushes / pulls are no-ops; second drain is a no-op
    r.push(np.zeros((0, 2), np.float32))
    r.drain()
    assert len(r.pull(10)) == 0
    r.close()
    # tiny input: fewer frames than one DFT block, drain must still deliver round(n*out/in) frames
    r = pkg.RateConverter(cfg, 2, "float")
    r.push(x[:10])
    assert len(r.pull(100)) == 0
    r.drain()
    y = r.pull(100)
    ref10, _ = oraclelib.resample(ocfg, x[:10], engine="float")
    assert np.array_equal(y, ref10) and len(y) == 9
    r.close()


def test_batch_device_resident_and_ranges():
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(48000, 44100, 50, 95, 0, 0)
    nstreams, nch, n = 5, 2, 48000
    xs = np.stack([signals.sweep_noise(48000, nch, n, stream=s) for s in range(nstreams)])
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="float", device=0)
    nout = b.frames_out(n)
    d_in = torch.from_numpy(xs).cuda()
    d_out = torch.zeros((nstreams, nout, nch), dtype=torch.float32, device="cuda")
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for s in range(nstreams):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
        assert ref.shape[0] == nout and np.array_equal(got[s], ref)
    # time-chunked: every range, computed from its halo'd input window only, equals the one-shot result
    for ob, oc in ((0, 5000), (12345, 7777), (nout - 4000, 4000)):
        f, c = b.input_window(n, ob, oc)
        win = d_in[:, f:f + c, :].contiguous()
        part = torch.zeros((nstreams, oc, nch), dtype=torch.float32, device="cuda")
        b.process_range(win.data_ptr(), f, c, n, ob, oc, part.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(part.cpu().numpy(), got[:, ob:ob + oc, :])
    b.close()


@pytest.mark.parametrize("case", BATCH_CASES, ids=lambda c: "%d-%d-q%d-%dch-x%d" % (c[0], c[1], c[5], c[6], c[7]))
def test_batch_cases_match_oracle(case):
    import torch
    import foo_dsp_resampler_b200 as pkg
    i, o, ph, bw, al, q, nch, nstreams = case
    cfg, ocfg = _cfgs(i, o, ph, bw, al, q)
    n = int(i * 0.4) + 13
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(nstreams)])
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="float", device=0)
    nout = b.frames_out(n)
    d_in = torch.from_numpy(xs).cuda()
    d_out = torch.zeros((nstreams, nout, nch), dtype=torch.float32, device="cuda")
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for s in range(nstreams):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
        assert ref.shape[0] == nout
        assert np.array_equal(got[s], ref), "stream %d max diff %g" % (s, np.abs(got[s] - ref).max())
    # a range in the middle from its halo'd window only (time-chunk primitive)
    ob, oc = nout // 3, nout // 4
    f, c = b.input_window(n, ob, oc)
    win = d_in[:, f:f + c, :].contiguous()
    part = torch.zeros((nstreams, oc, nch), dtype=torch.float32, device="cuda")
    b.process_range(win.data_ptr(), f, c, n, ob, oc, part.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(part.cpu().numpy(), got[:, ob:ob + oc, :])
    b.close()


# fp64 engine through the batch entry point: every block shape of dft64_kernel (x2 F-domain up-sampling, 1:1 with plain /
# zero-stuffed input and time-domain decimation, F-domain /2 and /4), the eight-output half-band and the two-slot
# polyphase kernel. (in, out, phase, bw, aliasing, quality, channels, streams, DFT kernel expected)
FP64_BATCH_CASES = [
    (44100, 48000, 50, 95, 0, 0, 2, 5, "dft64"), (48000, 44100, 25, 95, 0, 0, 2, 3, "dft64"), (44100, 96000, 50, 95, 0, 1, 1, 7, "dft64"),
    (192000, 44100, 25, 95, 0, 0, 8, 2, "dft64"), (384000, 48000, 50, 95, 0, 0, 8, 1, "dft64"), (44100, 11025, 50, 95, 0, 0, 2, 3, "dft64"),
    (32000, 24000, 50, 95, 0, 0, 2, 2, "dft64"), (32000, 24000, 50, 95, 1, 0, 3, 1, "dft64"), (48000, 32000, 50, 95, 0, 0, 2, 2, "dft64"),
    (96000, 48000, 75, 95, 0, 0, 1, 4, "dft64"), (50000, 40000, 50, 95, 0, 0, 2, 2, "dft64"), (8000, 48000, 50, 95, 0, 0, 1, 3, "dft"),      # zero-stuffed x3 (dft64) + x4 post stage (generic)
    (44100, 48000, 50, 97, 0, 0, 2, 2, "dft64"),                # 28-tap polyphase bank, N = 8192
    (44100, 176400, 50, 95, 0, 0, 2, 2, "dft64"), (48000, 192000, 25, 95, 0, 1, 1, 3, "dft64"),   # x4 F-domain up-sampling
    (44100, 192000, 50, 95, 0, 0, 2, 2, "dft64"),               # x2, polyphase, x4 post stage
    (8000, 384000, 50, 95, 0, 0, 1, 2, "dft64"),                # x8 post stage
    (192000, 48000, 50, 95, 0, 0, 4, 2, "dft64"), (96000, 22050, 50, 95, 0, 1, 3, 2, "dft64"),   # half-band with 4 / 3 channels
]


@pytest.mark.parametrize("case", FP64_BATCH_CASES, ids=lambda c: "%d-%d-p%d-q%d-%dch-x%d" % (c[0], c[1], c[2], c[5], c[6], c[7]))
def test_fp64_batch_kernels_match_oracle(case):
    import torch
    import foo_dsp_resampler_b200 as pkg
    i, o, ph, bw, al, q, nch, nstreams, dft_kernel = case
    cfg, ocfg = _cfgs(i, o, ph, bw, al, q)
    n = int(i * 0.5) + 29
    xs = np.stack([signals.sweep_noise(i, nch, n, stream=s) for s in range(nstreams)])
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="double", device=0)
    nout = b.frames_out(n)
    d_in = torch.from_numpy(xs).cuda()
    d_nat = torch.zeros((nstreams, nch, nout), dtype=torch.float64, device="cuda")
    b.process_native(d_in.data_ptr(), n, d_nat.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_nat.cpu().numpy()
    plan = b.plan()
    kinds = [st["kind"] for st in plan["stages"]]
    names = [b.stage_kernel(k) for k in range(len(kinds))]
    for k, kind in enumerate(kinds):
        if kind == 1:                                   # RR_STAGE_DFT: no silent fall-back to the generic kernel
            assert dft_kernel in names[k], names
    for s in range(nstreams):
        ref, _ = oraclelib.resample(ocfg, xs[s], engine="double", native=True)
        assert ref.shape[0] == nout
        err = np.abs(got[s].T - ref).max()
        assert err <= FP64_TOL, "stream %d max |err| %g (%s)" % (s, err, names)
    # the same through RR_pull's float rounding (interleaved float output written by the last stage's stores)
    d_out = torch.zeros((nstreams, nout, nch), dtype=torch.float32, device="cuda")
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), np.swapaxes(got, 1, 2).astype(np.float32))
    b.close()


@pytest.mark.parametrize("rates", [(48000, 44100, 2), (44100, 96000, 2), (384000, 48000, 8)], ids=lambda r: "%d-%d-%dch" % r)
def test_many_identical_streams_agree(rates):
    """Stress for the persistent multi-group kernels (named barriers, warp-local FFT phases, LDGSTS pipelines):
    hundreds of work items in flight on every SM, all fed the same input -- every stream must come out
    bit-identical to the oracle, twice in a row. (compute-sanitizer is not available on this pool.)"""
    import torch
    import foo_dsp_resampler_b200 as pkg
    i, o, nch = rates
    cfg, ocfg = _cfgs(i, o, 50, 95, 0, 0)
    nstreams, n = (600 if nch == 2 else 150), int(i * 0.5)
    x = signals.sweep_noise(i, nch, n)
    ref, _ = oraclelib.resample(ocfg, x, engine="float")
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="float", device=0)
    nout = b.frames_out(n)
    d_in = torch.from_numpy(x).cuda().unsqueeze(0).repeat(nstreams, 1, 1).contiguous()
    d_ref = torch.from_numpy(ref).cuda()
    for _ in range(2):
        d_out = torch.full((nstreams, nout, nch), 7.0, dtype=torch.float32, device="cuda")
        b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert bool((d_out == d_ref.unsqueeze(0)).all())
    b.close()


@pytest.mark.parametrize("rates", [(44100, 48000, 2, 50), (192000, 44100, 8, 25), (384000, 48000, 8, 50), (32000, 24000, 2, 50)],
                         ids=lambda r: "%d-%d-%dch-p%d" % r)
def test_many_identical_streams_agree_fp64(rates):
    """The same stress for the fp64 kernels (dft64_kernel's groups working in place on bit-reversed spectra behind named
    barriers, the register-pipelined half-band, the TMA-staged two-slot polyphase stage): every one of hundreds of
    identical streams must equal the first one bit for bit, twice in a row, and the first one the oracle within 1e-12."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    i, o, nch, ph = rates
    cfg, ocfg = _cfgs(i, o, ph, 95, 0, 0)
    nstreams, n = (400 if nch == 2 else 100), int(i * 0.5)
    x = signals.sweep_noise(i, nch, n)
    ref, _ = oraclelib.resample(ocfg, x, engine="double", native=True)
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="double", device=0)
    nout = b.frames_out(n)
    d_in = torch.from_numpy(x).cuda().unsqueeze(0).repeat(nstreams, 1, 1).contiguous()
    for _ in range(2):
        d_nat = torch.full((nstreams, nch, nout), 7.0, dtype=torch.float64, device="cuda")
        b.process_native(d_in.data_ptr(), n, d_nat.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert bool((d_nat == d_nat[0:1]).all())
        assert np.abs(d_nat[0].cpu().numpy().T - ref).max() <= FP64_TOL
    b.close()


def _fuzz_cases():
    import test_emulation
    return test_emulation.fuzz_cases(11, 20) + test_emulation.fuzz_cases(13, 30)


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: "%d-%d-p%d-b%d-a%d-q%d-%dch" % c[:7])
def test_fuzz_batch_matches_oracle(case):
    """Seeded random configurations (rates, phase, bandwidth, aliasing, quality, even channel counts) through the
    device-resident batch entry point: bit-exact against the oracle."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    i, o, ph, bw, al, q, nch, n, chunk = case
    cfg, ocfg = _cfgs(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, n)
    ref, _ = oraclelib.resample(ocfg, x, engine="float")
    b = pkg.BatchConverter(cfg, nch, 3, n, engine="float", device=0)
    nout = b.frames_out(n)
    assert ref.shape[0] == nout
    d_in = torch.from_numpy(np.stack([x, x * 0.5, x])).cuda()
    d_out = torch.zeros((3, nout, nch), dtype=torch.float32, device="cuda")
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert np.array_equal(got[0], ref) and np.array_equal(got[2], ref), "max diff %g" % np.abs(got[0] - ref).max()
    b.close()
    y, _ = pkg.resample(cfg, x, engine="float", chunk=chunk)
    assert np.array_equal(y, ref)


def test_handles_on_concurrent_host_threads():
    """Distinct handles driven from distinct host threads (the reference's threading contract, rate_uni.c:210):
    each handle has its own CUDA stream; every thread's output equals the oracle."""
    import threading
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(44100, 48000, 50, 95, 0, 0)
    x = signals.sweep_noise(44100, 2, 44100 * 3)
    ref, _ = oraclelib.resample(ocfg, x, engine="float", chunk=8192)
    out = [None] * 6

    def work(k):
        out[k], _ = pkg.resample(cfg, x * (1.0 if k % 2 == 0 else 0.5), engine="float", chunk=8192 + 512 * k)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(out))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k, y in enumerate(out):
        assert y is not None and np.array_equal(y, ref if k % 2 == 0 else ref * 0.5)


def test_full_size_config1_properties():
    """BASELINE config 1 at full size (60 s stereo): frame count, bit-exactness vs the oracle, linearity of
    the whole pipeline in the scaling-by-two sense (exact in binary floating point)."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(44100, 48000, 50, 95, 0, 0)
    n = 44100 * 60
    x = signals.sweep_noise(44100, 2, n)
    b = pkg.BatchConverter(cfg, 2, 2, n, engine="float", device=0)
    nout = b.frames_out(n)
    assert nout == 2880000
    d_in = torch.from_numpy(np.stack([x, x * 2.0])).cuda()
    d_out = torch.zeros((2, nout, 2), dtype=torch.float32, device="cuda")
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    ref, _ = oraclelib.resample(ocfg, x, engine="float", chunk=65536)
    assert np.array_equal(got[0], ref)
    assert np.array_equal(got[1], got[0] * 2.0)
    b.close()


def test_host_buffer_pipeline_matches_device_resident():
    """RRX_batch_process_host: pinned host buffers, more streams than one sub-batch, H2D / kernels / D2H
    overlapped on three streams -- same bits as the device-resident call and as the oracle."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(44100, 48000, 50, 95, 0, 0)
    total, sub, nch, n = 11, 4, 2, 30000
    xs = np.stack([signals.sweep_noise(44100, nch, n, stream=s) for s in range(total)])
    b = pkg.BatchConverter(cfg, nch, sub, n, engine="float", device=0)
    nout = b.frames_out(n)
    h_in = torch.from_numpy(xs).pin_memory()
    h_out = torch.zeros((total, nout, nch), dtype=torch.float32).pin_memory()
    for _ in range(2):                              # second call reuses the staging slots and streams
        h_out.zero_()
        b.process_host(h_in.data_ptr(), n, h_out.data_ptr(), total)
        got = h_out.numpy()
        for s in (0, 3, 4, 10):
            ref, _ = oraclelib.resample(ocfg, xs[s], engine="float")
            assert np.array_equal(got[s], ref)
    assert b.last_launches() == 2 * 3               # two stage kernels for each of the three sub-batches
    b.close()


def test_stage_timing_and_work_accounting():
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg, _ = _cfgs(48000, 44100, 50, 95, 0, 0)
    n = 48000
    b = pkg.BatchConverter(cfg, 2, 4, n, engine="float", device=0)
    x = torch.rand((4, n, 2), device="cuda") - 0.5
    y = torch.zeros((4, b.frames_out(n), 2), device="cuda")
    b.enable_timing(True)
    b.process(x.data_ptr(), n, y.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ms = b.stage_times()
    assert len(ms) == 2 and all(m > 0 for m in ms)
    w0, w1 = b.stage_work(n, 0), b.stage_work(n, 1)
    # DFT stage: 1748 new input samples read and 3496 samples written per block and lane
    blocks = w0["units"] / 8
    assert abs(w0["bytes"] / w0["units"] - 4 * (n / blocks + 3496)) < 64
    assert w1["flops"] == 48.0 * b.frames_out(n) * 8
    b.close()


def _golden():
    import test_emulation
    return test_emulation._golden()


@pytest.mark.parametrize("name", sorted(_golden()[0]))
def test_cuda_path_reproduces_reference_fixture(name):
    """The CUDA path against the committed outputs of the COMPILED REFERENCE (tests/golden, generated from
    oracle/_ref by tests/golden/make_golden.py) -- no oracle restatement in between: plan integers, frame counts
    after every push, samples (fp32 bit for bit vs rate_float.c, fp64 <= 1e-12 vs rate_double.c). Includes the
    ten-stage h10 / h13 plans (rate ratios above 500), F-domain / 4 and the stage-less identity."""
    import foo_dsp_resampler_b200 as pkg
    cases, outputs = _golden()
    g = cases[name]
    i, o, eng, ph, bw, al, q, nch, frames, chunk = g["case"]
    cfg = pkg.make_config(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, frames)
    r = pkg.RateConverter(cfg, nch, eng)
    assert r.plan() == g["plan"]
    r.close()
    y, counts = pkg.resample(cfg, x, engine=eng, chunk=chunk, native=True)
    assert counts == g["counts"] and y.shape[0] == g["out_frames"]
    if eng == "float":
        assert np.array_equal(y, outputs[name]), "max diff %g" % np.abs(y - outputs[name]).max()
    else:
        assert np.abs(y - outputs[name]).max() <= FP64_TOL


def _long_streams():
    import test_emulation
    return test_emulation.LONG_STREAMS


@pytest.mark.parametrize("case", _long_streams(), ids=lambda c: "%d-%d-%dh" % (c[0], c[1], c[3]))
def test_ranges_at_many_hour_offsets(case):
    """BASELINE config 5's real coordinates: time-chunk ranges of a 10-hour 384 kHz stream (input frame indices
    ~1.2e10, beyond 2^32) and of an 80-hour 48 -> 44.1 kHz stream (polyphase phase accumulator at ~1e10 outputs)."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    import test_emulation
    st = torch.cuda.current_stream().cuda_stream

    class Dev:                               # BatchConverter whose process_range runs on torch's current stream
        def __init__(self, cfg, nch, fmax):
            self.b = pkg.BatchConverter(cfg, nch, 1, fmax, engine="float", device=0)
        def __getattr__(self, k):
            return getattr(self.b, k)
        def process_range(self, *a):
            self.b.process_range(*a, st)

    def to_dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return (t, t.data_ptr())

    def from_dev(d):
        torch.cuda.synchronize()
        return d[0].cpu().numpy()

    test_emulation.long_offset_check(Dev, to_dev, from_dev, case)


def test_identity_conversion_passes_frames_through():
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg = pkg.make_config(48000, 48000)
    x = signals.sweep_noise(48000, 2, 70000)
    y, counts = pkg.resample(cfg, x, engine="float", chunk=30011)
    assert np.array_equal(y, x) and sum(counts) == 70000
    y, _ = pkg.resample(cfg, x, engine="auto", chunk=4096)
    assert np.array_equal(y, x)
    b = pkg.BatchConverter(cfg, 2, 2, 70000, engine="float", device=0)
    d_in = torch.from_numpy(np.stack([x, x * 0.5])).cuda()
    d_out = torch.zeros_like(d_in)
    b.process(d_in.data_ptr(), 70000, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool((d_out == d_in).all())
    b.close()


def test_push_from_pinned_buffer_that_is_refilled_at_once():
    """A page-locked caller buffer is transferred without staging; RR_push must not return while the copy out of
    it is in flight (the reference has consumed ibuf when RR_push returns, rate_base.h:616-636)."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    cfg, ocfg = _cfgs(44100, 48000, 50, 95, 0, 0)
    n, chunk = 44100 * 4, 65536
    x = signals.sweep_noise(44100, 2, n)
    ref, _ = oraclelib.resample(ocfg, x, engine="float", chunk=chunk)
    buf = torch.empty((chunk, 2), dtype=torch.float32).pin_memory()
    out = torch.empty((1 << 17, 2), dtype=torch.float32).pin_memory()
    r = pkg.RateConverter(cfg, 2, "float")
    got = []
    import ctypes as C
    ogen = C.c_size_t(0)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        buf[:m] = torch.from_numpy(x[s:s + m])
        assert r.lib.RR_push(r.h, buf.data_ptr(), m) == 0
        buf.fill_(123.0)                          # the caller reuses its buffer right away
        while True:
            assert r.lib.RR_pull(r.h, out.data_ptr(), out.shape[0], C.byref(ogen)) == 0
            if not ogen.value:
                break
            got.append(out[:ogen.value].numpy().copy())
    r.drain()
    while True:
        y = r.pull(1 << 16)
        if not len(y):
            break
        got.append(y.copy())
    r.close()
    assert np.array_equal(np.concatenate(got), ref)


_COLD_START = r"""
import sys, threading
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import foo_dsp_resampler_b200 as pkg, oraclelib, signals
cases = [(44100, 48000, 2, 0), (48000, 44100, 2, 0), (384000, 48000, 8, 0), (44100, 96000, 2, 0), (96000, 44100, 2, 0),
         (44100, 48000, 1, 1), (32000, 24000, 2, 0), (44100, 48001, 1, 0)]
work = [(cases[k %% len(cases)], k) for k in range(32)]
refs = {}
for c in cases:
    i, o, nch, q = c
    x = signals.sweep_noise(i, nch, int(i * 0.35))
    refs[c] = (x, oraclelib.resample(oraclelib.make_config(i, o, 50, 95, 0, q), x, engine="float", chunk=9000)[0])
out, start = [None] * len(work), threading.Barrier(len(work))
def run(k):
    (i, o, nch, q), _ = work[k]
    start.wait()                                   # every thread's FIRST launch happens at the same time
    out[k] = pkg.resample(pkg.make_config(i, o, 50, 95, 0, q), refs[work[k][0]][0], engine="float", chunk=9000 + 64 * k)[0]
th = [threading.Thread(target=run, args=(k,)) for k in range(len(work))]
[t.start() for t in th]; [t.join() for t in th]
bad = [k for k in range(len(work)) if out[k] is None or not np.array_equal(out[k], refs[work[k][0]][1])]
print("COLD_START_BAD", bad)
sys.exit(1 if bad else 0)
"""


def test_cold_start_concurrency_mixed_plans():
    """32 host threads x 8 different plans in a FRESH process, all starting at once: concurrent first launches of
    the same and of different kernels, with different shared-memory sizes, through the process-wide launch cache."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, "-c", _COLD_START % (os.path.dirname(here), here)], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def _fused_cases():
    import test_emulation
    return test_emulation.FUSED_CASES + [(48000, 44100, 80, 3.0), (44100, 48000, 150, 1.0)]


@pytest.mark.parametrize("case", _fused_cases(), ids=lambda c: "%d-%d-x%d" % c[:3])
def test_fused_dft_poly_kernel(case, monkeypatch):
    """The fused DFT + vpoly0 kernel on small batches (forced), and on batches large enough to be cut into several
    runs per lane pair: bit-exact vs the oracle, ranges equal the one-shot result."""
    import torch
    import foo_dsp_resampler_b200 as pkg
    import test_emulation
    monkeypatch.setenv("B200RATE_FUSE_MIN_PAIRS", "1")
    monkeypatch.setenv("B200RATE_FUSED", "1")
    st = torch.cuda.current_stream().cuda_stream

    class Dev:
        def __init__(self, cfg, nch, ns, n):
            self.b = pkg.BatchConverter(cfg, nch, ns, n, engine="float", device=0)
        def __getattr__(self, k):
            return getattr(self.b, k)
        def process(self, a, n, o):
            self.b.process(a, n, o, st)
        def process_range(self, *a):
            self.b.process_range(*a, st)

    def to_dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return (t, t.data_ptr())

    def from_dev(d):
        torch.cuda.synchronize()
        return d[0].cpu().numpy()

    test_emulation.fused_check(Dev, to_dev, from_dev, case)


def _multi_devices():
    import torch
    n = torch.cuda.device_count()
    return [[0], [0, 0, 0]] + ([list(range(n))] if n > 1 else [])


@pytest.mark.parametrize("devices", _multi_devices() if __import__("torch").cuda.is_available() else [[0]], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_device_layer(devices):
    """The multi-GPU entry points of the C ABI: stream sharding and time chunking over the listed devices (one device
    listed several times exercises the sharding on a single-GPU box; with more GPUs the gather runs over NCCL)."""
    import ctypes as C
    import torch
    import foo_dsp_resampler_b200 as pkg
    import test_emulation

    def to_dev(a, dev):
        t = torch.from_numpy(a).to("cuda:%d" % dev)
        return (t, t.data_ptr(), lambda: (torch.cuda.synchronize(dev), t.cpu().numpy())[1])

    test_emulation.multi_check(pkg.product(), devices, (48000, 44100, 2, 11, 30000), (384000, 48000, 8, 384000, 120000), to_dev=to_dev)
