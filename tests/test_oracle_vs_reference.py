"""CPU tier, build container only: the oracle restatement against the compiled reference checker
oracle/_ref/libref_rate.so (the unmodified reference sources, see oracle/Makefile), bit for bit.
Skipped where the checker has not been built (it travels to the GPU box prebuilt, so it normally exists)."""
import ctypes as C

import numpy as np
import pytest

import oraclelib
import reflib
import signals

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref/libref_rate.so not built")


@pytest.mark.parametrize("n", [16, 64, 512, 2048, 4096, 8192, 32768])
def test_real_ffts_bit_exact(n):
    R, O = reflib.lib(), oraclelib.lib()
    R.ref_rdft_f32.argtypes = [C.c_int, C.c_int, C.c_void_p]
    R.ref_rdft_f64.argtypes = [C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(n)
    for inv in (0, 1):
        x = rng.standard_normal(n).astype(np.float32)
        a, b = x.copy(), x.copy()
        R.ref_rdft_f32(n, inv, a.ctypes.data)
        O.orc_rdft_f32(n, inv, b.ctypes.data)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        xd = rng.standard_normal(n)
        ad, bd = xd.copy(), xd.copy()
        R.ref_rdft_f64(n, inv, ad.ctypes.data)
        O.orc_rdft_f64(n, inv, bd.ctypes.data)
        assert np.array_equal(ad.view(np.uint64), bd.view(np.uint64))


GRID = [(i, o, eng, ph, bw, al, q)
        for (i, o) in [(44100, 48000), (48000, 44100), (44100, 96000), (96000, 44100), (192000, 44100),
                       (384000, 48000), (44100, 22050), (22050, 44100), (8000, 44100), (44100, 8000),
                       (32000, 48000), (48000, 32000), (44100, 48001), (48000, 47999), (11025, 192000),
                       (88200, 96000), (44100, 176400), (176400, 48000), (37800, 44100), (44100, 5513)]
        for (eng, ph, bw, al, q) in [("float", 50, 95, 0, 0), ("double", 25, 95, 0, 0), ("float", 50, 99, 1, 1),
                                     ("double", 75, 90, 0, 1)]]


def test_plans_and_designs_over_a_grid():
    """Planner fidelity: every plan integer and every designed coefficient."""
    for i, o, eng, ph, bw, al, q in GRID:
        cfg = reflib.make_config(i, o, ph, bw, al, q)
        r = reflib.RefResampler(cfg, 1, eng)
        orc = oraclelib.OracleResampler(cfg, 1, eng)
        pr, po = r.plan(), orc.plan()
        assert pr == po, (i, o, eng, ph, bw, al, q)
        for inst in (0, 1):
            assert np.array_equal(r.dft_coefs(inst), orc.dft_coefs(inst)), (i, o, eng, ph, inst)
        pc = orc.poly_coefs()
        if len(pc):
            assert np.array_equal(r.poly_coefs(len(pc)), pc), (i, o, eng)
        r.close()
        orc.close()


@pytest.mark.parametrize("case", [(44100, 48000, "float", 50, 2), (192000, 44100, "double", 25, 3),
                                  (384000, 48000, "float", 50, 2), (44100, 48001, "double", 50, 1),
                                  (8000, 48000, "float", 30, 1), (48000, 32000, "double", 50, 1)])
def test_streaming_outputs_and_counts(case):
    i, o, eng, ph, nch = case
    cfg = reflib.make_config(i, o, phase=ph)
    x = signals.sweep_noise(i, nch, int(i * 0.4) + 5)
    for chunk in (65536, 1000):
        yr, cr = reflib.resample(cfg, x, engine=eng, chunk=chunk, native=True)
        yo, co = oraclelib.resample(cfg, x, engine=eng, chunk=chunk, native=True)
        assert cr == co
        assert np.array_equal(yr, yo)
