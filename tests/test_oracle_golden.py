"""CPU tier: the oracle restatement (oracle/rate_oracle.c) against the committed golden fixtures, which were
generated from the compiled reference by tests/golden/make_golden.py. Bit-exact everywhere: plans, designed
coefficient banks, frame counts after every push, and samples -- including the fp64 engine, because the oracle
restates the reference's own Ooura transform."""
import hashlib
import json
import os

import numpy as np
import pytest

import oraclelib
import signals

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_cases.json")) as f:
    CASES = json.load(f)
OUTPUTS = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_fixture(name):
    g = CASES[name]
    i, o, eng, ph, bw, al, q, nch, frames, chunk = g["case"]
    cfg = oraclelib.make_config(i, o, ph, bw, al, q)
    x = signals.sweep_noise(i, nch, frames)
    assert sha(x) == g["in_sha256"], "input generator drifted: fixtures no longer apply"
    orc = oraclelib.OracleResampler(cfg, nch, eng)
    assert orc.plan() == g["plan"]
    assert sha(orc.dft_coefs(0)) == g["coef_sha256"]["dft0"]
    assert sha(orc.dft_coefs(1)) == g["coef_sha256"]["dft1"]
    if "poly" in g["coef_sha256"]:
        assert sha(orc.poly_coefs()) == g["coef_sha256"]["poly"]
    orc.close()
    y, counts = oraclelib.resample(cfg, x, engine=eng, chunk=chunk, native=True)
    assert counts == g["counts"]
    assert y.shape[0] == g["out_frames"]
    assert np.array_equal(y, OUTPUTS[name])
    assert sha(y) == g["out_sha256"]


def test_reference_invariants_on_fixtures():
    """Properties the survey measured on the reference."""
    for name, g in CASES.items():
        i, o, eng, ph, bw, al, q, nch, frames, chunk = g["case"]
        # total frames after drain = round(n_in * out / in) (rate_flush, rate/rate_base.h:457)
        assert g["out_frames"] == int(frames / (i / o) + .5), name
        assert sum(g["counts"]) == g["out_frames"]
    # linear phase: an impulse at input frame p peaks at output round(p * out / in)
    cfg = oraclelib.make_config(44100, 48000)
    x = np.zeros((6000, 1), np.float32)
    x[1000] = 1.0
    y, _ = oraclelib.resample(cfg, x, engine="float")
    assert abs(int(np.argmax(np.abs(y[:, 0]))) - round(1000 * 48000 / 44100)) <= 1
