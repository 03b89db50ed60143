"""Generates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/libref_rate.so, i.e. the
unmodified /root/reference/rate sources -- see oracle/Makefile). Run in the build container, where
/root/reference exists:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these pin the oracle restatement
(and through it the CUDA path) to outputs of the reference itself: stage plans, designed coefficient banks
(as SHA-256 of their bytes), frame counts after every push, and the resampled output of short deterministic
inputs, for the five BASELINE.json configurations and the stage kinds they do not reach."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402
import signals  # noqa: E402

# name: (in_rate, out_rate, engine, phase, bandwidth, allow_aliasing, quality, channels, frames, chunk)
CASES = {
    "cfg1_44k1_48k_f32": (44100, 48000, "float", 50, 95, 0, 0, 2, 11025, 4096),
    "cfg2_44k1_96k_f32": (44100, 96000, "float", 50, 95, 0, 0, 2, 11025, 4096),
    "cfg3_192k_44k1_f64_ph25": (192000, 44100, "double", 25, 95, 0, 0, 8, 48000, 16384),
    "cfg4_48k_44k1_f32": (48000, 44100, "float", 50, 95, 0, 0, 2, 12000, 5000),
    "cfg5_384k_48k_f32": (384000, 48000, "float", 50, 95, 0, 0, 8, 96000, 30000),
    "vpoly2_44k1_48001_f32": (44100, 48001, "float", 50, 95, 0, 0, 1, 11025, 4096),
    "vpoly1_44k1_44101_norm_f32": (44100, 44101, "float", 50, 95, 0, 1, 1, 11025, 4096),
    "vpoly3_48k_47999_f64": (48000, 47999, "double", 50, 95, 0, 0, 1, 12000, 4096),
    "zerostuff_8k_48k_f32": (8000, 48000, "float", 50, 95, 0, 0, 1, 4000, 1000),
    "timedecim_48k_32k_f32": (48000, 32000, "float", 50, 95, 0, 0, 1, 12000, 4096),
    "minphase_norm_44k1_48k_f32": (44100, 48000, "float", 0, 90, 0, 1, 1, 11025, 4096),
    "phase75_alias_96k_44k1_f32": (96000, 44100, "float", 75, 99, 1, 0, 1, 24000, 8192),
    "three_stage_22k05_96k_f32": (22050, 96000, "float", 50, 95, 0, 0, 1, 5512, 2048),
    "fdomain_quarter_176k4_44k1_f64": (176400, 44100, "double", 50, 95, 0, 0, 2, 44100, 10000),
    # stage modes the BASELINE configurations do not reach (parity holes)
    "fdomain_quarter_l3_32k_24k_f32": (32000, 24000, "float", 50, 95, 0, 0, 2, 16000, 5000),     # L = 3, step -2
    "timedecim4_l3_32k_24k_alias_f32": (32000, 24000, "float", 50, 95, 1, 0, 2, 16000, 5000),    # L = 3, time-domain / 4
    "h9_norm_44k1_8k_f32": (44100, 8000, "float", 50, 95, 0, 1, 2, 22050, 8192),
    "h10_norm_49M152_44k1_f32": (49152000, 44100, "float", 50, 95, 0, 1, 1, 1200000, 400000),    # 11 stages
    "h13_best_24M576_44k1_f32": (24576000, 44100, "float", 50, 95, 0, 0, 2, 800000, 300000),     # 10 stages
    "h8_norm_32k_8k_f32": (32000, 8000, "float", 50, 95, 0, 1, 2, 16000, 5000),
    "h11_best_32k_8k_f64": (32000, 8000, "double", 50, 95, 0, 0, 1, 16000, 5000),
    "identity_48k_48k_f32": (48000, 48000, "float", 50, 95, 0, 0, 2, 5000, 1777),                 # no stages at all
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    meta = {}
    arrays = {}
    for name, (i, o, eng, ph, bw, al, q, nch, frames, chunk) in CASES.items():
        cfg = reflib.make_config(i, o, ph, bw, al, q)
        r = reflib.RefResampler(cfg, nch, eng)
        plan = r.plan()
        coefs = {"dft0": sha(r.dft_coefs(0)), "dft1": sha(r.dft_coefs(1))}
        poly = [s for s in plan["stages"] if s["kind"] == 2]
        if poly:
            s = poly[0]
            phases = s["L"] if s["interp_order"] == 0 else (1 << s["phase_bits"])
            coefs["poly"] = sha(r.poly_coefs(s["n"] * phases * (s["interp_order"] + 1)))
        r.close()
        x = signals.sweep_noise(i, nch, frames)
        y, counts = reflib.resample(cfg, x, engine=eng, chunk=chunk, native=True)
        meta[name] = {"case": [i, o, eng, ph, bw, al, q, nch, frames, chunk], "plan": plan, "coef_sha256": coefs,
                      "counts": counts, "out_frames": int(y.shape[0]), "out_sha256": sha(y), "in_sha256": sha(x)}
        arrays[name] = y
    with open(os.path.join(HERE, "reference_cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **arrays)
    print("wrote %d cases" % len(meta))


if __name__ == "__main__":
    main()
