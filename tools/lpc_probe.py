"""Times the track-edge extrapolation kernels (csrc/lpc.cu) with CUDA events on the launching stream, next to the
reference's own lpc/lpc.cpp (oracle/_ref/libref_lpc.so, one host thread, a bounded sample of the lanes), and checks
the sample bit for bit. Prints one JSON line per workload.

    python tools/lpc_probe.py [--streams 4096]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import foo_dsp_resampler_b200 as pkg  # noqa: E402
import lpclib  # noqa: E402
from foo_dsp_resampler_b200 import _capi  # noqa: E402


def run(name, in_rate, out_rate, nch, nstreams, track_frames, reps=5):
    lib = _capi.product()
    add, drop, prime, _ = pkg.track_edge_lengths(in_rate, out_rate)
    padded = track_frames + 2 * add
    rng = np.random.default_rng(7)
    one = lpclib.signal(0, track_frames, nch, seed=3)
    host = np.zeros((nstreams, padded, nch), np.float32)
    host[:, add:add + track_frames] = one[None] * rng.uniform(0.5, 1.0, (nstreams, 1, 1)).astype(np.float32)
    d = torch.from_numpy(host).cuda()
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for r in range(reps + 2):
        ev[0].record()
        rc = lib.RRX_lpc_extend_tracks(d.data_ptr(), nstreams, track_frames, prime, nch, 32, add, st)
        ev[1].record()
        assert rc == 0, lib.RRX_last_error()
        torch.cuda.synchronize()
        if r >= 2:
            ms.append(ev[0].elapsed_time(ev[1]))
    got = d[:8].cpu().numpy()
    # reference on a bounded sample of the streams, one host thread
    sample = min(nstreams, 8)
    want = host[:sample].copy()
    fn = lpclib.ref_extrapolate2 if lpclib.ref_available() else lpclib.oracle_extrapolate2
    t0 = time.perf_counter()
    for s in range(sample):
        fn(want[s], add, prime, add, 0)
        fn(want[s], add + track_frames - prime, prime, 0, add)
    cpu_s = (time.perf_counter() - t0) / sample
    exact = bool(np.array_equal(got[:sample].view(np.uint32), want.view(np.uint32)))
    t = float(np.median(ms)) * 1e-3
    lanes = nstreams * nch
    print(json.dumps({"workload": name, "rates": [in_rate, out_rate], "streams": nstreams, "channels": nch,
                      "prime": prime, "add": add, "gpu_ms": round(t * 1e3, 4),
                      "predicted_Msamples_per_s": round(2 * add * lanes / t / 1e6, 1),
                      "lane_edges_per_s": round(2 * lanes / t, 1),
                      "cpu_reference_ms_per_stream_1thread": round(cpu_s * 1e3, 4),
                      "cpu_kind": "reference" if lpclib.ref_available() else "port",
                      "speedup_vs_1thread": round(cpu_s * nstreams / t, 1), "bit_exact_sample": exact}))
    assert exact


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=4096)
    a = ap.parse_args()
    run("cfg4 tracks: 48 -> 44.1 kHz stereo, 10 s", 48000, 44100, 2, a.streams, 480000 // 10)   # edges only need the ends
    run("cfg1 tracks: 44.1 -> 48 kHz stereo", 44100, 48000, 2, a.streams, 44100)
    run("one stereo track 44.1 -> 48 kHz", 44100, 48000, 2, 1, 44100)
    run("cfg5-like: 384 -> 48 kHz 8 ch (prime 16384, add 8192)", 384000, 48000, 8, max(1, a.streams // 64), 65536)
