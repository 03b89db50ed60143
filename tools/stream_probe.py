"""Throughput of the reference-facing streaming protocol (RR_push / RR_pull / RR_drain with host buffers),
one handle, the way foo_dsp_rate.cpp drives it (65536-frame pushes, pull until empty)."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import foo_dsp_resampler_b200 as pkg
import signals

def run(i, o, nch, engine, secs=60, chunk=65536):
    cfg = pkg.make_config(i, o)
    x = signals.sweep_noise(i, nch, int(i * secs))
    r = pkg.RateConverter(cfg, nch, engine)
    t0 = time.perf_counter()
    n_out = 0
    for s in range(0, x.shape[0], chunk):
        r.push(x[s:s + chunk])
        while True:
            y = r.pull(chunk + 8192)
            if not len(y):
                break
            n_out += len(y)
    r.drain()
    while True:
        y = r.pull(chunk + 8192)
        if not len(y):
            break
        n_out += len(y)
    dt = time.perf_counter() - t0
    r.close()
    print("%d->%d %dch %s chunk %d: %.1f Msamples/s out (%.0fx real time), %d frames" % (i, o, nch, engine, chunk, n_out * nch / dt / 1e6, n_out / o / dt, n_out))

for eng in ("float", "double"):
    run(44100, 48000, 2, eng)
    run(44100, 48000, 2, eng)
run(192000, 44100, 8, "double", secs=20)
run(44100, 48000, 2, "float", chunk=4096)
