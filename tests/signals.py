"""Deterministic synthetic input: per channel c, frame i at rate fs
   x = 0.5*sin(2*pi*(f0*t + (f1-f0)*t^2/(2T)) + 0.3*c) + 0.05*u,  t = i/fs, f0 = 20 Hz, f1 = 0.45*fs,
with u uniform(-1,1) from a counter-based integer hash of (seed, stream, c, i), so any window of any
stream can be generated independently. Parity always runs on identical BYTES: generated once here
(numpy, float64 -> float32) and handed to both sides."""
import numpy as np

SEED = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def noise(stream, channel, start, count):
    with np.errstate(over="ignore"):
        i = np.arange(start, start + count, dtype=np.uint64)
        key = np.uint64((SEED + 0x632BE59BD9B4E019 * (stream * 64 + channel + 1)) & _M64)
        z = _mix(i * np.uint64(0x9E3779B97F4A7C15) + key)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def sweep_noise(fs, nch, frames, duration_frames=None, stream=0, start=0):
    """float32 [frames, nch] interleaved."""
    T = (duration_frames or frames) / float(fs)
    t = (np.arange(start, start + frames, dtype=np.float64)) / fs
    f0, f1 = 20.0, 0.45 * fs
    ph = 2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * T))
    x = np.empty((frames, nch), dtype=np.float32)
    for c in range(nch):
        x[:, c] = (0.5 * np.sin(ph + 0.3 * c) + 0.05 * noise(stream, c, start, frames)).astype(np.float32)
    return x
