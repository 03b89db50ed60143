"""Throughput of the reference-facing streaming protocol (RR_push / RR_pull / RR_drain with HOST buffers), the way
foo_dsp_rate.cpp drives it (65536-frame pushes, pull until empty): one handle, then one handle per host thread
(the shape of the CPU reference arm), pageable and page-locked caller buffers. Prints one JSON object per case."""
import ctypes as C, json, sys, threading, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import foo_dsp_resampler_b200 as pkg
import signals


def one_stream(cfg, nch, engine, x, chunk, pinned, out_buf, times=None):
    """Returns frames out; times gets (push seconds, pull seconds, seconds from first push to last pull: the handle's
    open / close -- plan design, table upload, ring allocation -- is not part of the stream's throughput)."""
    r = pkg.RateConverter(cfg, nch, engine)
    L, h = r.lib, r.h
    t_begin = time.perf_counter()
    ogen = C.c_size_t(0)
    n_out, t_push, t_pull = 0, 0.0, 0.0
    cap = out_buf.shape[0]
    optr = out_buf.data_ptr() if pinned else out_buf.ctypes.data
    for s in range(0, x.shape[0], chunk):
        blk = x[s:s + chunk]
        t0 = time.perf_counter()
        L.RR_push(h, blk.data_ptr() if pinned else blk.ctypes.data, blk.shape[0])
        t1 = time.perf_counter()
        while True:
            L.RR_pull(h, optr, cap, C.byref(ogen))
            if not ogen.value:
                break
            n_out += ogen.value
        t2 = time.perf_counter()
        t_push += t1 - t0; t_pull += t2 - t1
    L.RR_drain(h)
    while True:
        L.RR_pull(h, optr, cap, C.byref(ogen))
        if not ogen.value:
            break
        n_out += ogen.value
    t_end = time.perf_counter()
    r.close()
    if times is not None:
        times.append((t_push, t_pull, t_end - t_begin))
    return n_out


def run(i, o, nch, engine, secs=60, chunk=65536, pinned=False, threads=1, reps=4):
    cfg = pkg.make_config(i, o)
    xn = signals.sweep_noise(i, nch, int(i * secs))
    res = []
    for rep in range(reps):
        if pinned:
            xs = [torch.from_numpy(xn).pin_memory() for _ in range(threads)]
            outs = [torch.empty((chunk * 2 + 8192, nch), dtype=torch.float32).pin_memory() for _ in range(threads)]
        else:
            xs = [xn] * threads
            outs = [np.empty((chunk * 2 + 8192, nch), np.float32) for _ in range(threads)]
        times, tot = [], [0] * threads

        def work(k):
            tot[k] = one_stream(cfg, nch, engine, xs[k], chunk, pinned, outs[k], times)
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.perf_counter() - t0
        inner = max(t[2] for t in times)
        res.append({"Msamples_per_s": sum(tot) * nch / inner / 1e6, "with_open_close_Msamples_per_s": sum(tot) * nch / dt / 1e6,
                    "seconds": inner, "push_s": sum(t[0] for t in times) / threads, "pull_s": sum(t[1] for t in times) / threads})
    best = max(res, key=lambda r: r["Msamples_per_s"])
    print(json.dumps({"case": "%d->%d %dch %s chunk %d %s x%d threads" % (i, o, nch, engine, chunk, "pinned" if pinned else "pageable", threads),
                      "best": best, "all_Msamples_per_s": [round(r["Msamples_per_s"], 1) for r in res]}), flush=True)


if __name__ == "__main__":
    run(44100, 48000, 2, "float")
    run(44100, 48000, 2, "float", pinned=True)
    run(44100, 48000, 2, "double")                      # what RR_open(Best) selects
    run(44100, 48000, 2, "double", pinned=True)
    run(44100, 48000, 2, "float", chunk=4096)
    run(44100, 48000, 2, "float", chunk=1 << 20, pinned=True)
    run(192000, 44100, 8, "double", secs=20)
    run(44100, 48000, 2, "float", threads=16, secs=20)
    run(44100, 48000, 2, "float", threads=16, secs=20, pinned=True)
