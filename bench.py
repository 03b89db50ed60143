#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 rate engine (contract: see the task statement / DESIGN.md).

Metric (BASELINE.json): output Msamples/s, device-timed, whole job over all N GPUs
    = output frames x channels x streams (all ranks) / max-over-ranks device time per step.
A "step" is one pass of the hot path (every stage kernel of the plan) over one batch of synthetic input that
is already resident in HBM. Default workload = BASELINE config 4: 4096 independent stereo 48 kHz -> 44.1 kHz
fp32 Best streams of 10 s per GPU, sharded by stream (weak scaling: every rank converts its own 4096).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--impl reference]

For N > 1 launch with torchrun (one rank per GPU, NCCL only for the barrier / max-reduction / result gather).
`--impl reference` times the reference's own CPU implementation (oracle/_ref, SSE engines, one stream per host
thread) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (in_rate, out_rate, channels, engine, phase, seconds per stream, streams per GPU, description)
WORKLOADS = {
    "cfg1": (44100, 48000, 2, "float", 50.0, 60.0, 1,
             "BASELINE config 1: 44.1->48 kHz stereo fp32 Best linear-phase, one 60 s stream"),
    "cfg1x256": (44100, 48000, 2, "float", 50.0, 60.0, 256,
                 "config 1 conversion as a batch of 256 x 60 s stereo streams"),
    "cfg2": (44100, 96000, 2, "float", 50.0, 60.0, 256,
             "BASELINE config 2: 44.1->96 kHz stereo fp32 upsample, 256 x 60 s streams"),
    "cfg3": (192000, 44100, 8, "double", 25.0, 60.0, 16,
             "BASELINE config 3: 192->44.1 kHz 8-channel fp64 phase 25, 16 x 60 s streams"),
    "cfg4": (48000, 44100, 2, "float", 50.0, 10.0, 4096,
             "BASELINE config 4: 4096 independent stereo 48->44.1 kHz fp32 Best streams of 10 s per GPU"),
    "cfg5": (384000, 48000, 8, "float", 50.0, 4500.0, 1,
             "BASELINE config 5: one 8-channel 384->48 kHz fp32 stream, 10 h over 8 GPUs (1.25 h per GPU), "
             "time-chunked with filter-history halos and closed-form start phase"),
}


# what binds each dominant kernel, from its ncu --set full capture (profiles/r2_ncu_full_*.txt, profiles/README.md)
LIMITERS = {
    "dftp_kernel": "shared-memory pipe: l1tex 88 % busy (343 M wavefronts per 256-stream launch, 8 % of them bank conflicts) "
                   "moving FFT passes through shared memory; 20 warps per SM (96 registers x 640 threads, 185 KB shared memory), "
                   "issue 41 %, DRAM 19 % (profiles/r2_ncu_full_cfg4x256.txt) -- not HBM, although `bound` is reported against "
                   "the HBM roof as the contract asks; `alu` is the arithmetic fraction",
    "poly0_pair2_kernel": "shared-memory pipe: l1tex 84 % busy (24 x 8 B of window per output pair), DRAM 39 %",
    "poly0_pair_kernel": "shared-memory pipe (window reads), DRAM ~35 %",
    "halfband_pair_kernel": "shared-memory pipe: l1tex 86 % busy (8-byte window fills + 16-byte window reads), DRAM 41 %",
    "dft64_kernel": "latency of DFMA / LDS chains at 12 warps per SM (160 registers): issue 50 %, shared-memory pipe 61 %, DRAM 19 %",
    "poly0_dual_kernel": "shared-memory latency on 8-byte window reads; DRAM ~35 %",
    "halfband_kernel": "shared-memory pipe and latency (fp64, eight outputs per thread); DRAM ~45 %",
}

_JSON_FD = None     # saved stdout when the process-level stdout has been redirected (multi-rank runs)


def emit_json(line):
    """The one JSON line of the contract, on the real stdout."""
    text = json.dumps(line) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, text.encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every 10 ms (starts at
    once, so even a 100 ms region is covered); `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:                                   # CUDA index -> physical index
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                index = int(ids[index])
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag, self.t = None, [], False, None

    def _nvml_loop(self):
        n, h = self.nvml
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), {k for k, b in bits.items() if r & b}))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            h = n.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(s for s, _ in self.samples)
            reasons = set()
            for _, r in self.samples:
                reasons |= r
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------
# reference / CPU arm
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, budget_s=12.0, steps=1, threads=None):
    """Times oracle/_ref (the reference's own C sources, gcc -O3 -msse3 -ffp-contract=off, SSE engines:
    RR_ctor_SSE for fp32 -- with the C FFT standing in for the yasm one -- RR_ctor_SSE3 for fp64), one
    independent stream per host thread, 64 Ki-frame pushes, pull until empty, drain. Returns a list of
    (output samples, seconds) per step plus a description."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    in_rate, out_rate, nch, engine, phase, seconds, _, _ = WORKLOADS[wl]
    L = C.CDLL(reflib.REF_SO)
    L.ref_reinit.argtypes = [C.c_int, C.c_int, reflib._OOM]
    # fp32 SSE engine runs on sse==0 FFT tables because the C FFT stands in for the assembler one
    L.ref_reinit(1 if engine == "float" else 0, 0, reflib._oom_handler)
    ctor = L.RR_ctor_SSE if engine == "float" else L.RR_ctor_SSE3
    ctor.restype = C.c_void_p
    ctor.argtypes = [C.POINTER(reflib.RRConfig), C.c_int]
    L.RR_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.RR_pull.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.RR_drain.argtypes = [C.c_void_p]
    L.RR_close.argtypes = [C.POINTER(C.c_void_p)]
    cfg = reflib.make_config(in_rate, out_rate, phase=phase)
    sec = min(seconds, 10.0)                       # bounded sample: at most 10 s of each stream
    n = int(in_rate * sec)
    rng = np.random.default_rng(1)
    t = np.arange(n) / in_rate
    x = np.empty((n, nch), np.float32)
    for c in range(nch):
        x[:, c] = 0.5 * np.sin(2 * np.pi * (20 * t + (0.45 * in_rate - 20) * t * t / (2 * sec)) + 0.3 * c) \
            + 0.05 * rng.uniform(-1, 1, n)
    nthreads = threads or os.cpu_count() or 1

    def one_stream():
        h = C.c_void_p(ctor(C.byref(cfg), nch))
        out = np.empty((1 << 17, nch), np.float32)
        got, tot = C.c_size_t(0), 0
        for s in range(0, n, 65536):
            blk = x[s:s + 65536]
            L.RR_push(h, blk.ctypes.data, blk.shape[0])
            while True:
                L.RR_pull(h, out.ctypes.data, out.shape[0], C.byref(got))
                if not got.value:
                    break
                tot += got.value
        L.RR_drain(h)
        while True:
            L.RR_pull(h, out.ctypes.data, out.shape[0], C.byref(got))
            if not got.value:
                break
            tot += got.value
        L.RR_close(C.byref(h))
        return tot

    t0 = time.perf_counter()
    frames = one_stream()
    t_one = time.perf_counter() - t0
    reps = max(1, int(budget_s / max(t_one, 1e-3)))

    def worker(res, i):
        tot = 0
        for _ in range(reps):
            tot += one_stream()
        res[i] = tot

    results = []
    for _ in range(steps):
        res = [0] * nthreads
        ths = [threading.Thread(target=worker, args=(res, i)) for i in range(nthreads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        results.append((sum(res) * nch, dt))
    desc = ("%d host threads x %d streams each of %.0f s %d-ch %d->%d Hz through RR_ctor_%s/RR_push(64Ki)/RR_pull/RR_drain "
            "of oracle/_ref (reference sources, gcc -O3 -msse3 -ffp-contract=off%s)" %
            (nthreads, reps, sec, nch, in_rate, out_rate, "SSE" if engine == "float" else "SSE3",
             "; C FFT substituted for the yasm FFT" if engine == "float" else ""))
    assert frames > 0
    return results, nthreads, desc


def workload_config(wl, streams, seconds, world, scaling, stages=None, l2_note=None):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm of one workload."""
    in_rate, out_rate, nch, engine, phase, _, _, text = WORKLOADS[wl]
    chunked = wl == "cfg5"
    cfg = {"workload": wl, "description": text, "in_rate": in_rate, "out_rate": out_rate, "channels": nch, "engine": engine,
           "phase": phase, "quality": "Best", "bandwidth_pc": 95.0,
           "streams_total": streams * (world if scaling == "weak" and not chunked else 1) if not chunked else 1,
           "seconds_per_stream": seconds * (world if chunked else 1),
           "parallelism": ("one stream time-chunked over %d GPU(s) with filter-history halos" % world) if chunked else
                          ("streams sharded over %d GPU(s), no data-path collective" % world)}
    return cfg


def default_shape(args, wl, world):
    """(streams per rank, seconds per stream per rank, scaling) of a workload on `world` GPUs. BASELINE config 4 is
    4096 streams IN TOTAL, sharded by stream (strong scaling); config 5 is one 10-hour stream cut into `world` time
    chunks (1.25 h per GPU at 8 GPUs; a single GPU holds 20 minutes of it in HBM next to the result)."""
    in_rate, out_rate, nch, engine, phase, seconds, streams, text = WORKLOADS[wl]
    scaling = "strong" if wl == "cfg4" else "weak"
    if args.streams:
        streams = args.streams
    elif wl == "cfg4" and world > 1 and not args.weak:
        streams = streams // world
    if args.weak:
        scaling = "weak"
    if args.seconds:
        seconds = args.seconds
    elif wl == "cfg5":
        seconds = 36000.0 / 8 if world == 8 else 1200.0
    return streams, seconds, scaling


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl = args.workload
    results, nthreads, desc = cpu_reference_run(wl, budget_s=8.0, steps=args.warmup + args.steps)
    timed = results[args.warmup:]
    samples = sum(s for s, _ in timed)
    secs = sum(t for _, t in timed)
    value = samples / secs / 1e6
    in_rate, out_rate, nch, engine, phase, _, _, text = WORKLOADS[wl]
    streams, seconds, scaling = default_shape(args, wl, world)
    line = {
        "impl": "reference", "metric": "output Msamples/s", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(timed),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32" if engine == "float" else "f64", "data": "synthetic",
        "config": workload_config(wl, streams, seconds, world, scaling),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": nthreads, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def track_edges(torch, pkg, in_rate, out_rate, nch, nstreams, reps=5):
    """The step next to the path in the caller: LPC extrapolation of both edges of `nstreams`
    tracks (RRX_lpc_extend_tracks, csrc/lpc.cu), CUDA events on the launching stream; the reference's lpc/lpc.cpp
    (oracle/_ref/libref_lpc.so, one host thread) is timed on a sample of 8 tracks and compared bit for bit --
    this is the cpu_baseline leg of that step, the only place this function touches the checker."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import lpclib
    lib = pkg.product()
    add, drop, prime, _ = pkg.track_edge_lengths(in_rate, out_rate)
    frames = 2 * prime                               # the edges only touch the ends of a track
    padded = frames + 2 * add
    rng = np.random.default_rng(5)
    one = lpclib.signal(0, frames, nch, seed=3)
    host = np.zeros((nstreams, padded, nch), np.float32)
    host[:, add:add + frames] = one[None] * rng.uniform(0.5, 1.0, (nstreams, 1, 1)).astype(np.float32)
    d = torch.from_numpy(host).cuda()
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for r in range(reps + 3):
        e0.record()
        rc = lib.RRX_lpc_extend_tracks(d.data_ptr(), nstreams, frames, prime, nch, pkg.LPC_ORDER, add, st)
        e1.record()
        if rc != 0:
            raise RuntimeError("RRX_lpc_extend_tracks: %s" % lib.RRX_last_error())
        torch.cuda.synchronize()
        if r >= 3:
            ms.append(e0.elapsed_time(e1))
    sample = min(nstreams, 8)
    got = d[:sample].cpu().numpy()
    want = host[:sample].copy()
    kind = "reference" if lpclib.ref_available() else "port"
    fn = lpclib.ref_extrapolate2 if kind == "reference" else lpclib.oracle_extrapolate2
    t0 = time.perf_counter()
    for s in range(sample):
        fn(want[s], add, prime, add, 0)
        fn(want[s], add + frames - prime, prime, 0, add)
    cpu_ms = (time.perf_counter() - t0) / sample * 1e3
    t = sorted(ms)[len(ms) // 2]
    return {"what": "LPC extrapolation of both track edges (lpc/lpc.cpp), %d tracks x %d ch, prime %d, add %d frames per edge"
                    % (nstreams, nch, prime, add),
            "ms": t, "gpu_launches": 2, "predicted_Msamples_per_s": 2 * add * nstreams * nch / t / 1e3,
            "cpu_baseline": {"ms_per_track": cpu_ms, "cores": 1, "kind": kind, "sample": "%d tracks" % sample},
            "bit_exact_vs_cpu_sample": bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))}


def make_input(torch, nstreams, frames, nch, in_rate, seconds, seed):
    """Synthetic sweep + noise (SURVEY.md 8d shape), generated on the device in bounded slabs (both along streams
    and along time, so a 1.25-hour 8-channel window needs no multi-GB temporaries)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = torch.empty((nstreams, frames, nch), dtype=torch.float32, device="cuda")
    fslab = 1 << 22                                               # frames per time slab
    sslab = max(1, (1 << 26) // (min(frames, fslab) * nch))       # streams per slab
    chan = 0.3 * torch.arange(nch, device="cuda", dtype=torch.float64)
    for f0 in range(0, frames, fslab):
        f1 = min(frames, f0 + fslab)
        t = torch.arange(f0, f1, device="cuda", dtype=torch.float64) / in_rate
        ph = 2 * torch.pi * (20.0 * t + (0.45 * in_rate - 20.0) * t * t / (2 * seconds))
        base = (0.5 * torch.sin(ph.unsqueeze(1) + chan.unsqueeze(0))).to(torch.float32)      # [frames, nch]
        del t, ph
        for s0 in range(0, nstreams, sslab):
            s1 = min(nstreams, s0 + sslab)
            noise = torch.rand((s1 - s0, f1 - f0, nch), generator=g, device="cuda")
            x[s0:s1, f0:f1] = base.unsqueeze(0) + 0.05 * (2 * noise - 1)
            del noise
        del base
    return x


PEAKS_NOMINAL = {"f32": 74.4, "f64": 37.2}       # TFLOP/s FMA, 148 SMs x 128 (64) lanes x 2 x 1.965 GHz


def load_fp_peaks():
    p = os.path.join(ROOT, "profiles", "fp_peaks.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class Measurement:
    """One workload resident on this rank's GPU: step(), device-timed throughput, per-stage times, rooflines."""

    def __init__(self, torch, pkg, wl, streams, seconds, rank, world, local_rank):
        self.torch, self.pkg, self.wl, self.rank, self.world = torch, pkg, wl, rank, world
        in_rate, out_rate, nch, engine, phase, _, _, text = WORKLOADS[wl]
        self.in_rate, self.out_rate, self.nch, self.engine, self.streams, self.seconds = in_rate, out_rate, nch, engine, streams, seconds
        self.cfg = pkg.make_config(in_rate, out_rate, phase=phase)
        self.st = torch.cuda.current_stream().cuda_stream
        self.chunked = wl == "cfg5"
        if not self.chunked:
            frames = int(round(in_rate * seconds))
            frames -= frames & 1                                   # even: every stream starts on a 16-byte boundary
            self.b = pkg.BatchConverter(self.cfg, nch, streams, frames, engine=engine, device=local_rank)
            self.nout = self.b.frames_out(frames)
            self.x = make_input(torch, streams, frames, nch, in_rate, seconds, 1234 + rank)
            self.y = torch.empty((streams, self.nout, nch), dtype=torch.float32, device="cuda")
            self.frames = frames
            self.pieces = None
            self.out_samples_per_step = self.nout * nch * streams
        else:
            # one long stream, this rank's contiguous range of the OUTPUT timeline, processed as consecutive chunks
            # with halo'd input windows (RRX_batch_input_window / RRX_batch_process_range)
            frames_total = int(round(in_rate * seconds)) * world
            chunk_out = out_rate * 60                               # 60 s of output per call
            self.b = pkg.BatchConverter(self.cfg, nch, 1, int(chunk_out * in_rate / out_rate) + 65536, engine=engine,
                                        device=local_rank)
            nout_total = self.b.frames_out(frames_total)
            import foo_dsp_resampler_b200.sharding as sharding
            ranges = sharding.time_chunks(nout_total, world, sharding.last_stage_block(self.b.plan()))
            out_lo, cnt = ranges[rank]
            out_hi = out_lo + cnt
            f0, c0 = self.b.input_window(frames_total, out_lo, cnt)
            self.x = make_input(torch, 1, c0, nch, in_rate, seconds * world, 99 + rank)
            self.y = torch.empty((1, cnt, nch), dtype=torch.float32, device="cuda")
            self.pieces = []
            for ob in range(out_lo, out_hi, chunk_out):
                oc = min(chunk_out, out_hi - ob)
                f, c = self.b.input_window(frames_total, ob, oc)
                self.pieces.append((ob, oc, f, c))
            self.f0, self.out_lo, self.frames_total, self.frames = f0, out_lo, frames_total, c0
            self.nout = cnt
            self.out_samples_per_step = cnt * nch
        self.in_bytes = self.x.numel() * 4
        self.out_bytes = self.y.numel() * 4
        self.l2_note = ("inputs (%.1f GB) larger than L2" % (self.in_bytes / 1e9)) if self.in_bytes > 256e6 else \
            "L2 flushed between steps (256 MB write)"
        self.flush = None if self.in_bytes > 256e6 else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(self):
        b, x, y, st, nch = self.b, self.x, self.y, self.st, self.nch
        if not self.chunked:
            b.process(x.data_ptr(), self.frames, y.data_ptr(), st)
        else:
            for ob, oc, f, c in self.pieces:
                b.process_range(x.data_ptr() + (f - self.f0) * nch * 4, f, c, self.frames_total, ob, oc,
                                y.data_ptr() + (ob - self.out_lo) * nch * 4, st)

    def timed(self, steps, W, barrier, dist, sampler=None):
        torch, b = self.torch, self.b
        for _ in range(W):
            self.step()
        barrier()
        b.enable_timing(True)
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stage_acc, launches = None, 0
        barrier()
        for k in range(steps):
            if self.flush is not None:
                self.flush.fill_(k)                        # evict L2 between timed iterations
            ev[k][0].record()
            self.step()
            ev[k][1].record()
            launches += b.last_launches() * (len(self.pieces) if self.chunked else 1)
            if not self.chunked:
                tms = b.stage_times()                      # waits for this step's events (recorded on the launching stream)
                stage_acc = tms if stage_acc is None else [a + c for a, c in zip(stage_acc, tms)]
        barrier()
        self.clocks = sampler.stop() if sampler else None
        b.enable_timing(False)
        ms_total = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        self.ms_step = float(tt.item()) / steps
        tot = torch.tensor([float(self.out_samples_per_step)], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        self.total_out_samples = float(tot.item())
        self.value = self.total_out_samples / self.ms_step / 1e3          # Msamples/s, whole job
        self.launches = launches
        self.stage_ms = [m / steps for m in stage_acc] if stage_acc else None
        return self.value

    def rooflines(self):
        """SURVEY.md 8(d) accounting. Pipeline: algorithmic flops (split-radix count etc.) and bytes (fp32 in + fp32 out at
        the API, intermediates on chip) of one step over the CUDA-event time of the step. Dominant kernel: its
        algorithmic bytes -- the API input if it reads it plus the API output if it writes it -- over its own
        CUDA-event duration."""
        if self.chunked or not self.stage_ms:
            flops = None
        else:
            flops = self.b.flops(self.frames)
        peak_hbm, how = load_peaks()
        fpp = load_fp_peaks()
        dt = "f32" if self.engine == "float" else "f64"
        meas_fma = fpp.get("fp32_ffma_tflops" if dt == "f32" else "fp64_dfma_tflops")
        per_rank_ms = self.ms_step
        pipe = {"hbm_GBps": (self.in_bytes + self.out_bytes) / per_rank_ms / 1e6,
                "hbm_frac": (self.in_bytes + self.out_bytes) / per_rank_ms / 1e6 / peak_hbm,
                "bytes_per_out_sample": (self.in_bytes + self.out_bytes) / self.out_samples_per_step}
        if flops:
            pipe.update({"tflops": flops / per_rank_ms / 1e9, "flop_per_out_sample": flops / self.out_samples_per_step,
                         "fma_frac_nominal": flops / per_rank_ms / 1e9 / PEAKS_NOMINAL[dt], "fma_peak_nominal_tflops": PEAKS_NOMINAL[dt],
                         "fma_frac_measured": (flops / per_rank_ms / 1e9 / meas_fma) if meas_fma else None,
                         "fma_peak_measured_tflops": meas_fma,
                         "note": "bit-faithful fp32 issues un-fused multiplies and adds (SURVEY.md Appendix A): its attainable "
                                 "arithmetic peak is half the FMA peak" if dt == "f32" else "DFMA allowed (1e-12 contract)"})
        roof = None
        if self.stage_ms:
            n = len(self.stage_ms)
            dom = max(range(n), key=lambda i: self.stage_ms[i])
            dur = self.stage_ms[dom]
            kname = self.b.stage_kernel(dom) or "?"
            fused_next = dom + 1 < n and (self.b.stage_kernel(dom + 1) or "").startswith("(fused")
            last = dom + 1 if fused_next else dom
            alg = (self.in_bytes if dom == 0 else 0) + (self.out_bytes if last == n - 1 else 0)
            work = self.b.stage_work(self.frames, dom)
            kflops = work["flops"] + (self.b.stage_work(self.frames, dom + 1)["flops"] if fused_next else 0.0)
            kfamily = kname.split(" ")[0]
            packed = kfamily in ("dftp_kernel", "dft_poly_kernel", "poly0_pair_kernel", "poly0_pair2_kernel", "halfband_pair_kernel")
            if dt == "f32":
                alu_peak = fpp.get("fp32_fmul2_fadd2_tflops" if packed else "fp32_fmul_fadd_tflops")
                alu_how = ("measured, un-fused packed FMUL2+FADD2" if packed else "measured, un-fused FMUL+FADD") + " (profiles/fp_peaks.json)"
            else:
                alu_peak, alu_how = fpp.get("fp64_dfma_tflops"), "measured, DFMA (profiles/fp_peaks.json)"
            traffic = None
            tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
            if os.path.exists(tp):
                with open(tp) as f:
                    ent = json.load(f).get(kfamily + ":" + self.wl) or (json.load(open(tp)).get(kfamily) if self.wl == "cfg4" else None)
                if ent:
                    units = self.out_samples_per_step if ent.get("unit") == "output sample" else work["units"]
                    traffic = ent["bytes_per_unit"] * units
            roof = {"bound": "hbm", "kernel": kname, "stage": dom, "achieved": alg / dur / 1e6, "peak": peak_hbm, "unit": "GB/s",
                    "frac": alg / dur / 1e6 / peak_hbm, "traffic": traffic, "peak_source": how, "ms_per_launch": dur,
                    "algorithmic_bytes_per_launch": alg, "share_of_step": dur / self.ms_step,
                    "limiter": LIMITERS.get(kfamily, "see the kernel's ncu summary in profiles/README.md"),
                    "alu": {"achieved_tflops": kflops / dur / 1e9, "peak_tflops": alu_peak,
                            "frac": (kflops / dur / 1e9 / alu_peak) if alu_peak else None, "peak_source": alu_how},
                    "stage_kernels": [self.b.stage_kernel(i) for i in range(n)], "stage_ms": self.stage_ms}
        return roof, pipe

    def close(self):
        self.b.close()
        del self.x, self.y, self.flush
        self.torch.cuda.empty_cache()


def e2e_batch(torch, pkg, m, dist, barrier, local_rank):
    """End to end through the host-buffer entry point: ALL of this rank's streams from pinned host memory through the
    device and back (H2D + kernels + D2H inside the timed region, three CUDA streams, sub-batches of 64 streams)."""
    sub = min(64, m.streams)
    tot = m.streams
    bh = pkg.BatchConverter(m.cfg, m.nch, sub, m.frames, engine=m.engine, device=local_rank)
    h_in = torch.empty((tot, m.frames, m.nch), dtype=torch.float32, pin_memory=True)
    h_out = torch.empty((tot, m.nout, m.nch), dtype=torch.float32, pin_memory=True)
    h_in.copy_(m.x)
    torch.cuda.synchronize()
    bh.process_host(h_in.data_ptr(), m.frames, h_out.data_ptr(), tot)      # warm-up (allocates the slots)
    barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        bh.process_host(h_in.data_ptr(), m.frames, h_out.data_ptr(), tot)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    h2d, d2h = tot * m.frames * m.nch * 4, tot * m.nout * m.nch * 4
    e2e = {"value": m.total_out_samples / dt / 1e6, "unit": "Msamples/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "RRX_batch_process_host (pinned host buffers, all %d streams of the rank in sub-batches of %d, "
                  "H2D / kernels / D2H on three CUDA streams)" % (tot, sub),
           "streams": tot, "seconds_per_step": dt, "h2d_GBps": h2d / dt / 1e9, "d2h_GBps": d2h / dt / 1e9,
           "note": "bound by the host link: both directions of PCIe carry %.1f B per output sample" % ((h2d + d2h) / (tot * m.nout * m.nch))}
    # the ceiling of this box for the same bytes: H2D and D2H of the whole batch on two streams, no kernels
    try:
        d_in, d_out = m.x, m.y
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ct = torch.tensor([(time.perf_counter() - t0) / 2], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        ct = float(ct.item())
        e2e["copy_only"] = {"seconds_per_step": ct, "value": m.total_out_samples / ct / 1e6, "h2d_GBps_per_gpu": h2d / ct / 1e9,
                            "d2h_GBps_per_gpu": d2h / ct / 1e9,
                            "note": "same bytes, both directions at once, no kernels: what the host link of this box allows "
                                    "(all ranks at once); e2e / copy_only = %.2f" % (ct / dt)}
    except Exception as exc:  # noqa: BLE001
        e2e["copy_only"] = {"error": str(exc)}
    # check: the bits of the device-resident run (slices, to bound the comparison's memory)
    ok = True
    for s0 in range(0, tot, 256):
        ok = ok and bool(torch.equal(h_out[s0:s0 + 256].cuda(), m.y[s0:s0 + 256]))
    e2e["matches_device_resident_output"] = ok
    bh.close()
    del h_in, h_out
    return e2e


def e2e_streaming(torch, pkg, m, nthreads=16, secs=60.0):
    """The reference's own call shape (the CPU arm's): one handle per host thread, RR_push(64 Ki frames) / RR_pull until
    empty / RR_drain on page-locked host buffers, all threads at once. Timed from the first push to the last pull of
    every thread (the handles are opened before and closed after: plan design and table upload are once-per-open)."""
    import ctypes as C
    import threading
    frames = int(m.in_rate * secs)
    reps = -(-frames // m.frames)
    xs = m.x[0].cpu().repeat(reps, 1)[:frames].contiguous()        # the workload's signal, repeated to `secs` seconds
    bufs = [(xs.clone().pin_memory(), torch.empty((1 << 17, m.nch), dtype=torch.float32).pin_memory()) for _ in range(nthreads)]
    tot = [0] * nthreads
    best = None
    for _ in range(2):
        handles = [pkg.RateConverter(m.cfg, m.nch, m.engine) for _ in range(nthreads)]
        start = threading.Barrier(nthreads + 1)
        for k in range(nthreads):
            tot[k] = 0

        def work(k):
            r = handles[k]
            L, h = r.lib, r.h
            x, out = bufs[k]
            ogen = C.c_size_t(0)

            def pull_all():
                while True:
                    L.RR_pull(h, out.data_ptr(), out.shape[0], C.byref(ogen))
                    if not ogen.value:
                        return
                    tot[k] += ogen.value
            start.wait()
            for s in range(0, frames, 65536):
                n = min(65536, frames - s)
                L.RR_push(h, x.data_ptr() + s * m.nch * 4, n)
                pull_all()
            L.RR_drain(h)
            pull_all()

        th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
        [t.start() for t in th]
        start.wait()
        t0 = time.perf_counter()
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        for r in handles:
            r.close()
        v = sum(tot) * m.nch / dt / 1e6
        best = v if best is None else max(best, v)
    return {"value": best, "unit": "Msamples/s", "handles": nthreads, "host_threads": nthreads,
            "api": "RR_ctor_%s / RR_push(64 Ki frames) / RR_pull / RR_drain, one handle per host thread, page-locked caller buffers"
                   % ("float" if m.engine == "float" else "double"),
            "h2d_bytes_per_step": frames * m.nch * 4 * nthreads, "d2h_bytes_per_step": int(sum(tot)) * m.nch * 4,
            "seconds_of_audio_per_handle": frames / m.in_rate}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--seconds", type=float, default=0.0, help="override seconds per stream (per GPU for cfg5)")
    ap.add_argument("--weak", action="store_true", help="cfg4 with 4096 streams PER GPU instead of in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import foo_dsp_resampler_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to file descriptor 1 on the first collective; stdout must carry exactly
        # one JSON line, so the process-level stdout is pointed at stderr and the line is written to the saved fd
        global _JSON_FD
        if _JSON_FD is None:
            sys.stdout.flush()
            _JSON_FD = os.dup(1)
            os.dup2(2, 1)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)            # timing rule: at least 3 warm-up steps

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    wl = args.workload
    streams, seconds, scaling = default_shape(args, wl, world)
    in_rate, out_rate, nch, engine, phase, _, _, text = WORKLOADS[wl]
    m = Measurement(torch, pkg, wl, streams, seconds, rank, world, local_rank)
    value = m.timed(args.steps, W, barrier, dist, ClockSampler(local_rank))
    roofline, pipeline = m.rooflines()
    plan = m.b.plan()

    e2e, e2e_s = None, None
    if not args.no_e2e and not m.chunked:
        try:
            e2e = e2e_batch(torch, pkg, m, dist, barrier, local_rank)
        except Exception as exc:  # noqa: BLE001
            e2e = {"value": None, "unit": "Msamples/s", "error": str(exc)}
        if world == 1:
            try:
                e2e_s = e2e_streaming(torch, pkg, m)
            except Exception as exc:  # noqa: BLE001
                e2e_s = {"value": None, "unit": "Msamples/s", "error": str(exc)}

    # ---- NCCL gather of the real result shards (reported separately and inclusive; the only collective on the path) ----
    gather = None
    if dist:
        shard = m.y.reshape(-1)
        n_max = torch.tensor([shard.numel()], dtype=torch.int64, device="cuda")
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
        n_max = int(n_max.item())
        send = shard if shard.numel() == n_max else torch.cat([shard, shard.new_zeros(n_max - shard.numel())])
        bufs = torch.empty((world, n_max), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(bufs, send)                      # warm-up
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(bufs, send)
        g1.record()
        torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gms = float(gt.item())
        ok = bool(torch.equal(bufs[rank, :shard.numel()], shard))
        gather = {"collective": "ncclAllGather of every rank's whole result shard", "bytes_per_rank": n_max * 4, "ms": gms,
                  "GBps_per_rank": n_max * 4 * (world - 1) / gms / 1e6, "own_shard_intact": ok,
                  "value_including_gather": m.total_out_samples / (m.ms_step + gms) / 1e3}
        del bufs

    # ---- the other BASELINE configurations, device-timed the same way (N = 1: every one of them fits one GPU) ----
    configs = None
    if world == 1 and not args.no_configs and wl == "cfg4" and not args.streams:
        m.close()
        configs = []
        for w2 in ("cfg1", "cfg1x256", "cfg2", "cfg3", "cfg5"):
            try:
                s2, sec2, sc2 = default_shape(args, w2, 1)
                m2 = Measurement(torch, pkg, w2, s2, sec2, 0, 1, local_rank)
                v2 = m2.timed(max(3, min(args.steps, 5)), 3, barrier, None)
                r2, p2 = m2.rooflines()
                configs.append({"workload": w2, "description": WORKLOADS[w2][7], "value": v2, "unit": "Msamples/s",
                                "ms_per_step": m2.ms_step, "dtype": "f32" if WORKLOADS[w2][3] == "float" else "f64",
                                "streams": s2, "seconds_per_stream": sec2, "gpu_launches_per_step": m2.launches // max(3, min(args.steps, 5)),
                                "l2": m2.l2_note, "pipeline": p2,
                                "stage_kernels": r2["stage_kernels"] if r2 else None, "stage_ms": r2["stage_ms"] if r2 else None})
                m2.close()
            except Exception as exc:  # noqa: BLE001
                configs.append({"workload": w2, "error": str(exc)})

    edges = None
    if world == 1 and not args.no_configs and not args.streams:
        try:
            edges = track_edges(torch, pkg, WORKLOADS[wl][0], WORKLOADS[wl][1], WORKLOADS[wl][2], WORKLOADS[wl][6])
        except Exception as exc:  # noqa: BLE001
            edges = {"error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res, nthreads, desc = cpu_reference_run(wl, budget_s=10.0, steps=1)
            cpu = {"value": res[0][0] / res[0][1] / 1e6, "unit": "Msamples/s", "cores": nthreads, "kind": "reference",
                   "sample": desc}
        except Exception as exc:  # noqa: BLE001
            cpu = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % exc}

    if rank == 0:
        cfg = workload_config(wl, streams, seconds, world, scaling)
        cfg.update({"streams_per_gpu": streams, "stages": [{0: "halfband", 1: "dft", 2: "poly"}[s["kind"]] for s in plan["stages"]],
                    "l2": m.l2_note})
        line = {
            "metric": "output Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": m.ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32" if engine == "float" else "f64", "data": "synthetic", "config": cfg,
            "clocks": m.clocks, "gpu_launches": m.launches, "roofline": roofline, "pipeline": pipeline, "e2e": e2e,
            "cpu_baseline": cpu,
        }
        if e2e_s:
            line["e2e_streaming"] = e2e_s
        if gather:
            line["gather"] = gather
        if configs:
            line["configs"] = configs
        if edges:
            line["track_edges"] = edges
        emit_json(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
