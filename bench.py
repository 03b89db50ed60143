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


def run_reference_arm(args, rank):
    if rank != 0:
        return
    wl = args.workload
    results, nthreads, desc = cpu_reference_run(wl, budget_s=8.0, steps=args.warmup + args.steps)
    timed = results[args.warmup:]
    samples = sum(s for s, _ in timed)
    secs = sum(t for _, t in timed)
    value = samples / secs / 1e6
    in_rate, out_rate, nch, engine, phase, seconds, streams, text = WORKLOADS[wl]
    line = {
        "impl": "reference", "metric": "output Msamples/s", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if engine == "float" else "f64", "data": "synthetic",
        "config": {"workload": wl, "description": text, "in_rate": in_rate, "out_rate": out_rate, "channels": nch},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": nthreads, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--seconds", type=float, default=0.0, help="override seconds per stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
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

    in_rate, out_rate, nch, engine, phase, seconds, streams, text = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    if args.seconds:
        seconds = args.seconds
    cfg = pkg.make_config(in_rate, out_rate, phase=phase)
    st = torch.cuda.current_stream().cuda_stream
    chunked = args.workload == "cfg5"

    if not chunked:
        frames = int(round(in_rate * seconds))
        b = pkg.BatchConverter(cfg, nch, streams, frames, engine=engine, device=local_rank)
        nout = b.frames_out(frames)
        x = make_input(torch, streams, frames, nch, in_rate, seconds, 1234 + rank)
        y = torch.empty((streams, nout, nch), dtype=torch.float32, device="cuda")

        def step():
            b.process(x.data_ptr(), frames, y.data_ptr(), st)
        out_samples_per_step = nout * nch * streams
        in_bytes = x.numel() * 4
        l2_note = "inputs (%.1f GB) larger than L2" % (in_bytes / 1e9) if in_bytes > 256e6 else "L2 flushed between steps"
    else:
        # one long stream per rank-range, processed as consecutive output chunks with halo'd input windows
        frames_total = int(round(in_rate * seconds)) * world          # the whole stream, all ranks
        chunk_out = 48000 * 60                                          # 60 s of output per call
        b = pkg.BatchConverter(cfg, nch, 1, int(chunk_out * in_rate / out_rate) + 65536, engine=engine, device=local_rank)
        nout_total = b.frames_out(frames_total)
        per_rank = nout_total // world
        out_lo = rank * per_rank
        out_hi = nout_total if rank == world - 1 else out_lo + per_rank
        f0, c0 = b.input_window(frames_total, out_lo, out_hi - out_lo)
        x = make_input(torch, 1, c0, nch, in_rate, seconds * world, 99 + rank)   # this rank's halo'd window, resident
        y = torch.empty((1, out_hi - out_lo, nch), dtype=torch.float32, device="cuda")
        pieces = []
        for ob in range(out_lo, out_hi, chunk_out):
            oc = min(chunk_out, out_hi - ob)
            f, c = b.input_window(frames_total, ob, oc)
            pieces.append((ob, oc, f, c))

        def step():
            for ob, oc, f, c in pieces:
                b.process_range(x.data_ptr() + (f - f0) * nch * 4, f, c, frames_total, ob, oc,
                                y.data_ptr() + (ob - out_lo) * nch * 4, st)
        frames = c0
        out_samples_per_step = (out_hi - out_lo) * nch
        in_bytes = x.numel() * 4
        l2_note = "inputs (%.1f GB) larger than L2" % (in_bytes / 1e9)

    flush = None
    if in_bytes <= 256e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()

    b.enable_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms_acc, launches = None, 0
    barrier()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k)                          # evict L2 between timed iterations
        ev[k][0].record()
        step()
        ev[k][1].record()
        launches += b.last_launches() * (len(pieces) if chunked else 1)
        if not chunked:
            tms = b.stage_times()                   # waits for this step's events (recorded on the launching stream)
            stage_ms_acc = tms if stage_ms_acc is None else [a + c for a, c in zip(stage_ms_acc, tms)]
    barrier()
    clocks = sampler.stop()
    b.enable_timing(False)
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = out_samples_per_step * world / ms_step / 1e3            # Msamples/s, whole job

    # ---- roofline of the dominant kernel (measured live, CUDA events on the launching stream) ----
    roofline = None
    plan = b.plan()
    if stage_ms_acc:
        dom = max(range(len(stage_ms_acc)), key=lambda i: stage_ms_acc[i])
        dur_ms = stage_ms_acc[dom] / args.steps
        work = b.stage_work(frames, dom)
        peak, how = load_peaks()
        kname = b.stage_kernel(dom) or "?"
        kfamily = kname.split(" ")[0]
        # DRAM traffic of this launch from the committed ncu capture (bytes per launch unit x units launched);
        # captures exist for the cfg4 plan (48 -> 44.1 kHz fp32), other workloads report null
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tp) and engine == "float" and args.workload == "cfg4":
            with open(tp) as f:
                ent = json.load(f).get(kfamily)
            if ent:
                traffic = ent["bytes_per_unit"] * work["units"]
        # arithmetic peaks measured on this pool's B200 by tools/peak_fp.cu (profiles/fp_peaks.json): the
        # bit-faithful fp32 kernels issue un-fused multiplies and adds, packed two lanes per instruction in the
        # lane-pair kernels (FMUL2 / FADD2), scalar (FMUL / FADD) in the generic ones
        fpp = {}
        fp_path = os.path.join(ROOT, "profiles", "fp_peaks.json")
        if os.path.exists(fp_path):
            with open(fp_path) as f:
                fpp = json.load(f)
        packed = kfamily in ("dftp_kernel", "poly0_pair_kernel", "poly0_pair2_kernel", "halfband_pair_kernel")
        if engine == "float":
            alu_peak = fpp.get("fp32_fmul2_fadd2_tflops" if packed else "fp32_fmul_fadd_tflops")
            alu_how = ("measured, un-fused packed FMUL2+FADD2" if packed else "measured, un-fused FMUL+FADD") + " (profiles/fp_peaks.json)"
        else:
            alu_peak = fpp.get("fp64_dfma_tflops")
            alu_how = "measured, DFMA (profiles/fp_peaks.json)"
        ach_tflops = work["flops"] / dur_ms / 1e9
        roofline = {"bound": "hbm", "kernel": kname, "stage": dom,
                    "achieved": work["bytes"] / dur_ms / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": work["bytes"] / dur_ms / 1e6 / peak, "traffic": traffic, "peak_source": how,
                    "ms_per_launch": dur_ms, "algorithmic_bytes_per_launch": work["bytes"],
                    "share_of_step": dur_ms / ms_step,
                    "alu": {"achieved_tflops": ach_tflops, "peak_tflops": alu_peak,
                            "frac": (ach_tflops / alu_peak) if alu_peak else None, "peak_source": alu_how,
                            "note": "algorithmic flops (SURVEY.md 8d accounting) / kernel time; the kernel is bound by "
                                    "shared-memory bandwidth and latency, not by HBM (see profiles/README.md)"},
                    "stage_kernels": [b.stage_kernel(i) for i in range(len(stage_ms_acc))],
                    "stage_ms": [m / args.steps for m in stage_ms_acc]}

    # ---- end to end through the host-buffer entry point (H2D + kernels + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e and not chunked:
        try:
            sub = min(64, streams)
            tot = min(streams, 1024)
            bh = pkg.BatchConverter(cfg, nch, sub, frames, engine=engine, device=local_rank)
            h_in = torch.empty((tot, frames, nch), dtype=torch.float32, pin_memory=True)
            h_out = torch.empty((tot, nout, nch), dtype=torch.float32, pin_memory=True)
            h_in.copy_(x[:tot])
            torch.cuda.synchronize()
            bh.process_host(h_in.data_ptr(), frames, h_out.data_ptr(), tot)      # warm-up (allocates the slots)
            barrier()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                bh.process_host(h_in.data_ptr(), frames, h_out.data_ptr(), tot)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if dist:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": tot * nout * nch * world / dt / 1e6, "unit": "Msamples/s",
                   "h2d_bytes_per_step": tot * frames * nch * 4, "d2h_bytes_per_step": tot * nout * nch * 4,
                   "api": "RRX_batch_process_host (pinned host buffers, sub-batches of %d streams, 3-stream pipeline)" % sub,
                   "streams": tot, "seconds_per_step": dt}
            ok = bool(torch.equal(h_out.cuda(), y[:tot]))
            e2e["matches_device_resident_output"] = ok
            bh.close()
            del h_in, h_out
        except Exception as exc:  # noqa: BLE001
            e2e = {"value": None, "unit": "Msamples/s", "error": str(exc)}

    # ---- NCCL result gather (reported separately; the only collective on the path) ----
    gather = None
    if dist and not chunked:
        sl = y[:min(streams, 64)].contiguous()
        bufs = torch.empty((world,) + tuple(sl.shape), dtype=sl.dtype, device="cuda")
        dist.all_gather_into_tensor(bufs, sl)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(bufs, sl)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1)
        gather = {"collective": "ncclAllGather of output tiles", "bytes_per_rank": sl.numel() * 4, "ms": gms,
                  "GBps_per_rank": sl.numel() * 4 * (world - 1) / gms / 1e6}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res, nthreads, desc = cpu_reference_run(args.workload, budget_s=10.0, steps=1)
            cpu = {"value": res[0][0] / res[0][1] / 1e6, "unit": "Msamples/s", "cores": nthreads, "kind": "reference",
                   "sample": desc}
        except Exception as exc:  # noqa: BLE001
            cpu = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % exc}

    if rank == 0:
        line = {
            "metric": "output Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if engine == "float" else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": text, "in_rate": in_rate, "out_rate": out_rate,
                       "channels": nch, "streams_per_gpu": streams, "seconds_per_stream": seconds,
                       "stages": [{0: "halfband", 1: "dft", 2: "poly"}[s["kind"]] for s in plan["stages"]],
                       "l2": l2_note, "parallelism": "streams sharded over %d GPU(s), no data-path collective" % world
                       if not chunked else "one stream time-chunked over %d GPU(s) with halos" % world},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        }
        if gather:
            line["gather"] = gather
        emit_json(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
