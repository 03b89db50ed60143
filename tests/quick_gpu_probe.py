"""Scratch probe (not a test): quick device timing of a few workloads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import foo_dsp_resampler_b200 as pkg

def run(i, o, nch, nstreams, seconds, engine="float", phase=50, reps=5):
    cfg = pkg.make_config(i, o, phase=phase)
    n = int(i * seconds)
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine=engine, device=0)
    nout = b.frames_out(n)
    d_in = (torch.rand((nstreams, n, nch), device="cuda") - 0.5)
    d_out = torch.zeros((nstreams, nout, nch), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        b.process(d_in.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        b.process(d_in.data_ptr(), n, d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b.enable_timing(True)
    b.process(d_in.data_ptr(), n, d_out.data_ptr(), st)
    stage_ms = [round(x, 3) for x in b.stage_times()]
    b.enable_timing(False)
    samples = nout * nch * nstreams
    print(f"{i}->{o} {engine} nch={nch} streams={nstreams} {seconds}s: {ms:.3f} ms/step, {samples/ms/1e6:.2f} Gsamples/s, launches={b.last_launches()}, gflops={b.flops(n)/ms/1e6:.1f}, stage_ms={stage_ms}", flush=True)
    b.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(44100, 48000, 2, 1, 60)
    run(44100, 48000, 2, 64, 60)
    run(44100, 48000, 2, 512, 10)
    run(48000, 44100, 2, 512, 10)
    run(44100, 96000, 2, 64, 60)
    run(192000, 44100, 8, 8, 30, engine="double", phase=25)
    run(384000, 48000, 8, 8, 60)
