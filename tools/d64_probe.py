"""Per-stage CUDA-event times of the fp64 engine on the two double-precision probe cases (tools/stage_probe.py)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import foo_dsp_resampler_b200 as pkg
def probe(i, o, nch, nstreams, secs, **kw):
    cfg = pkg.make_config(i, o, **kw)
    n = int(i * secs)
    b = pkg.BatchConverter(cfg, nch, nstreams, n, engine="double", device=0)
    x = (torch.rand((nstreams, n, nch), device="cuda") - 0.5)
    nout = b.frames_out(n)
    y = torch.zeros((nstreams, nout, nch), device="cuda")
    b.enable_timing(True)
    for _ in range(3):
        b.process(x.data_ptr(), n, y.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ms = b.stage_times()
    print("%d->%d %dch x%d double: %s ms, %s, %.1f Gs/s out" % (i, o, nch, nstreams, ["%.3f" % m for m in ms],
          [b.stage_kernel(k).split(' ')[0] for k in range(len(ms))], nout * nch * nstreams / sum(ms) / 1e6))
    b.close()
probe(192000, 44100, 8, 16, 20, phase=25)
probe(44100, 48000, 2, 256, 10)
if len(sys.argv) > 1:
    probe(384000, 48000, 8, 8, 20)
    probe(44100, 96000, 2, 256, 10)
    probe(48000, 44100, 2, 256, 10)
if len(sys.argv) > 1:
    probe(44100, 176400, 2, 128, 10)                   # x4 F-domain up-sampling (one DFT stage)
    probe(44100, 192000, 2, 128, 10)                   # x2, polyphase, x4 post stage
