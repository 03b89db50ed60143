"""Scratch probe (not a test): one cfg5-shaped batch for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import foo_dsp_resampler_b200 as pkg
cfg = pkg.make_config(384000, 48000)
n, nch, ns = 384000 * 20, 8, 8
b = pkg.BatchConverter(cfg, nch, ns, n, engine="float", device=0)
nout = b.frames_out(n)
x = torch.rand((ns, n, nch), device="cuda") - 0.5
y = torch.zeros((ns, nout, nch), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    b.process(x.data_ptr(), n, y.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
