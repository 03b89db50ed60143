#!/usr/bin/env python
"""Group the per-line output of ncu_by_line.py into code regions of rate_kernels.cuh (by marker comments)."""
import subprocess
import sys

rep, cubin, mangled, demangled = sys.argv[1:5]
out = subprocess.run([sys.executable, "tools/ncu_by_line.py", rep, cubin, mangled, "2000", demangled], capture_output=True,
                     text=True).stdout.splitlines()
print(out[0])
src = open("foo_dsp_resampler_b200/csrc/rate_kernels.cuh").read().splitlines()


def find(pat, start=0):
    return next(i + 1 for i, l in enumerate(src) if i >= start and pat in l)


marks = [("arith", find("template <> struct Arith<float>"), find("template <> struct Arith<double>") + 6),
         ("cta_for/barrier", find("template <class F> RR_PROG void cta_for"), find("RR_PROG bool cta_leader")),
         ("view/ldg/async helpers", find("RR_HD long long lane_offset"), find("// Complex FFT of size M")),
         ("bfly+leaf math", find("RR_HD void sr_bfly"), find("// Leaf task `task`") - 1),
         ("leaf task (gather/store)", find("// Leaf task `task`"), find("// One butterfly of the combining pass") - 1),
         ("pass item", find("// One butterfly of the combining pass"), find("// Whole complex FFT for the CTA") - 1),
         ("cfft drivers", find("// Whole complex FFT for the CTA"), find("// Overlap-save DFT FIR stage") - 1),
         ("spec_freq_up", find("RR_HD C2<T> dft_spec_freq_up"), find("// Per-thread cache of the filter spectrum") - 1),
         ("coef cache", find("// Per-thread cache of the filter spectrum"), find("// Geometry of one work item") - 1),
         ("item geometry", find("// Geometry of one work item"), find("// Phase 0: bring the input tile") - 1),
         ("phase0 tile", find("// Phase 0: bring the input tile"), find("// `items` is a two-entry array") - 1),
         ("program head", find("// `items` is a two-entry array"), find("  // ---- phase 3:") - 1),
         ("phase3", find("  // ---- phase 3:"), find("  if (p.step == 0) {") - 1),
         ("phase4", find("  // ---- phase 4:"), find("  // ---- phases 5-6:") - 1),
         ("phase7", find("  // ---- phase 7:"), find("// Polyphase FIR stages") - 1),
         ("poly tile", find("struct Poly0Tile"), find("// Stage the input windows of the tile") - 1),
         ("poly load", find("// Stage the input windows of the tile"), find("RR_PROG void poly0_fast_compute") - 2),
         ("poly compute", find("RR_PROG void poly0_fast_compute") - 1, find("// vpoly1..3: 32.32 fixed-point") - 1)]
agg = {}
tot = [0, 0, 0]
for ln in out[2:]:
    f = ln.split()
    if len(f) < 6:
        continue
    fn, l = f[0].rsplit(":", 1)
    l = int(l)
    name = "other:" + fn
    if fn == "rate_kernels.cuh":
        for n, a, b in marks:
            if a <= l <= b:
                name = n
                break
    a = agg.setdefault(name, [0, 0, 0])
    vals = (float(f[1]), float(f[3]), float(f[5].rstrip("%")))
    for k in range(3):
        a[k] += vals[k]
        tot[k] += vals[k]
for n, (i, w, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s inst %5.1f%%  smem_wf %10.3g (%4.1f%%)  stall samples %5.1f%%" % (n, 100 * i / tot[0], w, 100 * w / max(tot[1], 1), s))
