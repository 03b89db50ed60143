"""One-off fuzz of the CUDA path against the oracle on a GPU box: python tools/gpu_fuzz.py [first_seed last_seed cases scale]
(seeded random configurations from tests/test_emulation.fuzz_cases; `scale` multiplies the signal length; B200RATE_FUZZ_ENGINE=double
fuzzes the fp64 engine at its un-cast tap with the 1e-12 contract instead of the fp32 engine bit for bit)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import os
import foo_dsp_resampler_b200 as pkg, oraclelib, signals, test_emulation
F64 = os.environ.get('B200RATE_FUZZ_ENGINE') == 'double'
worst = 0.0
bad = 0; n_ok = 0
A = [int(a) for a in sys.argv[1:]] + [100, 106, 50, 1][len(sys.argv) - 1:]
for seed in range(A[0], A[1]):
    for case in test_emulation.fuzz_cases(seed, A[2]):
        i, o, ph, bw, al, q, nch, n, chunk = case
        n *= A[3]
        try:
            cfg, ocfg = pkg.make_config(i, o, ph, bw, al, q), oraclelib.make_config(i, o, ph, bw, al, q)
            x = signals.sweep_noise(i, nch, n)
            if F64:
                ref, _ = oraclelib.resample(ocfg, x, engine="double", native=True)
                b = pkg.BatchConverter(cfg, nch, 2, n, engine="double", device=0)
                nout = b.frames_out(n)
                d_in = torch.from_numpy(np.stack([x, x])).cuda()
                d_nat = torch.zeros((2, nch, nout), dtype=torch.float64, device="cuda")
                b.process_native(d_in.data_ptr(), n, d_nat.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                got = d_nat.cpu().numpy()
                err = max(float(np.abs(got[k].T - ref).max()) if ref.size else 0.0 for k in range(2)) if ref.shape[0] == nout else 1e9
                worst = max(worst, err)
                b.close()
                if err <= 1e-12: n_ok += 1
                else: bad += 1; print("MISMATCH", case, err)
                continue
            ref, _ = oraclelib.resample(ocfg, x, engine="float")
            b = pkg.BatchConverter(cfg, nch, 2, n, engine="float", device=0)
            nout = b.frames_out(n)
            d_in = torch.from_numpy(np.stack([x, x])).cuda()
            d_out = torch.zeros((2, nout, nch), dtype=torch.float32, device="cuda")
            b.process(d_in.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()
            ok = ref.shape[0] == nout and np.array_equal(got[0], ref) and np.array_equal(got[1], ref)
            b.close()
            if not ok:
                bad += 1; print("MISMATCH", case)
            else:
                n_ok += 1
        except Exception as ex:
            bad += 1; print("EXC", case, repr(ex)[:160])
print("fuzz done ok", n_ok, "bad", bad, ("worst |err| %.2e" % worst) if F64 else "")
