O=gpurun_out/r2as; mkdir -p $O
B200RATE_FUZZ_ENGINE=double timeout 900 python tools/gpu_fuzz.py 100 104 50 1 > $O/fuzz_f64_a.txt 2>&1; tail -n 4 $O/fuzz_f64_a.txt
B200RATE_FUZZ_ENGINE=double timeout 900 python tools/gpu_fuzz.py 300 302 40 4 > $O/fuzz_f64_b.txt 2>&1; tail -n 4 $O/fuzz_f64_b.txt
