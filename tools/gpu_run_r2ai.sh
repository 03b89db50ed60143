O=gpurun_out/r2ai; mkdir -p $O
for c in 6 5 4 3; do echo CTAS=$c; B200RATE_PAIR_CTAS=$c python tools/stage_probe.py 2>&1 | grep -v "stage " | grep -E "poly0_pair" | cut -c1-64; done
