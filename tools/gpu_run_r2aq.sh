for b in 1536 2304 3072 4096; do echo BUDGET=$b; B200RATE_DUAL_BUDGET=$b python tools/stage_probe.py 2>&1 | grep -v "stage " | grep "poly0_dual" | cut -c1-64; done
