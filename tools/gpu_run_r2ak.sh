for t in 8192 4096; do echo TILE=$t; B200RATE_HALF_TILE=$t python tools/stage_probe.py 2>&1 | grep -v "stage " | grep -E "halfband_pair" | cut -c1-64; done
for t in 2048 4096 8192; do echo TILE=$t; B200RATE_HALF_TILE=$t python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | cut -c1-90; done
