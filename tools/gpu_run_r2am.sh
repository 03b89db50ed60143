for t in 8192 4096; do echo TILE_G=$t; B200RATE_HALF_TILE_G=$t python tools/stage_probe.py 2>&1 | grep -v "stage " | grep -E "halfband_kernel" | cut -c1-64; done
for t in 4096 8192; do B200RATE_HALF_TILE_G=$t python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | cut -c1-90; done
