O=gpurun_out/r2ae; mkdir -p $O
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
for nl in 3 2 1; do B200RATE_DUAL_NL=$nl $CMD > $O/bench_cfg3_nl$nl.json 2> $O/bench_cfg3.err; python -c "
import json; d=json.loads(open('$O/bench_cfg3_nl$nl.json').read().strip().splitlines()[-1]); print('NL=$nl', d['value'], d['roofline']['stage_ms'])"; done
for nl in 3 2 1; do echo NL=$nl; B200RATE_DUAL_NL=$nl python tools/stage_probe.py 2>&1 | grep -v "stage " | grep "poly0_dual" | cut -c1-120; done
