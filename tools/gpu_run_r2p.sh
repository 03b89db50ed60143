set -x
O=gpurun_out/r2p; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
B200RATE_HB_OPT4=1 python tools/stage_probe.py 2>&1 | grep "^384000" > $O/stage_probe_opt4.txt
python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_cfg5.json 2> $O/bench_cfg5.err
B200RATE_HB_OPT4=1 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_cfg5_opt4.json 2> $O/bench_cfg5_opt4.err
grep -v 'stage ' $O/stage_probe.txt; cat $O/stage_probe_opt4.txt; cut -c1-200 $O/bench_cfg5.json $O/bench_cfg5_opt4.json
