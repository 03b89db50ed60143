O=gpurun_out/r2ar; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; grep "poly0_dual" $O/stage_probe.txt | cut -c1-120
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_cfg3.json 2> $O/bench_cfg3.err; cut -c1-100 $O/bench_cfg3.json
