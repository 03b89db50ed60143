O=gpurun_out/r2ao; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; head -2 $O/stage_probe.txt | cut -c1-200; grep "384000->48000 8ch x8 double\|double" $O/stage_probe.txt | cut -c1-120 | head
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_cfg3.json 2> $O/bench_cfg3.err; cut -c1-100 $O/bench_cfg3.json
