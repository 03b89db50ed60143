set -x
O=gpurun_out/r2t; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs > $O/bench_cfg1.json 2> $O/bench_cfg1.err
grep -v "stage " $O/stage_probe.txt | tail -4; cut -c1-200 $O/bench_cfg1.json
