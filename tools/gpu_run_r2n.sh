set -x
O=gpurun_out/r2n; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in cfg3 cfg5; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
CMD="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv $CMD > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dft64|halfband_kernel|poly0_dual' -c 3 -s 12 -o $O/prof_cfg3 -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 2 "$f"; done | tail -n 12
grep -v 'stage ' $O/stage_probe.txt; for f in $O/bench_*.json; do cut -c1-220 $f; done
