O=gpurun_out/r2af; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; cut -c1-250 $O/stage_probe.txt
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/bench_cfg3.json 2> $O/bench_cfg3.err
ncu --set full --clock-control none -k regex:'poly0_dual' -c 1 -s 3 -o $O/prof_dual -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_dual.ncu-rep > $O/ncu_full_poly0_dual.txt 2>&1; rm -f $O/*.ncu-rep
