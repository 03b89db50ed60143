O=gpurun_out/r2aj; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; grep halfband $O/stage_probe.txt | cut -c1-200
CMD="python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/bench_cfg5.json 2> $O/bench_cfg5.err; cut -c1-100 $O/bench_cfg5.json
ncu --set full --clock-control none -k regex:'halfband_pair' -c 2 -s 6 -o $O/prof_hb -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_hb.ncu-rep > $O/ncu_full_cfg5_halfband.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|wavefronts|bank_conflicts|l1tex__throughput|issue_active|dram_throughput" $O/ncu_full_cfg5_halfband.txt
