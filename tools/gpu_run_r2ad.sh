O=gpurun_out/r2ad; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/bench_cfg3.json 2> $O/bench_cfg3.err; python -c "
import json; d=json.loads(open('$O/bench_cfg3.json').read().strip().splitlines()[-1]); print('group-fastest', d['value'], d['roofline']['stage_ms'])"
B200RATE_DUAL_TILE_FASTEST=1 $CMD > $O/bench_cfg3_tilefast.json 2> $O/bench_cfg3.err; python -c "
import json; d=json.loads(open('$O/bench_cfg3_tilefast.json').read().strip().splitlines()[-1]); print('tile-fastest', d['value'], d['roofline']['stage_ms'])"
ncu --set full --clock-control none -k regex:'poly0_dual' -c 1 -s 3 -o $O/prof_dual -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_dual.ncu-rep > $O/ncu_full_poly0_dual_shift.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|wavefronts|bank_conflicts|l1tex__throughput|issue_active|registers|stalled|dram__bytes" $O/ncu_full_poly0_dual_shift.txt
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; grep double $O/stage_probe.txt | cut -c1-200
