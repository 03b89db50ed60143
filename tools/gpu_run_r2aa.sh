O=gpurun_out/r2aa; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt
B200RATE_NO_PAIR_SHIFT=1 python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe_noshift.txt
echo shift; grep -E "48000->44100|44100->96000" $O/stage_probe.txt | cut -c1-200; echo noshift; grep -E "48000->44100|44100->96000" $O/stage_probe_noshift.txt | cut -c1-200
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err; cut -c1-120 $O/cfg4x256.json
ncu --set full --clock-control none -k regex:'poly0_pair2' -c 1 -s 4 -o $O/prof_poly -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_poly.ncu-rep > $O/ncu_full_poly0_pair2_shift.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|wavefronts|bank_conflicts|l1tex__throughput|issue_active|registers" $O/ncu_full_poly0_pair2_shift.txt
