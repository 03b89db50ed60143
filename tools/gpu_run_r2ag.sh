O=gpurun_out/r2ag; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; echo P2; grep poly0_pair2 $O/stage_probe.txt | cut -c1-70
B200RATE_PAIR2_P1=1 python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe_p1.txt; echo P1; grep poly0_pair2 $O/stage_probe_p1.txt | cut -c1-70
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
ncu --set full --clock-control none -k regex:'poly0_pair2' -c 1 -s 4 -o $O/prof_poly -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_poly.ncu-rep > $O/ncu_full_poly0_pair2_p2.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|grid_size|block_size|inst_executed|l1tex__throughput|issue_active|registers|stalled" $O/ncu_full_poly0_pair2_p2.txt
