set -x
O=gpurun_out/r2w; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt
B200RATE_NO_PAIR_DUP=1 python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe_nodup.txt
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err
ncu --set full --clock-control none -k regex:'poly0_pair2' -c 1 -s 4 -o $O/prof_poly -f $CMD > $O/ncu_f.log 2>&1
python tools/ncu_summary.py $O/prof_poly.ncu-rep > $O/ncu_full_poly0_pair2_dup.txt 2>&1; rm -f $O/*.ncu-rep
head -8 $O/stage_probe.txt; echo; head -8 $O/stage_probe_nodup.txt; cut -c1-160 $O/cfg4x256.json
