set -x
O=gpurun_out/r2d; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 6 $O/pytest.log
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err
(time python bench.py --steps 20 --warmup 5) > $O/cfg4_full.json 2> $O/cfg4_full.err
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
./examples/harness_multi batch 48000 44100 2 256 10 1 > $O/harness_multi.txt 2>&1
./examples/harness_multi stream 384000 48000 8 1 120 1 >> $O/harness_multi.txt 2>&1
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4x256.csv $CMD > $O/ncu_l.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'poly0_pair2' -c 1 -s 4 -o $O/prof_poly_tma -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 4 "$f"; done | tail -n 30
cut -c1-400 $O/cfg4x256.json; cat $O/stage_probe.txt | grep -v "stage "; cat $O/harness_multi.txt
