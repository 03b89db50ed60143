# round 2, second measurement: fused DFT + vpoly0 kernel
set -x
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 6 $O/pytest.log
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err
B200RATE_NO_FUSED=1 $CMD > $O/cfg4x256_unfused.json 2> $O/cfg4x256_unfused.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/cfg4.json 2> $O/cfg4.err
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
python tools/stream_probe.py > $O/stream_probe.txt 2>&1
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4x256.csv $CMD > $O/ncu_l.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dft_poly_kernel' -c 1 -s 4 -o $O/prof_cfg4x256 -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 2 "$f"; done | tail -n 20
cut -c1-600 $O/cfg4x256.json; cut -c1-300 $O/cfg4x256_unfused.json; cut -c1-300 $O/cfg4.json; cat $O/stage_probe.txt; cat $O/stream_probe.txt
