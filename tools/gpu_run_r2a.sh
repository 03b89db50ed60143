# round 2, first measurement: new parity tests, tile-through-registers + leaf holes, fadd variant A/B, profiles
set -x
O=gpurun_out/r2a; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 6 $O/pytest.log
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err
B200RATE_VARIANT=fadd $CMD > $O/cfg4x256_fadd.json 2> $O/cfg4x256_fadd.err
B200RATE_VARIANT=fadd python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_matches or batch_cases or fixture or many_identical" > $O/pytest_fadd.log 2>&1; echo "rc=$?" >> $O/pytest_fadd.log; tail -n 4 $O/pytest_fadd.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
python tools/stream_probe.py > $O/stream_probe.txt 2>&1
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4x256.csv $CMD > $O/ncu_l.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dftp_kernel|poly0_pair' -c 2 -s 8 -o $O/prof_cfg4x256 -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 2 "$f"; done | tail -n 20
cat $O/cfg4x256.json | cut -c1-400; cat $O/cfg4x256_fadd.json | cut -c1-400; cat $O/stage_probe.txt; cat $O/stream_probe.txt
