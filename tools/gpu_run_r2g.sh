set -x
O=gpurun_out/r2g; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 5 $O/pytest.log
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
B200RATE_NO_DUAL_POLY=1 python tools/stage_probe.py > $O/stage_probe_nodual.txt 2>&1
python bench.py --steps 10 --warmup 3 > $O/cfg4_full.json 2> $O/cfg4_full.err
python tools/stream_probe.py > $O/stream_probe.txt 2>&1
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'poly0_dual|halfband_kernel|dft_kernel' -c 3 -s 9 -o $O/prof_cfg3 -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 3 "$f"; done | tail -n 20
grep -v 'stage ' $O/stage_probe.txt; grep double $O/stage_probe_nodual.txt; cut -c1-300 $O/cfg4_full.json; cat $O/stream_probe.txt | cut -c1-250
