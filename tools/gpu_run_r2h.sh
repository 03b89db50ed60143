set -x
O=gpurun_out/r2h; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 5 $O/pytest.log
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
B200RATE_NO_DFT64=1 python tools/stage_probe.py > $O/stage_probe_nodft64.txt 2>&1
for G in 3 4 5; do B200RATE_D64_GROUPS=$G python tools/stage_probe.py 2>&1 | grep double > $O/stage_probe_g$G.txt; done
CMD="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/cfg3.json 2> $O/cfg3.err && ncu --set full --clock-control none --import-source on -k regex:'dft64' -c 1 -s 4 -o $O/prof_dft64_cfg3 -f $CMD > $O/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dft64' -c 1 -s 2 -o $O/prof_dft64_up2 -f python tools/stage_probe.py > $O/ncu_g.log 2>&1
grep -v 'stage ' $O/stage_probe.txt; grep double $O/stage_probe_nodft64.txt; cat $O/stage_probe_g*.txt; cut -c1-300 $O/cfg3.json
