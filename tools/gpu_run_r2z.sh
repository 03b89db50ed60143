set -x
O=gpurun_out/r2z; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
python tools/lpc_probe.py > $O/lpc_probe.txt 2> $O/lpc_probe.err; cat $O/lpc_probe.txt | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lpc_' -c 8 --csv --log-file $O/lpc_launches.csv python tools/lpc_probe.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'lpc_' -c 6 -o $O/prof_lpc -f python tools/lpc_probe.py > $O/ncu_lpc.log 2>&1
python tools/ncu_summary.py $O/prof_lpc.ncu-rep > $O/ncu_full_lpc.txt 2>&1
rm -f $O/*.ncu-rep
grep -E "^kernel|gpu__time|issue_active|l1tex__throughput|registers|no_inst|pipe_fp64" $O/ncu_full_lpc.txt
