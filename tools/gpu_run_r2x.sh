set -x
O=gpurun_out/r2x; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 12 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
python tools/lpc_probe.py > $O/lpc_probe.txt 2> $O/lpc_probe.err
ncu --set full --clock-control none --import-source on -k regex:'lpc_' -c 2 -o $O/prof_lpc -f python tools/lpc_probe.py > $O/ncu_lpc.log 2>&1
python tools/ncu_summary.py $O/prof_lpc.ncu-rep > $O/ncu_full_lpc.txt 2>&1
python tools/ncu_by_line.py $O/prof_lpc.ncu-rep > $O/ncu_by_line_lpc.txt 2>&1
rm -f $O/*.ncu-rep
head -70 $O/ncu_full_lpc.txt
