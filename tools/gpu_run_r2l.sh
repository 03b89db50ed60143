set -x
O=gpurun_out/r2m; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "double or native or fp64" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python tools/d64_probe.py all > $O/probe_default.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:halfband_kernel -c 2 -s 0 -o $O/prof_hb64 -f python tools/d64_probe.py > $O/ncu_f.log 2>&1
cat $O/probe_default.txt
