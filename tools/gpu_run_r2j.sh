set -x
O=gpurun_out/r2j; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "double or fp64 or f64 or native" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python tools/d64_probe.py all > $O/probe_default.txt 2>&1
for GT in 64 128; do for G in 2 3 4 6; do echo "GT=$GT G=$G" >> $O/probe_sweep.txt; B200RATE_D64_GT=$GT B200RATE_D64_GROUPS=$G python tools/d64_probe.py >> $O/probe_sweep.txt 2>&1; done; done
ncu --set full --clock-control none --import-source on -k regex:'dft64' -c 2 -s 2 -o $O/prof_dft64 -f python tools/d64_probe.py > $O/ncu_f.log 2>&1
cat $O/probe_default.txt; cat $O/probe_sweep.txt
