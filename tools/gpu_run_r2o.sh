set -x
O=gpurun_out/r2o; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
python tools/d64_probe.py all > $O/probe_default.txt 2>&1
cat $O/probe_default.txt
