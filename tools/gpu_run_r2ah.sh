O=gpurun_out/r2ah; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe.txt; echo P2; grep "poly0_pair_kernel" $O/stage_probe.txt | cut -c1-70
B200RATE_PAIR2_P1=1 python tools/stage_probe.py 2>&1 | grep -v "stage " > $O/stage_probe_p1.txt; echo P1; grep "poly0_pair_kernel" $O/stage_probe_p1.txt | cut -c1-70
