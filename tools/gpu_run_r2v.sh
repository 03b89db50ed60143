set -x
O=gpurun_out/r2v; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python tools/d64_probe.py all > $O/probe.txt 2>&1
B200RATE_NO_DFT64=1 python tools/d64_probe.py all 2>&1 | grep "176400\|192000->\|->192000" > $O/probe_nodft64.txt
python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/cfg4x256.json 2> $O/cfg4x256.err
cat $O/probe.txt $O/probe_nodft64.txt; cut -c1-160 $O/cfg4x256.json
