O=gpurun_out/r2au; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 2 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 3 --warmup 3 --no-configs > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
