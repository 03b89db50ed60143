set -x
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or many_identical" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 4 $O/pytest.log
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/cfg4x256.json 2> $O/cfg4x256.err
B200RATE_NO_FUSED=1 $CMD > $O/cfg4x256_unfused.json 2> $O/cfg4x256_unfused.err
python bench.py --steps 3 --warmup 3 > $O/cfg4_full.json 2> $O/cfg4_full.err
python tools/stream_probe.py > $O/stream_probe.txt 2>&1
for f in $O/*.err; do tail -n 3 "$f"; done | tail -n 30
cut -c1-700 $O/cfg4x256.json; cut -c1-300 $O/cfg4x256_unfused.json; cat $O/cfg4_full.json; cat $O/stream_probe.txt
