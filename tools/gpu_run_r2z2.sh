O=gpurun_out/r2z2; mkdir -p $O
( time python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err ) 2> $O/bench_time.txt; tail -3 $O/bench_time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z2/bench_cfg4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['track_edges']); print(d['cpu_baseline']); print(d['roofline'])
PY
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_lpc.py -x -q -k "batch_of_streams or inline or host_entry" > $O/sanitize_lpc.log 2>&1; tail -n 6 $O/sanitize_lpc.log
